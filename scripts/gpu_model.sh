#!/bin/bash
mkdir -p gpurun_out
echo "=== kernels (fixed tests)"; timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "gemm_epilogue or attention" -p no:cacheprovider 2>&1 | tail -15
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15
echo "=== model"; timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | tail -60 | tee gpurun_out/model_tests.log
