"""oracle/ -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under this package is imported, linked or executed by the product path
(`videocad_b200`).  Allowed consumers: `tests/`, `__graft_entry__.smoke()` (as the checker)
and `bench.py`'s `cpu_baseline` / `--impl reference` legs.

Contents
  shims/vit_pytorch, shims/timm  stand-ins for the two packages the reference imports but this
                                 image lacks (vit-pytorch is where the ViT arithmetic lives; see
                                 the header of shims/vit_pytorch/__init__.py: PARITY UNPINNED for
                                 that third-party dependency).
  reference_model.py             imports the UNMODIFIED reference from /root/reference with the
                                 shims on sys.path (works only in the build container).
  torch_oracle.py                self-contained functional restatement (plain torch fp32/fp64 ops on
                                 a state_dict); travels to the GPU box.  Pinned against the real
                                 reference by tests/test_oracle_vs_reference.py (here) and by the
                                 committed fixtures under tests/golden/ (everywhere).
  make_golden.py                 script that generated tests/golden/*.npz from the real reference.
  train_port.py                  CPU port of trainer._process_batch (loss, clip, Adam) used as the
                                 timed CPU baseline ("kind": "port").
  csrc/                          plain-C++ restatements of each CUDA kernel, used to build
                                 oracle/_build/libvc_emu.so: the same host orchestration as the
                                 product, with C loops instead of kernels, so that the host logic
                                 can be checked without a GPU.
"""
