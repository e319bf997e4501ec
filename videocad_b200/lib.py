"""ctypes binding of libvideocad_b200.so (the C ABI declared in include/videocad_b200.h).

This is the reference-side stub described in INTEGRATION.md: plain pointers, sizes and a cudaStream_t cross the
boundary; torch is used only for device memory and the current stream.  There is NO fallback: if the shared
library is missing or was not built from the CUDA sources, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VIDEOCAD_B200_LIB", os.path.join(_HERE, "libvideocad_b200.so"))

ACT_NONE, ACT_GELU, ACT_RELU, ACT_TANH, ACT_GELU_DSTORE, ACT_MUL_AUX = 0, 1, 2, 3, 4, 5
MASK_NONE, MASK_CAUSAL, MASK_WINDOW = 0, 1, 2

vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float


class Drop(C.Structure):
    _fields_ = [("p", C.c_float), ("site", C.c_uint32), ("seed", C.c_uint64), ("seed_ptr", C.c_void_p)]


class LossCfg(C.Structure):
    """mirror of vc_loss_cfg"""
    _fields_ = [("R", C.c_int), ("NC", C.c_int), ("NP", C.c_int), ("NV", C.c_int), ("cmd_w", C.c_float * 16),
                ("tolerance", C.c_int * 8), ("param_to_label", C.c_int * 8)]


class MetricsCfg(C.Structure):
    """mirror of vc_metrics_cfg"""
    _fields_ = [("above", C.c_int * 8), ("tolerance", C.c_int * 8), ("abs_tolerance", C.c_int), ("topk", C.c_int)]


# counter layout of vc_loss_metrics (VC_METRIC_* in include/videocad_b200.h)
METRIC_CORRECT, METRIC_TOTAL, METRIC_CMD_CORRECTS, METRIC_CMD_COUNTS = 0, 1, 2, 18
METRIC_PARAM_CORRECTS, METRIC_PARAM_COUNTS = 34, 42
METRIC_CMD_CORRECT_TOPK, METRIC_CMD_COUNTS_TOPK, METRIC_PARAM_CORRECT_TOPK, METRIC_PARAM_COUNTS_TOPK, METRIC_COUNT = 50, 51, 52, 53, 54


class AdamTensor(C.Structure):
    """mirror of vc_adam_tensor"""
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_int64), ("lr", C.c_float)]


class GemmDesc(C.Structure):
    _fields_ = [
        ("a_hi", vp), ("a_lo", vp), ("lda", i64), ("a_mn_major", i32),
        ("b_hi", vp), ("b_lo", vp), ("ldb", i64), ("b_mn_major", i32),
        ("M", i32), ("N", i32), ("K", i32), ("passes", i32), ("splitk", i32),
        ("bias", vp),
        ("rowadd", vp), ("ld_rowadd", i64), ("rowadd_div", i32), ("rowadd_mod", i32),
        ("preact", vp), ("ld_preact", i64),
        ("act", i32),
        ("drop", Drop),
        ("residual", vp), ("ld_res", i64),
        ("out_f32", vp), ("ldo", i64),
        ("out_hi", vp), ("out_lo", vp), ("ldo_split", i64),
        ("act_backward", i32), ("act_aux", vp), ("ld_act_aux", i64), ("act_aux_hi", vp), ("ld_act_aux_hi", i64), ("colsum", vp),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("q", vp), ("k", vp), ("v", vp), ("ldq", i64), ("ldk", i64), ("ldv", i64),
        ("q_hi", vp), ("q_lo", vp), ("k_hi", vp), ("k_lo", vp), ("v_hi", vp), ("v_lo", vp),
        ("B", i32), ("Tq", i32), ("Tk", i32), ("nh", i32), ("d", i32),
        ("mask", i32), ("window", i32),
        ("scale", C.c_float),
        ("drop", Drop),
        ("bsq", i64), ("bsk", i64), ("bsv", i64),
    ]


_PROTOS = {
    "vc_version": ([], i32),
    "vc_is_cuda_build": ([], i32),
    "vc_launch_count": ([], C.c_longlong),
    "vc_launch_count_reset": ([], None),
    "vc_gemm_pair_launch_count": ([], C.c_longlong),
    "vc_side_streams_enable": ([i32], None),
    "vc_gemm_profile": ([i32], None),
    "vc_gemm_profile_read": ([C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)], i32),
    "vc_gemm_profile_read_min": ([C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)], i32),
    "vc_gemm_profile_dump": ([C.c_char_p], i32),
    "vc_abi_sizeof": ([i32], C.c_size_t),
    "vc_gemm_desc_init": ([C.POINTER(GemmDesc)], None),
    "vc_gemm_pair_force_tile": ([i32], None),
    "vc_gemm": ([C.POINTER(GemmDesc), vp], i32),
    "vc_split_f32": ([vp, i64, i64, i64, vp, vp, i64, vp], i32),
    "vc_split_many": ([vp, i32, i64, vp], i32),
    "vc_layernorm_fwd": ([vp, i64, i64, i32, vp, vp, f32, vp, i64, vp, vp, i64, vp, vp, vp], i32),
    "vc_layernorm_bwd": ([vp, i64, vp, i64, vp, vp, vp, i64, i32, vp, i64, vp, i64, vp, vp, vp], i32),
    "vc_layernorm_bwd_fused": ([vp, i64, vp, i64, vp, vp, vp, i64, i32, vp, i64, vp, i64, vp, vp, Drop, vp, vp, i64, vp, vp], i32),
    "vc_attention_bwd_split": ([C.POINTER(AttnDesc), vp, vp, i64, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, i64, vp], i32),
    "vc_attention_bwd_split_bias": ([C.POINTER(AttnDesc), vp, vp, i64, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp], i32),
    "vc_attention_small_enable": ([i32], None),
    "vc_patch_layernorm_fwd": ([vp, i32, i32, vp, vp, f32, vp, vp, vp, vp, vp], i32),
    "vc_patch_layernorm_bwd_params": ([vp, i32, i32, vp, vp, vp, vp, vp, vp], i32),
    "vc_vit_assemble_fwd": ([vp, i32, i32, i32, vp, vp, Drop, vp, vp], i32),
    "vc_vit_assemble_bwd": ([vp, i32, i32, i32, Drop, vp, vp, vp, vp], i32),
    "vc_attention_fwd": ([C.POINTER(AttnDesc), vp, vp, i64, vp, vp], i32),
    "vc_attention_bwd": ([C.POINTER(AttnDesc), vp, vp, i64, vp, vp, i64, vp, i64, vp, i64, vp, i64, vp], i32),
    "vc_act_dropout_bwd": ([vp, i64, i64, i32, i32, vp, i64, vp, i64, Drop, vp, i64, vp, vp, i64, vp, vp], i32),
    "vc_row_reduce_mod": ([vp, i64, i64, i32, i32, i32, vp, vp], i32),
    "vc_broadcast_rows": ([vp, i64, i64, i32, i32, vp, i64, vp, vp, i64, vp], i32),
    "vc_embed_action_fwd": ([vp, i64, i32, i32, vp, vp, vp, i32, vp, vp, vp, vp], i32),
    "vc_embed_action_bwd": ([vp, vp, vp, i64, i32, i32, i32, vp, vp, vp, vp], i32),
    "vc_loss_workspace_floats": ([i32, i32], C.c_size_t),
    "vc_loss_forward": ([vp, vp, vp, vp, vp, vp, vp], i32),
    "vc_loss_backward": ([vp, vp, vp, vp, vp, vp, vp, vp, vp], i32),
    "vc_loss_metrics": ([vp, vp, vp, vp, i32, vp, vp], i32),
    "vc_clip_adam_scratch_floats": ([], C.c_size_t),
    "vc_clip_adam_step": ([vp, i32, C.c_double, C.c_double, C.c_double, C.c_double, i64, vp, vp, vp], i32),
    "vc_head_small_fwd": ([vp, i64, i32, vp, vp, i32, vp, vp], i32),
    "vc_head_small_bwd": ([vp, vp, i64, i32, vp, i32, vp, i32, vp, vp, vp], i32),
    "vc_linear_rows_fwd": ([vp, vp, vp, i64, i32, vp, vp, i32, i32, i32, vp, i64, vp, i64, vp, vp, i64, vp], i32),
    "vc_frames_u8_normalize": ([vp, i64, f32, f32, vp, vp], i32),
    "vc_frames_rgb_u8_ingest": ([vp, i64, i32, i32, i32, i32, vp, vp, i32, vp, vp, i32, vp, f32, f32, vp, vp], i32),
    "vc_add_f32": ([vp, vp, vp, i64, vp], i32),
    "vc_zero_f32": ([vp, i64, vp], i32),
    "vc_dropout_mask_debug": ([Drop, i64, vp, vp], i32),
}

# model-level entry points are registered by videocad_b200.model_abi (kept next to their struct definitions)
EXTRA_PROTOS: dict = {}

_lib: Optional[C.CDLL] = None


def exported_symbols() -> list:
    return sorted(list(_PROTOS.keys()) + list(EXTRA_PROTOS.keys()) + ["vc_last_error"])


def load(path: Optional[str] = None, require_cuda_build: bool = True) -> C.CDLL:
    """Load the shared library (once).  Raises if it is missing -- there is no CPU/eager fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    from . import model_abi  # noqa: F401  (registers the model-level prototypes in EXTRA_PROTOS)

    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"videocad_b200: native library not found at {p}; build it with `python -m videocad_b200.build` "
            "(nvcc, sm_100a).  There is no fallback path.")
    lib = C.CDLL(p)
    lib.vc_last_error.argtypes = []
    lib.vc_last_error.restype = C.c_char_p
    for name, (args, res) in list(_PROTOS.items()) + list(EXTRA_PROTOS.items()):
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = args
        fn.restype = res
    if require_cuda_build and lib.vc_is_cuda_build() != 1:
        raise RuntimeError(f"videocad_b200: {p} is not a CUDA build of the library")
    if path is None:
        _lib = lib
    return lib


def check(rc: int, lib: Optional[C.CDLL] = None) -> None:
    if rc != 0:
        l = lib or load()
        raise RuntimeError("videocad_b200: " + l.vc_last_error().decode("utf-8", "replace"))


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def cur_stream():
    return torch.cuda.current_stream().cuda_stream


def make_drop(p: float = 0.0, site: int = 0, seed: int = 0, seed_ptr=None) -> Drop:
    return Drop(float(p), int(site) & 0xFFFFFFFF, int(seed) & 0xFFFFFFFFFFFFFFFF, seed_ptr)


# ------------------------------------------------------------------------------------------------
# small tensor-level helpers (used by the unit parity tests; the model path goes through model_abi)
# ------------------------------------------------------------------------------------------------
def split(x: torch.Tensor):
    """fp32 [rows, cols] (contiguous) -> (hi, lo) bf16 tensors via the CUDA kernel."""
    assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()
    hi = torch.empty_like(x, dtype=torch.bfloat16)
    lo = torch.empty_like(x, dtype=torch.bfloat16)
    check(load().vc_split_f32(ptr(x), x.shape[1], x.shape[0], x.shape[1], ptr(hi), ptr(lo), x.shape[1], cur_stream()))
    return hi, lo


def gemm(a, b, M, N, K, *, a_mn=False, b_mn=False, passes=3, splitk=1, bias=None, rowadd=None, rowadd_div=1,
         rowadd_mod=1, preact=None, act=ACT_NONE, drop=None, residual=None, out_f32=None, out_split=None, act_backward=False,
         act_aux=None, act_aux_hi=None, colsum=None):
    """a, b: (hi, lo) tuples of contiguous bf16 matrices.  See vc_gemm_desc."""
    lib = load()
    d = GemmDesc()
    lib.vc_gemm_desc_init(C.byref(d))
    d.a_hi, d.a_lo, d.lda, d.a_mn_major = ptr(a[0]), ptr(a[1]), a[0].stride(0), int(a_mn)
    d.b_hi, d.b_lo, d.ldb, d.b_mn_major = ptr(b[0]), ptr(b[1]), b[0].stride(0), int(b_mn)
    d.M, d.N, d.K, d.passes, d.splitk = M, N, K, passes, splitk
    d.bias = ptr(bias)
    if rowadd is not None:
        d.rowadd, d.ld_rowadd, d.rowadd_div, d.rowadd_mod = ptr(rowadd), rowadd.stride(0), rowadd_div, rowadd_mod
    if preact is not None:
        d.preact, d.ld_preact = ptr(preact), preact.stride(0)
    d.act = act
    d.drop = drop if drop is not None else make_drop()
    if residual is not None:
        d.residual, d.ld_res = ptr(residual), residual.stride(0)
    if out_f32 is not None:
        d.out_f32, d.ldo = ptr(out_f32), out_f32.stride(0)
    if out_split is not None:
        d.out_hi, d.out_lo, d.ldo_split = ptr(out_split[0]), ptr(out_split[1]), out_split[0].stride(0)
    d.act_backward = int(act_backward)
    if act_aux is not None:
        d.act_aux, d.ld_act_aux = ptr(act_aux), act_aux.stride(0)
    if act_aux_hi is not None:
        d.act_aux_hi, d.ld_act_aux_hi = ptr(act_aux_hi), act_aux_hi.stride(0)
    d.colsum = ptr(colsum)
    check(lib.vc_gemm(C.byref(d), cur_stream()), lib)
