"""ctypes mirrors of the model-level structs of include/videocad_b200.h (vc_vit_*, vc_seq_*)."""
from __future__ import annotations

import ctypes as C

from . import lib as L

vp, i64, i32, f32, u32, u64, sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_uint32, C.c_uint64, C.c_size_t

VIT_DEPTH = 6
VIT_DIM = 512
VIT_HEADS = 16
VIT_DHEAD = 64
VIT_MLP = 512
PATCH = 32


class Linear(C.Structure):
    _fields_ = [("w", vp), ("b", vp), ("w_hi", vp), ("w_lo", vp), ("dw", vp), ("db", vp)]


class Norm(C.Structure):
    _fields_ = [("w", vp), ("b", vp), ("dw", vp), ("db", vp)]


class VitLayer(C.Structure):
    _fields_ = [("ln1", Norm), ("qkv", Linear), ("out", Linear), ("ln2", Norm), ("fc1", Linear), ("fc2", Linear)]


class VitWeights(C.Structure):
    _fields_ = [("pos", vp), ("cls", vp), ("dpos", vp), ("dcls", vp), ("pe_ln1", Norm), ("pe", Linear), ("pe_ln2", Norm),
                ("layer", VitLayer * VIT_DEPTH), ("norm", Norm)]


class VitCall(C.Structure):
    _fields_ = [("w", C.POINTER(VitWeights)), ("img", vp), ("F", i32), ("S", i32), ("dropout_p", f32), ("training", i32),
                ("seed", u64), ("site_base", u32), ("seed_dev", vp), ("passes", i32), ("aux_streams", i32), ("ws", vp), ("ws_bytes", sz),
                ("cls_out", vp)]


class DecLayer(C.Structure):
    _fields_ = [("sa_in", Linear), ("sa_out", Linear), ("ca_in", Linear), ("ca_out", Linear), ("lin1", Linear),
                ("lin2", Linear), ("n1", Norm), ("n2", Norm), ("n3", Norm)]


class SeqWeights(C.Structure):
    _fields_ = [("embed_state", Linear), ("embed_image", Linear), ("image_proj", Linear), ("head_params", Linear),
                ("embed_multiview", Linear),
                ("embed_action_w", vp), ("embed_action_b", vp), ("d_embed_action_w", vp), ("d_embed_action_b", vp),
                ("head_cmd_w", vp), ("head_cmd_b", vp), ("d_head_cmd_w", vp), ("d_head_cmd_b", vp),
                ("timestep_emb", vp), ("d_timestep_emb", vp),
                ("layers", C.POINTER(DecLayer)), ("num_layers", i32)]


class SeqCall(C.Structure):
    _fields_ = [("w", C.POINTER(SeqWeights)), ("B", i32), ("T", i32), ("H", i32), ("nhead", i32), ("Ff", i32), ("window", i32),
                ("past_actions", i32), ("past_states", i32), ("act_dim", i32), ("num_cmd", i32), ("num_param_out", i32),
                ("state_cls", vp), ("cad_cls", vp), ("actions", vp), ("num_views", i32), ("mv_cls", vp), ("dropout_p", f32), ("training", i32), ("seed", u64),
                ("site_base", u32), ("seed_dev", vp), ("passes", i32), ("ws", vp), ("ws_bytes", sz), ("cmds", vp), ("params", vp)]


L.EXTRA_PROTOS.update({
    "vc_vit_workspace_bytes": ([i32, i32], sz),
    "vc_vit_scratch_bytes": ([i32, i32], sz),
    "vc_vit_forward": ([C.POINTER(VitCall), vp], i32),
    "vc_vit_backward": ([C.POINTER(VitCall), vp, vp, sz, vp], i32),
    "vc_vit_backward_layers": ([C.POINTER(VitCall), vp, vp, sz, i32, i32, vp], i32),
    "vc_seq_workspace_bytes": ([i32, i32, i32, i32, i32, i32, i32, i32], sz),
    "vc_seq_scratch_bytes": ([i32, i32, i32, i32, i32, i32], sz),
    "vc_seq_forward": ([C.POINTER(SeqCall), vp], i32),
    "vc_seq_decode_scratch_bytes": ([i32, i32, i32, i32], sz),
    "vc_seq_decode_step": ([C.POINTER(SeqCall), i32, vp, vp, sz, vp, vp, vp], i32),
    "vc_seq_decode_dev_supported": ([C.POINTER(SeqCall)], i32),
    "vc_seq_decode_dev_scratch_bytes": ([i32, i32, i32, i32], sz),
    "vc_seq_decode_step_dev": ([C.POINTER(SeqCall), vp, vp, vp, sz, vp, vp, vp], i32),
    "vc_seq_backward": ([C.POINTER(SeqCall), vp, vp, vp, vp, vp, vp, sz, vp], i32),
})
