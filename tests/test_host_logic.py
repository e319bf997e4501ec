"""CPU tests of the host side: the C ABI surface, the ctypes binding, the drop-in module's contract, and the host
orchestration (model_vit.cpp / model_seq.cpp) executed against the CPU emulation of the kernels
(oracle/_build/libvc_emu.so -- test infrastructure, never loaded by the product path)."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from oracle import build_emu
from oracle import torch_oracle as to
from videocad_b200 import AutoRegressiveTransformer, ModelFactory, ModelType
from videocad_b200 import lib as L
from videocad_b200 import model_abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def emu():
    return L.load(build_emu.build(), require_cuda_build=False)


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "videocad_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vc_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_bound_and_exported(emu):
    syms = header_symbols()
    assert len(syms) >= 35
    assert set(syms) == set(L.exported_symbols()), set(syms) ^ set(L.exported_symbols())
    for s in syms:
        assert hasattr(emu, s), s
    native = os.path.join(ROOT, "videocad_b200", "libvideocad_b200.so")
    if not os.path.exists(native):
        from videocad_b200.build import build

        build()
    out = subprocess.run(["nm", "-D", "--defined-only", native], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (vc_[a-z0-9_]+)", out))
    assert set(syms) <= exported, set(syms) - exported


def test_ctypes_struct_layout_matches_c(emu):
    for which, struct in enumerate([L.Drop, L.GemmDesc, L.AttnDesc, A.Linear, A.Norm, A.VitWeights, A.VitCall, A.DecLayer,
                                    A.SeqWeights, A.SeqCall]):
        assert emu.vc_abi_sizeof(which) == C.sizeof(struct), (which, struct.__name__)


def test_product_loader_refuses_non_cuda_library():
    with pytest.raises(RuntimeError):
        L.load(build_emu.build(), require_cuda_build=True)
    with pytest.raises(RuntimeError):
        L.load("/nonexistent/libvideocad_b200.so")


def test_module_contract_and_errors():
    cfg = dict(hidden_size=128, nhead=4, num_decoder_layers=2, dim_feedforward=128, enable_past_actions=True,
               enable_past_states=True, enable_timestep_embedding=True, model_name="autoregressive", normalize=True,
               network_layers=[1, 2], enable_random=True)  # unknown keys must be tolerated, as in the reference
    m, mt = ModelFactory().create_model("whatever", dict(cfg, state_dim=1644, act_dim=7, encoder="vit"), "cpu")
    assert mt == ModelType.MULTI_CLASSES
    shapes = to.param_shapes(cfg)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == shapes
    # checkpoints saved from DDP / torch.compile wrappers load through the factory
    sd = {"module._orig_mod." + k: v for k, v in to.seeded_state_dict(cfg, 2).items()}
    sd["transformer.wte.weight"] = torch.zeros(1, 128)  # dead GPT-2 key of a reference checkpoint: ignored (strict=False)
    m2, _ = ModelFactory().create_model("x", dict(cfg, state_dim=1644, act_dim=7, encoder="vit"), "cpu", state_dict=sd)
    assert torch.equal(dict(m2.named_weights())["embed_action.weight"].data, sd["module._orig_mod.embed_action.weight"])
    # one flat parameter for the sequence transformer, two (split by depth) per image encoder -- not the reference's ~320 tensors
    # (registered bottom-up, encoders interleaved, sequence transformer last: DDP's bucket order, see model._ParamOrder)
    vit_parts = [(n,) for n in m2.state_embedding_model._spec.part_totals]
    assert len(vit_parts) == 2
    assert [tuple(p.shape) for p in m2.parameters()] == [vit_parts[0]] * 2 + [vit_parts[1]] * 2 + [(m2._spec.total,)]
    assert [tuple(p.shape) for p in m2.state_embedding_model.parameters()] == vit_parts  # trainer.py:244-245 (frozen lr groups)
    assert list(m.cad_embedding_model.parameters()) and list(m.state_embedding_model.parameters())
    with pytest.raises(ValueError):
        AutoRegressiveTransformer(state_dim=1644, act_dim=7, hidden_size=128, encoder="resnet")
    with pytest.raises(AssertionError):
        AutoRegressiveTransformer(state_dim=1644, act_dim=7, hidden_size=128, window_size=0)
    inp = to.model_inputs_from_batch(to.synthetic_batch(1, 3, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(inp)  # CPU tensors and no injected test library: the product path refuses


def build_emu_model(emu, cfg, seed=0, dropout=0.1):
    m = AutoRegressiveTransformer(state_dim=1644, act_dim=7, dropout=dropout, vit_dropout=dropout, encoder="vit", **cfg)
    sd = to.seeded_state_dict(cfg, seed)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not missing and not unexpected
    m._use_library_for_tests(emu)
    return m, sd


MODES = [dict(enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True),
         dict(enable_past_actions=False, enable_past_states=True, enable_timestep_embedding=True),
         dict(enable_past_actions=False, enable_past_states=False),
         dict(enable_past_actions=True, enable_past_states=False, enable_timestep_embedding=True)]


@pytest.mark.parametrize("mode", MODES)
def test_orchestration_forward_backward_vs_fp64_oracle(emu, mode):
    cfg = dict(hidden_size=128, nhead=4, num_decoder_layers=2, dim_feedforward=256, window_size=2, **mode)
    m, sd = build_emu_model(emu, cfg)
    m.eval()
    inp = to.model_inputs_from_batch(to.synthetic_batch(2, 4, 64))
    cmds, params = m(inp)
    sdd = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    oc, op = to.forward(sdd, cfg, {k: v.double() for k, v in inp.items()})
    assert (cmds.double() - oc).abs().max() < 1e-4 and (params.double() - op).abs().max() < 1e-4
    g = torch.Generator().manual_seed(5)
    wc, wp = torch.randn(cmds.shape, generator=g), torch.randn(params.shape, generator=g) * 0.05
    ((cmds * wc).sum() + (params * wp).sum()).backward()
    ((oc * wc.double()).sum() + (op * wp.double()).sum()).backward()
    for name, p in m.named_weights():
        ref = sdd[name].grad
        if p.grad is None:
            assert ref is None or ref.abs().max() == 0, f"{name}: missing gradient"
            continue
        ref = ref if ref is not None else torch.zeros_like(p.grad, dtype=torch.double)
        err = (p.grad.double() - ref).abs().max().item() / (ref.abs().max().item() + 1e-6)
        assert err < 2e-4, f"{name}: rel grad err {err:.3e}"


@pytest.mark.parametrize("mode", [MODES[0], MODES[2], MODES[3]])
def test_orchestration_multiview_vs_fp64_oracle(emu, mode):
    cfg = dict(hidden_size=128, nhead=4, num_decoder_layers=1, dim_feedforward=128, window_size=2, num_views=2, **mode)
    m, sd = build_emu_model(emu, cfg, dropout=0.0)
    m.eval()
    inp = to.model_inputs_from_batch(to.synthetic_batch(2, 4, 64))
    inp["multiview_images"] = to.synthetic_views(2, 2, 64)
    cmds, params = m(inp)
    sdd = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    oc, op = to.forward(sdd, cfg, {k: v.double() for k, v in inp.items()})
    assert (cmds.double() - oc).abs().max() < 1e-4 and (params.double() - op).abs().max() < 1e-4
    (cmds.sum() + params.sum() * 0.01).backward()
    (oc.sum() + op.sum() * 0.01).backward()
    for name, p in m.named_weights():
        ref = sdd[name].grad
        if p.grad is None:
            assert ref is None or ref.abs().max() == 0, name
            continue
        ref = ref if ref is not None else torch.zeros_like(p.grad, dtype=torch.double)  # unused weight: zero slice of the flat gradient
        assert (p.grad.double() - ref).abs().max() < 2e-4 * ref.abs().max() + 1e-9, name
    with pytest.raises(ValueError):
        m({k: v for k, v in inp.items() if k != "multiview_images"})
    with pytest.raises(ValueError):
        AutoRegressiveTransformer(state_dim=1644, act_dim=7, hidden_size=128, num_views=2, enable_past_states=True)


def test_orchestration_matches_reference_golden(emu):
    z = np.load(os.path.join(GOLDEN, "c0_states_actions.npz"))
    meta = json.loads(bytes(z["meta_json"]).decode())
    m, _ = build_emu_model(emu, meta["cfg"])
    m.eval()
    inp = to.model_inputs_from_batch(to.synthetic_batch(meta["B"], meta["T"] + 1, meta["S"], seed=meta["batch_seed"]))
    with torch.no_grad():
        cmds, params = m(inp)
    assert (cmds - torch.from_numpy(z["cmds"])).abs().max() < 2e-4
    assert (params - torch.from_numpy(z["params"])).abs().max() < 2e-4


def test_training_mode_dropout_bookkeeping(emu):
    """Forward and backward regenerate identical dropout masks: with dropout on, the analytic gradient must match a
    finite difference of the (fixed-seed) stochastic forward."""
    cfg = dict(hidden_size=128, nhead=4, num_decoder_layers=1, dim_feedforward=128, window_size=2,
               enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)
    m, _ = build_emu_model(emu, cfg, dropout=0.2)
    m.train()
    inp = to.model_inputs_from_batch(to.synthetic_batch(1, 3, 64))
    m._draw_seed = lambda: 4242  # same masks on every call
    c1, p1 = m(inp)
    c2, p2 = m(inp)
    assert torch.equal(c1, c2) and torch.equal(p1, p2)
    m.eval()
    ce, _ = m(inp)
    assert (ce - c1).abs().max() > 1e-3  # dropout really is active in training mode
    m.train()
    g = torch.Generator().manual_seed(1)
    wc = torch.randn(c1.shape, generator=g)
    (c1 * wc).sum().backward()
    W = dict(m.named_weights())
    for key, idx in (("embed_action.bias", 5), ("state_embedding_model.transformer.layers.3.1.net.1.bias", 17),
                     ("transformer_decoder.layers.0.norm2.weight", 3), ("cad_embedding_model.cls_token", 100)):
        base = W[key].data.view(-1)  # a view of the owning segment's flat parameter
        analytic = W[key].grad.view(-1)[idx].item()
        eps = 1e-2
        with torch.no_grad():
            old = base[idx].item()
            base[idx] = old + eps
            lp = (m(inp)[0] * wc).sum().item()
            base[idx] = old - eps
            lm = (m(inp)[0] * wc).sum().item()
            base[idx] = old
        fd = (lp - lm) / (2 * eps)
        assert abs(fd - analytic) < 2e-2 * max(1.0, abs(analytic)), (fd, analytic)


def _mask(emu, p, site, seed, n):
    out = np.empty(n, dtype=np.float32)
    L.check(emu.vc_dropout_mask_debug(L.make_drop(p, site, seed), n, out.ctypes.data, None), emu)
    return out


def test_dropout_mask_statistics(emu):
    """The dropout generator (csrc/dropout_rng.h, shared by the CUDA kernels and the emulation): keep-rate, uniformity of the
    underlying 16-bit draws (through the keep-rate at several p), independence across neighbouring elements, tensor rows /
    columns, sites and seeds.  Thresholds are ~5 sigma for the sample sizes used."""
    n = 1 << 21
    for p in (0.1, 0.2, 0.5):
        for site, seed in ((0, 0), (3, 1234), (40, 2 ** 40 + 7)):
            m = _mask(emu, p, site, seed, n)
            assert set(np.unique(m)) <= {0.0, np.float32(1.0 / (1.0 - p))}
            keep = (m != 0).astype(np.float64)
            assert abs(keep.mean() - (1.0 - p)) < 5 * np.sqrt(p * (1 - p) / n) + 1e-5, (p, site, seed, keep.mean())
            for lag in (1, 2, 3, 4, 50, 64, 512, 3072):
                c = np.corrcoef(keep[:-lag], keep[lag:])[0, 1]
                assert abs(c) < 5 / np.sqrt(n), (p, site, seed, lag, c)
    a = (_mask(emu, 0.1, 5, 77, n) != 0).astype(np.float64)
    b = (_mask(emu, 0.1, 6, 77, n) != 0).astype(np.float64)   # another site
    c = (_mask(emu, 0.1, 5, 78, n) != 0).astype(np.float64)   # the next seed
    assert abs(np.corrcoef(a, b)[0, 1]) < 5 / np.sqrt(n)
    assert abs(np.corrcoef(a, c)[0, 1]) < 5 / np.sqrt(n)
    assert (a != b).mean() > 0.15 and (a != c).mean() > 0.15
    k = a.reshape(-1, 512)  # [4096, 512]: column and row keep-rates scatter as a binomial would
    assert 0.85 < k.mean(0).std() / np.sqrt(0.09 / k.shape[0]) < 1.15
    assert 0.93 < k.mean(1).std() / np.sqrt(0.09 / 512) < 1.07


def _torch_ingest(u8):
    """transforms.ToTensor() + transforms.Normalize([0.5], [0.5]) on an already grey, already sized uint8 image (main.py:103-110)"""
    return u8.to(torch.float32).div(255).sub_(0.5).div_(0.5)


def test_frame_ingestion_u8_bit_exact_and_model_equivalence(emu):
    """uint8 frames normalised by vc_frames_u8_normalize == the reference loader's ToTensor + Normalize, bit for bit (all 256 grey
    levels, ragged length), and the model fed uint8 frames returns exactly what it returns for the normalised fp32 frames."""
    levels = torch.arange(256, dtype=torch.uint8).repeat(3)[:733]  # ragged: exercises the tail path of the CUDA kernel's twin
    dst = torch.empty(levels.numel(), dtype=torch.float32)
    L.check(emu.vc_frames_u8_normalize(levels.data_ptr(), levels.numel(), 0.5, 0.5, dst.data_ptr(), None), emu)
    assert torch.equal(dst, _torch_ingest(levels))
    assert dst.min() == -1.0 and dst.max() == 1.0
    cfg = dict(hidden_size=128, nhead=4, num_decoder_layers=1, dim_feedforward=128, window_size=2,
               enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)
    m, _ = build_emu_model(emu, cfg)
    m.eval()
    g = torch.Generator().manual_seed(3)
    acts = to.model_inputs_from_batch(to.synthetic_batch(2, 4, 64))["actions"]
    frames = torch.randint(0, 256, (2, acts.shape[1], 1, 64, 64), dtype=torch.uint8, generator=g)
    cad = torch.randint(0, 256, (2, 1, 64, 64), dtype=torch.uint8, generator=g)
    with torch.no_grad():
        c8, p8 = m({"frames": frames, "actions": acts, "cad_image": cad})
        cf, pf = m({"frames": _torch_ingest(frames), "actions": acts, "cad_image": _torch_ingest(cad)})
    assert torch.equal(c8, cf) and torch.equal(p8, pf)


# head dim 32 (nhead 4): the per-kernel step vc_seq_decode_step (17 sequences: its GEMM path, > 16 rows); head dim 64 (nhead 2): the
# device-resident step vc_seq_decode_step_dev (<= 8 and 9..16 sequences: both row-count instantiations; Ff != H)
@pytest.mark.parametrize("layers,window,L_,B_,nhead,ff", [(1, 2, 5, 2, 4, 128), (3, 3, 8, 2, 4, 128), (2, 2, 4, 17, 4, 128),
                                                          (3, 3, 9, 3, 2, 256), (2, 2, 6, 9, 2, 128), (1, 4, 5, 17, 2, 128)])
def test_sequential_inference_matches_oracle(emu, layers, window, L_, B_, nhead, ff):
    """Rollout with action feedback through the incremental decoder (one token per step against the key/value cache) against
    the oracle's O(T^2) recompute of the reference loop; three layers exercise the layer-output hand-over, T > window the
    clipping of the cross-attention window."""
    import videocad_b200.model as vm

    cfg = dict(hidden_size=128, nhead=nhead, num_decoder_layers=layers, dim_feedforward=ff, window_size=window,
               enable_past_actions=True, enable_past_states=True, enable_timestep_embedding=True)
    m, sd = build_emu_model(emu, cfg)
    m.eval()
    inp = to.model_inputs_from_batch(to.synthetic_batch(B_, L_, 64))
    used = []
    orig = vm._SeqRunner.decode_run
    vm._SeqRunner.decode_run = lambda self, dec, T: (used.append(T), orig(self, dec, T))[1]
    try:
        ac, ap = m.sequential_inference(inp["frames"], inp["cad_image"], action=True)
    finally:
        vm._SeqRunner.decode_run = orig
    assert bool(used) == (nhead == 2 and B_ <= 16), "which decode step ran"
    oc, op = to.rollout(sd, cfg, inp["frames"], inp["cad_image"], action=True)
    assert (ac - oc).abs().max() < 2e-4 and (ap - op).abs().max() < 2e-4
    assert (ac.argmax(-1) == oc.argmax(-1)).all()
    zc, zp = m.sequential_inference(inp["frames"], inp["cad_image"], action=False)
    fc, fp = m(dict(inp, actions=torch.zeros_like(inp["actions"])))
    assert torch.equal(zc, fc.detach()) and torch.equal(zp, fp.detach())


def test_clip_adam_step_is_seen_by_the_next_forward(emu):
    """ClipAdam writes the parameters through raw pointers; the split-bf16 weight mirror every GEMM reads is refreshed when the
    flat parameter's version counter moves.  After one step the drop-in must compute with the NEW weights everywhere
    (without the version bump the GEMM weights stayed stale while LayerNorm / embed_action read the fresh fp32 values)."""
    from videocad_b200.optim import ClipAdam

    cfg = dict(hidden_size=128, nhead=4, num_decoder_layers=1, dim_feedforward=128, window_size=2, **MODES[0])
    inp = to.model_inputs_from_batch(to.synthetic_batch(2, 3, 64))
    outs = []
    for which in ("torch", "fused"):
        m, _ = build_emu_model(emu, cfg, dropout=0.0)
        m.train()
        opt = (torch.optim.Adam(m.parameters(), lr=1e-2) if which == "torch"
               else ClipAdam(m.parameters(), lr=1e-2, max_norm=1.0, _lib=emu))
        for _ in range(2):
            opt.zero_grad()
            c, p = m(inp)
            (c.sum() + p.sum() * 0.01).backward()
            if which == "torch":
                torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
            opt.step()
        with torch.no_grad():
            outs.append(m(inp))
    assert (outs[0][0] - outs[1][0]).abs().max() < 1e-4 and (outs[0][1] - outs[1][1]).abs().max() < 1e-4


def test_image_input_gradient_is_refused_loudly(emu):
    cfg = dict(hidden_size=128, nhead=4, num_decoder_layers=1, dim_feedforward=128, window_size=2, **MODES[0])
    m, _ = build_emu_model(emu, cfg, dropout=0.0)
    m.eval()
    inp = to.model_inputs_from_batch(to.synthetic_batch(1, 2, 64))
    inp["cad_image"] = inp["cad_image"].clone().requires_grad_(True)  # trainer.generate_saliency_batch (trainer.py:621-645)
    with pytest.raises(RuntimeError, match="input images"):
        m(inp)


def test_ab_switches_keep_the_previous_paths_working():
    """VC_GELU_DSTORE=0 (fc2 dgrad recomputes erf + mask from the stored pre-activation) and VC_HEAD_DGRAD_SPLITK=1 (heads dgrad in
    one pass) are read once per process: the orchestration-vs-oracle check is repeated in a child process with both switched off."""
    import subprocess
    import sys

    env = dict(os.environ, VC_GELU_DSTORE="0", VC_HEAD_DGRAD_SPLITK="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-p", "no:cacheprovider",
                        "-k", "test_orchestration_forward_backward_vs_fp64_oracle and mode0"], env=env, capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), timeout=600)
    assert r.returncode == 0 and "1 passed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
