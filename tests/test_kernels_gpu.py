"""Unit parity tests of every CUDA kernel, through the C ABI, against plain torch (fp64 where it matters).

Run on the B200 box: python -m pytest tests -m gpu
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from videocad_b200 import lib as L  # noqa: E402
import kwrap as K  # noqa: E402


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, generator=g, device="cuda") * scale


def _relerr(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


def test_library_is_cuda_build():
    lib = L.load()
    assert lib.vc_is_cuda_build() == 1


@pytest.mark.parametrize("rows,cols", [(7, 8), (300, 512), (1000, 1024)])
def test_split_bit_exact(rows, cols):
    x = _rand(rows, cols, seed=1) * 3.0
    hi, lo = L.split(x)
    rhi, rlo = K.ref_split(x)
    assert torch.equal(hi.view(torch.int16), rhi.view(torch.int16))
    assert torch.equal(lo.view(torch.int16), rlo.view(torch.int16))
    assert ((hi.float() + lo.float()) - x).abs().max() <= x.abs().max() * 2.0 ** -16


GEMM_SHAPES = [(128, 128, 64), (256, 256, 128), (300, 136, 200), (50, 8, 512), (1000, 3072, 512), (14400, 512, 1024),
               (2560, 1024, 192), (12800, 512, 512)]  # the last two take the 2-SM (CTA-pair) 256x256 kernel


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_majors(M, N, K, a_mn, b_mn):
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("MN-major operand needs a 16-byte aligned leading dimension")
    A = _rand(M, K, seed=2)
    B = _rand(N, K, seed=3)
    a_st = A.t().contiguous() if a_mn else A
    b_st = B.t().contiguous() if b_mn else B
    out = torch.full((M, N), float("nan"), device="cuda")
    L.gemm(L.split(a_st), L.split(b_st), M, N, K, a_mn=a_mn, b_mn=b_mn, out_f32=out)
    ref = A.double() @ B.double().t()
    err = _relerr(out, ref)
    assert err < 3e-5, f"gemm M{M} N{N} K{K} a_mn={a_mn} b_mn={b_mn}: rel err {err:.3e}"


def test_gemm_single_pass_is_bf16_grade():
    M, N, K = 512, 256, 512
    A, B = _rand(M, K, seed=4), _rand(N, K, seed=5)
    out = torch.empty(M, N, device="cuda")
    L.gemm(L.split(A), L.split(B), M, N, K, passes=1, out_f32=out)
    ref_bf = A.to(torch.bfloat16).double() @ B.to(torch.bfloat16).double().t()
    assert _relerr(out, ref_bf) < 1e-5
    ref = A.double() @ B.double().t()
    assert 1e-4 < _relerr(out, ref) < 2e-2


@pytest.mark.parametrize("act", [L.ACT_NONE, L.ACT_GELU, L.ACT_RELU, L.ACT_TANH])
def test_gemm_epilogue(act):
    M, N, Kd, T = 520, 264, 192, 8
    A, B = _rand(M, Kd, seed=6, scale=0.5), _rand(N, Kd, seed=7, scale=0.2)
    bias, rowadd, res = _rand(N, seed=8), _rand(T, N, seed=9), _rand(M, N, seed=10)
    drop = L.make_drop(0.1, 17, 1234)
    pre = torch.empty(M, N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    outs = K.bf16_pair((M, N))
    L.gemm(L.split(A), L.split(B), M, N, Kd, bias=bias, rowadd=rowadd, rowadd_div=1, rowadd_mod=T, preact=pre, act=act,
           drop=drop, residual=res, out_f32=out, out_split=outs)
    mask = K.dropout_mask(drop, M * N).reshape(M, N)
    rows = torch.arange(M, device="cuda") % T
    z = A.double() @ B.double().t() + bias.double() + rowadd.double()[rows]
    f = {L.ACT_NONE: lambda t: t, L.ACT_GELU: lambda t: torch.nn.functional.gelu(t), L.ACT_RELU: torch.relu,
         L.ACT_TANH: torch.tanh}[act]
    ref = f(z) * mask.double() + res.double()
    assert _relerr(pre, z) < 3e-5
    assert _relerr(out, ref) < 3e-5
    assert (K.join(outs) - out).abs().max() <= out.abs().max() * 2.0 ** -15
    keep = (mask > 0).float().mean().item()
    assert abs(keep - 0.9) < 0.01


@pytest.mark.parametrize("act", [L.ACT_NONE, L.ACT_GELU, L.ACT_RELU, L.ACT_TANH])
def test_gemm_backward_activation_epilogue(act):
    """dgrad GEMM with the fused activation/dropout backward and bias-gradient column sums."""
    M, N, Kd = 700, 264, 320
    G, W = _rand(M, Kd, seed=70, scale=0.5), _rand(Kd, N, seed=71, scale=0.2)  # dX[M,N] = G[M,Kd] W[Kd,N]
    pre = _rand(M, N, seed=72)
    drop = L.make_drop(0.1, 33, 4321)
    mask = K.dropout_mask(drop, M * N).reshape(M, N)
    aux, aux_hi, deriv = None, None, torch.ones_like(pre)
    if act == L.ACT_GELU:
        aux = pre
        t = pre.double().requires_grad_(True)
        torch.nn.functional.gelu(t).sum().backward()
        deriv = t.grad.float()
    elif act == L.ACT_TANH:
        aux = torch.tanh(pre)
        deriv = 1 - aux * aux
    elif act == L.ACT_RELU:
        aux_hi = (torch.relu(pre) * mask).to(torch.bfloat16)
        deriv = (pre > 0).float()
    outs = K.bf16_pair((M, N))
    cs = torch.zeros(N, device="cuda")
    L.gemm(L.split(G), L.split(W), M, N, Kd, b_mn=True, act=act, act_backward=True, act_aux=aux, act_aux_hi=aux_hi, drop=drop,
           out_split=outs, colsum=cs)
    ref = (G.double() @ W.double()) * mask.double() * deriv.double()
    assert _relerr(K.join(outs), ref) < 3e-5
    assert _relerr(cs, ref.sum(0)) < 3e-5


@pytest.mark.parametrize("M,N,Kd,p_drop", [(520, 264, 192, 0.1), (520, 264, 192, 0.0), (2560, 1024, 512, 0.1)])
def test_gemm_gelu_derivative_store_and_mul_aux(M, N, Kd, p_drop):
    """VC_ACT_GELU_DSTORE / VC_ACT_MUL_AUX: the forward epilogue leaves D = gelu'(z) * mask * scale in `preact`, the dgrad epilogue
    multiplies by it -- same outputs as VC_ACT_GELU forward, same gradient as the act_backward GELU mode that recomputes erf and
    the mask (single-CTA and 2-SM kernels)."""
    A, B = _rand(M, Kd, seed=6, scale=0.5), _rand(N, Kd, seed=7, scale=0.2)
    bias = _rand(N, seed=8)
    drop = L.make_drop(p_drop, 17, 1234) if p_drop > 0 else None
    mask = K.dropout_mask(drop, M * N).reshape(M, N).double() if drop is not None else torch.ones(M, N, device="cuda").double()
    z = A.double() @ B.double().t() + bias.double()
    zt = z.clone().requires_grad_(True)
    torch.nn.functional.gelu(zt).sum().backward()
    D = torch.full((M, N), float("nan"), device="cuda")
    outs, outs_ref = K.bf16_pair((M, N)), K.bf16_pair((M, N))
    pre = torch.empty(M, N, device="cuda")
    L.gemm(L.split(A), L.split(B), M, N, Kd, bias=bias, preact=D, act=L.ACT_GELU_DSTORE, drop=drop, out_split=outs)
    L.gemm(L.split(A), L.split(B), M, N, Kd, bias=bias, preact=pre, act=L.ACT_GELU, drop=drop, out_split=outs_ref)
    fo, fr = K.join(outs), K.join(outs_ref)
    assert (fo - fr).abs().max() <= 2.0 ** -15 * fr.abs().max()  # the forward output does not change (one split-bf16 rounding step)
    assert _relerr(D, zt.grad * mask) < 1e-4  # gelu' of the fp32-grade GEMM result against gelu' of the exact product
    pt = pre.double().requires_grad_(True)     # against gelu' of the kernel's own pre-activation: fp32 erf / exp rounding only
    torch.nn.functional.gelu(pt).sum().backward()
    assert _relerr(D, pt.grad * mask) < 3e-6
    assert (D[mask == 0] == 0).all()
    # backward: dX = (G W) * D, column sums for the bias gradient
    G, W = _rand(M, Kd, seed=70, scale=0.5), _rand(Kd, N, seed=71, scale=0.2)
    got, ref = K.bf16_pair((M, N)), K.bf16_pair((M, N))
    cs, cs_ref = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
    L.gemm(L.split(G), L.split(W), M, N, Kd, b_mn=True, act=L.ACT_MUL_AUX, act_backward=True, act_aux=D, out_split=got, colsum=cs)
    L.gemm(L.split(G), L.split(W), M, N, Kd, b_mn=True, act=L.ACT_GELU, act_backward=True, act_aux=pre, drop=drop, out_split=ref,
           colsum=cs_ref)
    exact = (G.double() @ W.double()) * mask * zt.grad
    assert _relerr(K.join(got), exact) < 1e-4 and _relerr(K.join(ref), exact) < 3e-5
    assert _relerr(K.join(got), K.join(ref)) < 2.0 ** -15  # the two backward modes agree to one split-bf16 rounding step
    assert _relerr(cs, exact.sum(0)) < 1e-4 and _relerr(cs, cs_ref) < 3e-6
    # misuse is refused
    with pytest.raises(RuntimeError):
        L.gemm(L.split(A), L.split(B), M, N, Kd, act=L.ACT_GELU_DSTORE, out_split=outs)                      # no preact
    with pytest.raises(RuntimeError):
        L.gemm(L.split(G), L.split(W), M, N, Kd, b_mn=True, act=L.ACT_MUL_AUX, act_backward=True, out_split=got)  # no act_aux
    with pytest.raises(RuntimeError):
        L.gemm(L.split(A), L.split(B), M, N, Kd, act=L.ACT_MUL_AUX, out_split=outs)                          # forward


def test_gemm_rowadd_div():
    M, N, K, T = 64, 128, 64, 8
    A, B = _rand(M, K, seed=11), _rand(N, K, seed=12)
    rowadd = _rand(M // T, N, seed=13)
    out = torch.empty(M, N, device="cuda")
    L.gemm(L.split(A), L.split(B), M, N, K, rowadd=rowadd, rowadd_div=T, rowadd_mod=M // T, out_f32=out)
    ref = A.double() @ B.double().t() + rowadd.double().repeat_interleave(T, dim=0)
    assert _relerr(out, ref) < 3e-5


@pytest.mark.parametrize("splitk", [2, 4, 16])
def test_gemm_splitk_wgrad_shape(splitk):
    # wgrad: dW[N,K] = dY^T[N,M] X[M,K] -> both operands MN-major, contraction over the 3000 rows
    rows, Nw, Kw = 3000, 512, 1024
    dY, X = _rand(rows, Nw, seed=14), _rand(rows, Kw, seed=15)
    bias = _rand(Kw, seed=16)
    out = torch.zeros(Nw, Kw, device="cuda")
    L.gemm(L.split(dY), L.split(X), Nw, Kw, rows, a_mn=True, b_mn=True, splitk=splitk, bias=bias, out_f32=out)
    ref = dY.double().t() @ X.double() + bias.double()
    assert _relerr(out, ref) < 3e-5


def _pair_count():
    return L.load().vc_gemm_pair_launch_count()


PAIR_SHAPE = (2560, 1024, 320)  # 10 x 4 tiles of 256 x 256: taken by the CTA-pair kernel


def test_gemm_pair_kernel_is_used_for_encoder_shapes():
    M, N, Kd = PAIR_SHAPE
    A, B = _rand(M, Kd, seed=40), _rand(N, Kd, seed=41)
    out = torch.full((M, N), float("nan"), device="cuda")
    before = _pair_count()
    L.gemm(L.split(A), L.split(B), M, N, Kd, out_f32=out)
    assert _pair_count() == before + 1
    assert _relerr(out, A.double() @ B.double().t()) < 3e-5
    # a decoder-sized problem stays on the single-CTA kernel
    out2 = torch.empty(256, 512, device="cuda")
    L.gemm(L.split(A[:256]), L.split(B[:512]), 256, 512, Kd, out_f32=out2)
    assert _pair_count() == before + 1
    assert _relerr(out2, A[:256].double() @ B[:512].double().t()) < 3e-5


@pytest.fixture(params=[0, 128, 256, -2], ids=["default", "bn128", "bn256", "wave-model"])
def pair_tile(request):
    """both tile widths of the 2-SM kernel (256 x 256 and 256 x 128), and the automatic choice"""
    lib = L.load()
    lib.vc_gemm_pair_force_tile(request.param)
    yield request.param
    lib.vc_gemm_pair_force_tile(0)


@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("M,N,Kd", [(12800, 512, 512), (5120, 384, 192), (5120, 1024, 64)])
def test_gemm_pair_tile_widths(pair_tile, M, N, Kd, b_mn):
    if pair_tile == 256 and N % 256 != 0:
        pytest.skip("N is not a multiple of 256")
    A = _rand(M, Kd, seed=38)
    B = _rand(Kd, N, seed=39) if b_mn else _rand(N, Kd, seed=39)
    out = torch.full((M, N), float("nan"), device="cuda")
    before = _pair_count()
    L.gemm(L.split(A), L.split(B), M, N, Kd, b_mn=b_mn, out_f32=out)
    assert _pair_count() == before + 1
    assert _relerr(out, A.double() @ (B.double() if b_mn else B.double().t())) < 3e-5


@pytest.mark.parametrize("variant", ["split_bias", "f32_drop_res", "split_act", "all"])
def test_gemm_pair_epilogues(variant, pair_tile):
    M, N, Kd = PAIR_SHAPE
    T = 8
    A, B = _rand(M, Kd, seed=42, scale=0.5), _rand(N, Kd, seed=43, scale=0.2)
    bias, res, rowadd = _rand(N, seed=44), _rand(M, N, seed=45), _rand(T, N, seed=46)
    drop = L.make_drop(0.1, 21, 777)
    mask = K.dropout_mask(drop, M * N).reshape(M, N).double()
    z = A.double() @ B.double().t() + bias.double()
    before = _pair_count()
    if variant == "split_bias":
        outs = K.bf16_pair((M, N))
        L.gemm(L.split(A), L.split(B), M, N, Kd, bias=bias, out_split=outs)
        assert _relerr(K.join(outs), z) < 3e-5
    elif variant == "f32_drop_res":
        out = torch.empty(M, N, device="cuda")
        L.gemm(L.split(A), L.split(B), M, N, Kd, bias=bias, drop=drop, residual=res, out_f32=out)
        assert _relerr(out, z * mask + res.double()) < 3e-5
    elif variant == "split_act":
        outs = K.bf16_pair((M, N))
        pre = torch.empty(M, N, device="cuda")
        L.gemm(L.split(A), L.split(B), M, N, Kd, bias=bias, preact=pre, act=L.ACT_GELU, drop=drop, out_split=outs)
        assert _relerr(pre, z) < 3e-5
        assert _relerr(K.join(outs), torch.nn.functional.gelu(z) * mask) < 3e-5
    else:
        out = torch.empty(M, N, device="cuda")
        outs = K.bf16_pair((M, N))
        pre = torch.empty(M, N, device="cuda")
        L.gemm(L.split(A), L.split(B), M, N, Kd, bias=bias, rowadd=rowadd, rowadd_div=1, rowadd_mod=T, preact=pre, act=L.ACT_TANH,
               drop=drop, residual=res, out_f32=out, out_split=outs)
        rows = torch.arange(M, device="cuda") % T
        zz = z + rowadd.double()[rows]
        assert _relerr(pre, zz) < 3e-5
        assert _relerr(out, torch.tanh(zz) * mask + res.double()) < 3e-5
        assert (K.join(outs) - out).abs().max() <= out.abs().max() * 2.0 ** -15
    assert _pair_count() == before + 1


def test_gemm_pair_backward_activation_and_colsum(pair_tile):
    M, N, Kd = 2560, 1024, 512
    G, W = _rand(M, Kd, seed=47, scale=0.5), _rand(Kd, N, seed=48, scale=0.2)
    pre = _rand(M, N, seed=49)
    drop = L.make_drop(0.1, 34, 99)
    mask = K.dropout_mask(drop, M * N).reshape(M, N)
    t = pre.double().requires_grad_(True)
    torch.nn.functional.gelu(t).sum().backward()
    outs = K.bf16_pair((M, N))
    cs = torch.zeros(N, device="cuda")
    before = _pair_count()
    L.gemm(L.split(G), L.split(W), M, N, Kd, b_mn=True, act=L.ACT_GELU, act_backward=True, act_aux=pre, drop=drop, out_split=outs,
           colsum=cs)
    assert _pair_count() == before + 1
    ref = (G.double() @ W.double()) * mask.double() * t.grad
    assert _relerr(K.join(outs), ref) < 3e-5
    assert _relerr(cs, ref.sum(0)) < 3e-5


@pytest.mark.parametrize("rows,Nw,Kw", [(6400, 512, 512), (12800, 3072, 512), (3200, 512, 1024)])
def test_gemm_pair_splitk_wgrad(rows, Nw, Kw, pair_tile):
    dY, X = _rand(rows, Nw, seed=50), _rand(rows, Kw, seed=51)
    out = torch.zeros(Nw, Kw, device="cuda")
    before = _pair_count()
    L.gemm(L.split(dY), L.split(X), Nw, Kw, rows, a_mn=True, b_mn=True, splitk=8, out_f32=out)
    assert _pair_count() == before + 1
    assert _relerr(out, dY.double().t() @ X.double()) < 3e-5


def test_gemm_strided_outputs_and_views():
    # output written into the left half of a wider buffer; B operand is a row-slice of a packed weight
    M, H = 256, 256
    A = _rand(M, 512, seed=17)
    Wfull = _rand(3 * H, 512, seed=18)
    wide = torch.zeros(M, 2 * H, device="cuda")
    wide_s = (torch.zeros(M, 2 * H, dtype=torch.bfloat16, device="cuda"), torch.zeros(M, 2 * H, dtype=torch.bfloat16, device="cuda"))
    whi, wlo = L.split(Wfull)
    L.gemm(L.split(A), (whi[H:2 * H], wlo[H:2 * H]), M, H, 512, out_f32=wide, out_split=wide_s)
    ref = A.double() @ Wfull[H:2 * H].double().t()
    assert _relerr(wide[:, :H], ref) < 3e-5
    assert wide[:, H:].abs().max() == 0
    assert (K.join(wide_s)[:, :H] - wide[:, :H]).abs().max() <= wide.abs().max() * 2.0 ** -15


@pytest.mark.parametrize("rows,C", [(5, 256), (1000, 512), (333, 1024)])
def test_layernorm_fwd_bwd(rows, C):
    x = _rand(rows, C, seed=20) * 2 + 0.3
    gamma, beta = 1 + 0.1 * _rand(C, seed=21), 0.1 * _rand(C, seed=22)
    y, ys, mean, rstd = K.layernorm_fwd(x, gamma, beta)
    xd = x.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xd, (C,), gd, bd, 1e-5)
    assert (y.double() - ref).abs().max() < 2e-5
    assert (K.join(ys) - y).abs().max() <= y.abs().max() * 2.0 ** -15
    dy, dres = _rand(rows, C, seed=23), _rand(rows, C, seed=24)
    ref.backward(dy.double())
    dx, dg, db = K.layernorm_bwd(dy, x, mean, rstd, gamma, dres)
    assert (dx.double() - (xd.grad + dres.double())).abs().max() < 5e-5
    assert _relerr(dg, gd.grad) < 2e-5
    assert _relerr(db, bd.grad) < 2e-5


@pytest.mark.parametrize("rows,C", [(77, 256), (1000, 512), (40, 1024)])
def test_layernorm_bwd_fused_second_output(rows, C):
    x = _rand(rows, C, seed=25)
    gamma, beta = 1 + 0.1 * _rand(C, seed=26), 0.1 * _rand(C, seed=27)
    _, _, mean, rstd = K.layernorm_fwd(x, gamma, beta)
    dy, dres = _rand(rows, C, seed=28), _rand(rows, C, seed=29)
    drop = L.make_drop(0.1, 21, 31337)
    dx_ref, dg_ref, db_ref = K.layernorm_bwd(dy, x, mean, rstd, gamma, dres)
    dx, dg, db, gs, cs = K.layernorm_bwd_fused(dy, x, mean, rstd, gamma, dres, drop)
    assert (dx - dx_ref).abs().max() <= 2e-6 * dx_ref.abs().max()  # two instantiations of one template: FMA contraction may differ
    assert _relerr(dg, dg_ref) < 1e-5 and _relerr(db, db_ref) < 1e-5
    mask = K.dropout_mask(drop, rows * C).reshape(rows, C)
    g = dx * mask
    assert (K.join(gs) - g).abs().max() <= g.abs().max() * 2.0 ** -15
    assert _relerr(cs, g.double().sum(0)) < 1e-5


@pytest.mark.parametrize("F,S", [(3, 64), (5, 224)])
def test_patch_layernorm(F, S):
    img = _rand(F, 1, S, S, seed=30).clamp(-1, 1)
    gamma, beta = 1 + 0.1 * _rand(1024, seed=31), 0.1 * _rand(1024, seed=32)
    ys, mean, rstd = K.patch_layernorm_fwd(img, gamma, beta)
    h = S // 32
    patches = img.reshape(F, 1, h, 32, h, 32).permute(0, 2, 4, 3, 5, 1).reshape(F * h * h, 1024)
    pd = patches.double()
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(pd, (1024,), gd, bd, 1e-5)
    assert (K.join(ys).double() - ref).abs().max() < 1e-4
    dy = _rand(F * h * h, 1024, seed=33)
    ref.backward(dy.double())
    dg, db = K.patch_layernorm_bwd_params(img, mean, rstd, dy)
    assert _relerr(dg, gd.grad) < 2e-5
    assert _relerr(db, bd.grad) < 2e-5


@pytest.mark.parametrize("p", [0.0, 0.1])
def test_vit_assemble(p):
    F, N, C = 9, 49, 512
    e, cls, pos = _rand(F * N, C, seed=40), _rand(C, seed=41), _rand(N + 1, C, seed=42)
    drop = L.make_drop(p, 3, 99)
    x = K.vit_assemble_fwd(e, F, N, C, cls, pos, drop)
    mask = K.dropout_mask(drop, F * (N + 1) * C).reshape(F, N + 1, C) if p > 0 else torch.ones(F, N + 1, C, device="cuda")
    ref = torch.cat([cls.expand(F, 1, C), e.reshape(F, N, C)], 1) + pos
    ref = ref * mask
    assert (x.reshape(F, N + 1, C) - ref).abs().max() < 1e-6
    dx = _rand(F * (N + 1), C, seed=43)
    de, dcls, dpos = K.vit_assemble_bwd(dx, F, N, C, drop)
    g = dx.reshape(F, N + 1, C) * mask
    assert (de.reshape(F, N, C) - g[:, 1:]).abs().max() < 1e-6
    assert (dpos - g.sum(0)).abs().max() < 1e-4
    assert (dcls - g[:, 0].sum(0)).abs().max() < 1e-4


def _attn_ref(q, k, v, B, Tq, Tk, nh, d, mask, window, pmask=None):
    qh = q.reshape(B, Tq, nh, d).permute(0, 2, 1, 3)
    kh = k.reshape(B, Tk, nh, d).permute(0, 2, 1, 3)
    vh = v.reshape(B, Tk, nh, d).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(d)
    if mask != L.MASK_NONE:
        i = torch.arange(Tq, device=q.device)[:, None]
        j = torch.arange(Tk, device=q.device)[None, :]
        bad = j > i
        if mask == L.MASK_WINDOW:
            bad = bad | (j <= i - window)
        s = s.masked_fill(bad, float("-inf"))
    p = torch.softmax(s, -1)
    if pmask is not None:
        p = p * pmask
    o = (p @ vh).permute(0, 2, 1, 3).reshape(B * Tq, nh * d)
    return o, torch.logsumexp(s, -1)


ATTN_CASES = [
    # B, T, nh, d, mask, window, p
    (6, 50, 16, 64, L.MASK_NONE, 1, 0.0),
    (6, 50, 16, 64, L.MASK_NONE, 1, 0.1),
    (3, 5, 16, 64, L.MASK_NONE, 1, 0.0),
    (4, 8, 4, 128, L.MASK_CAUSAL, 1, 0.0),
    (4, 8, 4, 128, L.MASK_WINDOW, 3, 0.1),
    (2, 186, 4, 256, L.MASK_CAUSAL, 1, 0.0),
    (2, 186, 4, 256, L.MASK_WINDOW, 10, 0.1),
    (2, 100, 4, 64, L.MASK_WINDOW, 1, 0.0),
    (2, 33, 4, 64, L.MASK_CAUSAL, 1, 0.1),
    # short sequences: csrc/attention_small.cu (Tq == Tk <= 32)
    (32, 8, 4, 128, L.MASK_CAUSAL, 1, 0.1),
    (32, 8, 4, 128, L.MASK_WINDOW, 10, 0.1),
    (3, 32, 4, 256, L.MASK_CAUSAL, 1, 0.1),
    (3, 32, 4, 256, L.MASK_WINDOW, 10, 0.1),
    (2, 2, 4, 64, L.MASK_CAUSAL, 1, 0.0),
    (2, 1, 4, 64, L.MASK_WINDOW, 1, 0.0),
    (2, 17, 4, 64, L.MASK_WINDOW, 3, 0.1),
    (2, 31, 2, 192, L.MASK_NONE, 1, 0.1),
]


@pytest.mark.parametrize("B,T,nh,d,mask,window,p", ATTN_CASES)
def test_attention_fwd_bwd(B, T, nh, d, mask, window, p):
    H = nh * d
    qkv = _rand(B * T, 3 * H, seed=50, scale=1.0)
    q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
    drop = L.make_drop(p, 5, 77)
    a = K.attn_desc(q, k, v, B, T, T, nh, d, mask=mask, window=window, drop=drop)
    o, lse = K.attention_fwd(a, B, T, nh, d)
    pmask = K.dropout_mask(drop, B * nh * T * T).reshape(B, nh, T, T).double() if p > 0 else None
    qd = q.double().contiguous().requires_grad_(True)
    kd = k.double().contiguous().requires_grad_(True)
    vd = v.double().contiguous().requires_grad_(True)
    ref, ref_lse = _attn_ref(qd, kd, vd, B, T, T, nh, d, mask, window, pmask)
    assert (K.join(o).double() - ref).abs().max() < 5e-5, "forward output"
    assert (lse.double() - ref_lse).abs().max() < 1e-4, "lse"
    dout = _rand(B * T, H, seed=51)
    ref.backward(dout.double())
    dq, dk, dv = K.attention_bwd(a, o, lse, dout, B, T, T, nh, d)
    for name, got, want in (("dq", dq, qd.grad), ("dk", dk, kd.grad), ("dv", dv, vd.grad)):
        assert torch.isfinite(got).all(), name
        err = (got.double() - want).abs().max().item() / max(1.0, want.abs().max().item())
        assert err < 1e-4, f"{name}: {err:.3e}"


@pytest.mark.parametrize("B,T,nh,d,mask,window,p", [(32, 8, 4, 128, L.MASK_CAUSAL, 1, 0.1), (4, 8, 4, 128, L.MASK_WINDOW, 10, 0.1),
                                                    (3, 32, 4, 256, L.MASK_WINDOW, 10, 0.1), (2, 17, 4, 64, L.MASK_CAUSAL, 1, 0.0),
                                                    (2, 40, 4, 64, L.MASK_WINDOW, 3, 0.1)])
def test_attention_bwd_split_bias(B, T, nh, d, mask, window, p):
    """Decoder attention backward delivering dq|dk|dv as one split-bf16 [R, 3W] operand plus the in_proj bias gradient: the fused
    short-sequence kernel (T <= 32) and the generic fallback (T = 40) against fp64, and the short kernel against the generic one."""
    H = nh * d
    qkv = _rand(B * T, 3 * H, seed=54, scale=1.0)
    q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
    drop = L.make_drop(p, 5, 79)
    a = K.attn_desc(q, k, v, B, T, T, nh, d, mask=mask, window=window, drop=drop)
    o, lse = K.attention_fwd(a, B, T, nh, d)
    pmask = K.dropout_mask(drop, B * nh * T * T).reshape(B, nh, T, T).double() if p > 0 else None
    qd, kd, vd = (t.double().contiguous().requires_grad_(True) for t in (q, k, v))
    ref, _ = _attn_ref(qd, kd, vd, B, T, T, nh, d, mask, window, pmask)
    dout = _rand(B * T, H, seed=55)
    ref.backward(dout.double())
    g, db = K.attention_bwd_split_bias(a, o, lse, dout, B, T, nh, d)
    want = torch.cat([qd.grad, kd.grad, vd.grad], 1)
    got = K.join(g).double()
    assert torch.isfinite(got).all()
    assert (got - want).abs().max().item() / max(1.0, want.abs().max().item()) < 1e-4
    assert _relerr(db, want.sum(0)) < 1e-4
    if T <= 32:
        lib = L.load()
        lib.vc_attention_small_enable(0)
        try:
            o2, lse2 = K.attention_fwd(a, B, T, nh, d)
            g2, db2 = K.attention_bwd_split_bias(a, o2, lse2, dout, B, T, nh, d)
        finally:
            lib.vc_attention_small_enable(1)
        # both round their fp32 results to split-bf16 (2^-16 relative): agreement to a few units of that rounding
        assert (K.join(o2) - K.join(o)).abs().max() < 1e-4
        assert (lse2 - lse).abs().max() < 2e-5
        assert (K.join(g2) - K.join(g)).abs().max().item() <= 1e-4 * max(1.0, want.abs().max().item())
        assert _relerr(db2, db.double()) < 1e-4


@pytest.mark.parametrize("B,T,p", [(5, 50, 0.0), (5, 50, 0.1), (3, 5, 0.1), (2, 64, 0.0)])
def test_vit_attention_split_inputs(B, T, p):
    """The tensor-core ViT kernels fed with split-bf16 q/k/v/dO (as the GEMM epilogues write them) and delivering
    split-bf16 dq/dk/dv: same answers as the fp64 reference."""
    nh, d = 16, 64
    H = nh * d
    qkv = _rand(B * T, 3 * H, seed=52)
    hi, lo = L.split(qkv)
    drop = L.make_drop(p, 6, 78)
    a = K.attn_desc_split((hi[:, :H], lo[:, :H]), (hi[:, H:2 * H], lo[:, H:2 * H]), (hi[:, 2 * H:], lo[:, 2 * H:]), B, T, nh, d, drop=drop)
    o, lse = K.attention_fwd(a, B, T, nh, d)
    joined = (hi.float() + lo.float()).double()
    qd = joined[:, :H].contiguous().requires_grad_(True)
    kd = joined[:, H:2 * H].contiguous().requires_grad_(True)
    vd = joined[:, 2 * H:].contiguous().requires_grad_(True)
    pmask = K.dropout_mask(drop, B * nh * T * T).reshape(B, nh, T, T).double() if p > 0 else None
    ref, ref_lse = _attn_ref(qd, kd, vd, B, T, T, nh, d, L.MASK_NONE, 1, pmask)
    assert (K.join(o).double() - ref).abs().max() < 5e-5
    assert (lse.double() - ref_lse).abs().max() < 1e-4
    dout = _rand(B * T, H, seed=53)
    dsplit = L.split(dout)
    ref.backward((dsplit[0].float() + dsplit[1].float()).double())
    dq, dk, dv = K.attention_bwd_split(a, o, lse, dsplit, B, T, nh, d)
    for name, got, want in (("dq", dq, qd.grad), ("dk", dk, kd.grad), ("dv", dv, vd.grad)):
        err = (K.join(got).double() - want).abs().max().item() / max(1.0, want.abs().max().item())
        assert err < 1e-4, f"{name}: {err:.3e}"


@pytest.mark.parametrize("M,N,Kd,act,split_in", [(8, 3072, 1024, L.ACT_NONE, False), (8, 512, 512, L.ACT_RELU, True), (3, 6000, 256, L.ACT_NONE, False),
                                                   (16, 1024, 1024, L.ACT_TANH, True), (1, 40, 128, L.ACT_NONE, False)])
def test_linear_rows_fwd(M, N, Kd, act, split_in):
    """nn.Linear on a handful of rows (decoding step): exact fp32 against fp64, fp32 or split-bf16 input, bias / activation /
    residual, fp32 and split outputs, strided output rows (the key/value cache slot)."""
    x, Wt, bias, res = _rand(M, Kd, seed=70), _rand(N, Kd, seed=71, scale=0.1), _rand(N, seed=72), _rand(M, N, seed=73)
    xs = L.split(x) if split_in else None
    xin = (xs[0].float() + xs[1].float()) if split_in else x
    ldo = 3 * N  # rows of the output live 3N apart
    out = torch.full((M, ldo), float("nan"), device="cuda")
    outs = K.bf16_pair((M, N))
    L.check(L.load().vc_linear_rows_fwd(None if split_in else x.data_ptr(), xs[0].data_ptr() if split_in else None,
                                       xs[1].data_ptr() if split_in else None, Kd, M, Wt.data_ptr(), bias.data_ptr(), N, Kd, act, res.data_ptr(), N,
                                       out.data_ptr(), ldo, outs[0].data_ptr(), outs[1].data_ptr(), N, L.cur_stream()))
    z = xin.double() @ Wt.double().t() + bias.double()
    if act == L.ACT_RELU:
        z = torch.relu(z)
    elif act == L.ACT_TANH:
        z = torch.tanh(z)
    z = z + res.double()
    assert _relerr(out[:, :N], z) < 2e-6
    assert torch.isnan(out[:, N:]).all()
    assert (K.join(outs) - out[:, :N]).abs().max() <= out[:, :N].abs().max() * 2.0 ** -15


@pytest.mark.parametrize("n", [16, 733, 8 * 224 * 224 + 5])
def test_frames_u8_normalize_bit_exact(n):
    """uint8 frame ingestion == ToTensor() + Normalize([0.5], [0.5]) of the reference's loader (main.py:103-110), bit for bit"""
    g = torch.Generator(device="cuda").manual_seed(n)
    u8 = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda", generator=g)
    u8[:min(n, 256)] = torch.arange(min(n, 256), dtype=torch.uint8, device="cuda")
    dst = torch.full((n,), float("nan"), device="cuda")
    L.check(L.load().vc_frames_u8_normalize(u8.data_ptr(), n, 0.5, 0.5, dst.data_ptr(), L.cur_stream()))
    # the reference's loader runs these ops on the CPU, where div is a true fp32 division (torch's CUDA kernel multiplies by the
    # rounded reciprocal of a scalar divisor instead, which differs in the last bit for some grey levels)
    assert torch.equal(dst.cpu(), u8.cpu().to(torch.float32).div(255).sub_(0.5).div_(0.5))


@pytest.mark.parametrize("act", [L.ACT_NONE, L.ACT_GELU, L.ACT_RELU, L.ACT_TANH])
def test_act_dropout_bwd(act):
    M, N = 777, 512
    dy, pre = _rand(M, N, seed=60), _rand(M, N, seed=61)
    drop = L.make_drop(0.1, 9, 5)
    mask = K.dropout_mask(drop, M * N).reshape(M, N)
    aux, aux_hi, deriv = None, None, torch.ones_like(pre)
    if act == L.ACT_GELU:
        aux = pre
        t = pre.double().requires_grad_(True)
        torch.nn.functional.gelu(t).sum().backward()
        deriv = t.grad.float()
    elif act == L.ACT_TANH:
        aux = torch.tanh(pre)
        deriv = 1 - aux * aux
    elif act == L.ACT_RELU:
        out = torch.relu(pre) * mask
        aux_hi = out.to(torch.bfloat16)
        deriv = (pre > 0).float()
    g, gs, cs = K.act_dropout_bwd(dy, act, aux=aux, aux_hi=aux_hi, drop=drop)
    ref = dy * mask * deriv
    assert (g - ref).abs().max() < 1e-5
    assert (K.join(gs) - g).abs().max() <= g.abs().max() * 2.0 ** -15
    assert _relerr(cs, ref.double().sum(0)) < 1e-5
