"""GPU: the tcgen05/TMEM/TMA GEMM against an fp64 reference computed from the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

from helpers import max_abs
from l3ac_b200 import ops
from test_kernels_gpu import _gemm_ref, rnd

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def bf(x):
    return x.to(torch.bfloat16)


@pytest.mark.parametrize("B,T,K,N,taps,shift0,step", [
    (1, 128, 64, 32, 1, 0, 1),            # one tile, one k-block
    (1, 300, 128, 64, 1, 0, 1),           # M tail
    (1, 1000, 512, 2048, 1, 0, 1),        # pw_conv1 of the C=512 ConvUnit (BN=256, 8 N tiles)
    (1, 700, 2048, 512, 1, 0, 1),         # pw_conv2 (32 k-blocks, deep pipeline wrap)
    (2, 333, 96, 384, 1, 0, 1),           # K tail (96 = 64 + 32)
    (2, 500, 24, 96, 1, 0, 1),            # K < 64
    (2, 131, 192, 48, 1, 0, 1),           # N tail inside a 64-wide tile
    (3, 150, 128, 512, 3, -1, 1),         # k3 conv, per-sample zero padding through TMA OOB fill
    (2, 400, 24, 24, 7, -27, 9),          # dilated k7 LegacyUnit conv, K not a multiple of 64 per tap
    (1, 260, 128, 576, 1, 0, 1),          # to_qkv: N = 576 -> BN 192
    (1, 129, 352, 128, 1, 0, 1),          # FeedForward out, K = 352
    (40, 1779, 256, 1024, 1, 0, 1),       # > 148 tiles per CTA wave: persistent loop + TMEM double buffering
])
def test_gemm_tc_matches_fp64(cuda_lib, B, T, K, N, taps, shift0, step):
    a, w, bias = bf(rnd(B, T, K, seed=1)), bf(rnd(N, taps * K, seed=2, scale=0.1)), rnd(N, seed=3)
    got = ops.gemm(a.to(DEV), w.to(DEV), B=B, T=T, K=K, taps=taps, tap_shift0=shift0, tap_step=step, bias=bias.to(DEV))
    torch.cuda.synchronize()
    if B * T * N > 2e7:      # big case: check a random subset of rows against the reference
        rows = torch.randint(0, T, (64,))
        want = _gemm_ref(a.float(), w.float(), bias, taps, shift0, step)[:, rows] if taps > 1 else \
            (a.double()[:, rows] @ w.double().t() + bias.double())
        got = got.cpu()[:, rows]
    else:
        want = _gemm_ref(a.float(), w.float(), bias, taps, shift0, step)
        got = got.cpu()
    tol = 2e-5 * max(1.0, float(want.abs().max())) * max(1.0, (taps * K) ** 0.5 / 8)
    assert max_abs(got, want) < tol


def test_gemm_tc_epilogues(cuda_lib):
    B, T, K, N = 2, 300, 256, 1024
    a, w, bias = bf(rnd(B, T, K, seed=1)), bf(rnd(N, K, seed=2, scale=0.05)), rnd(N, seed=3, scale=0.1)
    alpha, gamma, beta = 0.5 + torch.rand(N), rnd(N, seed=4, scale=0.1), rnd(N, seed=5, scale=0.1)
    res = rnd(B, T, N, seed=6)
    lin = (a.double() @ w.double().t() + bias.double())
    sn = lin + (alpha.double() + 1e-8).reciprocal() * torch.sin(alpha.double() * lin).pow(2)
    want = sn * (1 + gamma.double()) + beta.double() + res.double()
    got = ops.gemm(a.to(DEV), w.to(DEV), B=B, T=T, K=K, bias=bias.to(DEV), act=ops.ACT_SNAKE, alpha=alpha.to(DEV),
                   scale=(1 + gamma).to(DEV), shift=beta.to(DEV), residual=res.to(DEV))
    assert max_abs(got.cpu(), want) < 2e-4                      # __sinf in the fast epilogue
    got16 = ops.gemm(a.to(DEV), w.to(DEV), B=B, T=T, K=K, bias=bias.to(DEV), act=ops.ACT_SNAKE, alpha=alpha.to(DEV),
                     scale=(1 + gamma).to(DEV), shift=beta.to(DEV), out_dtype=torch.bfloat16)
    want16 = sn * (1 + gamma.double()) + beta.double()
    assert max_abs(got16.float().cpu(), want16) < 1e-2 * max(1.0, float(want16.abs().max()))
    # GEGLU (interleaved columns), fp32 and bf16 outputs, N/2 = 352-style tail
    N2 = 704
    w2, = (bf(rnd(N2, K, seed=7, scale=0.05)),)
    lin2 = a.double() @ w2.double().t()
    want = lin2[..., 0::2] * F.gelu(lin2[..., 1::2])
    got = ops.gemm(a.to(DEV), w2.to(DEV), B=B, T=T, K=K, act=ops.ACT_GEGLU)
    assert max_abs(got.cpu(), want) < 2e-5 * max(1.0, float(want.abs().max())) * 4
    got16 = ops.gemm(a.to(DEV), w2.to(DEV), B=B, T=T, K=K, act=ops.ACT_GEGLU, out_dtype=torch.bfloat16)
    assert max_abs(got16.float().cpu(), want) < 1e-2 * max(1.0, float(want.abs().max()))


def test_gemm_tc_linearity_at_full_size(cuda_lib):
    """Size-independent property at the bench shape (B=16 x 8895 rows, C=256 MLP): GEMM(a1 + a2) == GEMM(a1) + GEMM(a2)."""
    B, T, K, N = 16, 8895, 256, 1024
    g = torch.Generator(device=DEV).manual_seed(0)
    a1 = (torch.randint(-8, 9, (B, T, K), generator=g, device=DEV).float() / 8).to(torch.bfloat16)   # exactly representable
    a2 = (torch.randint(-8, 9, (B, T, K), generator=g, device=DEV).float() / 8).to(torch.bfloat16)
    w = (torch.randint(-4, 5, (N, K), generator=g, device=DEV).float() / 16).to(torch.bfloat16)
    y1 = ops.gemm(a1, w, B=B, T=T, K=K)
    y2 = ops.gemm(a2, w, B=B, T=T, K=K)
    y12 = ops.gemm((a1.float() + a2.float()).to(torch.bfloat16), w, B=B, T=T, K=K)
    assert torch.equal(y12, y1 + y2)          # all products/sums are exact in fp32 for these dyadic operands


def _split(x):
    hi = x.to(torch.bfloat16)
    return ops.Split(hi.to(DEV).contiguous(), (x - hi.float()).to(torch.bfloat16).to(DEV).contiguous())


@pytest.mark.parametrize("B,T,K,N,taps,shift0,step", [
    (2, 300, 24, 96, 1, 0, 1), (2, 300, 96, 24, 1, 0, 1), (1, 200, 144, 48, 1, 0, 1), (2, 150, 192, 128, 3, -1, 1),
    (1, 500, 192, 768, 1, 0, 1), (1, 500, 768, 192, 1, 0, 1), (1, 260, 128, 576, 1, 0, 1)])
def test_gemm_tc_split_is_fp32_class(cuda_lib, B, T, K, N, taps, shift0, step):
    """3-term split-bf16 GEMM against fp64 on the *unrounded* fp32 operands: error at the 2^-16 level, not 2^-8."""
    a, w, bias = rnd(B, T, K, seed=1), rnd(N, taps * K, seed=2, scale=0.1), rnd(N, seed=3)
    want = _gemm_ref(a, w, bias, taps, shift0, step)
    got = ops.gemm(_split(a), _split(w), B=B, T=T, K=K, taps=taps, tap_shift0=shift0, tap_step=step, bias=bias.to(DEV))
    scale = max(1.0, float(want.abs().max()))
    assert max_abs(got.cpu(), want) < 3e-5 * scale
    plain = ops.gemm(bf(a).to(DEV), bf(w).to(DEV), B=B, T=T, K=K, taps=taps, tap_shift0=shift0, tap_step=step, bias=bias.to(DEV))
    assert max_abs(got.cpu(), want) < 0.05 * max_abs(plain.cpu(), want) + 1e-6     # >= 20x closer than plain bf16


def test_split_outputs_and_producers(cuda_lib):
    x = rnd(3, 100, 96, seed=1)
    sp = ops.split_bf16(x.to(DEV))
    assert max_abs(sp.float().cpu(), x) < 2e-5 and torch.equal(sp.hi.cpu(), x.to(torch.bfloat16))
    w, b = 1 + rnd(96, seed=2, scale=0.1), rnd(96, seed=3, scale=0.1)
    ln32 = ops.layernorm(x.to(DEV), w.to(DEV), b.to(DEV), 1e-5)
    lnsp = ops.layernorm(x.to(DEV), w.to(DEV), b.to(DEV), 1e-5, out_dtype=ops.SPLIT)
    assert max_abs(lnsp.float(), ln32) < 3e-5
    # GEMM epilogue emitting a split pair (snake + affine), then consumed by a second split GEMM
    a, w1, w2 = rnd(2, 200, 48, seed=4), rnd(192, 48, seed=5, scale=0.2), rnd(48, 192, seed=6, scale=0.1)
    alpha = 0.5 + torch.rand(192)
    h = ops.gemm(_split(a), _split(w1), B=2, T=200, K=48, act=ops.ACT_SNAKE, alpha=alpha.to(DEV), out_dtype=ops.SPLIT)
    lin = a.double() @ w1.double().t()
    want_h = lin + (alpha.double() + 1e-8).reciprocal() * torch.sin(alpha.double() * lin).pow(2)
    assert max_abs(h.float().cpu(), want_h) < 1e-4              # __sinf (~5e-7) + pair representation error |x| * 2^-17
    y = ops.gemm(h, _split(w2), B=2, T=200, K=192, residual=a.to(DEV))
    want_y = h.float().cpu().double() @ w2.double().t() + a.double()
    assert max_abs(y.cpu(), want_y) < 3e-5 * max(1.0, float(want_y.abs().max()))


@pytest.mark.parametrize("M,C", [(128, 256), (1000, 256), (333, 96), (2000, 48), (5000, 96), (40000, 256), (300, 192), (130, 64),
                                  (128 * 148 * 5 + 77, 48), (128 * 148 * 3 + 5, 96), (128 * 148 * 2 + 1, 128)])
def test_fused_convunit_mlp(cuda_lib, M, C):
    """Fused MLP kernel vs (a) the two-GEMM tcgen05 path it replaces and (b) fp64 on the same bf16 operands."""
    a = bf(rnd(1, M, C, seed=1))
    w1, w2 = bf(rnd(4 * C, C, seed=2, scale=C ** -0.5)), bf(rnd(C, 4 * C, seed=3, scale=(4 * C) ** -0.5))
    b1, b2 = rnd(4 * C, seed=4, scale=0.1), rnd(C, seed=5, scale=0.1)
    alpha, gamma, beta = 0.5 + torch.rand(4 * C), rnd(4 * C, seed=6, scale=0.1), rnd(4 * C, seed=7, scale=0.1)
    x = rnd(1, M, C, seed=8)
    dev = lambda t: t.to(DEV)
    got = ops.convunit_mlp(dev(a), dev(w1), dev(b1), dev(alpha), dev(1 + gamma), dev(beta), dev(w2), dev(b2), dev(x))
    h = ops.gemm(dev(a), dev(w1), B=1, T=M, K=C, bias=dev(b1), act=ops.ACT_SNAKE, alpha=dev(alpha), scale=dev(1 + gamma),
                 shift=dev(beta), out_dtype=torch.bfloat16)
    two = ops.gemm(h, dev(w2), B=1, T=M, K=4 * C, bias=dev(b2), residual=dev(x))
    lin = a.double() @ w1.double().t() + b1.double()
    hid = (lin + (alpha.double() + 1e-8).reciprocal() * torch.sin(alpha.double() * lin).pow(2)) * (1 + gamma.double()) + beta.double()
    want = hid.to(torch.bfloat16).double() @ w2.double().t() + b2.double() + x.double()
    e_two, e_fused = max_abs(two.cpu(), want), max_abs(got.cpu(), want)
    print(f"[mlp M={M} C={C}] max-abs vs fp64(bf16 hidden): fused {e_fused:.2e}  two-GEMM {e_two:.2e}; fused-vs-two {max_abs(got, two):.2e}")
    assert e_fused < 2e-2 * max(1.0, float(want.abs().max())) and e_fused < 3 * e_two + 1e-3


@pytest.mark.parametrize("M,C", [(128 * 148 * 6 + 33, 48), (128 * 148 * 4 + 90, 96), (128 * 148 * 3 + 7, 256)])
def test_fused_convunit_mlp_is_deterministic(cuda_lib, M, C):
    """The fused kernel hands buffers between TMA, two MMA issuers and 16 epilogue warps through ~50 mbarriers: a missed
    hand-over would show up as run-to-run differences.  Many tiles per CTA, repeated launches, bitwise comparison."""
    a = bf(rnd(1, M, C, seed=11)).to(DEV)
    w1, w2 = bf(rnd(4 * C, C, seed=12, scale=C ** -0.5)).to(DEV), bf(rnd(C, 4 * C, seed=13, scale=(4 * C) ** -0.5)).to(DEV)
    b1, b2 = rnd(4 * C, seed=14, scale=0.1).to(DEV), rnd(C, seed=15, scale=0.1).to(DEV)
    alpha, scale, shift = (0.5 + torch.rand(4 * C)).to(DEV), (1 + rnd(4 * C, seed=16, scale=0.1)).to(DEV), rnd(4 * C, seed=17, scale=0.1).to(DEV)
    x = rnd(1, M, C, seed=18).to(DEV)
    first = ops.convunit_mlp(a, w1, b1, alpha, scale, shift, w2, b2, x)
    with_ch0, ch0 = ops.convunit_mlp(a, w1, b1, alpha, scale, shift, w2, b2, x, want_ch0=True)      # + channel 0 as a compact plane
    assert torch.equal(with_ch0, first) and ch0.shape == first.shape[:-1] and torch.equal(ch0, first[..., 0])
    for _ in range(8):
        again = ops.convunit_mlp(a, w1, b1, alpha, scale, shift, w2, b2, x)
        assert torch.equal(first, again)
    # identical rows in different tiles / CTAs must give identical results (tile-position independence)
    a2, x2 = a.clone(), x.clone()
    a2[0, 128 * 200 + 5] = a2[0, 3]
    x2[0, 128 * 200 + 5] = x2[0, 3]
    out = ops.convunit_mlp(a2, w1, b1, alpha, scale, shift, w2, b2, x2)
    assert torch.equal(out[0, 128 * 200 + 5], out[0, 3])
