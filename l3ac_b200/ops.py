"""Tensor-level wrappers over the C ABI (``include/l3ac_b200.h``).

PyTorch is plumbing here: it owns device memory and the current stream; every computation below is a
call into ``libl3ac_b200.so``.  Inputs must be contiguous CUDA tensors; outputs are fresh tensors on
the same device.  Activations are channels-last ``(B, T, C)``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import ACT_GEGLU, ACT_GELU, ACT_NONE, ACT_SNAKE, ACT_TANH, BF16, BF16X2, F32, GemmDesc, check  # noqa: F401

SPLIT = "split_bf16"        # out_dtype marker: emit a (hi, lo) bf16 pair, the operand format of the 3-term GEMM
_DT = {torch.float32: F32, torch.bfloat16: BF16, SPLIT: BF16X2}


class Split:
    """x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi): two bf16 planes of identical shape."""
    __slots__ = ("hi", "lo")

    def __init__(self, hi: torch.Tensor, lo: torch.Tensor):
        self.hi, self.lo = hi, lo

    @property
    def shape(self):
        return self.hi.shape

    @property
    def device(self):
        return self.hi.device

    def float(self) -> torch.Tensor:
        return self.hi.float() + self.lo.float()

    def view(self, *shape) -> "Split":
        return Split(self.hi.view(*shape), self.lo.view(*shape))


def _empty_act(shape, device, out_dtype):
    """Allocates an activation of the requested kind; returns (object, hi_ptr_tensor, lo_ptr_tensor_or_None)."""
    if out_dtype == SPLIT:
        hi = torch.empty(shape, device=device, dtype=torch.bfloat16)
        lo = torch.empty(shape, device=device, dtype=torch.bfloat16)
        return Split(hi, lo), hi, lo
    t = torch.empty(shape, device=device, dtype=out_dtype)
    return t, t, None

# Instrumentation used by bench.py: number of kernels this library launched, and an optional hook that brackets
# every launch with CUDA events (``OP_HOOK(name, flops, bytes) -> context manager``).  Off by default.
LAUNCHES = 0
OP_HOOK = None


def _count(n: int = 1):
    global LAUNCHES
    LAUNCHES += n


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NULL = _NullCtx()


def _hook(name: str, nbytes: float = 0.0, flops: float = 0.0):
    return _NULL if OP_HOOK is None else OP_HOOK(name, flops, nbytes)


def _nbytes(*tensors) -> int:
    n = 0
    for t in tensors:
        if t is None:
            continue
        if isinstance(t, Split):
            n += 2 * t.hi.numel() * 2
        else:
            n += t.numel() * t.element_size()
    return n


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def _chk(t: torch.Tensor, dtype=torch.float32, name="tensor"):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (l3ac_b200 has no CPU path)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def _levels(levels: Sequence[int]):
    return (C.c_int * len(levels))(*[int(v) for v in levels])


def stem(audio: torch.Tensor, branch_w, branch_b, w1, b1, w2, b2) -> torch.Tensor:
    _chk(audio, name="audio")
    B, T = audio.shape
    Cout = w2.shape[0]
    out = torch.empty((B, T, Cout), device=audio.device, dtype=torch.float32)
    _count()
    with _hook("stem", _nbytes(audio, out), 2.0 * audio.numel() * 3700), torch.cuda.device(audio.device):
        check(_lib.load().l3ac_stem(_ptr(audio), B, T, _ptr(branch_w), _ptr(branch_b), _ptr(w1), _ptr(b1), _ptr(w2),
                                    _ptr(b2), Cout, _ptr(out), _stream(audio)), "l3ac_stem")
    return out


def stem_tc(audio: torch.Tensor, branch_w, branch_b, w1, b1, w2, b2) -> torch.Tensor:
    """Encoder stem with the two 1x1 convs as 3-term split-bf16 tensor-core MMAs (fp32-class)."""
    _chk(audio, name="audio")
    B, T = audio.shape
    Cout = w2.shape[0]
    out = torch.empty((B, T, Cout), device=audio.device, dtype=torch.float32)
    _count()
    with _hook("stem_tc", _nbytes(audio, out), 2.0 * audio.numel() * 3700), torch.cuda.device(audio.device):
        check(_lib.load().l3ac_stem_tc(_ptr(audio), B, T, _ptr(branch_w), _ptr(branch_b), _ptr(w1), _ptr(b1), _ptr(w2),
                                       _ptr(b2), Cout, _ptr(out), _stream(audio)), "l3ac_stem_tc")
    return out


def dwconv7_ln(x, dw_w, dw_b, ln_w, ln_b, eps: float, out_dtype=torch.float32) -> torch.Tensor:
    _chk(x, name="x")
    B, T, Cc = x.shape
    out, hi, lo = _empty_act((B, T, Cc), x.device, out_dtype)
    _count()
    with _hook("dwconv7_ln", _nbytes(x, out)), torch.cuda.device(x.device):
        check(_lib.load().l3ac_dwconv7_ln(_ptr(x), B, T, Cc, _ptr(dw_w), _ptr(dw_b), _ptr(ln_w), _ptr(ln_b), eps,
                                          _ptr(hi), _ptr(lo), _DT[out_dtype], _stream(x)), "l3ac_dwconv7_ln")
    return out


class DwconvPlan:
    """Host-side parameters of the thread-per-row dwconv7 + LayerNorm kernel (``l3ac_dwconv_plan``, C = 48 / 96, bf16 out)."""

    def __init__(self, dw_w, dw_b, ln_w, ln_b, eps: float):
        host = lambda t: t.detach().to("cpu", torch.float32).contiguous()
        dw, db, lw, lb = (host(t) for t in (dw_w, dw_b, ln_w, ln_b))
        self.C = int(db.numel())
        if tuple(dw.shape) != (7, self.C):
            raise ValueError("DwconvPlan: dw_w (7, C) expected")
        self.handle = C.c_void_p()
        check(_lib.load().l3ac_dwconv_plan_create(self.C, dw.data_ptr(), db.data_ptr(), lw.data_ptr(), lb.data_ptr(), float(eps),
                                                  C.byref(self.handle)), "l3ac_dwconv_plan_create")

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                _lib.load().l3ac_dwconv_plan_destroy(h)
            except Exception:
                pass


def dwconv7_ln_plan(x: torch.Tensor, plan: DwconvPlan) -> torch.Tensor:
    """dwconv7 + LayerNorm, thread-per-row kernel: x (B, T, C) fp32 -> (B, T, C) bf16."""
    _chk(x, name="x")
    B, T, Cc = x.shape
    if Cc != plan.C:
        raise ValueError(f"dwconv7_ln_plan: plan is for C = {plan.C}, got {Cc}")
    out = torch.empty((B, T, Cc), device=x.device, dtype=torch.bfloat16)
    _count()
    with _hook("dwconv7_ln", _nbytes(x, out)), torch.cuda.device(x.device):
        check(_lib.load().l3ac_dwconv7_ln_plan(plan.handle, _ptr(x), B, T, _ptr(out), _stream(x)), "l3ac_dwconv7_ln_plan")
    return out


def layernorm(x, w, b, eps: float, out_dtype=torch.float32) -> torch.Tensor:
    _chk(x, name="x")
    Cc = x.shape[-1]
    M = x.numel() // Cc
    out, hi, lo = _empty_act(x.shape, x.device, out_dtype)
    _count()
    with _hook("layernorm", _nbytes(x, out)), torch.cuda.device(x.device):
        check(_lib.load().l3ac_layernorm(_ptr(x), M, Cc, _ptr(w), _ptr(b), eps, _ptr(hi), _ptr(lo), _DT[out_dtype],
                                         _stream(x)), "l3ac_layernorm")
    return out


def split_bf16(x: torch.Tensor) -> Split:
    """fp32 -> (hi, lo) bf16 pair."""
    _chk(x, name="x")
    out, hi, lo = _empty_act(x.shape, x.device, SPLIT)
    _count()
    with _hook("split_bf16", _nbytes(x, out)), torch.cuda.device(x.device):
        check(_lib.load().l3ac_split_bf16(_ptr(x), x.numel(), _ptr(hi), _ptr(lo), _stream(x)), "l3ac_split_bf16")
    return out


def snake(x, alpha, out_dtype=torch.float32) -> torch.Tensor:
    _chk(x, name="x")
    Cc = x.shape[-1]
    out = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    _count()
    with _hook("snake", _nbytes(x, out)), torch.cuda.device(x.device):
        check(_lib.load().l3ac_snake(_ptr(x), x.numel() // Cc, Cc, _ptr(alpha), _ptr(out), _DT[out_dtype], _stream(x)),
              "l3ac_snake")
    return out


def gemm(a, w, *, B: int, T: int, K: int, taps: int = 1, tap_shift0: int = 0, tap_step: int = 1, bias=None,
         act: int = ACT_NONE, alpha=None, scale=None, shift=None, residual=None, out_dtype=torch.float32,
         lda: Optional[int] = None, out=None):
    """out[(b,t), n] = epi(bias[n] + sum_s sum_k a[b, t + shift_s, k] w[n, s*K + k])  (see the header).

    ``a`` is any contiguous tensor whose memory is ``B*T`` rows of pitch ``lda`` (default ``K``).  fp32 operands take
    the SIMT path, bf16 operands the tcgen05 path, ``Split`` operands the 3-term split tcgen05 path.
    ``out_dtype``: torch.float32 | torch.bfloat16 | ops.SPLIT.  Returns ``(B, T, N_out)``.
    """
    split = isinstance(a, Split)
    if split != isinstance(w, Split):
        raise ValueError("gemm operands must both be Split pairs or both plain tensors")
    a_hi, w_hi = (a.hi, w.hi) if split else (a, w)
    if a_hi.dtype not in (torch.float32, torch.bfloat16) or a_hi.dtype != w_hi.dtype:
        raise ValueError(f"gemm operands must both be fp32 or both bf16, got {a_hi.dtype} / {w_hi.dtype}")
    for t, name in ((a_hi, "a"), (w_hi, "w")) + (((a.lo, "a.lo"), (w.lo, "w.lo")) if split else ()):
        _chk(t, a_hi.dtype, name)
    lda = K if lda is None else lda
    if a_hi.numel() != B * T * lda:
        raise ValueError(f"a has {a_hi.numel()} elements, expected B*T*lda = {B * T * lda}")
    N = w_hi.shape[0]
    if w_hi.shape[1] != taps * K:
        raise ValueError(f"w must be (N, taps*K) = (N, {taps * K}), got {tuple(w_hi.shape)}")
    n_out = N // 2 if act == ACT_GEGLU else N
    if out is None:
        out, o_hi, o_lo = _empty_act((B, T, n_out), a_hi.device, out_dtype)
    else:       # caller-provided destination (a contiguous row range of a larger activation)
        o_hi, o_lo = (out.hi, out.lo) if isinstance(out, Split) else (out, None)
        want = torch.bfloat16 if out_dtype == SPLIT else out_dtype
        if (out_dtype == SPLIT) != isinstance(out, Split) or o_hi.dtype != want or o_hi.numel() != B * T * n_out or \
                not o_hi.is_contiguous():
            raise ValueError("out does not match the requested output kind / shape")
    if residual is not None:
        _chk(residual, name="residual")
        if residual.numel() != B * T * n_out:
            raise ValueError("residual shape mismatch")
    d = GemmDesc(A=_ptr(a_hi), W=_ptr(w_hi), bias=_ptr(bias), alpha=_ptr(alpha), scale=_ptr(scale), shift=_ptr(shift),
                 residual=_ptr(residual), out=_ptr(o_hi), A_lo=_ptr(a.lo) if split else None,
                 W_lo=_ptr(w.lo) if split else None, out_lo=_ptr(o_lo), lda=lda, ldr=n_out, ldo=n_out, B=B, T=T, K=K, N=N,
                 taps=taps, tap_shift0=tap_shift0, tap_step=tap_step, act=act, out_dtype=_DT[out_dtype])
    lib = _lib.load()
    fn, what = (lib.l3ac_gemm_f32, "l3ac_gemm_f32") if a_hi.dtype == torch.float32 else (lib.l3ac_gemm_bf16_tc,
                                                                                         "l3ac_gemm_bf16_tc")
    _count()
    with torch.cuda.device(a_hi.device):
        if OP_HOOK is None:
            check(fn(C.byref(d), _stream(a_hi)), what)
        else:
            flops = 2.0 * B * T * N * K * taps
            esz = a_hi.element_size() * (2 if split else 1)
            osz = {torch.float32: 4, torch.bfloat16: 2, SPLIT: 4}[out_dtype]
            nbytes = esz * (B * T * K + N * K * taps) + osz * B * T * n_out + (0 if residual is None else 4 * residual.numel())
            kind = "gemm_f32" if a_hi.dtype == torch.float32 else ("gemm_tc_split" if split else "gemm_tc")
            with OP_HOOK(kind, flops, nbytes):
                check(fn(C.byref(d), _stream(a_hi)), what)
    return out


def convunit_mlp(a: torch.Tensor, w1, b1, alpha, scale, shift, w2, b2, residual: torch.Tensor, ialpha=None, want_ch0: bool = False):
    """Fused pw_conv1 -> snake/GRN -> pw_conv2 -> +residual (tcgen05, hidden activation on chip).  a (…, C) bf16.
    ``want_ch0``: also return channel 0 of the result as a compact fp32 plane (for the EnhanceBlock statistics pass)."""
    _chk(a, torch.bfloat16, "a")
    _chk(residual, name="residual")
    Cc = a.shape[-1]
    M = a.numel() // Cc
    if tuple(w1.shape) != (4 * Cc, Cc) or tuple(w2.shape) != (Cc, 4 * Cc) or residual.numel() != a.numel():
        raise ValueError("convunit_mlp shape mismatch")
    out = torch.empty(residual.shape, device=a.device, dtype=torch.float32)
    ch0 = torch.empty(residual.shape[:-1], device=a.device, dtype=torch.float32) if want_ch0 else None
    if ialpha is None:
        ialpha = 1.0 / (alpha + 1e-8)
    _count()
    with _hook("convunit_mlp_tc", _nbytes(a, residual, out, w1, w2), 2.0 * M * 2 * 4 * Cc * Cc), torch.cuda.device(a.device):
        check(_lib.load().l3ac_convunit_mlp_tc_ch0(_ptr(a), _ptr(w1), _ptr(b1), _ptr(alpha), _ptr(ialpha), _ptr(scale), _ptr(shift), _ptr(w2),
                                                   _ptr(b2), _ptr(residual), _ptr(out), _ptr(ch0), M, Cc, _stream(a)), "l3ac_convunit_mlp_tc")
    return (out, ch0) if want_ch0 else out


def convunit_thin(x: torch.Tensor, dw_w, dw_b, ln_w, ln_b, eps: float, w1, b1, alpha, scale, shift, w2, b2, out_dtype=torch.float32):
    """Fused fp32 Residual(ConvUnit) for C = 24 (x (B, T, 24) fp32 -> same shape; ``out_dtype=SPLIT`` returns the split pair)."""
    _chk(x, name="x")
    B, T, Cc = x.shape
    if out_dtype not in (torch.float32, SPLIT):
        raise ValueError("convunit_thin emits fp32 or the split-bf16 pair")
    out, hi, lo = _empty_act(tuple(x.shape), x.device, out_dtype)
    _count()
    with _hook("convunit_thin_f32", _nbytes(x) + B * T * Cc * 4, 2.0 * B * T * (7 * Cc + 8 * Cc * Cc)), torch.cuda.device(x.device):
        check(_lib.load().l3ac_convunit_thin_f32(_ptr(x), B, T, Cc, _ptr(dw_w), _ptr(dw_b), _ptr(ln_w), _ptr(ln_b), eps, _ptr(w1),
                                                 _ptr(b1), _ptr(alpha), _ptr(scale), _ptr(shift), _ptr(w2), _ptr(b2), _ptr(hi),
                                                 _ptr(lo), _DT[out_dtype], _stream(x)), "l3ac_convunit_thin_f32")
    return out


def convunit_thin_tc(x: torch.Tensor, dw_w, dw_b, ln_w, ln_b, eps: float, w1, b1, alpha, scale, shift, w2, b2, out_dtype=torch.float32,
                     operands=None):
    """Fused Residual(ConvUnit) for C = 24 / 48 on the tensor cores: x (B, T, C) fp32 -> same shape; ``out_dtype=SPLIT``
    returns the split pair.  ``operands``: SPLIT (default) = 3-term split-bf16 products (fp32-class, encode side);
    torch.bfloat16 = plain bf16 operands with fp32 accumulation (decode side)."""
    operands = SPLIT if operands is None else operands
    _chk(x, name="x")
    B, T, Cc = x.shape
    if out_dtype not in (torch.float32, SPLIT):
        raise ValueError("convunit_thin_tc emits fp32 or the split-bf16 pair")
    out, hi, lo = _empty_act(tuple(x.shape), x.device, out_dtype)
    _count()
    kind = "convunit_thin_tc" if operands == SPLIT else "convunit_thin_tc_bf16"
    with _hook(kind, _nbytes(x) + B * T * Cc * 4, 2.0 * B * T * (7 * Cc + 8 * Cc * Cc)), torch.cuda.device(x.device):
        check(_lib.load().l3ac_convunit_thin_tc(_ptr(x), B, T, Cc, _ptr(dw_w), _ptr(dw_b), _ptr(ln_w), _ptr(ln_b), eps, _ptr(w1),
                                                _ptr(b1), _ptr(alpha), _ptr(scale), _ptr(shift), _ptr(w2), _ptr(b2), _ptr(hi),
                                                _ptr(lo), _DT[out_dtype], _DT[operands], _stream(x)), "l3ac_convunit_thin_tc")
    return out


def local_attention(qkv: torch.Tensor, bias_table: torch.Tensor, heads: int, window: int) -> torch.Tensor:
    _chk(qkv, name="qkv")
    _chk(bias_table, name="bias_table")
    B, T, three_hd = qkv.shape
    D = three_hd // (3 * heads)
    if tuple(bias_table.shape) != (heads, 2 * window):
        raise ValueError(f"bias_table must be (heads, 2*window), got {tuple(bias_table.shape)}")
    out = torch.empty((B, T, heads * D), device=qkv.device, dtype=torch.float32)
    _count()
    with _hook("local_attention", _nbytes(qkv, out)), torch.cuda.device(qkv.device):
        check(_lib.load().l3ac_local_attention_f32(_ptr(qkv), _ptr(bias_table), B, T, heads, D, window, _ptr(out),
                                                   _stream(qkv)), "l3ac_local_attention_f32")
    return out


# "tcgen05": l3ac_local_attention_umma everywhere (the product path).  "auto" keeps the token-rate layers (T < 897: 1-5 key tiles
# per CTA, set-up bound; 35 vs 37 us per 24 clips) on the register-level mma.sync kernel; "mma_sync" forces that kernel.
ATTENTION_IMPL = os.environ.get("L3AC_ATT_IMPL", "tcgen05")


def local_attention_tc(qkv, bias_table: torch.Tensor, heads: int, window: int, out_dtype=torch.float32, impl: Optional[str] = None):
    """Tensor-core attention over bf16 q/k/v: ``qkv`` is a bf16 tensor or a ``Split`` pair of shape (B, T, 3*heads*32).
    impl = "tcgen05": l3ac_local_attention_umma (TMEM accumulators); "mma_sync": the register-level l3ac_local_attention_tc."""
    split = isinstance(qkv, Split)
    hi = qkv.hi if split else qkv
    _chk(hi, torch.bfloat16, "qkv")
    if split:
        _chk(qkv.lo, torch.bfloat16, "qkv.lo")
    _chk(bias_table, name="bias_table")
    B, T, three_hd = hi.shape
    D = three_hd // (3 * heads)
    if tuple(bias_table.shape) != (heads, 2 * window):
        raise ValueError(f"bias_table must be (heads, 2*window), got {tuple(bias_table.shape)}")
    out, o_hi, o_lo = _empty_act((B, T, heads * D), hi.device, out_dtype)
    _count()
    # useful MACs: query p sees (w if p >= w else 0) + (p mod w) + 1 keys; two products of D MACs each
    keys = sum((window if p >= window else 0) + (p % window) + 1 for p in range(T)) if OP_HOOK is not None else 0
    which = impl or ATTENTION_IMPL
    if which == "auto":
        which = "tcgen05" if T >= 897 else "mma_sync"
    with _hook(("local_attention_tc_split" if split else "local_attention_tc") + ("_umma" if which == "tcgen05" else ""), _nbytes(qkv, out),
               2.0 * 2 * B * heads * keys * D * (1 if not split else 1)), torch.cuda.device(hi.device):
        lib = _lib.load()
        fn = lib.l3ac_local_attention_umma if which == "tcgen05" else lib.l3ac_local_attention_tc
        check(fn(_ptr(hi), _ptr(qkv.lo) if split else None, _ptr(bias_table), B, T, heads, D,
                 window, _ptr(o_hi), _ptr(o_lo), _DT[out_dtype], _stream(hi)), "l3ac_local_attention_tc")
    return out


def rotary_pack(qkv: torch.Tensor, heads: int, window: int, cos_table: torch.Tensor, sin_table: torch.Tensor,
                out_dtype=torch.float32):
    """fp32 (B, T, 3*heads*32) q|k|v -> rotated per-window segments (B*ceil(T/window), 2*window, 3*heads*32) of the requested kind."""
    _chk(qkv, name="qkv")
    _chk(cos_table, name="cos_table")
    _chk(sin_table, name="sin_table")
    B, T, three_hd = qkv.shape
    D = three_hd // (3 * heads)
    if tuple(cos_table.shape) != (2 * window, D) or tuple(sin_table.shape) != (2 * window, D):
        raise ValueError(f"rotary tables must be (2*window, {D})")
    nw = -(-T // window)
    out, hi, lo = _empty_act((B * nw, 2 * window, three_hd), qkv.device, out_dtype)
    _count()
    with _hook("rotary_pack", _nbytes(qkv, out)), torch.cuda.device(qkv.device):
        check(_lib.load().l3ac_rotary_pack(_ptr(qkv), B, T, heads, D, window, _ptr(cos_table), _ptr(sin_table), _ptr(hi),
                                           _ptr(lo), _DT[out_dtype], _stream(qkv)), "l3ac_rotary_pack")
    return out


def rotary_unpack(seg, B: int, T: int, window: int):
    """Segment attention output (B*ceil(T/window), 2*window, C) -> (B, T, C); same kind as ``seg`` (tensor or Split)."""
    planes = (seg.hi, seg.lo) if isinstance(seg, Split) else (seg,)
    outs = []
    for pl in planes:
        if not pl.is_cuda or not pl.is_contiguous():
            raise ValueError("seg must be a contiguous CUDA tensor")
        o = torch.empty((B, T, pl.shape[-1]), device=pl.device, dtype=pl.dtype)
        _count()
        with _hook("rotary_unpack", _nbytes(o, o)), torch.cuda.device(pl.device):
            check(_lib.load().l3ac_rotary_unpack(_ptr(pl), _ptr(o), B, T, window, pl.shape[-1] * pl.element_size(),
                                                 _stream(pl)), "l3ac_rotary_unpack")
        outs.append(o)
    return Split(*outs) if isinstance(seg, Split) else outs[0]


def fsq_quantize(x: torch.Tensor, w_in, b_in, w_out, b_out, levels: Sequence[int], want_z: bool = False):
    _chk(x, name="x")
    F = x.shape[-1]
    lead = x.shape[:-1]
    M = x.numel() // F
    D = len(levels)
    q = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    idx = torch.empty(lead, device=x.device, dtype=torch.int32)
    lvl = torch.empty((*lead, D), device=x.device, dtype=torch.float32)
    z = torch.empty((*lead, D), device=x.device, dtype=torch.float32) if want_z else None
    _count()
    with _hook("fsq_quantize", _nbytes(x, q, idx, lvl)), torch.cuda.device(x.device):
        check(_lib.load().l3ac_fsq_quantize(_ptr(x), M, F, _ptr(w_in), _ptr(b_in), _ptr(w_out), _ptr(b_out),
                                            _levels(levels), D, _ptr(q), _ptr(idx), _ptr(lvl), _ptr(z), _stream(x)),
              "l3ac_fsq_quantize")
    return q, idx, lvl, z


def fsq_quantize_latents(z: torch.Tensor, levels: Sequence[int]):
    _chk(z, name="z")
    D = len(levels)
    if z.shape[-1] != D:
        raise ValueError("last dim of z must equal len(levels)")
    lead = z.shape[:-1]
    q = torch.empty(z.shape, device=z.device, dtype=torch.float32)
    idx = torch.empty(lead, device=z.device, dtype=torch.int32)
    lvl = torch.empty(z.shape, device=z.device, dtype=torch.float32)
    _count()
    with _hook("fsq_quantize_latents", _nbytes(z, q, idx, lvl)), torch.cuda.device(z.device):
        check(_lib.load().l3ac_fsq_quantize_latents(_ptr(z), z.numel() // D, _levels(levels), D, _ptr(q), _ptr(idx),
                                                    _ptr(lvl), _stream(z)), "l3ac_fsq_quantize_latents")
    return q, idx, lvl


def fsq_dequantize(indices: torch.Tensor, w_out, b_out, levels: Sequence[int]) -> torch.Tensor:
    if indices.dtype not in (torch.int32, torch.int64):
        raise ValueError(f"indices must be int32 or int64, got {indices.dtype}")
    _chk(indices, indices.dtype, "indices")
    F = w_out.shape[0]
    out = torch.empty((*indices.shape, F), device=indices.device, dtype=torch.float32)
    _count()
    with _hook("fsq_dequantize", _nbytes(indices, out)), torch.cuda.device(indices.device):
        check(_lib.load().l3ac_fsq_dequantize(_ptr(indices), int(indices.dtype == torch.int64), indices.numel(), F,
                                              _ptr(w_out), _ptr(b_out), _levels(levels), len(levels), _ptr(out),
                                              _stream(indices)), "l3ac_fsq_dequantize")
    return out


def upsample_linear_cn(x: torch.Tensor, scale: int, cn_w=None, cn_b=None, eps: float = 1e-8) -> torch.Tensor:
    _chk(x, name="x")
    B, T, Cc = x.shape
    out = torch.empty((B, T * scale, Cc), device=x.device, dtype=torch.float32)
    _count()
    with _hook("upsample_linear_cn", _nbytes(x, out)), torch.cuda.device(x.device):
        check(_lib.load().l3ac_upsample_linear_cn(_ptr(x), B, T, Cc, scale, _ptr(cn_w), _ptr(cn_b), eps, _ptr(out),
                                                  _stream(x)), "l3ac_upsample_linear_cn")
    return out


class UpDwPlan:
    """Host-side parameters of the fused upsample + ChannelNorm + dwconv7 + LayerNorm kernel (``l3ac_updw_plan``)."""

    def __init__(self, scale: int, cn_w, cn_b, cn_eps: float, dw_w, dw_b, ln_w, ln_b, ln_eps: float):
        host = lambda t: t.detach().to("cpu", torch.float32).contiguous()
        cw, cb, dw, db, lw, lb = (host(t) for t in (cn_w, cn_b, dw_w, dw_b, ln_w, ln_b))
        self.C, self.scale = int(db.numel()), int(scale)
        if tuple(dw.shape) != (7, self.C) or cw.numel() != self.C:
            raise ValueError("UpDwPlan: dw_w (7, C) and cn_w (C) expected")
        self.handle = C.c_void_p()
        check(_lib.load().l3ac_updw_plan_create(self.C, self.scale, cw.data_ptr(), cb.data_ptr(), float(cn_eps), dw.data_ptr(), db.data_ptr(),
                                                lw.data_ptr(), lb.data_ptr(), float(ln_eps), C.byref(self.handle)), "l3ac_updw_plan_create")

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                _lib.load().l3ac_updw_plan_destroy(h)
            except Exception:
                pass


def upsample_cn_dwconv7_ln(y: torch.Tensor, plan: UpDwPlan):
    """y (B, T, C) fp32 -> (x_up (B, T*scale, C) fp32, a (B, T*scale, C) bf16): Upsample + ChannelNorm, then dwconv7 + LayerNorm."""
    _chk(y, name="y")
    B, T, Cc = y.shape
    if Cc != plan.C:
        raise ValueError(f"upsample_cn_dwconv7_ln: plan is for C = {plan.C}, got {Cc}")
    xup = torch.empty((B, T * plan.scale, Cc), device=y.device, dtype=torch.float32)
    a = torch.empty((B, T * plan.scale, Cc), device=y.device, dtype=torch.bfloat16)
    _count()
    with _hook("upsample_cn_dwconv7_ln", _nbytes(y, xup, a)), torch.cuda.device(y.device):
        check(_lib.load().l3ac_upsample_cn_dwconv7_ln(plan.handle, _ptr(y), B, T, _ptr(xup), _ptr(a), _stream(y)), "l3ac_upsample_cn_dwconv7_ln")
    return xup, a


def enhance(x: torch.Tensor, conv_w, conv_b, in_w, in_b, merge_w, merge_b, out_dtype=torch.float32,
            stream_branches: bool = True, ch0: Optional[torch.Tensor] = None) -> torch.Tensor:
    """EnhanceBlock.  ``stream_branches``: the stats pass stores the four branch signals (B, T, 4) and the apply pass streams
    them back (default); False recomputes them per tile in the apply pass.  ``ch0``: channel 0 of x as a compact (B, T) plane
    (emitted by the producing kernel); the stats pass then reads it instead of one 128-byte line of x per row."""
    _chk(x, name="x")
    B, T, Cc = x.shape
    lib = _lib.load()
    partials = torch.empty(lib.l3ac_enhance_partials_floats(B, T), device=x.device, dtype=torch.float32)
    branches = torch.empty((B, T, 4), device=x.device, dtype=torch.float32) if stream_branches else None
    out = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    _count(2)
    with _hook("enhance", 2 * x.numel() // x.shape[-1] * 4 + _nbytes(x, out)), torch.cuda.device(x.device):
        st = _stream(x)
        if ch0 is not None and stream_branches:
            _chk(ch0, name="ch0")
            check(lib.l3ac_enhance_stats(_ptr(ch0), B, T, 1, _ptr(conv_w), _ptr(conv_b), _ptr(partials), _ptr(branches), st),
                  "l3ac_enhance_stats")
        else:
            check(lib.l3ac_enhance_stats(_ptr(x), B, T, Cc, _ptr(conv_w), _ptr(conv_b), _ptr(partials), _ptr(branches), st),
                  "l3ac_enhance_stats")
        check(lib.l3ac_enhance_apply(_ptr(x), B, T, Cc, _ptr(conv_w), _ptr(conv_b), _ptr(in_w), _ptr(in_b),
                                     _ptr(merge_w), _ptr(merge_b), _ptr(partials), _ptr(branches), _ptr(out), _DT[out_dtype], st),
              "l3ac_enhance_apply")
    return out


class EnhUpPlan:
    """Weights of the fused EnhanceBlock gate + up 1x1 conv kernel (``l3ac_enhup_plan``; (C_in, C_out) = (48, 24) / (96, 48))."""

    def __init__(self, in_w, in_b, merge_w, merge_b, up_w, up_b, device):
        host = lambda t: t.detach().to("cpu", torch.float32).contiguous()
        iw, ib, mw, mb, uw, ub = (host(t) for t in (in_w, in_b, merge_w, merge_b, up_w, up_b))
        self.C_out, self.C_in = int(uw.shape[0]), int(uw.shape[1])
        if tuple(mw.shape) != (self.C_in, 4) or mb.numel() != self.C_in or ub.numel() != self.C_out:
            raise ValueError("EnhUpPlan: merge_w (C_in, 4), merge_b (C_in), up_w (C_out, C_in), up_b (C_out) expected")
        self.handle = C.c_void_p()
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            check(_lib.load().l3ac_enhup_plan_create(self.C_in, self.C_out, iw.data_ptr(), ib.data_ptr(), mw.data_ptr(), mb.data_ptr(),
                                                     uw.data_ptr(), ub.data_ptr(), C.byref(self.handle)), "l3ac_enhup_plan_create")

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                _lib.load().l3ac_enhup_plan_destroy(h)
            except Exception:
                pass


def enhance_up(x: torch.Tensor, conv_w, conv_b, plan: EnhUpPlan, ch0: Optional[torch.Tensor] = None) -> torch.Tensor:
    """EnhanceBlock + the up layer's 1x1 conv: stats pass, then the fused gate + conv kernel.  x (B, T, C_in) fp32 -> (B, T, C_out) fp32."""
    _chk(x, name="x")
    B, T, Cc = x.shape
    if Cc != plan.C_in:
        raise ValueError(f"enhance_up: plan is for C_in = {plan.C_in}, got {Cc}")
    lib = _lib.load()
    partials = torch.empty(lib.l3ac_enhance_partials_floats(B, T), device=x.device, dtype=torch.float32)
    branches = torch.empty((B, T, 4), device=x.device, dtype=torch.float32)
    out = torch.empty((B, T, plan.C_out), device=x.device, dtype=torch.float32)
    _count(2)
    with _hook("enhance_up", 2 * B * T * 4 * 4 + _nbytes(x, out), 2.0 * B * T * plan.C_in * plan.C_out), torch.cuda.device(x.device):
        st = _stream(x)
        if ch0 is not None:
            check(lib.l3ac_enhance_stats(_ptr(ch0), B, T, 1, _ptr(conv_w), _ptr(conv_b), _ptr(partials), _ptr(branches), st), "l3ac_enhance_stats")
        else:
            check(lib.l3ac_enhance_stats(_ptr(x), B, T, Cc, _ptr(conv_w), _ptr(conv_b), _ptr(partials), _ptr(branches), st), "l3ac_enhance_stats")
        check(lib.l3ac_enhance_up(plan.handle, _ptr(x), B, T, _ptr(partials), _ptr(branches), _ptr(out), st), "l3ac_enhance_up")
    return out


def tail_conv_tanh(x: torch.Tensor, alpha, w, bias: float) -> torch.Tensor:
    _chk(x, name="x")
    B, T, Cc = x.shape
    out = torch.empty((B, T), device=x.device, dtype=torch.float32)
    _count()
    with _hook("tail_conv_tanh", _nbytes(x, out)), torch.cuda.device(x.device):
        check(_lib.load().l3ac_tail_conv_tanh(_ptr(x), B, T, Cc, _ptr(alpha), _ptr(w), float(bias), _ptr(out),
                                              _stream(x)), "l3ac_tail_conv_tanh")
    return out


def pack_mma_b_fragments(w: torch.Tensor, k_pad: int = 32) -> torch.Tensor:
    """(N, K) fp32 weight -> bf16 mma.m16n8k16 B fragments [K_pad/16 ksteps][N/8 ntiles][32 lanes][4] (see the header)."""
    N, K = w.shape
    if N % 8 or k_pad % 16 or K > k_pad:
        raise ValueError("pack_mma_b_fragments needs N % 8 == 0 and K <= k_pad, k_pad % 16 == 0")
    wp = torch.zeros((N, k_pad), dtype=torch.float32, device=w.device)
    wp[:, :K] = w
    lane = torch.arange(32, device=w.device)
    n = (torch.arange(N // 8, device=w.device)[:, None] * 8 + lane[None, :] // 4)              # (ntiles, 32)
    k0 = (torch.arange(k_pad // 16, device=w.device)[:, None] * 16 + (lane[None, :] % 4) * 2)  # (ksteps, 32)
    offs = torch.tensor([0, 1, 8, 9], device=w.device)
    kk = k0[:, None, :, None] + offs                                                           # (ksteps, 1, 32, 4)
    nn = n[None, :, :, None].expand(k_pad // 16, N // 8, 32, 4)
    return wp[nn, kk.expand_as(nn)].to(torch.bfloat16).contiguous()


def decoder_tail(x: torch.Tensor, conv_frags, conv_bias, pw_frags, pw_bias, alpha0, alpha1, dilations, alpha_f, w_f,
                 bias_f: float) -> torch.Tensor:
    _chk(x, name="x")
    B, T, Cc = x.shape
    out = torch.empty((B, T), device=x.device, dtype=torch.float32)
    dil = (C.c_int * 3)(*[int(d) for d in dilations])
    _count()
    with _hook("decoder_tail", _nbytes(x, out), 2.0 * B * T * (3 * (7 * 24 * 24 + 24 * 24) + 7 * 24)), torch.cuda.device(x.device):
        check(_lib.load().l3ac_decoder_tail(_ptr(x), B, T, Cc, _ptr(conv_frags), _ptr(conv_bias), _ptr(pw_frags), _ptr(pw_bias),
                                            _ptr(alpha0), _ptr(alpha1), dil, _ptr(alpha_f), _ptr(w_f), float(bias_f), _ptr(out),
                                            _stream(x)), "l3ac_decoder_tail")
    return out


class TailPlan:
    """Packed weights of the tcgen05 decoder tail (``l3ac_tail_plan``): built once from folded fp32 host arrays."""

    def __init__(self, conv_w, conv_b, pw_w, pw_b, alpha0, alpha1, dilations, alpha_f, w_f, bias_f: float, device):
        host = lambda t: t.detach().to("cpu", torch.float32).contiguous()
        self._keep = [host(t) for t in (conv_w, conv_b, pw_w, pw_b, alpha0, alpha1, alpha_f, w_f)]
        cw, cb, pw, pb, a0, a1, af, wf = self._keep
        if cw.shape != (3, 24, 24, 7) or pw.shape != (3, 24, 24) or wf.shape != (7, 24):
            raise ValueError("TailPlan: conv_w (3,24,24,7), pw_w (3,24,24), w_f (7,24) expected")
        dil = (C.c_int * 3)(*[int(d) for d in dilations])
        self.handle = C.c_void_p()
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            check(_lib.load().l3ac_tail_plan_create(cw.data_ptr(), cb.data_ptr(), pw.data_ptr(), pb.data_ptr(), a0.data_ptr(),
                                                    a1.data_ptr(), dil, af.data_ptr(), wf.data_ptr(), float(bias_f), 24,
                                                    C.byref(self.handle)), "l3ac_tail_plan_create")

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                _lib.load().l3ac_tail_plan_destroy(h)
            except Exception:
                pass


def decoder_tail_tc(x: torch.Tensor, plan: TailPlan, split: bool = False) -> torch.Tensor:
    """Fused decoder tail on tcgen05.  ``split``: 3-term split-bf16 operands (fp32-class) instead of plain bf16."""
    _chk(x, name="x")
    B, T, Cc = x.shape
    if Cc != 24:
        raise ValueError("decoder_tail_tc: 24 channels expected")
    out = torch.empty((B, T), device=x.device, dtype=torch.float32)
    _count()
    name = "decoder_tail_tc_split" if split else "decoder_tail_tc"
    with _hook(name, _nbytes(x, out), 2.0 * B * T * (3 * (7 * 24 * 24 + 24 * 24) + 7 * 24)), torch.cuda.device(x.device):
        fn = _lib.load().l3ac_decoder_tail_tc_split if split else _lib.load().l3ac_decoder_tail_tc
        check(fn(plan.handle, _ptr(x), B, T, _ptr(out), _stream(x)), "l3" + "ac_" + name)
    return out


class StemPlan:
    """Packed weights of the tcgen05 encoder stem (``l3ac_stem_plan``): built once from folded fp32 tensors."""

    def __init__(self, branch_w, branch_b, w1, b1, w2, b2, device):
        host = lambda t: t.detach().to("cpu", torch.float32).contiguous()
        self._keep = [host(t) for t in (branch_w, branch_b, w1, b1, w2, b2)]
        bw, bb, w1h, b1h, w2h, b2h = self._keep
        if bw.numel() != 140 or bb.numel() != 20 or tuple(w1h.shape) != (80, 20) or tuple(w2h.shape) != (24, 81):
            raise ValueError("StemPlan: branch_w (5,4,7), branch_b (20), w1 (80,20), w2 (24,81) expected")
        self.handle = C.c_void_p()
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            check(_lib.load().l3ac_stem_plan_create(bw.data_ptr(), bb.data_ptr(), w1h.data_ptr(), b1h.data_ptr(), w2h.data_ptr(),
                                                    b2h.data_ptr(), 24, C.byref(self.handle)), "l3ac_stem_plan_create")

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                _lib.load().l3ac_stem_plan_destroy(h)
            except Exception:
                pass


def stem_umma(audio: torch.Tensor, plan: StemPlan) -> torch.Tensor:
    _chk(audio, name="audio")
    B, T = audio.shape
    out = torch.empty((B, T, 24), device=audio.device, dtype=torch.float32)
    _count()
    with _hook("stem_umma", _nbytes(audio, out), 2.0 * audio.numel() * 3700), torch.cuda.device(audio.device):
        check(_lib.load().l3ac_stem_umma(plan.handle, _ptr(audio), B, T, _ptr(out), _stream(audio)), "l3ac_stem_umma")
    return out


class ConvUnitPlan:
    """Packed weights of the tcgen05 thin ConvUnit (``l3ac_convunit_plan``, C = 24 / 48): built once from folded fp32 tensors."""

    def __init__(self, dw_w, dw_b, ln_w, ln_b, eps: float, w1, b1, alpha, scale, shift, w2, b2, device):
        host = lambda t: t.detach().to("cpu", torch.float32).contiguous()
        self._keep = [host(t) for t in (dw_w, dw_b, ln_w, ln_b, w1, b1, alpha, scale, shift, w2, b2)]
        dw, db, lw, lb, w1h, b1h, al, sc, sh, w2h, b2h = self._keep
        self.C = int(db.numel())
        if tuple(dw.shape) != (7, self.C) or tuple(w1h.shape) != (4 * self.C, self.C) or tuple(w2h.shape) != (self.C, 4 * self.C):
            raise ValueError("ConvUnitPlan: dw_w (7,C), w1 (4C,C), w2 (C,4C) expected")
        self.handle = C.c_void_p()
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            check(_lib.load().l3ac_convunit_plan_create(self.C, dw.data_ptr(), db.data_ptr(), lw.data_ptr(), lb.data_ptr(), float(eps),
                                                        w1h.data_ptr(), b1h.data_ptr(), al.data_ptr(), sc.data_ptr(), sh.data_ptr(),
                                                        w2h.data_ptr(), b2h.data_ptr(), C.byref(self.handle)), "l3ac_convunit_plan_create")

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                _lib.load().l3ac_convunit_plan_destroy(h)
            except Exception:
                pass


def convunit_umma(x: torch.Tensor, plan: ConvUnitPlan, out_dtype=torch.float32):
    """Fused Residual(ConvUnit) for C = 24 / 48 on tcgen05: x (B, T, C) fp32 -> same shape (fp32, or the split pair)."""
    _chk(x, name="x")
    B, T, Cc = x.shape
    if Cc != plan.C:
        raise ValueError(f"convunit_umma: plan is for C = {plan.C}, got {Cc}")
    if out_dtype not in (torch.float32, SPLIT):
        raise ValueError("convunit_umma emits fp32 or the split-bf16 pair")
    out, hi, lo = _empty_act(tuple(x.shape), x.device, out_dtype)
    _count()
    with _hook("convunit_thin_umma", _nbytes(x) + B * T * Cc * 4, 2.0 * B * T * (7 * Cc + 8 * Cc * Cc)), torch.cuda.device(x.device):
        check(_lib.load().l3ac_convunit_umma(plan.handle, _ptr(x), B, T, _ptr(hi), _ptr(lo), _DT[out_dtype], _stream(x)), "l3ac_convunit_umma")
    return out


class NativeCodec:
    """Step-level handle (``l3ac_codec``, csrc/codec.cu): the library packs the reference's checkpoint tensors and owns the
    launch sequence of ``encode_audio`` / ``decode_audio`` in the product precision.  Torch provides the device buffers
    (inputs, outputs, one workspace per call) and the stream; everything else happens behind ``l3ac_encode`` / ``l3ac_decode``."""

    def __init__(self, mc, weights, device, precision: str = "bf16"):
        lib = _lib.load()
        cfg = _lib.CodecConfig()
        cfg.precision = {"bf16": 0, "split": 1}[precision]
        cfg.feature_dim = mc.feature_dim
        cfg.n_encoder_stages = len(mc.encoder_dims)
        cfg.n_decoder_stages = len(mc.decoder_dims)
        if max(cfg.n_encoder_stages, cfg.n_decoder_stages) > _lib.MAX_STAGES or len(mc.levels) > 8:
            raise ValueError("NativeCodec: too many stages / levels")
        for name in ("encoder_dims", "encoder_depths", "compress_rates", "decoder_dims", "decoder_depths", "decode_rates"):
            for i, v in enumerate(getattr(mc, name)):
                getattr(cfg, name)[i] = int(v)
        cfg.en_coder_depth, cfg.en_coder_window_size = mc.en_coder_depth, mc.en_coder_window_size
        cfg.en_coder_compress_rate, cfg.en_coder_dynamic_pos = mc.en_coder_compress_rate, int(mc.en_coder_dynamic_pos)
        cfg.n_levels = len(mc.levels)
        for i, v in enumerate(mc.levels):
            cfg.levels[i] = int(v)
        keep, names = [], []
        for mod, sd in weights.items():
            for key, t in sd.items():
                keep.append(t.detach().to("cpu", torch.float32).contiguous())
                names.append(f"{mod}.{key}".encode())
        arr = (_lib.Tensor * len(keep))()
        for i, (n, t) in enumerate(zip(names, keep)):
            arr[i].name, arr[i].data, arr[i].numel = n, t.data_ptr(), t.numel()
        self.handle = C.c_void_p()
        self.device = torch.device(device)
        self.feature_dim, self.n_levels = mc.feature_dim, len(mc.levels)
        with torch.cuda.device(self.device):
            rc = lib.l3ac_create(C.byref(cfg), arr, len(keep), C.byref(self.handle))
        if rc != 0:
            raise (ValueError if rc in (-1, -2) else RuntimeError)(
                f"l3ac_create failed: {lib.l3ac_last_error().decode()} ({lib.l3ac_error_string(rc).decode()}, code {rc})")
        self.hop = lib.l3ac_hop_length(self.handle)
        self._ws = {}

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                _lib.load().l3ac_destroy(h)
            except Exception:
                pass

    def _workspace(self, B: int, T: int) -> torch.Tensor:
        n = self._ws.get((B, T))
        if n is None:
            n = _lib.load().l3ac_workspace_bytes(self.handle, B, T)
            if n < 0:
                raise RuntimeError(f"l3ac_workspace_bytes failed: {_lib.load().l3ac_last_error().decode()}")
            if len(self._ws) > 4096:
                self._ws.clear()
            self._ws[(B, T)] = n
        return torch.empty(n, device=self.device, dtype=torch.uint8)

    def _call(self, what, fn, *args):
        lib = _lib.load()
        n0 = lib.l3ac_launch_count(self.handle)
        with torch.cuda.device(self.device):
            rc = fn(*args)
        _count(lib.l3ac_launch_count(self.handle) - n0)
        if rc != 0:
            raise (ValueError if rc in (-1, -2) else RuntimeError)(
                f"{what} failed: {lib.l3ac_last_error().decode()} ({lib.l3ac_error_string(rc).decode()}, code {rc})")

    def encode(self, audio: torch.Tensor):
        """audio (B, T) fp32 on the device -> (q_feature (B, T_tok, F), indices (B, T_tok) int32, level_indices (B, T_tok, D))."""
        _chk(audio, name="audio")
        B, T = audio.shape
        t_tok = -(-T // self.hop)
        q = torch.empty((B, t_tok, self.feature_dim), device=audio.device, dtype=torch.float32)
        idx = torch.empty((B, t_tok), device=audio.device, dtype=torch.int32)
        lvl = torch.empty((B, t_tok, self.n_levels), device=audio.device, dtype=torch.float32)
        ws = self._workspace(B, T)
        self._call("l3ac_encode", _lib.load().l3ac_encode, self.handle, _ptr(audio), B, T, _ptr(ws), ws.numel(), _ptr(q), _ptr(idx),
                   _ptr(lvl), _stream(audio))
        return q, idx, lvl

    def decode(self, feat: Optional[torch.Tensor] = None, indices: Optional[torch.Tensor] = None) -> torch.Tensor:
        """q_feature (B, T_tok, F) fp32, or indices (B, T_tok) int32 / int64 -> audio (B, T_tok * hop) fp32."""
        src = feat if feat is not None else indices
        if feat is not None:
            _chk(feat, name="audio_feature")
        else:
            if indices.dtype not in (torch.int32, torch.int64):
                raise ValueError(f"indices must be int32 or int64, got {indices.dtype}")
            _chk(indices, indices.dtype, "indices")
        B, t_tok = src.shape[0], src.shape[1]
        wav = torch.empty((B, t_tok * self.hop), device=src.device, dtype=torch.float32)
        ws = self._workspace(B, t_tok * self.hop)
        self._call("l3ac_decode", _lib.load().l3ac_decode, self.handle, _ptr(indices) if feat is None else None,
                   int(feat is None and indices.dtype == torch.int64), _ptr(feat), B, t_tok, _ptr(ws), ws.numel(), _ptr(wav), _stream(src))
        return wav
