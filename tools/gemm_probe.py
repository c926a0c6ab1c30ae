"""Times individual tcgen05 GEMM shapes with CUDA events (development probe; not part of the product path)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from l3ac_b200 import ops  # noqa: E402

DEV = "cuda:0"


def run(M, K, N, out_dtype=torch.float32, residual=False, act=ops.ACT_NONE, taps=1, iters=5, flush=True, tag="", split=False):
    a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = (torch.randn(N, taps * K, device=DEV) * 0.05).to(torch.bfloat16)
    if split:
        a, w = ops.Split(a, a.clone()), ops.Split(w, w.clone())
    n_out = N // 2 if act == ops.ACT_GEGLU else N
    res = torch.randn(M, n_out, device=DEV) if residual else None
    alpha = torch.ones(N, device=DEV) if act == ops.ACT_SNAKE else None
    junk = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    times = []
    for i in range(iters + 2):
        if flush:
            junk.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(a, w, B=1, T=M, K=K, taps=taps, tap_shift0=-(taps // 2), act=act, alpha=alpha, residual=res,
                 out_dtype=out_dtype)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            times.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(times)[len(times) // 2]
    flops = 2.0 * M * K * N * taps
    osz = 2 if out_dtype == torch.bfloat16 else 4
    nbytes = 2 * (M * K + N * K * taps) + osz * M * n_out + (4 * M * n_out if residual else 0)
    print(f"{tag:28s} M={M:8d} K={K:5d} N={N:5d} taps={taps} out={'f32' if osz == 4 else 'bf16'} res={int(residual)} act={act} "
          f"{t:8.1f} us  {flops / t / 1e6:7.1f} TF/s  {nbytes / t / 1e3:7.0f} GB/s", flush=True)


def run_mlp(M, C, iters=5):
    a = torch.randn(M, C, device=DEV).to(torch.bfloat16)
    w1 = (torch.randn(4 * C, C, device=DEV) * C ** -0.5).to(torch.bfloat16)
    w2 = (torch.randn(C, 4 * C, device=DEV) * (4 * C) ** -0.5).to(torch.bfloat16)
    b1, b2 = torch.randn(4 * C, device=DEV) * 0.1, torch.randn(C, device=DEV) * 0.1
    alpha, scale, shift = torch.ones(4 * C, device=DEV), torch.ones(4 * C, device=DEV), torch.zeros(4 * C, device=DEV)
    x = torch.randn(M, C, device=DEV)
    junk = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    res = {}
    for name in ("fused", "two_gemm"):
        times = []
        for i in range(iters + 2):
            junk.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if name == "fused":
                ops.convunit_mlp(a, w1, b1, alpha, scale, shift, w2, b2, x)
            else:
                h = ops.gemm(a, w1, B=1, T=M, K=C, bias=b1, act=ops.ACT_SNAKE, alpha=alpha, scale=scale, shift=shift,
                             out_dtype=torch.bfloat16)
                ops.gemm(h, w2, B=1, T=M, K=4 * C, bias=b2, residual=x)
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                times.append(e0.elapsed_time(e1) * 1e3)
        res[name] = sorted(times)[len(times) // 2]
    flops = 2.0 * M * 2 * 4 * C * C
    print(f"mlp M={M:8d} C={C:4d}  fused {res['fused']:8.1f} us ({flops / res['fused'] / 1e6:6.1f} TF/s)   two-GEMM {res['two_gemm']:8.1f} us "
          f"({flops / res['two_gemm'] / 1e6:6.1f} TF/s)", flush=True)


if __name__ == "__main__":
    only = sys.argv[1] if len(sys.argv) > 1 else None
    if only == "mlp1":      # one shape, for ncu: python tools/gemm_probe.py mlp1 M C
        run_mlp(int(sys.argv[2]), int(sys.argv[3]), iters=1)
        sys.exit(0)
    if only == "mlp":
        for M, C in ((142320, 256), (426960, 96), (1280880, 48), (28464, 256), (2561760, 48)):
            run_mlp(M, C)
        sys.exit(0)
    thin = {
        "t_k96_n24": dict(M=2401650, K=96, N=24),
        "t_k96_n24_res": dict(M=2401650, K=96, N=24, residual=True),
        "t_k96_n24_split": dict(M=2401650, K=96, N=24, split=True),
        "t_k96_n24_split_res": dict(M=2401650, K=96, N=24, split=True, residual=True),
        "t_k64_n24": dict(M=2401650, K=64, N=24),
        "t_k128_n24": dict(M=2401650, K=128, N=24),
        "t_k192_n24": dict(M=2401650, K=192, N=24),
        "t_k96_n32": dict(M=2401650, K=96, N=32),
        "t_k96_n64": dict(M=2401650, K=96, N=64),
        "t_k96_n96": dict(M=2401650, K=96, N=96),
        "t_k24_n96_split": dict(M=2401650, K=24, N=96, split=True, out_dtype=torch.bfloat16),
        "t_k192_n48_res": dict(M=400275, K=192, N=48, residual=True),
        "t_k192_n48_split_res": dict(M=400275, K=192, N=48, residual=True, split=True),
        "t_k32_n24_k7": dict(M=2401650, K=32, N=24, taps=7, out_dtype=torch.bfloat16),
    }
    if only == "thin":
        for name, kw in thin.items():
            run(tag=name, **kw)
        sys.exit(0)
    cases = {
        "c256_pw2": dict(M=142320, K=1024, N=256, residual=True),
        "c256_pw2_nores": dict(M=142320, K=1024, N=256),
        "c256_pw2_bf16": dict(M=142320, K=1024, N=256, out_dtype=torch.bfloat16),
        "c256_pw2_smallM": dict(M=18944, K=1024, N=256, residual=True),
        "c256_pw1": dict(M=142320, K=256, N=1024, out_dtype=torch.bfloat16, act=ops.ACT_SNAKE),
        "c256_pw1_noact": dict(M=142320, K=256, N=1024, out_dtype=torch.bfloat16),
        "c512_pw1": dict(M=28464, K=512, N=2048, out_dtype=torch.bfloat16, act=ops.ACT_SNAKE),
        "c512_pw2": dict(M=28464, K=2048, N=512, residual=True),
        "c96_pw1": dict(M=426960, K=96, N=384, out_dtype=torch.bfloat16, act=ops.ACT_SNAKE),
        "c96_pw2": dict(M=426960, K=384, N=96, residual=True),
        "c48_pw1": dict(M=1280880, K=48, N=192, out_dtype=torch.bfloat16, act=ops.ACT_SNAKE),
        "c24_pw1_split": dict(M=2561760, K=24, N=96, split=True, out_dtype=ops.SPLIT, act=ops.ACT_SNAKE),
        "c24_1x1": dict(M=2561760, K=24, N=24, residual=True),
        "c24_k7": dict(M=2561760, K=24, N=24, taps=7, out_dtype=torch.bfloat16, act=ops.ACT_SNAKE),
        "square_4k": dict(M=4096, K=4096, N=4096, out_dtype=torch.bfloat16),
        "square_8k": dict(M=8192, K=8192, N=8192, out_dtype=torch.bfloat16),
    }
    for name, kw in cases.items():
        if only is None or only == name:
            run(tag=name, **kw)
