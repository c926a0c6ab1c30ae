"""Diagnostic: per-128-row-tile error of the fused MLP vs the two-GEMM path (development aid)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from l3ac_b200 import ops
DEV = "cuda:0"
for M, C in ((40000, 256), (40000, 96), (60000, 48), (18944 * 2, 256)):
    torch.manual_seed(0)
    a = torch.randn(M, C, device=DEV).to(torch.bfloat16)
    w1 = (torch.randn(4 * C, C, device=DEV) * C ** -0.5).to(torch.bfloat16)
    w2 = (torch.randn(C, 4 * C, device=DEV) * (4 * C) ** -0.5).to(torch.bfloat16)
    b1, b2 = torch.randn(4 * C, device=DEV) * 0.1, torch.randn(C, device=DEV) * 0.1
    alpha, scale, shift = 0.5 + torch.rand(4 * C, device=DEV), torch.ones(4 * C, device=DEV), torch.zeros(4 * C, device=DEV)
    x = torch.randn(M, C, device=DEV)
    for rep in range(2):
        got = ops.convunit_mlp(a, w1, b1, alpha, scale, shift, w2, b2, x)
        h = ops.gemm(a, w1, B=1, T=M, K=C, bias=b1, act=ops.ACT_SNAKE, alpha=alpha, scale=scale, shift=shift, out_dtype=torch.bfloat16)
        two = ops.gemm(h, w2, B=1, T=M, K=4 * C, bias=b2, residual=x)
        torch.cuda.synchronize()
        err = (got - two).abs().reshape(M, C)
        nt = (M + 127) // 128
        pad = torch.zeros(nt * 128 - M, C, device=DEV)
        per_tile = torch.cat([err, pad]).view(nt, 128, C).amax(dim=(1, 2)).cpu()
        bad = (per_tile > 1e-2).nonzero().flatten().tolist()
        print(f"M={M} C={C} rep={rep}: max {float(err.max()):.3e}; bad tiles {len(bad)}/{nt}; first bad {bad[:12]}; "
              f"bad tile idx mod 148: {sorted(set(b % 148 for b in bad))[:10]} ; bad//148: {sorted(set(b // 148 for b in bad))}")
        if bad:
            t = bad[0]
            e = err[t * 128:(t + 1) * 128]
            rows = (e.amax(dim=1) > 1e-2).nonzero().flatten().tolist()
            cols = (e.amax(dim=0) > 1e-2).nonzero().flatten().tolist()
            print(f"   tile {t}: bad rows {rows[:8]}..{rows[-3:]} (n={len(rows)}), bad cols {cols[:8]}..{cols[-3:]} (n={len(cols)})")
