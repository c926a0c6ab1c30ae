"""CUDA-event timing of dwconv7 + LayerNorm (bf16 out) at the decode side's thin-stage shapes: lane-group kernel vs the
thread-per-row plan kernel (L2 flushed between iterations by the 0.5-0.7 GB working set itself)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from l3ac_b200 import ops

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g)
B = 32
for C, T in ((48, 80000), (96, 26667)):
    xs = [r(B, T, C) for _ in range(3)]
    w, b, lw, lb = r(7, C), r(C), r(C), r(C)
    plan = ops.DwconvPlan(w, b, lw, lb, 1e-8)
    for name, fn in (("lane-group", lambda x: ops.dwconv7_ln(x, w, b, lw, lb, 1e-8, out_dtype=torch.bfloat16)),
                     ("thread-per-row", lambda x: ops.dwconv7_ln_plan(x, plan))):
        for x in xs:
            fn(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 12
        e0.record()
        for i in range(n):
            fn(xs[i % 3])
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        print(f"C={C} T={T} {name:15s} {us:7.1f} us  {B * T * C * 6 / us / 1e3:6.0f} GB/s")

# the up-layer tail fused with the next unit's prologue vs the two kernels it replaces
for C, T, S in ((48, 26667, 3), (96, 8889, 3)):
    ys = [r(B, T, C) for _ in range(3)]
    cw, cb, w, b, lw, lb = r(C), r(C), r(7, C), r(C), r(C), r(C)
    dplan = ops.DwconvPlan(w, b, lw, lb, 1e-8)
    uplan = ops.UpDwPlan(S, cw, cb, 1e-8, w, b, lw, lb, 1e-8)
    for name, fn in (("upsample_cn + dwconv_plan", lambda y: ops.dwconv7_ln_plan(ops.upsample_linear_cn(y, S, cw, cb, 1e-8), dplan)),
                     ("fused", lambda y: ops.upsample_cn_dwconv7_ln(y, uplan))):
        for y in ys:
            fn(y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 12
        e0.record()
        for i in range(n):
            fn(ys[i % 3])
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        print(f"C={C} T={T} x{S} {name:26s} {us:7.1f} us  {B * T * C * (4 + S * 6) / us / 1e3:6.0f} GB/s (algorithmic bytes of the fused form)")
