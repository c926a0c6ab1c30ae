"""Times the fused thin ConvUnit kernels (fp32 SIMT vs tensor-core split) on one 24-clip micro-batch (development probe)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from l3ac_b200 import ops  # noqa: E402

DEV = "cuda:0"
g = torch.Generator().manual_seed(0)
rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).to(DEV)
junk = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timed(fn, n=6):
    ts = []
    for i in range(n + 2):
        junk.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2], out


for C, T in ((24, 160110), (48, 26685), (48, 80055)):
    B = 24
    x = rnd(B, T, C)
    args = (rnd(7, C, scale=0.3), rnd(C, scale=0.1), 1 + rnd(C, scale=0.1), rnd(C, scale=0.1), 1e-8, rnd(4 * C, C, scale=0.2),
            rnd(4 * C, scale=0.1), (0.5 + torch.rand(4 * C, generator=g)).to(DEV), 1 + rnd(4 * C, scale=0.1), rnd(4 * C, scale=0.1),
            rnd(C, 4 * C, scale=0.1), rnd(C, scale=0.1))
    plan = ops.ConvUnitPlan(*args, DEV)
    for kind in (torch.float32, ops.SPLIT):
        tu, outu = timed(lambda: ops.convunit_umma(x, plan, out_dtype=kind))
        print(f"thin_umma (tcgen05) C={C} rows={B * T} out={'split' if kind == ops.SPLIT else 'f32'}: {tu:.1f} us", flush=True)
        t, out = timed(lambda: ops.convunit_thin_tc(x, *args, out_dtype=kind))
        o = out if kind == torch.float32 else out.hi.float() + out.lo.float()
        msg = f"thin_tc C={C} rows={B * T} out={'split' if kind == ops.SPLIT else 'f32'}: {t:.1f} us ({B * T * C * 8 / t / 1e3:.0f} GB/s)"
        if C == 24:
            t2, ref = timed(lambda: ops.convunit_thin(x, *args, out_dtype=kind))
            r = ref if kind == torch.float32 else ref.hi.float() + ref.lo.float()
            msg += f" | fp32 SIMT kernel {t2:.1f} us, max-abs difference {(o - r).abs().max().item():.2e}"
        print(msg, flush=True)
    t, out = timed(lambda: ops.convunit_thin_tc(x, *args, operands=torch.bfloat16))
    print(f"thin_tc C={C} rows={B * T} bf16 operands, f32 out: {t:.1f} us", flush=True)
