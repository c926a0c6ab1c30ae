"""Times the encoder stem kernels (fp32 SIMT vs tensor-core split) on one 24-clip micro-batch (development probe)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from l3ac_b200 import ops  # noqa: E402

DEV = "cuda:0"
g = torch.Generator().manual_seed(0)
rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).to(DEV)
junk = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
B, T = 24, 160110
args = (rnd(B, T, scale=0.1), rnd(5, 4, 7, scale=0.3), rnd(20, scale=0.1), rnd(80, 20, scale=0.2), rnd(80, scale=0.1),
        rnd(24, 81, scale=0.2), rnd(24, scale=0.1))
res = {}
plan = ops.StemPlan(*args[1:], DEV)
for name, fn in (("stem_umma", lambda *a: ops.stem_umma(a[0], plan)), ("stem_tc", ops.stem_tc), ("stem", ops.stem)):
    ts = []
    for i in range(8):
        junk.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res[name] = fn(*args)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"{name}: {sorted(ts)[len(ts) // 2]:.1f} us for {B * T} samples", flush=True)
print(f"umma vs fp32 SIMT: {(res['stem_umma'] - res['stem']).abs().max().item():.2e}")
print(f"max-abs difference {(res['stem_tc'] - res['stem']).abs().max().item():.2e} (values up to {res['stem'].abs().max().item():.2f})")
