"""Times the fused decoder tail kernel on one 24-clip micro-batch (development probe; not part of the product path).
Variants are selected with L3AC_TAIL_VARIANT (read once per process)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from l3ac_b200 import ops  # noqa: E402

DEV = "cuda:0"
B, T, C = 24, 160110, 24
g = torch.Generator().manual_seed(0)
rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).to(DEV)
convs = torch.stack([ops.pack_mma_b_fragments(rnd(C, 7 * C, scale=0.08), k_pad=176) for _ in range(3)]).contiguous()
pws = torch.stack([ops.pack_mma_b_fragments(rnd(C, C, scale=0.15)) for _ in range(3)]).contiguous()
cb, pb = rnd(3, C, scale=0.05), rnd(3, C, scale=0.05)
a0, a1 = (0.5 + torch.rand(3, C, generator=g)).to(DEV), (0.5 + torch.rand(3, C, generator=g)).to(DEV)
af, wf = (0.5 + torch.rand(C, generator=g)).to(DEV), rnd(7, C, scale=0.1)
x = rnd(B, T, C, scale=0.7)
cw_ref, pw_ref = rnd(3, C, C, 7, scale=0.08), rnd(3, C, C, scale=0.15)
plan = ops.TailPlan(cw_ref, cb, pw_ref, pb, a0, a1, (1, 3, 9), af, wf, 0.01, DEV)
junk = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
times = []
for i in range(8):
    junk.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ops.decoder_tail(x, convs, cb, pws, pb, a0, a1, (1, 3, 9), af, wf, 0.01)
    e1.record()
    torch.cuda.synchronize()
    if i >= 2:
        times.append(e0.elapsed_time(e1) * 1e3)
t = sorted(times)[len(times) // 2]
times_tc = []
for i in range(8):
    junk.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out_tc = ops.decoder_tail_tc(x, plan)
    e1.record()
    torch.cuda.synchronize()
    if i >= 2:
        times_tc.append(e0.elapsed_time(e1) * 1e3)
t_tc = sorted(times_tc)[len(times_tc) // 2]
print(f"decoder_tail_tc (tcgen05) B={B} T={T}: {t_tc:.1f} us ({B * T * 24 * 4 / t_tc / 1e3:.0f} GB/s of input, "
      f"{B * T * 168 / t_tc / 1e6:.2f} T sin/s, checksum {out_tc.double().sum().item():.6f})", flush=True)
print(f"decoder_tail variant={os.environ.get('L3AC_TAIL_VARIANT', '0')} B={B} T={T}: {t:.1f} us "
      f"({B * T * 24 * 4 / t / 1e3:.0f} GB/s of input, checksum {out.double().sum().item():.6f})", flush=True)
