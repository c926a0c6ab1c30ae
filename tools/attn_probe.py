"""Times the tensor-core local attention on the 1kbps shapes of one 24-clip micro-batch (development probe)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from l3ac_b200 import ops  # noqa: E402

DEV = "cuda:0"
g = torch.Generator().manual_seed(0)
junk = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for T, w in ((1779, 750), (593, 250)):
    B, H, D = 24, 6, 32
    qkv = (torch.randn(B, T, 3 * H * D, generator=g)).to(DEV)
    table = (torch.randn(H, 2 * w, generator=g) * 0.5).to(DEV)
    hi = qkv.to(torch.bfloat16)
    sp = ops.Split(hi, (qkv - hi.float()).to(torch.bfloat16))
    for name, arg in (("bf16", hi), ("split", sp)):
      for impl in ("tcgen05", "mma_sync"):
        ts = []
        for i in range(8):
            junk.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = ops.local_attention_tc(arg, table, H, w, out_dtype=torch.bfloat16 if name == "bf16" else ops.SPLIT, impl=impl)
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"local_attention {impl:8s} {name} B={B} T={T} w={w}: {sorted(ts)[len(ts) // 2]:.1f} us", flush=True)
