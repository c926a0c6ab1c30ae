"""How long does the host take to ISSUE one encode+decode step (no sync) vs the GPU time of the step?  (development probe)"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import l3ac_b200  # noqa: E402

codec = l3ac_b200.get_model("1kbps", pretrained=False)
codec.network.cuda()
audio = (0.1 * torch.randn(64, 160000)).clamp(-1, 1).cuda()
with torch.inference_mode():
    for _ in range(3):
        q, idx = codec.encode_audio(audio)
        codec.decode_audio(indices=idx["indices"])
    torch.cuda.synchronize()
    for _ in range(3):
        t0 = time.perf_counter()
        q, idx = codec.encode_audio(audio)
        t1 = time.perf_counter()
        wav = codec.decode_audio(indices=idx["indices"])
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        print(f"issue encode {1e3 * (t1 - t0):.2f} ms, issue decode {1e3 * (t2 - t1):.2f} ms, drain {1e3 * (t3 - t2):.2f} ms, "
              f"total {1e3 * (t3 - t0):.2f} ms", flush=True)
