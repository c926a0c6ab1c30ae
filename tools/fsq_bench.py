"""FSQ quantizer / dequantizer throughput (north-star item 3: achieved HBM GB/s of the bottleneck kernels).
Sizes: one 24-clip micro-batch of BASELINE config #2 (14 k tokens) and BASELINE config #4 (256 x 30 s at 3kbps = 1.28 M tokens).
Algorithmic bytes: quantize 1 076 B/token (x in, q_feature + indices + level_indices out),
dequantize 516 B/token.  Prints one JSON object; run on the GPU box:  python tools/fsq_bench.py > gpurun_out/fsq.json"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from l3ac_b200 import ops          # noqa: E402


def time_ms(fn, iters=20, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()                      # > L2: the next launch reads its input from HBM
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = "cuda:0"
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    g = torch.Generator().manual_seed(0)
    out = {"peak_gbs": peak, "cases": []}
    for levels in ((7, 7, 7, 7, 7, 7), (9, 9, 9, 7, 7, 7)):
        w_in = (torch.randn(6, 128, generator=g) * 0.1).to(dev)
        b_in = (torch.randn(6, generator=g) * 0.1).to(dev)
        w_out = (torch.randn(128, 6, generator=g) * 0.4).to(dev)
        b_out = (torch.randn(128, generator=g) * 0.1).to(dev)
        for tokens in (24 * 593, 256 * 5000):
            x = torch.randn(tokens, 128, generator=g).to(dev).view(1, tokens, 128)
            q, idx, lvl, _ = ops.fsq_quantize(x, w_in, b_in, w_out, b_out, levels)
            tq = time_ms(lambda: ops.fsq_quantize(x, w_in, b_in, w_out, b_out, levels), flush=flush)
            td = time_ms(lambda: ops.fsq_dequantize(idx, w_out, b_out, levels), flush=flush)
            assert torch.equal(ops.fsq_dequantize(idx, w_out, b_out, levels), q)
            out["cases"].append({"levels": list(levels), "tokens": tokens,
                                 "quantize_us": tq * 1e3, "quantize_gbs": tokens * 1076 / tq / 1e6, "quantize_frac": tokens * 1076 / tq / 1e6 / peak,
                                 "dequantize_us": td * 1e3, "dequantize_gbs": tokens * 516 / td / 1e6, "dequantize_frac": tokens * 516 / td / 1e6 / peak})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
