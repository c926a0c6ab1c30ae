"""Runs BASELINE.json configs #2-#5 through bench.py on N GPUs of this box and collects the JSON lines.

    python tools/run_configs.py --gpus N [--out gpurun_out/r02_configs_nN.json] [--only 3a,3b,4,5]

#2  1kbps,  64 x 10 s per GPU (weak; the driver's headline)          #3a/#3b  0k75bps / 1k5bps, 256 x 10 s in total (strong)
#4  3kbps, 256 x 30 s in total (strong)                               #5  decode-from-indices sweep, 1-60 s x 1-1024 clips (strong)
N > 1 launches torchrun (one rank per GPU, NCCL), as the driver does."""
import argparse
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CASES = {
    "2": ["--config", "1kbps", "--batch", "64", "--seconds", "10"],
    "3a": ["--config", "0k75bps", "--batch", "256", "--seconds", "10", "--scaling", "strong"],
    "3b": ["--config", "1k5bps", "--batch", "256", "--seconds", "10", "--scaling", "strong"],
    "4": ["--config", "3kbps", "--batch", "256", "--seconds", "30", "--scaling", "strong"],
    "5": ["--config", "1kbps", "--mode", "decode-sweep"],
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default="3a,3b,4,5")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    out = {"n_gpus": a.gpus, "lines": {}}
    for i, key in enumerate(a.only.split(",")):
        base = [sys.executable] if a.gpus == 1 else [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
                                                     "--master-addr", "127.0.0.1", "--master-port", str(29540 + i)]
        cmd = base + [str(ROOT / "bench.py"), "--gpus", str(a.gpus), "--steps", str(a.steps), "--warmup", str(a.warmup),
                      "--no-cpu-baseline"] + CASES[key]
        r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
        line = None
        for ln in r.stdout.splitlines():
            if ln.startswith("{"):
                try:
                    line = json.loads(ln)
                except json.JSONDecodeError:
                    pass
        if line is None:
            line = {"error": r.stderr[-2000:], "rc": r.returncode}
        else:       # keep the table small: the per-kernel breakdowns live in the bench line of the headline run
            for k in ("roofline_hbm", "roofline_mma_sync", "fp32_simt_gemm"):
                line.pop(k, None)
        out["lines"][key] = line
        print(key, json.dumps({k: line.get(k) for k in ("value", "ms_per_step", "n_gpus", "scaling", "error")}), flush=True)
    path = Path(a.out or ROOT / "gpurun_out" / f"r02_configs_n{a.gpus}.json")
    path.parent.mkdir(exist_ok=True)
    path.write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
