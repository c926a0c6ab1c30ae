"""Pipeline timeline of the fused ConvUnit-MLP kernel (debug tool, not part of the product path).

Builds a private copy of mlp_fused.cu with -DL3AC_MLP_TRACE (CTA 0 stamps clock64() at every barrier hand-over of two
steady-state tiles), runs one launch and prints the events in time order, in microseconds at the measured SM clock.

    python tools/mlp_trace.py M C
"""
import ctypes
import pathlib
import subprocess
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
CSRC = ROOT / "l3ac_b200" / "csrc"
OUT = ROOT / "gpurun_out"

EVENTS = {
    1: "prod1 a_empty acquired (A load issued)",
    10: "mma1  a_full acquired",
    12: "mma2  a2_full (+ d2_empty) acquired",
    13: "mma2  GEMM2 issued + commit",
    15: "mma1  d1_empty acquired",
    16: "mma1  GEMM1 issued + commit",
    17: "mma1    slot acquired, before MMAs",
    18: "mma1    MMAs issued, before commits",
}
for g in range(4):
    EVENTS[20 + g] = f"epi{g}  wait d1_full ..."
    EVENTS[30 + g] = f"epi{g}  d1_full acquired"
    EVENTS[40 + g] = f"epi{g}  chunk written (a2_full arrive)"
    EVENTS[50 + g] = f"epi{g}  wait d2_full ..."
    EVENTS[60 + g] = f"epi{g}  d2_full acquired"
    EVENTS[70 + g] = f"epi{g}  output done (d2_empty arrive)"
    EVENTS[80 + g] = f"epi{g}    tmem_ld done, pass j"
    EVENTS[90 + g] = f"epi{g}    math + st.shared done"


def build():
    OUT.mkdir(exist_ok=True)
    lib = OUT / "libmlp_trace.so"
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DL3AC_MLP_TRACE", "-Xcompiler", "-fPIC",
           "-shared", "-I", str(ROOT / "include"), str(CSRC / "mlp_fused.cu"), str(CSRC / "api.cu"), "-o", str(lib), "-lcuda"]
    subprocess.run(cmd, check=True)
    return ctypes.CDLL(str(lib))


def main():
    M, C = int(sys.argv[1]), int(sys.argv[2])
    lib = build()
    dev = "cuda:0"
    a = torch.randn(M, C, device=dev).to(torch.bfloat16)
    w1 = (torch.randn(4 * C, C, device=dev) * C ** -0.5).to(torch.bfloat16)
    w2 = (torch.randn(C, 4 * C, device=dev) * (4 * C) ** -0.5).to(torch.bfloat16)
    f = lambda n, v: torch.full((n,), v, device=dev)
    b1, alpha, ialpha, scale, shift, b2 = f(4 * C, 0.1), f(4 * C, 1.0), f(4 * C, 1.0), f(4 * C, 1.0), f(4 * C, 0.0), f(C, 0.1)
    x = torch.randn(M, C, device=dev)
    out = torch.empty_like(x)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    buf = (ctypes.c_ulonglong * (8 * 512))()
    for rep in range(3):
        rc = lib.l3ac_convunit_mlp_tc(P(a), P(w1), P(b1), P(alpha), P(ialpha), P(scale), P(shift), P(w2), P(b2), P(x), P(out),
                                      ctypes.c_longlong(M), ctypes.c_int(C), ctypes.c_void_p(0))
        assert rc == 0, rc
        lib.l3ac_debug_mlp_trace(buf)
    raw = [buf[r * 512 + 1 + i] for r in range(8) for i in range(min(int(buf[r * 512]), 511))]
    n = len(raw)
    ev = sorted((v >> 20, (v >> 16) & 15, (v >> 8) & 255, v & 255) for v in raw)
    t0 = ev[0][0]
    mhz = 1900.0
    print(f"# M={M} C={C}: {n} events, times in us at {mhz:.0f} MHz")
    for t, it, e, j in ev:
        print(f"{(t - t0) / mhz:9.3f}  tile {it}  {EVENTS.get(e, e):44s} j={j}")


if __name__ == "__main__":
    main()
