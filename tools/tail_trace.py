"""Pipeline timeline of the tcgen05 decoder-tail kernel (debug tool, not part of the product path).

Builds a private copy of tail_tc.cu with -DL3AC_TAIL_TRACE (CTA 0 stamps clock64() at the hand-overs of its second
tile), runs one launch on 24 x 10 s clips and prints the events in time order (SM cycles since the first event)."""
import ctypes
import pathlib
import subprocess

import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
CSRC = ROOT / "l3ac_b200" / "csrc"
OUT = ROOT / "gpurun_out"
EV = {1: "S1 start", 2: "S1 done (a_ready arrive)", 3: "d_ready acquired", 4: "S2 done (h_ready arrive)", 5: "o_ready acquired",
      6: "S3 done", 7: "final P acquired", 8: "y stored", 10: "mma: a_ready acquired, issue conv", 11: "mma:   conv MMAs issued", 13: "mma:   commit issued", 12: "mma: h_ready acquired, issue pw"}


def main():
    OUT.mkdir(exist_ok=True)
    lib_path = OUT / "libtail_trace.so"
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DL3AC_TAIL_TRACE", "-Xcompiler", "-fPIC",
                    "--expt-relaxed-constexpr", "-shared", "-I", str(ROOT / "include"), str(CSRC / "tail_tc.cu"), str(CSRC / "api.cu"),
                    "-o", str(lib_path)], check=True)
    lib = ctypes.CDLL(str(lib_path))
    g = torch.Generator().manual_seed(0)
    C, B, T = 24, 24, 160110
    r = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).contiguous()
    cw, cb, pw, pb = r(3, C, C, 7, scale=0.08), r(3, C, scale=0.05), r(3, C, C, scale=0.15), r(3, C, scale=0.05)
    a0, a1, af, wf = 0.5 + torch.rand(3, C, generator=g), 0.5 + torch.rand(3, C, generator=g), 0.5 + torch.rand(C, generator=g), r(7, C, scale=0.1)
    dil = (ctypes.c_int * 3)(1, 3, 9)
    plan = ctypes.c_void_p()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.l3ac_tail_plan_create(P(cw), P(cb), P(pw), P(pb), P(a0), P(a1), dil, P(af), P(wf), ctypes.c_float(0.01), 24, ctypes.byref(plan))
    assert rc == 0, rc
    x = (torch.randn(B, T, C, generator=g) * 0.7).cuda()
    out = torch.empty(B, T, device="cuda")
    for _ in range(2):
        rc = lib.l3ac_decoder_tail_tc(plan, P(x), B, T, P(out), None)
        assert rc == 0, rc
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (6 * 256))()
    lib.l3ac_debug_tail_trace(buf)
    events = []
    for role in range(6):
        n = buf[role * 256]
        for i in range(n):
            v = buf[role * 256 + 1 + i]
            events.append((v >> 16, role, (v >> 8) & 0xff, (v >> 4) & 0xf, v & 0xf))
    events.sort()
    t0 = events[0][0]
    for t, role, ev, unit, blk in events:
        who = f"wg{role}" if role < 4 else ("MMA" if role == 4 else "MMA2")
        print(f"{t - t0:8d}  {who:4s} u{unit} b{blk}  {EV.get(ev, ev)}")


if __name__ == "__main__":
    main()
