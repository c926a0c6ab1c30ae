"""Debug: runs the tcgen05 decoder tail built with -DL3AC_MBAR_WATCHDOG (a stuck mbarrier wait prints its source line and traps)."""
import ctypes, pathlib, subprocess, sys
import torch
ROOT = pathlib.Path(__file__).resolve().parent.parent
CSRC = ROOT / "l3ac_b200" / "csrc"
OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
lib_path = OUT / "libtail_wd.so"
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DL3AC_MBAR_WATCHDOG", "-Xcompiler", "-fPIC",
                "--expt-relaxed-constexpr", "-shared", "-I", str(ROOT / "include"), str(CSRC / "tail_tc.cu"), str(CSRC / "api.cu"),
                "-o", str(lib_path)], check=True)
lib = ctypes.CDLL(str(lib_path))
g = torch.Generator().manual_seed(0)
C = 24
B, T = int(sys.argv[1]), int(sys.argv[2])
r = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).contiguous()
cw, cb, pw, pb = r(3, C, C, 7, scale=0.08), r(3, C, scale=0.05), r(3, C, C, scale=0.15), r(3, C, scale=0.05)
a0, a1, af, wf = 0.5 + torch.rand(3, C, generator=g), 0.5 + torch.rand(3, C, generator=g), 0.5 + torch.rand(C, generator=g), r(7, C, scale=0.1)
dil = (ctypes.c_int * 3)(1, 3, 9)
plan = ctypes.c_void_p()
P = lambda t: ctypes.c_void_p(t.data_ptr())
assert lib.l3ac_tail_plan_create(P(cw), P(cb), P(pw), P(pb), P(a0), P(a1), dil, P(af), P(wf), ctypes.c_float(0.01), 24, ctypes.byref(plan)) == 0
x = (torch.randn(B, T, C, generator=g) * 0.7).cuda()
out = torch.empty(B, T, device="cuda")
print("rc", lib.l3ac_decoder_tail_tc(plan, P(x), B, T, P(out), None), flush=True)
torch.cuda.synchronize()
print("ok", out.abs().mean().item())
