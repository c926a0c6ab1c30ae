"""Pipeline timeline of the tcgen05 attention kernel (debug tool): builds attention_umma.cu with -DL3AC_ATTU_TRACE, runs the
1kbps frame-rate shape (24 clips, T = 1779, w = 750) and prints the events of CTA (11, 0, 0) -- a 12-tile CTA -- in time order."""
import ctypes, pathlib, subprocess, sys
import torch
ROOT = pathlib.Path(__file__).resolve().parent.parent
CSRC = ROOT / "l3ac_b200" / "csrc"
OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
EV = {1: "kernel start", 2: "set-up done", 9: "softmax: wait s_full", 10: "softmax: s_full acquired", 11: "softmax: pass 1 done", 12: "softmax: max exchanged, corr",
      13: "softmax: pv_done + fold done", 14: "softmax: pass 2 done, p_full arrive", 20: "mma: S issued", 21: "mma: P V issued",
      30: "loader: tile issued", 31: "loader: stage free"}
lib_path = OUT / "libattu_trace.so"
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DL3AC_ATTU_TRACE", "-Xcompiler", "-fPIC",
                "--expt-relaxed-constexpr", "-shared", "-I", str(ROOT / "include"), str(CSRC / "attention_umma.cu"), str(CSRC / "api.cu"),
                "-o", str(lib_path)], check=True)
lib = ctypes.CDLL(str(lib_path))
g = torch.Generator().manual_seed(0)
B, T, H, D, w = 24, 1779, 6, 32, 750
SPLIT = len(sys.argv) > 1 and sys.argv[1] == "split"
qkv32 = torch.randn(B, T, 3 * H * D, generator=g).cuda()
qkv = qkv32.to(torch.bfloat16)
qlo = (qkv32 - qkv.float()).to(torch.bfloat16)
out_lo = torch.empty(B, T, H * D, device="cuda", dtype=torch.bfloat16)
table = (torch.randn(H, 2 * w, generator=g) * 0.5).cuda()
out = torch.empty(B, T, H * D, device="cuda", dtype=torch.bfloat16)
P = lambda t: ctypes.c_void_p(t.data_ptr())
for _ in range(2):
    rc = (lib.l3ac_local_attention_umma(P(qkv), P(qlo), P(table), B, T, H, D, w, P(out), P(out_lo), 2, None) if SPLIT else
          lib.l3ac_local_attention_umma(P(qkv), None, P(table), B, T, H, D, w, P(out), None, 1, None))
    assert rc == 0, rc
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (4 * 512))()
lib.l3ac_debug_attu_trace(buf)
ev = []
for role in range(4):
    n = buf[role * 512]
    # the buffer holds both launches back to back: keep the second half
    items = [buf[role * 512 + 1 + i] for i in range(n)]
    items = items[len(items) // 2:]
    for v in items:
        ev.append((v >> 16, role, (v >> 8) & 0xff, v & 0xff))
ev.sort()
t0 = ev[0][0]
for t, role, e, j in ev:
    print(f"{t - t0:8d}  {['softmax', 'mma', 'loader', 'cta'][role]:8s} tile {j:2d}  {EV.get(e, e)}")
