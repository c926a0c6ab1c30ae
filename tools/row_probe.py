"""Runs the HBM-bound row kernels of the decode side once at their 32-clip x 10 s shapes (for an ncu capture):
dwconv7_ln (C = 48 / 96 / 256, bf16 out), enhance (C = 48), upsample_linear_cn (24 channels x2, 48 channels x3), layernorm (128)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from l3ac_b200 import ops

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g)
B = 32
for C, T in ((48, 80000), (96, 26667), (256, 8889)):
    x = r(B, T, C)
    for _ in range(1):
        ops.dwconv7_ln(x, r(7, C), r(C), r(C), r(C), 1e-8, out_dtype=torch.bfloat16)
for C, T in ((48, 80000), (96, 26667)):
    x = r(B, T, C)
    for _ in range(1):
        ops.enhance(x, r(4, 7), r(4), r(4), r(4), r(C, 4), r(C), out_dtype=torch.bfloat16)
for C, T, s in ((24, 80000, 2), (48, 26667, 3)):
    x = r(B, T, C)
    for _ in range(1):
        ops.upsample_linear_cn(x, s, r(C), r(C), 1e-8)
x = r(B, 1778, 128)
for _ in range(1):
    ops.layernorm(x, r(128), r(128), 1e-5, out_dtype=torch.bfloat16)
torch.cuda.synchronize()
