"""A few launches of nfdpm_gemm_nt at one shape, for ncu:  python tools/run_gemm.py M N K [split|bf16] [f32|same]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200"))
import torch
from normalizing_flow import _native as N
M, Nn, K = (int(v) for v in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "split"
out = sys.argv[5] if len(sys.argv) > 5 else "same"
dev = torch.device("cuda")
dt = N.SPLIT if mode == "split" else torch.bfloat16


def operand(rows, cols, scale):
    v = torch.randn(rows, cols, device=dev) * scale
    if dt != N.SPLIT:
        return v.to(dt)
    hi = v.bfloat16()
    lo = (v - hi.float()).bfloat16()
    w = torch.stack([hi.reshape(rows, cols // 32, 32), lo.reshape(rows, cols // 32, 32)], dim=2).contiguous()
    return w.view(torch.int32).reshape(rows, cols)


a, w = operand(M, K, 0.5), operand(Nn, K, 0.05)
d = torch.empty(M, Nn, dtype=torch.float32 if out == "f32" else dt, device=dev)
es, eb = torch.zeros(Nn, device=dev), torch.zeros(Nn, device=dev)
args = (a, K, w, K, d, Nn, M, Nn, K) + (() if out == "f32" else (N.EPI_ACTNORM_RELU, es, eb))
for _ in range(6):
    N.gemm_nt(*args)
torch.cuda.synchronize()
print("ok")
