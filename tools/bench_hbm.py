"""HBM roofline of the memory-bound kernels at sizes larger than L2 (CUDA events, median of 20, L2 flushed between
repetitions): achieved = ALGORITHMIC bytes / time vs the measured copy bandwidth in MEASURED_PEAKS.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200"))
import torch
from normalizing_flow import _native as N

dev = torch.device("cuda")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for s, e in ev:
        flush.zero_()
        s.record(); fn(); e.record()
    torch.cuda.synchronize()
    t = sorted(s.elapsed_time(e) for s, e in ev)
    return t[len(t) // 2] * 1e-3


rows = []


def report(name, nbytes, t):
    gbs = nbytes / t / 1e9
    rows.append(dict(kernel=name, bytes=nbytes, us=t * 1e6, gbs=gbs, frac=gbs / peak))
    print(f"{name:58s} {nbytes / 1e6:9.1f} MB {t * 1e6:9.1f} us {gbs:8.0f} GB/s  {100 * gbs / peak:5.1f} % of {peak:.0f}")


f32 = dict(dtype=torch.float32, device=dev)
# K-A: fused ActNorm + 1x1 conv (8*C*P bytes per image): config-3 batch (1024 x 12 x 16x16) and config-4 level 0 (64 x 12 x 64x64)
for (B, C, P) in [(1024, 12, 256), (8192, 12, 256), (64, 12, 4096), (2048, 24, 64), (4096, 48, 16)]:
    x, y = torch.randn(B, C, P, **f32), torch.empty(B, C, P, **f32)
    mt, beta = torch.randn(C * C, **f32) * 0.3, torch.randn(C, **f32)
    t = timeit(lambda: N.channel_mix(x, y, mt, beta, B, C, P, C * P, C * P))
    report(f"channel_mix (K-A) B={B} C={C} P={P}", 8.0 * B * C * P, t)
# squeeze / unsqueeze
B, C, H, W = 2048, 3, 32, 32
x, y = torch.randn(B, C, H, W, **f32), torch.empty(B, 4 * C, H // 2, W // 2, **f32)
report("squeeze B=2048 3x32x32", 8.0 * x.numel(), timeit(lambda: N.squeeze(x, y, B, C, H, W, C * H * W, C * H * W)))
# ActNorm + ReLU backward on rows (bf16 in/out): 6 bytes per element
M, F = 262144, 512
dh = torch.randn(M, F, device=dev).to(torch.bfloat16); h = torch.randn(M, F, device=dev).clamp_min(0).to(torch.bfloat16)
dpre = torch.empty(M, F, dtype=torch.bfloat16, device=dev); sc = torch.zeros(F, **f32)
part = torch.empty((M // 64) * 2 * F, **f32)
report("actnorm_relu_bwd M=262144 N=512 (bf16)", 6.0 * M * F, timeit(lambda: N.actnorm_relu_bwd(dh, F, h, F, sc, dpre, F, part, M, F, 64)))
# Gaussian prior log-p (read-only reduction)
B, C, P = 16384, 48, 16
z = torch.randn(B, C, P, **f32); out = torch.empty(B, **f32)
bias, logs = torch.zeros(2 * C, **f32), torch.zeros(2 * C, **f32)
report("gauss_logp_const B=16384 48x4x4", 4.0 * z.numel(), timeit(lambda: N.gauss_logp_const(z, bias, logs, out, B, C, P)))
print(json.dumps(rows))
