"""In-graph time (16 back-to-back launches per replay) of the coupling-network GEMM shapes of BASELINE config 2.
Run under different switches (NFDPM_TC_BN=..., NFDPM_TC_SPLIT=0, NFDPM_TC_DEEP=0|2, NFDPM_TC_NARROW=...) to compare GEMM variants."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200"))
import torch
from normalizing_flow import _native as N

dev = torch.device("cuda")
dt = torch.bfloat16


def graph_time(fn, reps=16, iters=20):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / (iters * reps)


out = {}
for name, M, Nn, K, out_dt, epi in [("L0 gemm1", 32768, 512, 64, dt, 1), ("L0 gemm2", 32768, 512, 512, dt, 1),
                                    ("L0 gemm3", 32768, 112, 512, torch.float32, 0), ("L1 gemm1", 8192, 512, 128, dt, 1),
                                    ("L1 gemm2", 8192, 512, 512, dt, 1), ("L1 gemm3", 8192, 224, 512, torch.float32, 0),
                                    ("L2 gemm1", 2048, 512, 256, dt, 1), ("L2 gemm2", 2048, 512, 512, dt, 1),
                                    ("L2 gemm3", 2048, 432, 512, torch.float32, 0)]:
    a = (torch.randn(M, K, device=dev) * 0.5).to(dt)
    w = (torch.randn(Nn, K, device=dev) * 0.05).to(dt)
    d = torch.empty(M, Nn, dtype=out_dt, device=dev)
    es, eb = torch.zeros(Nn, device=dev), torch.zeros(Nn, device=dev)
    args = (a, K, w, K, d, Nn, M, Nn, K) + ((N.EPI_ACTNORM_RELU, es, eb) if epi else ())
    us = graph_time(lambda: N.gemm_nt(*args))
    ref = (a.float() @ w.float().T)
    if epi:
        ref = ref.clamp_min(0)
    err = float((d.float() - ref).abs().max() / ref.abs().max())
    out[name] = round(us, 2)
    print(f"{name:9s} M={M:6d} N={Nn:4d} K={K:4d}  {us:7.2f} us  {2.0 * M * Nn * K / us / 1e6:7.1f} TFLOP/s  rel err {err:.1e}")
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("NFDPM_")}, "us": out}))
