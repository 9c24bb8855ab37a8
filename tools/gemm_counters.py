"""Where the tcgen05 GEMM spends its time: per-CTA wait / work cycle counters (nfdpm_gemm_debug) for the coupling-network
shapes of BASELINE config 2, launched back to back the way the product runs them (A operand L2-resident)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200"))
import torch
from normalizing_flow import _native as N

dev = torch.device("cuda")
dt = N.SPLIT if os.environ.get("MODE", "bf16") == "split" else torch.bfloat16       # MODE=split: fp32-faithful operand pairs
F = 512


def operand(rows, cols, scale):
    v = torch.randn(rows, cols, device=dev) * scale
    if dt != N.SPLIT:
        return v.to(dt)
    hi = v.bfloat16()
    lo = (v - hi.float()).bfloat16()
    w = torch.stack([hi.reshape(rows, cols // 32, 32), lo.reshape(rows, cols // 32, 32)], dim=2).contiguous()
    return w.view(torch.int32).reshape(rows, cols)


for name, M, Nn, K, out_dt, epi in [("L0 gemm1", 32768, 512, 64, dt, 1), ("L0 gemm2", 32768, 512, 512, dt, 1),
                                    ("L0 gemm3", 32768, 112, 512, torch.float32, 0), ("L1 gemm2", 8192, 512, 512, dt, 1),
                                    ("L2 gemm2", 2048, 512, 512, dt, 1)]:
    a = operand(M, K, 0.5)
    w = operand(Nn, K, 0.05)
    d = torch.empty(M, Nn, dtype=out_dt, device=dev)
    es, eb = torch.zeros(Nn, device=dev), torch.zeros(Nn, device=dev)
    args = (a, K, w, K, d, Nn, M, Nn, K) + ((N.EPI_ACTNORM_RELU, es, eb) if epi else ())
    for _ in range(3):
        N.gemm_nt(*args)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        N.gemm_nt(*args)
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1e3 / 20
    buf = torch.zeros(148, 16, dtype=torch.int64, device=dev)
    N.lib.nfdpm_gemm_debug(buf.data_ptr())
    N.gemm_nt(*args)
    N.gemm_nt(*args)
    torch.cuda.synchronize()
    N.lib.nfdpm_gemm_debug(None)
    t = buf.cpu().double()
    t = t[t[:, 2] > 0]
    us_c = lambda c: float(c.mean()) / 1.9e3
    print(f"{name} M={M} N={Nn} K={K}: {us:.2f} us/launch (eager back-to-back); CTAs {t.shape[0]}, tiles/CTA {float(t[:, 10].mean()):.2f}")
    print(f"   lifetime: entry->pdl_wait done {us_c(t[:, 11]):.2f} us, setup {us_c(t[:, 12]):.2f} us, epilogue-warp total {us_c(t[:, 5]):.2f} us; "
          f"first CTA entry -> last CTA end (globaltimer) {(float(t[:, 14].max()) - float(t[:, 13].min())) / 1e3:.2f} us, "
          f"entry spread {(float(t[:, 13].max()) - float(t[:, 13].min())) / 1e3:.2f} us, end spread {(float(t[:, 14].max()) - float(t[:, 14].min())) / 1e3:.2f} us")
    print(f"   CTA cycles as us: total {us_c(t[:, 2]):.2f} | producer waits for free stage {us_c(t[:, 1]):.2f} | "
          f"MMA waits for TMA bytes {us_c(t[:, 3]):.2f}, for drained accumulator {us_c(t[:, 4]):.2f} | epilogue: waits tfull "
          f"{us_c(t[:, 6]):.2f}, waits staging free {us_c(t[:, 7]):.2f}, TMEM->smem {us_c(t[:, 8]):.2f}, barrier {us_c(t[:, 9]):.2f}")
