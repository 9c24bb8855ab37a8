"""Per-kernel timing (CUDA events, L2 flushed between repetitions) of the hot-path kernels at the BASELINE config-2
shapes: the three coupling GEMMs and the fused boundary kernel at every level.  Prints a table + JSON."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200"))
import torch
from normalizing_flow import _native as N

dev = torch.device("cuda")
B = int(os.environ.get("B", 128))
mode = os.environ.get("NFDPM_PRECISION", "bf16")
dt = torch.bfloat16 if mode == "bf16" else torch.float32
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
WARM = os.environ.get("WARM", "0") == "1"     # WARM=1: no flush (operands L2-resident when they fit)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for s, e in ev:
        if not WARM:
            flush.zero_()
        s.record(); fn(); e.record()
    torch.cuda.synchronize()
    t = sorted(s.elapsed_time(e) for s, e in ev)
    return t[len(t) // 2] * 1e3   # median us


rows = []
F = 512
for lvl, (C, hw) in enumerate([(12, 16), (24, 8), (48, 4)]):
    P = hw * hw
    M = B * P
    K1p = (9 * (C // 2) + 63) // 64 * 64
    ldp = (9 * C + 15) // 16 * 16
    a1 = (torch.randn(M, K1p, device=dev) * 0.5).to(dt)
    w1 = (torch.randn(F, K1p, device=dev) * 0.1).to(dt)
    h1 = torch.empty(M, F, dtype=dt, device=dev)
    w2 = (torch.randn(F, F, device=dev) * 0.05).to(dt)
    h2 = torch.empty(M, F, dtype=dt, device=dev)
    w3 = (torch.randn(ldp, F, device=dev) * 0.05).to(dt)
    pm = torch.empty(M, ldp, device=dev)
    es, eb = torch.zeros(F, device=dev), torch.zeros(F, device=dev)
    t1 = timeit(lambda: N.gemm_nt(a1, K1p, w1, K1p, h1, F, M, F, K1p, N.EPI_ACTNORM_RELU, es, eb))
    t2 = timeit(lambda: N.gemm_nt(h1, F, w2, F, h2, F, M, F, F, N.EPI_ACTNORM_RELU, es, eb))
    t3 = timeit(lambda: N.gemm_nt(h2, F, w3, F, pm, ldp, M, ldp, F))
    x = torch.randn(B, C, hw, hw, device=dev)
    mt, beta = torch.randn(C, C, device=dev) * 0.3, torch.randn(C, device=dev)
    b3, l3 = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    part = torch.empty(B, device=dev)
    tb = timeit(lambda: N.flow_boundary(x, C * P, False, pm, ldp, b3, l3, part, mt, beta, x, C * P, a1, K1p, B, C, hw, hw, False))
    tf = float("nan")          # (the single-kernel coupling network of round 1 left the library)
    fl = lambda k, n: 2.0 * M * n * k
    rows.append(dict(level=lvl, C=C, P=P, M=M, K1p=K1p, ldp=ldp,
                     gemm1_us=t1, gemm1_tflops=fl(K1p, F) / t1 / 1e6,
                     gemm2_us=t2, gemm2_tflops=fl(F, F) / t2 / 1e6,
                     gemm3_us=t3, gemm3_tflops=fl(F, ldp) / t3 / 1e6,
                     fused_us=tf, fused_tflops=(fl(K1p, F) + fl(F, F) + fl(F, ldp)) / tf / 1e6,
                     boundary_us=tb,
                     boundary_gbs=(M * ldp * 4 + 2 * B * C * P * 4 + M * K1p * a1.element_size()) / tb / 1e3))
print(f"# B={B} mode={mode} warm={WARM}")
print(f"{'lvl':>3} {'M':>6} {'gemm1':>8} {'TF':>6} {'gemm2':>8} {'TF':>6} {'gemm3':>8} {'TF':>6} {'fused':>8} {'TF':>6} {'bound':>8} {'GB/s':>6}  (us)")
tot = 0
for r in rows:
    print(f"{r['level']:>3} {r['M']:>6} {r['gemm1_us']:8.1f} {r['gemm1_tflops']:6.0f} {r['gemm2_us']:8.1f} {r['gemm2_tflops']:6.0f} "
          f"{r['gemm3_us']:8.1f} {r['gemm3_tflops']:6.0f} {r['fused_us']:8.1f} {r['fused_tflops']:6.0f} "
          f"{r['boundary_us']:8.1f} {r['boundary_gbs']:6.0f}")
    tot += (r['fused_us'] if r['fused_us'] == r['fused_us'] else r['gemm1_us'] + r['gemm2_us'] + r['gemm3_us']) + r['boundary_us']
print(f"# sum per StepFlow over levels = {tot:.1f} us  ->  x16 steps x2 directions = {tot * 32 / 1e3:.2f} ms")
print(json.dumps(rows))
