"""In-situ time of one StepFlow kernel chain per level (BASELINE config 2 shapes): 16 back-to-back repetitions of
GEMM1 -> GEMM2 -> GEMM3 -> boundary captured in a CUDA graph (as the product runs them), CUDA-event timed, per kernel
and for the whole chain.  Complements tools/bench_kernels.py (isolated, cold-L2 launches)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200"))
import torch
from normalizing_flow import _native as N

dev = torch.device("cuda")
B = int(os.environ.get("B", 128))
dt = torch.bfloat16
F, REP = 512, 16


def graph_time(fn, reps=REP, iters=20):
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
    return s.elapsed_time(e) * 1e3 / (iters * reps)     # us per call of fn


out = []
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
    x = torch.randn(B, C, hw, hw, device=dev)
    mt, beta = torch.randn(C, C, device=dev) * 0.3, torch.randn(C, device=dev)
    b3, l3 = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    part = torch.empty(B, device=dev)
    ep = torch.zeros(4 * F, device=dev); ep[:F] = 1; ep[2 * F:3 * F] = 1
    g1 = lambda: N.gemm_nt(a1, K1p, w1, K1p, h1, F, M, F, K1p, N.EPI_ACTNORM_RELU, es, eb)
    g2 = lambda: N.gemm_nt(h1, F, w2, F, h2, F, M, F, F, N.EPI_ACTNORM_RELU, es, eb)
    g3 = lambda: N.gemm_nt(h2, F, w3, F, pm, ldp, M, ldp, F)
    bd = lambda: N.flow_boundary(x, C * P, False, pm, ldp, b3, l3, part, mt, beta, x, C * P, a1, K1p, B, C, hw, hw, False)
    fu = lambda: N.coupling_fused(a1, K1p, w1, w2, w3, pm, ldp, M, K1p, ep)

    def chain():
        g1(); g2(); g3(); bd()

    def chain_f():
        fu(); bd()
    g3b = lambda: N.gemm3_boundary(h2, F, w3, None, 0, x, C * P, b3, l3, part, mt, beta, x, C * P, None, 0, a1, K1p, B, C,
                                   hw, hw, F, ldp, False)

    def chain_g3():
        g1(); g2(); g3b()
    r = dict(level=lvl, M=M, gemm1=graph_time(g1), gemm2=graph_time(g2), gemm3=graph_time(g3), boundary=graph_time(bd),
             fused=graph_time(fu), g3_boundary=graph_time(g3b), chain=graph_time(chain), chain_fused=graph_time(chain_f),
             chain_g3=graph_time(chain_g3))
    out.append(r)
    print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items()})
tot = sum(r["chain"] for r in out)
print(f"# chain with gemm3+boundary fused: {sum(r['chain_g3'] for r in out):.1f} us -> x32 = {sum(r['chain_g3'] for r in out) * 32 / 1e3:.2f} ms")
print(f"# chain sum over levels {tot:.1f} us -> x32 = {tot * 32 / 1e3:.2f} ms; fused chain {sum(r['chain_fused'] for r in out) * 32 / 1e3:.2f} ms")
print(json.dumps(out))
