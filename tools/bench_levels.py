"""In-situ time of one StepFlow kernel chain per level (BASELINE config 2 shapes): 16 back-to-back repetitions of
GEMM1 -> GEMM2 -> GEMM3 -> boundary captured in a CUDA graph (as the product runs them), CUDA-event timed, per kernel
and for the whole chain, for both tensor-core operand formats (MODE=bf16 | split | both).  Complements
tools/bench_kernels.py (isolated, cold-L2 launches)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200"))
import torch
from normalizing_flow import _native as N

dev = torch.device("cuda")
B = int(os.environ.get("B", 128))
MODE = os.environ.get("MODE", "both")
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


def operand(rows, cols, scale, dt):
    """Random [rows, cols] operand in format dt (any bit pattern of finite bf16 values is a valid split pair)."""
    v = torch.randn(rows, cols, device=dev) * scale
    if dt == torch.bfloat16:
        return v.to(dt)
    hi = v.bfloat16()
    lo = (v - hi.float()).bfloat16()
    w = torch.stack([hi.reshape(rows, cols // 32, 32), lo.reshape(rows, cols // 32, 32)], dim=2).contiguous()
    return w.view(torch.int32).reshape(rows, cols)


def run(dt, tag):
    out = []
    mmas = 3 if dt == N.SPLIT else 1
    for lvl, (C, hw) in enumerate([(12, 16), (24, 8), (48, 4)]):
        P = hw * hw
        M = B * P
        K1p = (9 * (C // 2) + 63) // 64 * 64
        ldp = (9 * C + 15) // 16 * 16
        a1 = operand(M, K1p, 0.5, dt)
        w1 = operand(F, K1p, 0.1, dt)
        h1 = torch.empty(M, F, dtype=dt, device=dev)
        w2 = operand(F, F, 0.05, dt)
        h2 = torch.empty(M, F, dtype=dt, device=dev)
        w3 = operand(ldp, F, 0.05, dt)
        pm = torch.empty(M, ldp, device=dev)
        es, eb = torch.zeros(F, device=dev), torch.zeros(F, device=dev)
        x = torch.randn(B, C, hw, hw, device=dev)
        mt, beta = torch.randn(C, C, device=dev) * 0.3, torch.randn(C, device=dev)
        b3, l3 = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        part = torch.empty(B, device=dev)
        g1 = lambda: N.gemm_nt(a1, K1p, w1, K1p, h1, F, M, F, K1p, N.EPI_ACTNORM_RELU, es, eb)
        g2 = lambda: N.gemm_nt(h1, F, w2, F, h2, F, M, F, F, N.EPI_ACTNORM_RELU, es, eb)
        g3 = lambda: N.gemm_nt(h2, F, w3, F, pm, ldp, M, ldp, F)
        bd = lambda: N.flow_boundary(x, C * P, False, pm, ldp, b3, l3, part, mt, beta, x, C * P, a1, K1p, B, C, hw, hw, False)

        def chain():
            g1(); g2(); g3(); bd()
        r = dict(mode=tag, level=lvl, M=M, gemm1=graph_time(g1), gemm2=graph_time(g2), gemm3=graph_time(g3),
                 boundary=graph_time(bd), chain=graph_time(chain))
        r["gemm2_tflops_alg"] = 2.0 * M * F * F / r["gemm2"] / 1e6          # algorithmic (fp32-equivalent) FLOPs
        r["gemm2_tflops_issued"] = mmas * r["gemm2_tflops_alg"]            # bf16 MMA FLOPs actually issued
        if N.gemm3_boundary_ok(B, C, hw, hw, F * (2 if dt == N.SPLIT else 1), ldp):
            g3b = lambda: N.gemm3_boundary(h2, F, w3, None, 0, x, C * P, b3, l3, part, mt, beta, x, C * P, None, 0, a1, K1p,
                                           B, C, hw, hw, F, ldp, False)

            def chain_g3():
                g1(); g2(); g3b()
            r["g3_boundary"] = graph_time(g3b)
            r["chain_g3"] = graph_time(chain_g3)
        out.append(r)
        print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items()}, flush=True)
    best = sum(min(r["chain"], r.get("chain_g3", 1e9)) for r in out)
    print(f"# [{tag}] best chain sum over levels {best:.1f} us -> x16 steps x2 directions = {best * 32 / 1e3:.2f} ms")
    return out


res = []
if MODE in ("bf16", "both"):
    res += run(torch.bfloat16, "bf16")
if MODE in ("split", "both"):
    res += run(N.SPLIT, "split")
print(json.dumps(res))
