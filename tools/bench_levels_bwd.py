"""In-situ time (CUDA graph of 16 back-to-back launches, CUDA events) of every backward kernel of one StepFlow at the three
level shapes of BASELINE config 2 (B=128).  Complements tools/bench_levels.py (forward chain)."""
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
    return s.elapsed_time(e) * 1e3 / (iters * reps)


out = []
for lvl, (C, hw) in enumerate([(12, 16), (24, 8), (48, 4)]):
    P, Ch = hw * hw, C // 2
    M = B * P
    K1p = (9 * Ch + 63) // 64 * 64
    ldp = (9 * C + 15) // 16 * 16
    Kp3 = (9 * C + 63) // 64 * 64
    f32 = dict(dtype=torch.float32, device=dev)
    rb = lambda *s: (torch.randn(*s, device=dev) * 0.3).to(dt)
    dy, u, x = torch.randn(B, C, P, **f32), torch.randn(B, C, P, **f32), torch.randn(B, C, P, **f32)
    pm = torch.randn(M, ldp, **f32) * 0.1
    b3, l3 = torch.zeros(C, **f32), torch.zeros(C, **f32)
    dld = torch.ones(B, **f32)
    du, dxb = torch.empty(B, C, P, **f32), torch.empty(B, C, P, **f32)
    dpm = torch.empty(M * Kp3, dtype=dt, device=dev)
    dpar = torch.empty(B * 2 * C, **f32)
    gb, gl = torch.empty(C, **f32), torch.empty(C, **f32)
    h1, h2, dh, dpre = rb(M, F), rb(M, F), rb(M, F), rb(M, F)
    a1 = rb(M, K1p)
    w3t, w2t, w1t = rb(F, Kp3), rb(F, F), rb(K1p, F)
    dA1 = torch.empty(M * K1p, **f32)
    scale = torch.zeros(F, **f32)
    rows = 64
    while rows > 8 and (M + rows - 1) // rows < 256:
        rows //= 2
    n_cta = (M + rows - 1) // rows
    an_part = torch.empty(n_cta * 2 * F, **f32)
    gs, gbb = torch.empty(F, **f32), torch.empty(F, **f32)
    ws = torch.empty(max(N.gemm_tn_workspace(M, F, F), N.gemm_tn_workspace(M, ldp, F), N.gemm_tn_workspace(M, F, K1p)), **f32)
    dw2, dw3, dw1 = torch.empty(F * F, **f32), torch.empty(C * F * 9, **f32), torch.empty(F * Ch * 9, **f32)
    mt = torch.randn(C * C, **f32) * 0.3
    Tm = N.mix_bwd_tiles(C, hw, hw)
    part = torch.empty(B * Tm * (C * C + C), **f32)
    r = dict(level=lvl, M=M)
    r["coupling_bwd"] = graph_time(lambda: N.coupling_bwd(dy, C * P, dld, u, C * P, pm, ldp, b3, l3, du, C * P, dpm, Kp3, dpar,
                                                          B, C, hw, hw, gb, gl))
    r["wgrad3"] = graph_time(lambda: N.gemm_tn(dpm, Kp3, h2, F, dw3, M, ldp, F, ws, out_mode=N.TN_OUT_TAPS, out_c=C))
    r["dgrad3"] = graph_time(lambda: N.gemm_nt(dpm, Kp3, w3t, Kp3, dh, F, M, F, Kp3))
    r["actnorm_relu_bwd"] = graph_time(lambda: N.actnorm_relu_bwd(dh, F, h2, F, scale, dpre, F, an_part, M, F, rows))
    n_mt = (M + 127) // 128
    part2 = torch.empty(n_mt * 2 * F, **f32)
    r["reduce_rows2"] = graph_time(lambda: N.reduce_rows2(an_part, gs, gbb, n_cta, F, F, 2 * F))
    r["wgrad2"] = graph_time(lambda: N.gemm_tn(dpre, F, h1, F, dw2, M, F, F, ws))
    r["dgrad2"] = graph_time(lambda: N.gemm_nt(dpre, F, w2t, F, dh, F, M, F, F))
    r["wgrad1"] = graph_time(lambda: N.gemm_tn(dpre, F, a1, K1p, dw1, M, F, K1p, ws, out_mode=N.TN_OUT_STRIP, out_c=Ch * 9))
    r["dgrad1"] = graph_time(lambda: N.gemm_nt(dpre, F, w1t, F, dA1, K1p, M, K1p, F))
    r["mix_bwd"] = graph_time(lambda: N.mix_bwd(du, C * P, dA1, K1p, x, C * P, mt, dxb, C * P, part, B, C, hw, hw))
    r["sum"] = (r["coupling_bwd"] + r["wgrad3"] + r["dgrad3_fused"] + r["wgrad2"] + r["dgrad2_fused"] + r["wgrad1"] +
                r["dgrad1"] + r["mix_bwd"] + 2 * r["reduce_rows2"])
    r["sum_unfused"] = (r["sum"] - r["dgrad3_fused"] - r["dgrad2_fused"] + r["dgrad3"] + r["dgrad2"] +
                        2 * r["actnorm_relu_bwd"])
    out.append(r)
    print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items()})
print(f"# backward chain per StepFlow, sum over levels: {sum(r['sum'] for r in out):.1f} us -> x16 = {sum(r['sum'] for r in out) * 16 / 1e3:.2f} ms")
print(json.dumps(out))
