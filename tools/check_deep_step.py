"""nfdpm_deep_step (cluster-fused StepFlow of a deep level) against the four-kernel chain it replaces
(3 x nfdpm_gemm_nt + nfdpm_flow_boundary): parity on ragged batches, then the in-graph time of both at the BASELINE
config-2 shapes (B = 128)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200"))
import torch
from normalizing_flow import _native as N

dev = torch.device("cuda")
dt = torch.bfloat16
F = 512


def setup(B, C, hw, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s, sc=1.0: torch.randn(*s, device=dev, generator=g) * sc
    P = hw * hw
    M = B * P
    K1p = (9 * (C // 2) + 63) // 64 * 64
    ldp = (9 * C + 15) // 16 * 16
    d = dict(B=B, C=C, hw=hw, P=P, M=M, K1p=K1p, ldp=ldp)
    d["a1"] = r(M, K1p, sc=0.5).to(dt)
    d["a1"][:, 9 * (C // 2):] = 0
    d["w1"] = r(F, K1p, sc=0.1).to(dt)
    d["w2"] = r(F, F, sc=0.05).to(dt)
    d["w3"] = r(ldp, F, sc=0.02).to(dt)
    d["w3"][9 * C:] = 0
    d["s1"], d["b1"], d["s2"], d["b2"] = r(F, sc=0.1), r(F, sc=0.3), r(F, sc=0.1), r(F, sc=0.3)
    d["x"] = r(B, C, hw, hw)
    d["mt"], d["beta"] = r(C, C, sc=0.3), r(C)
    d["b3"], d["l3"] = r(C, sc=0.1), r(C, sc=0.1)
    d["ld_pm"] = ldp
    return d


def run_ref(d, inverse, mix, a1_dt):
    B, C, hw, P, M, K1p, ldp = (d[k] for k in ("B", "C", "hw", "P", "M", "K1p", "ldp"))
    h1 = torch.empty(M, F, dtype=dt, device=dev)
    h2 = torch.empty(M, F, dtype=dt, device=dev)
    pm = torch.empty(M, ldp, device=dev)
    N.gemm_nt(d["a1"], K1p, d["w1"], K1p, h1, F, M, F, K1p, N.EPI_ACTNORM_RELU, d["s1"], d["b1"])
    N.gemm_nt(h1, F, d["w2"], F, h2, F, M, F, F, N.EPI_ACTNORM_RELU, d["s2"], d["b2"])
    N.gemm_nt(h2, F, d["w3"], F, pm, ldp, M, ldp, F)
    y, xs = torch.empty_like(d["x"]), torch.empty_like(d["x"])
    a1n = torch.full((M, K1p), 3.0, dtype=a1_dt, device=dev) if mix else None
    part = torch.zeros(B, device=dev)
    m_, b_ = (d["mt"], d["beta"]) if mix else (None, None)
    if inverse:
        N.flow_boundary(d["x"], C * P, False, pm, ldp, d["b3"], d["l3"], None, m_, b_, y, C * P, a1n, K1p if mix else 0,
                        B, C, hw, hw, True)
    else:
        N.flow_boundary_stash(d["x"], C * P, False, pm, ldp, d["b3"], d["l3"], part, m_, b_, y, C * P, xs, C * P, a1n,
                              K1p if mix else 0, B, C, hw, hw)
    return h1, h2, pm, y, xs, a1n, part


def run_deep(d, inverse, mix, a1_dt):
    B, C, hw, P, M, K1p, ldp, ld_pm = (d[k] for k in ("B", "C", "hw", "P", "M", "K1p", "ldp", "ld_pm"))
    h1 = torch.full((M, F), float("nan"), dtype=dt, device=dev)
    h2 = torch.full((M, F), float("nan"), dtype=dt, device=dev)
    pm = torch.full((M, ld_pm), float("nan"), device=dev)
    y, xs = torch.empty_like(d["x"]), torch.empty_like(d["x"])
    a1n = torch.full((M, K1p), 5.0, dtype=a1_dt, device=dev) if mix else None
    part = torch.zeros(B, device=dev)
    m_, b_ = (d["mt"], d["beta"]) if mix else (None, None)
    N.deep_step(d["a1"], d["w1"], d["w2"], d["w3"], d["s1"], d["b1"], d["s2"], d["b2"], h1, h2, pm, ld_pm, d["x"], C * P,
                d["b3"], d["l3"], None if inverse else part, m_, b_, y, C * P, None if inverse else xs,
                0 if inverse else C * P, a1n, K1p if mix else 0, B, C, hw, hw, F, K1p, ldp, inverse)
    return h1, h2, pm, y, xs, a1n, part


def parity():
    worst = {}
    for (B, C, hw) in [(5, 24, 8), (4, 8, 8), (13, 48, 4), (8, 16, 4), (128, 24, 8), (128, 48, 4)]:
        assert N.deep_step_ok(B, C, hw, hw, F, (9 * (C // 2) + 63) // 64 * 64, (9 * C + 15) // 16 * 16)
        d = setup(B, C, hw, seed=B + C)
        for inverse in (False, True):
            for mix in (True, False):
                for a1_dt in (torch.bfloat16, torch.float32):
                    ref = run_ref(d, inverse, mix, a1_dt)
                    got = run_deep(d, inverse, mix, a1_dt)
                    torch.cuda.synchronize()
                    names = ["h1", "h2", "pm", "y", "xs", "a1", "part"]
                    for n, r_, g_ in zip(names, ref, got):
                        if r_ is None or (inverse and n in ("xs", "part")):
                            continue
                        if n == "pm":
                            g_ = g_[:, :d["ldp"]]
                        rf, gf = r_.float(), g_.float()
                        err = float((rf - gf).abs().max() / (rf.abs().max() + 1e-30))
                        exact = bool(torch.equal(rf, gf))
                        key = (n,)
                        worst[n] = max(worst.get(n, 0.0), err)
                        if not exact and err > 1e-2:
                            print("MISMATCH", B, C, hw, inverse, mix, a1_dt, n, err)
                            return False
        print("case", B, C, hw, "ok", {k: f"{v:.2e}" for k, v in worst.items()})
    return True


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


def timing():
    out = []
    for lvl, (C, hw) in [(1, (24, 8)), (2, (48, 4))]:
        d = setup(128, C, hw)
        B, P, M, K1p, ldp, ld_pm = (d[k] for k in ("B", "P", "M", "K1p", "ldp", "ld_pm"))
        h1 = torch.empty(M, F, dtype=dt, device=dev)
        h2 = torch.empty(M, F, dtype=dt, device=dev)
        pm = torch.empty(M, ldp, device=dev)
        pmd = torch.empty(M, ld_pm, device=dev)
        x = d["x"].clone()
        part = torch.empty(B, device=dev)
        a1 = d["a1"]

        def chain():
            N.gemm_nt(a1, K1p, d["w1"], K1p, h1, F, M, F, K1p, N.EPI_ACTNORM_RELU, d["s1"], d["b1"])
            N.gemm_nt(h1, F, d["w2"], F, h2, F, M, F, F, N.EPI_ACTNORM_RELU, d["s2"], d["b2"])
            N.gemm_nt(h2, F, d["w3"], F, pm, ldp, M, ldp, F)
            N.flow_boundary(x, C * P, False, pm, ldp, d["b3"], d["l3"], part, d["mt"], d["beta"], x, C * P, a1, K1p, B, C,
                            hw, hw, False)

        def deep():
            N.deep_step(a1, d["w1"], d["w2"], d["w3"], d["s1"], d["b1"], d["s2"], d["b2"], None, None, None, 0, x, C * P,
                        d["b3"], d["l3"], part, d["mt"], d["beta"], x, C * P, None, 0, a1, K1p, B, C, hw, hw, F, K1p, ldp,
                        False)

        def deep_stash():
            N.deep_step(a1, d["w1"], d["w2"], d["w3"], d["s1"], d["b1"], d["s2"], d["b2"], h1, h2, pmd, ld_pm, x, C * P,
                        d["b3"], d["l3"], part, d["mt"], d["beta"], x, C * P, d["x"], C * P, a1, K1p, B, C, hw, hw, F, K1p,
                        ldp, False)
        r = dict(level=lvl, M=M, chain=graph_time(chain), deep_step=graph_time(deep), deep_step_stash=graph_time(deep_stash))
        out.append(r)
        print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items()})
    print(json.dumps(out))


if __name__ == "__main__":
    ok = parity()
    print("PARITY", "OK" if ok else "FAILED")
    if ok or os.environ.get("FORCE_TIMING"):
        timing()


def timeline():
    """Per-CTA phase timeline (clock64 deltas in us at 1.9 GHz nominal; globaltimer for the whole CTA)."""
    for lvl, (C, hw) in [(1, (24, 8)), (2, (48, 4))]:
        d = setup(128, C, hw)
        B, P, M, K1p, ldp, ld_pm = (d[k] for k in ("B", "P", "M", "K1p", "ldp", "ld_pm"))
        h1 = torch.empty(M, F, dtype=dt, device=dev)
        h2 = torch.empty(M, F, dtype=dt, device=dev)
        pmd = torch.empty(M, ld_pm, device=dev)
        x = d["x"].clone()
        part = torch.empty(B, device=dev)
        grid = (M // 128) * min(8, 128 // P)
        buf = torch.zeros(grid, 16, dtype=torch.int64, device=dev)
        N.lib.nfdpm_deep_step_debug(buf.data_ptr())
        for _ in range(3):
            N.deep_step(d["a1"], d["w1"], d["w2"], d["w3"], d["s1"], d["b1"], d["s2"], d["b2"], None, None, None, 0, x, C * P,
                        d["b3"], d["l3"], part, d["mt"], d["beta"], x, C * P, None, 0, d["a1"], K1p, B, C, hw, hw, F, K1p,
                        ldp, False)
        torch.cuda.synchronize()
        N.lib.nfdpm_deep_step_debug(None)
        t = buf.cpu().double()
        ck = (t[:, 1:15] - t[:, 1:2]) / 1.9e3          # us since CTA start
        names = ["start", "pdl_wait", "setup", "p0_xbar", "p0_epi", "p0_arrive", "p1_xbar", "p1_epi", "p1_arrive", "p2_xbar",
                 "p2_epi", "p2_arrive", "pbar", "body"]
        print(f"level {lvl}: grid {grid}; CTA wall (globaltimer) mean {float((t[:, 15] - t[:, 0]).mean()) / 1e3:.2f} us, "
              f"first start -> last end {float(t[:, 15].max() - t[:, 0].min()) / 1e3:.2f} us, "
              f"start spread {float(t[:, 0].max() - t[:, 0].min()) / 1e3:.2f} us")
        print("   mean:", {n: round(float(ck[:, i].mean()), 2) for i, n in enumerate(names)})
        print("   max :", {n: round(float(ck[:, i].max()), 2) for i, n in enumerate(names)})


if __name__ == "__main__" and os.environ.get("TIMELINE"):
    timeline()
