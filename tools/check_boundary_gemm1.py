"""nfdpm_boundary_gemm1 (step boundary + first coupling GEMM of the next StepFlow in one launch) against
nfdpm_flow_boundary(_stash) + nfdpm_gemm_nt: parity on small ragged cases, then the in-graph time of one StepFlow chain per
level (BASELINE config-2 shapes, B = 128) with and without the fusion."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200"))
import torch
from normalizing_flow import _native as N

dev = torch.device("cuda")
dt = torch.bfloat16
F = 512


def setup(B, C, h, w, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s, sc=1.0: torch.randn(*s, device=dev, generator=g) * sc
    P = h * w
    M = B * P
    K1p = (9 * (C // 2) + 63) // 64 * 64
    ldp = (9 * C + 15) // 16 * 16
    d = dict(B=B, C=C, h=h, w=w, P=P, M=M, K1p=K1p, ldp=ldp)
    d["pm"] = r(M, ldp, sc=0.3)
    d["w1"] = r(F, K1p, sc=0.1).to(dt)
    d["w1"][:, 9 * (C // 2):] = 0
    d["w2"] = r(F, F, sc=0.05).to(dt)
    d["w3"] = r(ldp, F, sc=0.02).to(dt)
    d["s1"], d["b1"] = r(F, sc=0.1), r(F, sc=0.3)
    d["x"] = r(B, C, h, w)
    d["xsq"] = r(B, C // 4, 2 * h, 2 * w) if C % 4 == 0 else None
    d["mt"], d["beta"] = r(C, C, sc=0.3), r(C)
    d["b3"], d["l3"] = r(C, sc=0.1), r(C, sc=0.1)
    return d


def parity():
    for (B, C, h, w) in [(3, 12, 16, 16), (5, 24, 8, 8), (7, 48, 4, 4), (4, 8, 8, 8), (6, 32, 4, 2), (3, 4, 16, 16), (2, 16, 8, 16)]:
        d = setup(B, C, h, w, seed=B + C)
        P, M, K1p, ldp = d["P"], d["M"], d["K1p"], d["ldp"]
        assert N.boundary_gemm1_ok(C, h, w, F, K1p), (C, h, w)
        for mode in ("coupling_fwd", "coupling_inv", "plain", "squeeze"):
            if mode == "squeeze" and d["xsq"] is None:
                continue
            for stash in (True, False):
                cp = mode.startswith("coupling")
                inv = mode == "coupling_inv"
                src = d["xsq"] if mode == "squeeze" else d["x"]
                src_bs = C * P
                pm = d["pm"] if cp else None
                y_ref, xs_ref = torch.empty_like(d["x"]), torch.empty_like(d["x"])
                a_ref = torch.empty(M, K1p, dtype=dt, device=dev)
                part_ref = torch.zeros(B, device=dev)
                if inv:
                    N.flow_boundary(src, src_bs, False, pm, ldp, d["b3"], d["l3"], None, d["mt"], d["beta"], y_ref, C * P,
                                    a_ref, K1p, B, C, h, w, True)
                else:
                    N.flow_boundary_stash(src, src_bs, mode == "squeeze", pm, ldp if cp else 0, d["b3"] if cp else None,
                                          d["l3"] if cp else None, part_ref if cp else None, d["mt"], d["beta"], y_ref, C * P,
                                          xs_ref, C * P, a_ref, K1p, B, C, h, w)
                h1_ref = torch.empty(M, F, dtype=dt, device=dev)
                N.gemm_nt(a_ref, K1p, d["w1"], K1p, h1_ref, F, M, F, K1p, N.EPI_ACTNORM_RELU, d["s1"], d["b1"])
                y, xs = torch.empty_like(d["x"]), torch.empty_like(d["x"])
                a1 = torch.full((M, K1p), 7.0, dtype=dt, device=dev) if stash else None
                part = torch.zeros(B, device=dev)
                h1 = torch.full((M, F), float("nan"), dtype=dt, device=dev)
                N.boundary_gemm1(src, src_bs, mode == "squeeze", pm, ldp if cp else 0, d["b3"] if cp else None,
                                 d["l3"] if cp else None, part if (cp and not inv) else None, d["mt"], d["beta"], y, C * P,
                                 None if inv else xs, 0 if inv else C * P, a1, d["w1"], d["s1"], d["b1"], h1, B, C, h, w, F,
                                 K1p, inv)
                torch.cuda.synchronize()
                ok = torch.equal(y, y_ref) and torch.equal(h1, h1_ref)
                if stash:
                    ok = ok and torch.equal(a1, a_ref)
                if not inv:
                    ok = ok and torch.equal(xs, xs_ref) and torch.equal(part, part_ref)
                if not ok:
                    print("MISMATCH", (B, C, h, w), mode, stash, float((h1.float() - h1_ref.float()).abs().max()),
                          float((y - y_ref).abs().max()))
                    return False
        print("case", (B, C, h, w), "ok")
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
    for lvl, (C, hw) in enumerate([(12, 16), (24, 8), (48, 4)]):
        d = setup(128, C, hw, hw)
        B, P, M, K1p, ldp = d["B"], d["P"], d["M"], d["K1p"], d["ldp"]
        a1 = torch.zeros(M, K1p, dtype=dt, device=dev)
        h1 = torch.empty(M, F, dtype=dt, device=dev)
        h2 = torch.empty(M, F, dtype=dt, device=dev)
        pm = torch.empty(M, ldp, device=dev)
        x = d["x"].clone()
        part = torch.empty(B, device=dev)
        s2, b2 = torch.zeros(F, device=dev), torch.zeros(F, device=dev)
        g1 = lambda: N.gemm_nt(a1, K1p, d["w1"], K1p, h1, F, M, F, K1p, N.EPI_ACTNORM_RELU, d["s1"], d["b1"])
        g2 = lambda: N.gemm_nt(h1, F, d["w2"], F, h2, F, M, F, F, N.EPI_ACTNORM_RELU, s2, b2)
        g3 = lambda: N.gemm_nt(h2, F, d["w3"], F, pm, ldp, M, ldp, F)
        bd = lambda: N.flow_boundary(x, C * P, False, pm, ldp, d["b3"], d["l3"], part, d["mt"], d["beta"], x, C * P, a1, K1p,
                                     B, C, hw, hw, False)
        bg = lambda: N.boundary_gemm1(x, C * P, False, pm, ldp, d["b3"], d["l3"], part, d["mt"], d["beta"], x, C * P, None, 0,
                                      None, d["w1"], d["s1"], d["b1"], h1, B, C, hw, hw, F, K1p, False)
        g3b = lambda: N.gemm3_boundary(h2, F, d["w3"], None, 0, x, C * P, d["b3"], d["l3"], part, d["mt"], d["beta"], x, C * P,
                                       None, 0, a1, K1p, B, C, hw, hw, F, ldp, False)

        def chain():
            g1(); g2(); g3(); bd()

        def chain_g3():
            g1(); g2(); g3b()

        def chain_bg():
            g2(); g3(); bg()
        r = dict(level=lvl, M=M, boundary=graph_time(bd), gemm1=graph_time(g1), boundary_gemm1=graph_time(bg),
                 chain=graph_time(chain), chain_bg1=graph_time(chain_bg))
        if N.gemm3_boundary_ok(B, C, hw, hw, F, ldp):
            r["chain_g3"] = graph_time(chain_g3)
        out.append(r)
        print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items()})
    print(json.dumps(out))


if __name__ == "__main__":
    ok = parity()
    print("PARITY", "OK" if ok else "FAILED")
    if ok or os.environ.get("FORCE_TIMING"):
        timing()
