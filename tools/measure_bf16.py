"""Measure bf16 (tcgen05) coupling-net mode against the fp32 oracle: numbers behind the stated bf16 tolerance."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200"))
import torch
import normalizing_flow as nf
from oracle import glow_oracle as O
torch.set_grad_enabled(False)
DEV = torch.device("cuda")


def rel(a, b):
    a, b = a.cpu().double(), b.double()
    return float((a - b).norm() / b.norm())


def run(c, L, K, B, S, seed, mode, zero_sigma=0.02):
    os.environ["NFDPM_PRECISION"] = mode
    sd, psd = O.seeded_state(c, L, K, seed, zero_sigma=zero_sigma)
    flow = nf.Glow(c, L, K).to(DEV); flow.load_state_dict(sd)
    prior = nf.GaussianPrior(2 ** (L + 1) * c).to(DEV); prior.load_state_dict(psd)
    x = O.seeded_input((B, c, S, S), seed + 1)
    ld = torch.zeros(B, dtype=torch.float64, device=DEV); lp = torch.zeros(B, dtype=torch.float64, device=DEV)
    zs, ld, lp = flow.transform(x.to(DEV), ld, lp)
    pl = prior.compute_log_prob(zs[-1])
    xr = flow.invert(zs)
    ld_o = torch.zeros(B, dtype=torch.float64); lp_o = torch.zeros(B, dtype=torch.float64)
    zo, ld_o, lp_o = O.glow_transform(sd, x, L, K, ld_o, lp_o)
    pl_o = O.gaussian_prior_logp(psd, zo[-1])
    n_pix = float(c * S * S)
    bpd = float(O.bpd_loss(ld.cpu() + lp.cpu() + pl.cpu().double(), 32.0, n_pix))
    bpd_o = float(O.bpd_loss(ld_o + lp_o + pl_o.double(), 32.0, n_pix))
    xo = O.glow_invert(sd, zo, L, K)
    xr_from_oracle_z = flow.invert([t.to(DEV) for t in zo])
    return {"cfg": f"c{c} L{L} K{K} B{B} S{S} zs={zero_sigma}", "mode": mode,
            "z_rel": [round(rel(a, b), 7) for a, b in zip(zs, zo)],
            "ld_rel": float(((ld.cpu() - ld_o).abs() / ld_o.abs()).max()),
            "lp_rel": float(((lp.cpu() - lp_o).abs() / lp_o.abs()).max()),
            "bpd_abs": abs(bpd - bpd_o), "bpd": bpd_o,
            "recon_own": float((xr.cpu() - x).abs().max()), "recon_oracle": float((xo - x).abs().max()),
            "inv_vs_oracle": float((xr_from_oracle_z.cpu() - xo).abs().max())}


for cfg in [(1, 3, 2, 3, 32, 11), (3, 3, 1, 2, 32, 12), (3, 3, 4, 16, 32, 13), (3, 3, 16, 16, 32, 14), (1, 3, 4, 64, 32, 15)]:
    for mode in ("fp32", "bf16"):
        print(json.dumps(run(*cfg, mode)))
print(json.dumps(run(3, 3, 16, 16, 32, 14, "bf16", zero_sigma=0.005)))
print(json.dumps(run(3, 3, 16, 16, 32, 14, "fp32", zero_sigma=0.005)))
