"""Per-tensor relative L2 error of the training gradients against the CPU oracle's autograd, per precision mode."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200"))
import torch
import normalizing_flow as nf
from oracle import glow_oracle as O
DEV = torch.device("cuda")
for (c, L, K, B, S, seed) in [(3, 2, 2, 3, 16, 7), (3, 3, 2, 4, 32, 5), (3, 3, 4, 16, 32, 13)]:
    sd, psd = O.seeded_state(c, L, K, seed)
    x = O.seeded_input((B, c, S, S), seed + 1)
    lo, g_o, _ = O.train_grads(sd, psd, x, L, K, 32.0, S * S * 3.0)
    for mode in ("fp32_simt", "fp32", "bf16"):
        os.environ["NFDPM_PRECISION"] = mode
        flow = nf.Glow(c, L, K).to(DEV); flow.load_state_dict(sd)
        prior = nf.GaussianPrior(2 ** (L + 1) * c).to(DEV); prior.load_state_dict(psd)
        ld, lp = nf.initialize_with_zeros(2, B, DEV)
        zs, ld, lp = flow.transform(x.to(DEV), ld, lp)
        lp += prior.compute_log_prob(zs[-1])
        loss = nf.calculate_loss(ld + lp, 32.0, S * S * 3.0)
        loss.backward()
        errs = {}
        for k, p in flow.named_parameters():
            errs[k] = float((p.grad.cpu() - g_o[k]).norm() / (g_o[k].norm() + 1e-30))
        kinds = {}
        for k, e in errs.items():
            kind = k.split(".")[-1] if "net" not in k else "net." + ".".join(k.split(".net.")[1].split(".")[:1]) + "." + k.split(".")[-1]
            kinds[kind] = max(kinds.get(kind, 0.0), e)
        worst = max(errs, key=errs.get)
        tot = float(torch.sqrt(sum(((p.grad.cpu() - g_o[k]) ** 2).sum() for k, p in flow.named_parameters()) /
                               sum((g_o[k] ** 2).sum() for k, p in flow.named_parameters())))
        print(json.dumps({"cfg": f"c{c} L{L} K{K} B{B} S{S}", "mode": mode, "loss_abs_err": abs(float(loss) - float(lo)),
                          "all_params_rel_l2": tot, "worst": [worst, errs[worst]],
                          "worst_by_kind": {k: float(f"{v:.2e}") for k, v in kinds.items()}}))
