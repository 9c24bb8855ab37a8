"""First eager training steps of a config: loss per step, gradient norm, non-finite counts (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200"))
import torch
import normalizing_flow as nf
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_configs as BC
DEV = torch.device("cuda")
cfg = int(sys.argv[1]); mode = sys.argv[2]; nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
lr = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-4
os.environ["NFDPM_PRECISION"] = mode
c, L, K, B, S, _ = BC.CONFIGS[cfg]
flow, prior, x = BC.build(c, L, K, B, S)
params = list(flow.parameters()) + list(prior.parameters())
opt = nf.FusedClipAdam(params, lr=lr, clip_params=list(flow.parameters()), clip_value=1.0, max_norm=1.0)
torch.manual_seed(3)
for it in range(nsteps):
    opt.zero_grad(set_to_none=True)
    ld, lp = nf.initialize_with_zeros(2, B, DEV)
    zs, ld, lp = flow.transform(x + torch.rand_like(x) / 32.0, ld, lp)
    lp = lp + prior.compute_log_prob(zs[-1])
    loss = nf.calculate_loss(ld + lp, 32.0, S * S * 3.0)
    loss.backward()
    gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in flow.parameters() if p.grad is not None))
    bad = sum(int((~torch.isfinite(p.grad)).sum()) for p in params if p.grad is not None)
    zmax = max(float(z.abs().max()) for z in zs)
    print(f"cfg{cfg} {mode} step {it}: loss={float(loss.detach()):.4f} ld={float(ld.mean()):.1f} lp={float(lp.mean()):.1f} "
          f"|g|={float(gn):.3e} nonfinite_grads={bad} max|z|={zmax:.2f}", flush=True)
    opt.step()
