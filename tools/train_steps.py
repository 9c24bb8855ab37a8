"""Two eager (no CUDA graph) training steps at BASELINE config 2 — the command profiled by the ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200"))
import torch
import normalizing_flow as nf
import synthetic as O

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda")
c, L, K, S = 3, 3, 16, 32
sd, psd = O.seeded_state(c, L, K, 0)
flow = nf.Glow(c, L, K).to(dev); flow.load_state_dict(sd)
prior = nf.GaussianPrior(2 ** (L + 1) * c).to(dev); prior.load_state_dict(psd)
params = list(flow.parameters()) + list(prior.parameters())
opt = nf.FusedClipAdam(params, lr=1e-4, clip_params=list(flow.parameters()))
x = O.seeded_input((B, c, S, S), 1).to(dev)
for it in range(steps):
    if it == steps - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()       # ncu --profile-from-start off: only the last step is listed
    opt.zero_grad(set_to_none=True)
    ld, lp = nf.initialize_with_zeros(2, B, dev)
    zs, ld, lp = flow.transform(x + torch.rand_like(x) / 32.0, ld, lp)
    lp += prior.compute_log_prob(zs[-1])
    loss = nf.calculate_loss(ld + lp, 32.0, S * S * 3.0)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss.detach()))
