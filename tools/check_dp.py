"""Data-parallel correctness on real GPUs (run under torchrun, N >= 2): every rank trains on ITS shard of a global batch
with the NCCL gradient all-reduce (nf.GradAllReduce); rank 0 also computes the gradient of the WHOLE batch alone.  The
averaged sharded gradients must equal the single-process gradients (the loss is a batch mean, utils.py:256), and after one
FusedClipAdam step the parameters of all ranks must be bit-identical."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200"))
import torch
import torch.distributed as dist
import normalizing_flow as nf
import synthetic as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
c, L, K, S, Bg = 3, 3, 4, 32, 8 * world
sd, psd = O.seeded_state(c, L, K, 3)
xg = O.seeded_input((Bg, c, S, S), 4).to(dev)


def make():
    f = nf.Glow(c, L, K).to(dev); f.load_state_dict(sd)
    p = nf.GaussianPrior(2 ** (L + 1) * c).to(dev); p.load_state_dict(psd)
    return f, p


def step(f, p, x):
    ld, lp = nf.initialize_with_zeros(2, x.shape[0], dev)
    zs, ld, lp = f.transform(x, ld, lp)
    lp += p.compute_log_prob(zs[-1])
    loss = nf.calculate_loss(ld + lp, 32.0, S * S * 3.0)
    loss.backward()
    return loss


flow, prior = make()
dp = nf.GradAllReduce(flow, prior)
dp.broadcast_parameters(src=0)
x = xg[nf.shard(Bg, rank, world)]
loss = step(flow, prior, x)
dp.finish()
torch.cuda.synchronize()
ok = True
if rank == 0:
    dp.detach()
    ref_f, ref_p = make()
    step(ref_f, ref_p, xg)
    worst = 0.0
    for (k, a), (_, b) in zip(list(flow.named_parameters()) + list(prior.named_parameters()),
                              list(ref_f.named_parameters()) + list(ref_p.named_parameters())):
        d = float((a.grad - b.grad).norm() / (b.grad.norm() + 1e-30))
        worst = max(worst, d)
    print(f"world={world} sharded+allreduced vs whole-batch gradients: worst relative difference {worst:.3e}")
    ok = worst < 5e-2          # bf16 GEMMs: different batch split -> different rounding; fp32 mode gives ~1e-6
opt = nf.FusedClipAdam(list(flow.parameters()) + list(prior.parameters()), lr=1e-4, clip_params=list(flow.parameters()))
opt.step()
chk = torch.stack([p.detach().double().sum() for p in flow.parameters()]).sum().reshape(1)
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(float(a) == float(allc[0]) for a in allc)
if rank == 0:
    print("parameters bit-identical across ranks after the optimiser step:", same)
    print("DP CHECK", "OK" if (ok and same) else "FAILED")

# --- data-dependent ActNorm initialisation over the GLOBAL batch (dp.global_initialization): every rank runs the init pass
# on its shard, statistics are combined layer by layer; must equal one process initialising on the whole batch
os.environ["NFDPM_PRECISION"] = "fp32"
torch.manual_seed(0)
f_dp = nf.Glow(c, L, K).to(dev)
dp2 = nf.GradAllReduce(f_dp)
dp2.broadcast_parameters(src=0)
with torch.no_grad(), dp2.global_initialization():
    ld, lp = nf.initialize_with_zeros(2, x.shape[0], dev)
    f_dp.transform(x, ld, lp)
sd_dp = {k: v.clone() for k, v in f_dp.state_dict().items()}
flat = torch.cat([v.double().reshape(-1) for k, v in sd_dp.items() if "actnorm" in k and v.dtype == torch.float32])
allf = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(allf, flat)
same_init = all(torch.equal(a, allf[0]) for a in allf)
if rank == 0:
    torch.manual_seed(0)
    f_one = nf.Glow(c, L, K).to(dev)
    with torch.no_grad():
        ld, lp = nf.initialize_with_zeros(2, Bg, dev)
        f_one.transform(xg, ld, lp)
    worst = 0.0
    for k, v in f_one.state_dict().items():
        if "actnorm" in k and v.dtype == torch.float32:
            worst = max(worst, float((sd_dp[k] - v).abs().max()))
    inited = all(int(v) == 1 for k, v in sd_dp.items() if k.endswith("is_initialized"))
    print(f"global data-dependent init over {world} shards vs whole-batch init: worst |d scale/bias| {worst:.3e}; "
          f"replicas bit-identical: {same_init}; all flags set: {inited}")
    print("DP INIT CHECK", "OK" if (worst < 1e-4 and same_init and inited) else "FAILED")
dist.destroy_process_group()
