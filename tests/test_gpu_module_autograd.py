"""Stand-alone transforms are ordinary autograd modules in the reference (normalizing_flow/transforms.py:56-309,
glow.py:46-48,107-111).  Here their ``transform`` goes through normalizing_flow/_modgrad.py (one autograd Function per
call, composed from the backward kernels of the training path).  Every module — ActNorm, InvConv2d, AffineCoupling, Squeeze,
Split, and StepFlow / GlowBlock which compose them — against torch autograd over the CPU oracle's functional restatement:
outputs, the in-place accumulators, d/dx and every parameter gradient.  Exact fp32 mode (fp32_simt): 2e-4 relative L2 per
tensor; the tensor-core modes are covered through the whole-Glow tests."""
import numpy as np
import pytest
import torch

import normalizing_flow as nf
from oracle import glow_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")
TOL = 2e-4


@pytest.fixture(autouse=True)
def _exact(monkeypatch):
    monkeypatch.setenv("NFDPM_PRECISION", "fp32_simt")
    torch.set_grad_enabled(True)


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def rnd(*shape, seed=0, scale=1.0):
    return torch.from_numpy((scale * np.random.default_rng(seed).standard_normal(shape)).astype(np.float32))


def step_state(C, seed):
    """state_dict of one StepFlow with seeded, non-degenerate weights (prefix '')."""
    sd, _ = O.seeded_state(C // 4, 2, 1, seed)          # level 0 of a Glow whose squeezed width is C
    pre = "blocks.0.flows.0."
    return {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}


def leafify(sd):
    return {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in sd.items()}


def check_params(mod, sd_g, prefix=""):
    for k, p in mod.named_parameters():
        ref = sd_g[prefix + k].grad
        assert p.grad is not None, k
        if ref is None or float(ref.abs().max()) == 0.0:
            assert float(p.grad.abs().max()) < 1e-6, k
        else:
            assert rel(p.grad, ref) <= TOL, (k, rel(p.grad, ref))


@pytest.mark.parametrize("B,C,H,W", [(3, 12, 8, 8), (2, 4, 16, 16), (5, 48, 4, 4)])
def test_stepflow_and_its_parts(B, C, H, W):
    sd = step_state(C, 31)
    x = rnd(B, C, H, W, seed=32)
    wy, wl = rnd(B, C, H, W, seed=33), rnd(B, seed=34).double()
    # ---- whole StepFlow
    step = nf.StepFlow(C).to(DEV)
    step.load_state_dict(sd)
    xg = x.to(DEV).requires_grad_(True)
    ld = torch.zeros(B, dtype=torch.float64, device=DEV)
    y, ld, _ = step.transform(xg, ld, None)
    loss = (y * wy.to(DEV)).sum() + (ld * wl.to(DEV)).sum()
    loss.backward()
    sd_g = leafify(sd)
    xo = x.clone().requires_grad_(True)
    ld_o = torch.zeros(B, dtype=torch.float64)
    yo = O.step_fwd(xo, sd_g, "", ld_o)
    ((yo * wy).sum() + (ld_o * wl).sum()).backward()
    assert rel(y, yo) <= 1e-5 and rel(ld, ld_o) <= 1e-5
    assert rel(xg.grad, xo.grad) <= TOL
    check_params(step, sd_g)
    # ---- the three parts on their own
    an = nf.ActNorm(C).to(DEV)
    an.load_state_dict({k[len("actnorm."):]: v for k, v in sd.items() if k.startswith("actnorm.")})
    xg = x.to(DEV).requires_grad_(True)
    ld = torch.zeros(B, dtype=torch.float32, device=DEV)
    y, ld, _ = an.transform(xg, ld, None)
    ((y * wy.to(DEV)).sum() + (ld * wl.float().to(DEV)).sum()).backward()
    sg = leafify(sd)
    xo = x.clone().requires_grad_(True)
    yo, d = O.actnorm_fwd(xo, sg["actnorm.scale"], sg["actnorm.bias"])
    ((yo * wy).sum() + (d.expand(B) * wl.float()).sum()).backward()
    assert rel(y, yo) <= 1e-6 and rel(xg.grad, xo.grad) <= TOL
    assert rel(an.scale.grad, sg["actnorm.scale"].grad) <= TOL and rel(an.bias.grad, sg["actnorm.bias"].grad) <= TOL
    ic = nf.InvConv2d(C).to(DEV)
    ic.load_state_dict({"weight": sd["invconv2d.weight"]})
    xg = x.to(DEV).requires_grad_(True)
    ld = torch.zeros(B, dtype=torch.float64, device=DEV)
    y, ld, _ = ic.transform(xg, ld, None)
    ((y * wy.to(DEV)).sum() + (ld * wl.to(DEV)).sum()).backward()
    sg = leafify(sd)
    xo = x.clone().requires_grad_(True)
    yo, d = O.invconv_fwd(xo, sg["invconv2d.weight"])
    ((yo * wy).sum() + (d.double().expand(B) * wl).sum()).backward()
    assert rel(y, yo) <= 1e-5 and rel(xg.grad, xo.grad) <= TOL
    assert rel(ic.weight.grad, sg["invconv2d.weight"].grad) <= TOL
    cp = nf.AffineCoupling(C).to(DEV)
    cp.load_state_dict({k[len("affcoupling."):]: v for k, v in sd.items() if k.startswith("affcoupling.")})
    xg = x.to(DEV).requires_grad_(True)
    ld = torch.zeros(B, dtype=torch.float64, device=DEV)
    y, ld, _ = cp.transform(xg, ld, None)
    ((y * wy.to(DEV)).sum() + (ld * wl.to(DEV)).sum()).backward()
    sg = leafify(sd)
    xo = x.clone().requires_grad_(True)
    yo, d = O.coupling_fwd(xo, sg, "affcoupling.net.")
    ((yo * wy).sum() + (d.double() * wl).sum()).backward()
    assert rel(y, yo) <= 1e-5 and rel(ld, d) <= 1e-5 and rel(xg.grad, xo.grad) <= TOL
    check_params(cp, sg, "affcoupling.")


@pytest.mark.parametrize("learn", [True, False])
def test_squeeze_split_and_glowblock(learn):
    """GlowBlock = Squeeze -> K StepFlows -> Split (glow.py:107-111): outputs (kept half, z), both accumulators, d/dx and
    every parameter gradient incl. the Split prior's ZeroConv."""
    c, K, B, S = 3, 2, 3, 16
    sd_all, _ = O.seeded_state(c, 2, K, 41, learn_prior=learn)
    pre = "blocks.0."
    sd = {k[len(pre):]: v for k, v in sd_all.items() if k.startswith(pre)}
    blk = nf.GlowBlock(c, K, learn_prior_mean_logs=learn).to(DEV)
    blk.load_state_dict(sd)
    x = O.seeded_input((B, c, S, S), 42)
    C, h = 4 * c, S // 2
    wy, wz = rnd(B, C // 2, h, h, seed=43), rnd(B, C // 2, h, h, seed=44)
    wl, wp = rnd(B, seed=45).double(), rnd(B, seed=46).double()
    xg = x.to(DEV).requires_grad_(True)
    ld = torch.zeros(B, dtype=torch.float64, device=DEV)
    lp = torch.zeros(B, dtype=torch.float64, device=DEV)
    y, ld, z, lp = blk.transform(xg, ld, lp)
    ((y * wy.to(DEV)).sum() + (z * wz.to(DEV)).sum() + (ld * wl.to(DEV)).sum() + (lp * wp.to(DEV)).sum()).backward()
    sg = leafify(sd)
    xo = x.clone().requires_grad_(True)
    ld_o = torch.zeros(B, dtype=torch.float64)
    t = O.squeeze2x2(xo)
    for k in range(K):
        t = O.step_fwd(t, sg, f"flows.{k}.", ld_o)
    yo, zo = t.chunk(2, dim=1)
    mean, logs = O.split_prior_params(yo, sg, "split.")
    lp_o = O.gaussian_logp(zo, mean, logs).double()
    ((yo * wy).sum() + (zo * wz).sum() + (ld_o * wl).sum() + (lp_o * wp).sum()).backward()
    assert rel(y, yo) <= 1e-5 and rel(z, zo) <= 1e-5 and rel(ld, ld_o) <= 1e-5 and rel(lp, lp_o) <= 1e-5
    assert rel(xg.grad, xo.grad) <= TOL
    check_params(blk, sg)
    # Squeeze on its own: a permutation, so the gradient is the inverse permutation of the upstream gradient
    xs = rnd(2, 3, 8, 6, seed=47).to(DEV).requires_grad_(True)
    ys, _, _ = nf.Squeeze().transform(xs, None, None)
    w = rnd(2, 12, 4, 3, seed=48)
    (ys * w.to(DEV)).sum().backward()
    assert torch.equal(xs.grad.cpu(), O.unsqueeze2x2(w))


def test_accumulation_and_no_grad_inputs():
    """Gradients accumulate over two backward passes like any autograd module; inputs that do not require grad get none."""
    C, B = 12, 2
    sd = step_state(C, 51)
    step = nf.StepFlow(C).to(DEV)
    step.load_state_dict(sd)
    x = rnd(B, C, 8, 8, seed=52).to(DEV)
    for _ in range(2):
        ld = torch.zeros(B, dtype=torch.float64, device=DEV)
        y, ld, _ = step.transform(x, ld, None)
        (y.sum() + ld.sum()).backward()
    g2 = {k: p.grad.clone() for k, p in step.named_parameters()}
    step.zero_grad(set_to_none=True)
    ld = torch.zeros(B, dtype=torch.float64, device=DEV)
    y, ld, _ = step.transform(x, ld, None)
    (y.sum() + ld.sum()).backward()
    for k, p in step.named_parameters():
        assert torch.allclose(g2[k], 2 * p.grad, rtol=1e-5, atol=1e-7), k
    assert x.grad is None
