"""Training path (GPU): loss and parameter gradients of one training step through the drop-in module API
(``Glow.transform`` + ``GaussianPrior.compute_log_prob`` under autograd -> hand-written backward kernels) against
(a) the unmodified reference's autograd (tests/golden/glow_grad_*.npz) and (b) the CPU oracle on fresh seeded inputs.

Tolerances: exact fp32 (NFDPM_PRECISION=fp32_simt, CUDA-core GEMMs) — every gradient tensor within 2e-4 relative L2 of
the reference (fp32 summation order differs); fp32-faithful tensor-core mode (NFDPM_PRECISION=fp32, split bf16 pairs in the
stash, the dgrad and the weight-gradient GEMMs) — loss within 1e-5, the whole gradient within 1e-4 relative L2 (measured
1.4e-5), single tensors within 1e-2 (X3_TOL below); bf16 tensor-core mode (stated) — within 8e-2 relative L2 per tensor
(measured worst 4.2e-2: the first conv's weight at the deepest level, three bf16 GEMMs downstream of the loss), loss within
1e-3 bits/dim.
"""
import os

import numpy as np
import pytest
import torch

import normalizing_flow as nf
from oracle import glow_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")
GRAD_CASES = ["glow_grad_c3_L2_K2_b3_s16", "glow_grad_c1_L3_K1_b2_s32"]


def _build(c, L, K, seed):
    sd, psd = O.seeded_state(c, L, K, seed)
    flow = nf.Glow(c, L, K).to(DEV)
    flow.load_state_dict(sd, strict=True)
    prior = nf.GaussianPrior(2 ** (L + 1) * c).to(DEV)
    prior.load_state_dict(psd, strict=True)
    return flow, prior, sd, psd


def _train_step(flow, prior, x, S):
    """normalizing_flow/trainer.py:154-164"""
    B = x.shape[0]
    ld, lp = nf.initialize_with_zeros(2, B, DEV)
    zs, ld, lp = flow.transform(x, ld, lp)
    lp += prior.compute_log_prob(zs[-1])          # in place, like trainer.py:156
    loss = nf.calculate_loss(ld + lp, 32.0, S * S * 3.0)
    loss.backward()
    return loss


#: fp32-faithful (split bf16 pair) training on the tensor cores: per-tensor relative L2 of the gradients.  The error of a
#: tensor is cancellation-dominated (entries ~1e-5 that are sums of ~1e-3 terms): measured worst 5.6e-3 in this file's
#: cases against 2.8e-4 for exact fp32 FMA arithmetic on such tensors; the WHOLE gradient is within 1.4e-5 relative L2
#: (exact fp32: 9.2e-6, bf16: 1.3e-3) — tools/measure_grad_parity.py, profiles/r02_grad_parity.jsonl.
X3_TOL = 1e-2
X3_TOTAL_TOL = 1e-4


def _rel(a, b):
    a, b = a.detach().cpu().double().reshape(-1), torch.as_tensor(b).detach().cpu().double().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("name", GRAD_CASES)
@pytest.mark.parametrize("mode,tol", [("fp32_simt", 2e-4), ("fp32", X3_TOL), ("bf16", 8e-2)])
def test_gradients_against_reference_golden(golden_dir, name, mode, tol, monkeypatch):
    monkeypatch.setenv("NFDPM_PRECISION", mode)
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    c, L, K, B, S, seed, _ = [int(v) for v in g["cfg"]]
    flow, prior, sd, psd = _build(c, L, K, seed)
    assert np.allclose(np.array(O.state_checksum(sd)), g["checksum"], atol=1e-9)
    x = torch.from_numpy(g["x"]).to(DEV)
    loss = _train_step(flow, prior, x, S)
    assert abs(float(loss) - float(g["loss"])) < (1e-5 if mode != "bf16" else 1e-3)
    params = dict(flow.named_parameters())
    params.update({"prior/" + k: p for k, p in prior.named_parameters()})
    for i, k in enumerate(g["names"]):
        k = str(k)
        p = params[k]
        assert p.grad is not None, k
        assert p.grad.shape == p.shape and p.grad.dtype == torch.float32, k
        ref_n, ref_p = g["sig/" + k]
        nrm, proj = O.grad_signature(p.grad.cpu(), 5000 + i)
        if ref_n == 0.0:
            assert nrm == 0.0, k
            continue
        # |proj - ref_p| <= ||g - g_ref|| * ||r|| ~ tol * ref_n * sqrt(n)
        assert abs(nrm - ref_n) <= tol * ref_n, (k, nrm, ref_n)
        assert abs(proj - ref_p) <= 4 * tol * ref_n * np.sqrt(p.numel()), (k, proj, ref_p)
        if ("grad/" + k) in g.files:
            assert _rel(p.grad, g["grad/" + k]) <= tol, (k, _rel(p.grad, g["grad/" + k]))


@pytest.mark.parametrize("cfg", [(3, 3, 2, 4, 32, 71), (1, 3, 2, 5, 32, 72)])
@pytest.mark.parametrize("mode,tol", [("fp32_simt", 2e-4), ("fp32", X3_TOL), ("bf16", 8e-2)])
def test_gradients_against_oracle(cfg, mode, tol, monkeypatch):
    """Every gradient tensor, full comparison, on seeded inputs at a three-level shape (incl. d loss / d x)."""
    monkeypatch.setenv("NFDPM_PRECISION", mode)
    c, L, K, B, S, seed = cfg
    flow, prior, sd, psd = _build(c, L, K, seed)
    x_cpu = O.seeded_input((B, c, S, S), seed + 1)
    loss_o, g_o, pg_o = O.train_grads(sd, psd, x_cpu, L, K, 32.0, S * S * 3.0)
    x = x_cpu.to(DEV).requires_grad_(True)
    loss = _train_step(flow, prior, x, S)
    assert abs(float(loss) - float(loss_o)) < (1e-5 if mode != "bf16" else 1e-3)
    worst = ("", 0.0)
    for k, p in flow.named_parameters():
        r = _rel(p.grad, g_o[k])
        if r > worst[1]:
            worst = (k, r)
        assert r <= tol, (k, r)
    for k, p in prior.named_parameters():
        if float(pg_o[k].abs().max()) == 0.0:
            assert float(p.grad.abs().max()) == 0.0
        else:
            assert _rel(p.grad, pg_o[k]) <= tol, k
    # input gradient: oracle via autograd on x
    xo = x_cpu.clone().requires_grad_(True)
    with torch.enable_grad():
        O.nll_bpd(sd, psd, xo, L, K, 32.0, S * S * 3.0).backward()
    assert _rel(x.grad, xo.grad) <= tol
    if mode == "fp32":                       # the whole gradient vector, where cancellation inside single tensors averages out
        num = sum(float((p.grad.cpu().double() - g_o[k].double()).pow(2).sum()) for k, p in flow.named_parameters())
        den = sum(float(g_o[k].double().pow(2).sum()) for k, p in flow.named_parameters())
        assert (num / den) ** 0.5 <= X3_TOTAL_TOL, (num / den) ** 0.5
    print("worst parameter", worst)


@pytest.mark.parametrize("cfg", [(3, 2, 1, 2, 64, 75), (3, 3, 1, 1, 128, 76)])
@pytest.mark.parametrize("mode,tol", [("fp32_simt", 2e-3), ("fp32", X3_TOL), ("bf16", 8e-2)])
def test_gradients_large_images(cfg, mode, tol, monkeypatch):
    """Images larger than one CTA (BASELINE config 4 geometry: 64x64 / 32x32 / 16x16 levels with 12 / 24 / 48 channels):
    unfused forward with stash, pixel-tiled coupling / K-A backward kernels.  fp32 tolerance 2e-3: the weight-gradient
    sums run over up to 16 384 pixels in a different order than the reference's (measured worst 8.2e-4)."""
    monkeypatch.setenv("NFDPM_PRECISION", mode)
    c, L, K, B, S, seed = cfg
    flow, prior, sd, psd = _build(c, L, K, seed)
    x_cpu = O.seeded_input((B, c, S, S), seed + 1)
    loss_o, g_o, pg_o = O.train_grads(sd, psd, x_cpu, L, K, 32.0, S * S * 3.0)
    loss = _train_step(flow, prior, x_cpu.to(DEV), S)
    assert abs(float(loss) - float(loss_o)) < (1e-5 if mode != "bf16" else 1e-3)
    for k, p in flow.named_parameters():
        assert _rel(p.grad, g_o[k]) <= tol, (k, _rel(p.grad, g_o[k]))


@pytest.mark.parametrize("mode,tol,ztol", [("fp32_simt", 2e-4, 1e-4), ("fp32", X3_TOL, 1e-4), ("bf16", 8e-2, 5e-3)])
def test_ragged_shapes_mnist_28(mode, tol, ztol, monkeypatch):
    """Non-power-of-two geometry (MNIST 1x28x28, L=2: 14x14 = 196 and 7x7 = 49 pixels per image, batch 5, K=3): forward,
    inverse and every gradient against the oracle — neither level fits the whole-image GEMM tiling, ragged tail tiles."""
    monkeypatch.setenv("NFDPM_PRECISION", mode)
    c, L, K, B, S, seed = 1, 2, 3, 5, 28, 83
    flow, prior, sd, psd = _build(c, L, K, seed)
    x_cpu = O.seeded_input((B, c, S, S), seed + 1)
    with torch.no_grad():
        ld = torch.zeros(B, dtype=torch.float64, device=DEV)
        lp = torch.zeros(B, dtype=torch.float64, device=DEV)
        zs, ld, lp = flow.transform(x_cpu.to(DEV), ld, lp)
        xr = flow.invert(zs)
    ld_o, lp_o = torch.zeros(B, dtype=torch.float64), torch.zeros(B, dtype=torch.float64)
    zo, ld_o, lp_o = O.glow_transform(sd, x_cpu, L, K, ld_o, lp_o)
    for a, b in zip(zs, zo):
        assert _rel(a, b) < ztol
    assert torch.allclose(ld.cpu(), ld_o, rtol=1e-4 if mode != "bf16" else 2e-4)
    assert torch.allclose(lp.cpu(), lp_o, rtol=1e-4 if mode != "bf16" else 2e-3)
    assert float((xr.cpu() - x_cpu).abs().max()) < (1e-4 if mode != "bf16" else 5e-2)
    loss_o, g_o, pg_o = O.train_grads(sd, psd, x_cpu, L, K, 32.0, S * S * 3.0)
    loss = _train_step(flow, prior, x_cpu.to(DEV), S)
    assert abs(float(loss) - float(loss_o)) < (1e-5 if mode != "bf16" else 1e-3)
    for k, p in flow.named_parameters():
        assert _rel(p.grad, g_o[k]) <= tol, (k, _rel(p.grad, g_o[k]))


def test_batch_of_one_and_repeated_backward(monkeypatch):
    """B = 1 (every per-image reduction degenerates) and two training steps in a row on the same module (the stash of
    the first step is released, caches are refreshed after the parameter update)."""
    monkeypatch.setenv("NFDPM_PRECISION", "fp32_simt")
    c, L, K, B, S, seed = 3, 3, 1, 1, 32, 85
    flow, prior, sd, psd = _build(c, L, K, seed)
    x_cpu = O.seeded_input((B, c, S, S), seed + 1)
    loss_o, g_o, _ = O.train_grads(sd, psd, x_cpu, L, K, 32.0, S * S * 3.0)
    for it in range(2):
        flow.zero_grad(set_to_none=True)
        prior.zero_grad(set_to_none=True)
        loss = _train_step(flow, prior, x_cpu.to(DEV), S)
        assert abs(float(loss) - float(loss_o)) < 1e-5
        for k, p in flow.named_parameters():
            assert _rel(p.grad, g_o[k]) <= 2e-4, (it, k, _rel(p.grad, g_o[k]))
    # gradient ACCUMULATION (p.grad already populated): a third backward adds to the existing gradients
    loss = _train_step(flow, prior, x_cpu.to(DEV), S)
    for k, p in flow.named_parameters():
        assert _rel(p.grad, 2 * g_o[k]) <= 2e-4, (k, _rel(p.grad, 2 * g_o[k]))


def test_logp_none_and_latent_gradients(monkeypatch):
    """NFBackbone-style call (logp=None, diffusion_prior/trainer.py:139): gradients arrive through the latents."""
    monkeypatch.setenv("NFDPM_PRECISION", "fp32_simt")
    c, L, K, B, S, seed = 3, 3, 1, 3, 16, 81
    flow, _, sd, _ = _build(c, L, K, seed)
    x_cpu = O.seeded_input((B, c, S, S), seed + 1)
    rng = np.random.default_rng(5)
    ld = torch.zeros(B, dtype=torch.float64, device=DEV)
    zs, ld, none = flow.transform(x_cpu.to(DEV), ld, None)
    assert none is None
    ws = [torch.from_numpy(rng.standard_normal(tuple(z.shape)).astype(np.float32)) for z in zs]
    loss = sum((z * w.to(DEV)).sum() for z, w in zip(zs, ws)) + 0.3 * ld.sum()
    loss.backward()
    sd_g = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in sd.items()}
    with torch.enable_grad():
        ld_o = torch.zeros(B, dtype=torch.float64)
        zo, ld_o, _ = O.glow_transform(sd_g, x_cpu, L, K, ld_o, None)
        lo = sum((z * w).sum() for z, w in zip(zo, ws)) + 0.3 * ld_o.sum()
        lo.backward()
    assert abs(float(loss) - float(lo)) <= 1e-4 * abs(float(lo))
    for k, p in flow.named_parameters():
        ref = sd_g[k].grad
        if ref is None:
            assert float(p.grad.abs().max()) == 0.0, k
        else:
            assert _rel(p.grad, ref) <= 2e-4, (k, _rel(p.grad, ref))


def test_adam_steps_track_the_oracle(monkeypatch):
    """Three optimiser steps of the reference recipe (clip value 1, clip norm 1, Adam 1e-4; trainer.py:161-167):
    the loss trajectory follows the oracle's."""
    monkeypatch.setenv("NFDPM_PRECISION", "fp32_simt")
    c, L, K, B, S, seed = 3, 3, 1, 4, 16, 91
    flow, prior, sd, psd = _build(c, L, K, seed)
    x_cpu = O.seeded_input((B, c, S, S), seed + 1)
    x = x_cpu.to(DEV)
    params = list(flow.parameters()) + list(prior.parameters())
    opt = torch.optim.Adam(params, lr=1e-4)
    sd_g = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in sd.items()}
    psd_g = {k: v.clone().requires_grad_(True) for k, v in psd.items()}
    po = [v for v in sd_g.values() if v.dtype.is_floating_point] + list(psd_g.values())
    opt_o = torch.optim.Adam(po, lr=1e-4)
    for it in range(3):
        opt.zero_grad(set_to_none=True)
        loss = _train_step(flow, prior, x, S)
        torch.nn.utils.clip_grad_value_(params, 1.0)
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        opt_o.zero_grad(set_to_none=True)
        with torch.enable_grad():
            lo = O.nll_bpd(sd_g, psd_g, x_cpu, L, K, 32.0, S * S * 3.0)
            lo.backward()
        torch.nn.utils.clip_grad_value_(po, 1.0)
        torch.nn.utils.clip_grad_norm_(po, 1.0)
        opt_o.step()
        assert abs(float(loss) - float(lo)) < 2e-4, (it, float(loss), float(lo))
    assert float(loss) < 1e9


@pytest.mark.parametrize("decoupled,wd", [(False, 0.0), (True, 1e-2)])
def test_fused_clip_adam_matches_torch(decoupled, wd):
    """FusedClipAdam.step() == clip_grad_value_(1) + clip_grad_norm_(1) + torch Adam/AdamW step (trainer.py:165-167,
    utils.py:120-137) on identical gradients: parameters, moments and the in-place clipped gradients."""
    rng = np.random.default_rng(3)
    shapes = [(512, 6, 3, 3), (512, 1, 1), (12, 12, 1, 1), (5000,), (1, 7, 1, 1), (3,), (96, 96, 3, 3)]
    ps = [torch.nn.Parameter(torch.from_numpy(rng.standard_normal(s).astype(np.float32)).to(DEV)) for s in shapes]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    n_clip = 5                                  # the last two tensors play the prior's parameters: not clipped
    opt = nf.FusedClipAdam(ps, lr=1e-3, weight_decay=wd, decoupled_weight_decay=decoupled, clip_params=ps[:n_clip])
    ref = (torch.optim.AdamW(qs, lr=1e-3, weight_decay=wd) if decoupled else torch.optim.Adam(qs, lr=1e-3))
    for it in range(4):
        scale = [5.0, 0.01, 1.0, 0.3][it]       # exercise both clips, then neither
        for p, q in zip(ps, qs):
            g = torch.from_numpy((scale * rng.standard_normal(tuple(p.shape))).astype(np.float32)).to(DEV)
            p.grad, q.grad = g.clone(), g.clone()
        v0 = ps[0]._version
        opt.step()
        assert ps[0]._version > v0
        torch.nn.utils.clip_grad_value_(qs[:n_clip], 1.0)
        total = torch.nn.utils.clip_grad_norm_(qs[:n_clip], 1.0)
        ref.step()
        assert abs(float(opt.grad_norm) - float(total)) <= 1e-5 * float(total)
        for p, q in zip(ps, qs):
            assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-8)
            assert torch.allclose(p, q, rtol=1e-5, atol=1e-6), float((p - q).abs().max())
            assert torch.allclose(opt.state[p]["exp_avg"], ref.state[q]["exp_avg"], rtol=1e-5, atol=1e-8)
            assert torch.allclose(opt.state[p]["exp_avg_sq"], ref.state[q]["exp_avg_sq"], rtol=2e-5, atol=1e-10)
    assert float(opt.state[ps[0]]["step"]) == 4.0
    # a torch Adam state_dict loads into the fused optimiser
    opt2 = nf.FusedClipAdam(ps, lr=1e-3)
    opt2.load_state_dict(ref.state_dict())
    assert torch.allclose(opt2.state[ps[3]]["exp_avg"], ref.state[qs[3]]["exp_avg"])


def test_checkpoint_resume_is_bit_identical(tmp_path, monkeypatch):
    """save_model -> load (reference wire format, prior.py:102-115) in the middle of training: the resumed run's next step
    equals the uninterrupted run's bit for bit (parameters, moments, step counter), also when the checkpoint is loaded
    into a model whose cached LU / packed-weight state was built from OTHER weights (cache invalidation on
    load_state_dict)."""
    monkeypatch.setenv("NFDPM_PRECISION", "bf16")
    c, L, K, B, S = 3, 2, 2, 4, 16
    x = O.seeded_input((B, c, S, S), 21).to(DEV)

    def make(seed):
        flow, prior, _, _ = _build(c, L, K, seed)
        params = list(flow.parameters()) + list(prior.parameters())
        return flow, prior, nf.FusedClipAdam(params, lr=1e-3, clip_params=list(flow.parameters()))

    def step(flow, prior, opt):
        opt.zero_grad(set_to_none=False)
        loss = _train_step(flow, prior, x, S)
        opt.step()
        return float(loss)

    flow, prior, opt = make(31)
    for _ in range(2):
        step(flow, prior, opt)
    path = nf.save_model(None, flow, prior, opt, 2, 2, str(tmp_path))
    l3 = step(flow, prior, opt)

    flow2, prior2, opt2 = make(32)                     # other weights; one step and one inference call warm every
    step(flow2, prior2, opt2)                          # cache (LU, packed weights, pack plans, graphs) with them
    with torch.no_grad():
        flow2.invert(flow2.transform(x, nf.initialize_with_zeros(1, B, DEV), None)[0])
    ck = torch.load(path, map_location="cpu")
    flow2.load_state_dict(ck["flow"], strict=True)
    prior2.load_state_dict(ck["prior_dist"], strict=True)
    opt2.load_state_dict(ck["optimizer"])
    assert ck["current_iter"] == 2 and float(opt2.state[next(iter(flow2.parameters()))]["step"]) == 2.0
    l3b = step(flow2, prior2, opt2)
    assert l3 == l3b
    for (k, a), b in zip(flow.state_dict().items(), flow2.state_dict().values()):
        assert torch.equal(a, b), k
    for a, b in zip(prior.parameters(), prior2.parameters()):
        assert torch.equal(a, b)
    for p, q in zip(flow.parameters(), flow2.parameters()):
        assert torch.equal(opt.state[p]["exp_avg"], opt2.state[q]["exp_avg"])
        assert torch.equal(opt.state[p]["exp_avg_sq"], opt2.state[q]["exp_avg_sq"])
    with torch.no_grad():                              # the inference path (its own pack plan and graphs) follows too
        za = flow.transform(x, nf.initialize_with_zeros(1, B, DEV), None)[0]
        zb = flow2.transform(x, nf.initialize_with_zeros(1, B, DEV), None)[0]
        assert all(torch.equal(a, b) for a, b in zip(za, zb))
        assert torch.equal(flow.invert(za), flow2.invert(zb))
    # and a torch Adam (what the reference's trainer constructs, utils.py:120-137) accepts the same optimiser state
    ref_opt = nf.init_optimizer("adam", list(flow2.parameters()) + list(prior2.parameters()), 1e-3)
    ref_opt.load_state_dict(ck["optimizer"])
    assert float(ref_opt.state[next(iter(flow2.parameters()))]["step"]) == 2.0


def test_captured_training_chains_equal_eager(monkeypatch):
    """NFDPM_TRAIN_GRAPHS=1: the autograd Function replays captured stash-forward / backward chains.  Same losses,
    gradients and parameters as the eager chain, bit for bit, over several optimiser steps, with zero_grad in both
    flavours and with a forward whose backward never runs in between."""
    monkeypatch.setenv("NFDPM_PRECISION", "bf16")
    c, L, K, B, S = 3, 3, 2, 8, 32
    x = O.seeded_input((B, c, S, S), 61).to(DEV)

    def run(graphs: bool):
        monkeypatch.setenv("NFDPM_TRAIN_GRAPHS", "1" if graphs else "0")
        flow, prior, _, _ = _build(c, L, K, 62)
        params = list(flow.parameters()) + list(prior.parameters())
        opt = nf.FusedClipAdam(params, lr=1e-3, clip_params=list(flow.parameters()))
        losses = []
        for it in range(4):
            opt.zero_grad(set_to_none=(it % 2 == 0))
            if it == 2:                                  # a forward under autograd that is never back-propagated
                ld, lp = nf.initialize_with_zeros(2, B, DEV)
                flow.transform(x, ld, lp)
            losses.append(float(_train_step(flow, prior, x, S).detach()))
            opt.step()
        return losses, [p.detach().clone() for p in params], [p.grad.detach().clone() for p in params]

    le, pe, ge = run(False)
    lg, pg, gg = run(True)
    assert le == lg
    for a, b in zip(ge, gg):
        assert torch.equal(a, b)
    for a, b in zip(pe, pg):
        assert torch.equal(a, b)
