"""The drop-in claim, executed (GPU): INTEGRATION.md's Option-A overlay is built in a temporary directory from a copy of
the reference tree (baseline/_ref, staged by __graft_entry__.build(); skipped when no copy exists) and the reference's
OWN, unmodified ``normalizing_flow.trainer.train`` / ``calculate_bpd`` and ``Glow.sample`` run over this repository's hot
path on cuda:0 (tests/dropin_driver.py, own process: the overlay has the same import name as the product mirror).

Checked against the CPU oracle: (a) every training loss the reference trainer saw equals the oracle's loss on the SAME
input and the SAME parameters (snapshots before each step), (b) the loss trajectory follows the oracle's own run of the
reference recipe — clip value 1, clip norm 1, Adam — from the initial snapshot (trainer.py:150-167), (c) the bits/dim
calculate_bpd reported (trainer.py:22-55) equal the oracle's on the inputs it was fed, (d) Glow.sample at temperature 0
equals the oracle's mean decode, (e) the reference's checkpoint file was written.  fp32 precision mode."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import glow_oracle as O
from oracle import reference_module as RM

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200")


def test_reference_trainer_runs_unchanged_over_the_overlay(tmp_path):
    ref = RM.find()
    if ref is None:
        pytest.skip("no copy of the reference tree (baseline/_ref is staged by __graft_entry__.build())")
    env = dict(os.environ, NFDPM_PRECISION="fp32")
    env.pop("PYTHONPATH", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_driver.py"), ref, PKG,
                        os.path.join(PKG, "lib", "libnfdpm_b200.so"), str(tmp_path), ROOT],
                       capture_output=True, text=True, env=env, timeout=900, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-4000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    obs = torch.load(os.path.join(tmp_path, "observed.pt"), weights_only=False)
    c, L, K, S, n_bins = 3, 3, 2, 32, 32.0
    n_pixel = S * S * 3.0
    assert out["n_inputs"] == 3 and len(out["losses"]) == 3 and out["n_bpd_inputs"] == 3      # 2 test + 1 train batches
    assert out["checkpoints"] == ["model_gaussian_001.pt"], out["checkpoints"]
    assert out["img1_finite"] and out["img1_shape"] == [4, c, S, S] and out["train_mode_after_sample"], out
    # (a) loss of every step == oracle on the same input and the same parameters
    for t in range(3):
        sd, psd = obs["snapshots"][t]
        lo = float(O.nll_bpd(sd, psd, obs["inputs"][t], L, K, n_bins, n_pixel))
        assert abs(out["losses"][t] - lo) < 2e-5, (t, out["losses"][t], lo)
    # (b) the oracle's own trajectory of the reference recipe from the first snapshot, on the same inputs
    sd0, psd0 = obs["snapshots"][0]
    sd_g = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in sd0.items()}
    psd_g = {k: v.clone().requires_grad_(True) for k, v in psd0.items()}
    po = [v for v in sd_g.values() if v.dtype.is_floating_point]          # the trainer optimises flow.parameters() only
    opt = torch.optim.Adam(po, lr=1e-4)
    for t in range(3):
        opt.zero_grad(set_to_none=True)
        lo = O.nll_bpd(sd_g, psd_g, obs["inputs"][t], L, K, n_bins, n_pixel)
        lo.backward()
        torch.nn.utils.clip_grad_value_(po, 1.0)
        torch.nn.utils.clip_grad_norm_(po, 1.0)
        opt.step()
        assert abs(out["losses"][t] - float(lo)) < 5e-4, (t, out["losses"][t], float(lo))
    fin = obs["final"][0]
    # 3 Adam steps of lr 1e-4: the two parameter trajectories agree up to sign flips of near-zero gradients
    num = sum(float((fin[k] - sd_g[k].detach()).double().pow(2).sum()) for k in fin if fin[k].dtype.is_floating_point)
    den = sum(float((sd0[k] - sd_g[k].detach()).double().pow(2).sum()) for k in fin if fin[k].dtype.is_floating_point)
    assert den > 0 and (num / den) ** 0.5 < 0.05, (num, den)
    # (c) calculate_bpd: test loader (2 batches) then train loader (1 batch), final parameters
    fsd, fpsd = obs["final"]
    per = [float(O.nll_bpd(fsd, fpsd, x, L, K, n_bins, n_pixel)) for x in obs["bpd_inputs"]]
    assert abs(out["bpds"][0] - float(np.mean(per[:2]))) < 2e-5 and abs(out["bpds"][1] - per[2]) < 2e-5, (out["bpds"], per)
    # (d) Glow.sample([last], T=0) == oracle inverse with every Split at its prior mean
    x0 = O.glow_invert(fsd, [obs["last"]], L, K, temperature=0.0)
    assert float((obs["img0"] - x0).abs().max()) < 1e-4
