"""Runs INSIDE ITS OWN PROCESS (tests/test_gpu_dropin_overlay.py): builds INTEGRATION.md's Option-A overlay — the
reference's own ``normalizing_flow/__init__.py`` and ``trainer.py`` on top of this repository's hot-path modules — and
drives the reference's UNCHANGED training / evaluation / sampling code over it on cuda:0, the way
``run_baseline_experiment.py:42-49,89-101`` does.  Prints one JSON line with what it observed; the pytest side checks
it against the CPU oracle's own trajectory.

usage: dropin_driver.py <reference tree> <product package dir> <libnfdpm_b200.so> <work dir> <repo root>
"""
import json
import logging
import os
import shutil
import sys
from unittest.mock import MagicMock

ref, pkg, lib, work, root = sys.argv[1:6]

# ---- INTEGRATION.md §2, Option A, step by step
ov = os.path.join(work, "overlay")
nf_dir = os.path.join(ov, "normalizing_flow")
shutil.copytree(os.path.join(ref, "normalizing_flow"), nf_dir)                      # cp -r $REF/normalizing_flow overlay_nf
shutil.copy(os.path.join(nf_dir, "utils.py"), os.path.join(nf_dir, "_ref_utils.py"))  # keep the non-hot-path helpers
for f in ("_native.py", "_engine.py", "_train.py", "_modgrad.py", "_dp.py", "_optim.py", "base.py", "transforms.py", "glow.py", "prior.py",
          "utils.py"):
    shutil.copy(os.path.join(pkg, "normalizing_flow", f), os.path.join(nf_dir, f))
with open(os.path.join(nf_dir, "utils.py"), "a") as fh:
    fh.write("\nfrom ._ref_utils import track_images, save_images, get_data_transforms  # noqa\n")
os.environ["NFDPM_B200_LIB"] = lib
# __init__.py and trainer.py are the reference's own, untouched
for f in ("__init__.py", "trainer.py"):
    assert open(os.path.join(nf_dir, f)).read() == open(os.path.join(ref, "normalizing_flow", f)).read()

for m in ["aim", "skimage", "skimage.transform", "cleanfid", "cleanfid.fid", "cleanfid.features", "cleanfid.utils",
          "cleanfid.resize", "ignite", "ignite.metrics"]:                           # logging / dataset / metric packages
    sys.modules.setdefault(m, MagicMock())
sys.path[:0] = [ov, ref, root]

import torch  # noqa: E402
import normalizing_flow as nf  # noqa: E402
from normalizing_flow import trainer as T  # noqa: E402

assert os.path.realpath(nf.__file__).startswith(os.path.realpath(ov))
assert hasattr(nf.glow, "E") and hasattr(nf.transforms, "N"), "the overlaid hot path is not the B200 one"
assert "aim" in open(T.__file__).read() and T.train.__module__ == "normalizing_flow.trainer"

torch.manual_seed(0)
dev = torch.device("cuda")
c, L, K, B, S, n_bits = 3, 3, 2, 8, 32, 5
g = torch.Generator().manual_seed(3)
batches = [torch.rand(B, c, S, S, generator=g) for _ in range(3)]                   # images in [0, 1] (ToTensor output)
test_batches = [torch.rand(B, c, S, S, generator=g) for _ in range(2)]

# the data loaders are outside the hot path: synthetic ones (a list iterates like a DataLoader of tensors)
T.read_dataset = lambda **kw: (batches, None, test_batches, batches[:1])

rec = {"losses": [], "inputs": [], "snapshots": [], "bpd_inputs": [], "phase": "train"}
orig_loss = T.calculate_loss


def loss_spy(ll, n_bins, n_pixel):
    out = orig_loss(ll, n_bins, n_pixel)
    rec["losses"].append(float(out.detach()))
    return out


T.calculate_loss = loss_spy

# run_baseline_experiment.py:42-49
flow = nf.Glow(in_channel=c, L=L, K=K, learn_prior_mean_logs=True)
flow.to(flow.device)
prior = nf.GaussianPrior(in_channels=2 ** (L + 1) * c, learn_prior_mean_logs=True)
# non-zero ZeroConvs so that the coupling networks matter from the first step (the reference initialises them to zero)
with torch.no_grad():
    gz = torch.Generator().manual_seed(4)
    for name, p in list(flow.named_parameters()) + list(prior.named_parameters()):
        if ".net.4." in name or ".split.conv." in name or "_GaussianPrior__conv." in name:
            p.add_((0.003 * torch.randn(p.shape, generator=gz)).to(p.device))

orig_transform = flow.transform


def transform_spy(x, ld, lp):
    if torch.is_grad_enabled():                     # a training step (data-dependent init and bpd evaluation are no_grad)
        rec["inputs"].append(x.detach().cpu().clone())
        rec["snapshots"].append(({k: v.detach().cpu().clone() for k, v in flow.state_dict().items()},
                                 {k: v.detach().cpu().clone() for k, v in prior.state_dict().items()}))
    elif rec["phase"] == "bpd":
        rec["bpd_inputs"].append(x.detach().cpu().clone())
    return orig_transform(x, ld, lp)


flow.transform = transform_spy
logger = logging.getLogger("dropin")
ckpt, res = os.path.join(work, "checkpoints"), os.path.join(work, "results")
os.makedirs(ckpt), os.makedirs(res)
# the reference's save_images / track_images helpers write PNGs / Aim records: not part of the path
nf.trainer.save_images = lambda *a, **k: None
nf.trainer.track_images = lambda *a, **k: None

# run_baseline_experiment.py:89-101 -> the reference's own train(): data-dependent init, 3 iterations of
# transform / prior / bpd loss / backward / clip / Adam, checkpoint, calculate_bpd on the test and train loaders
orig_bpd = T.calculate_bpd


def bpd_spy(*a, **k):
    rec["phase"] = "bpd"
    out = orig_bpd(*a, **k)
    rec.setdefault("bpds", []).append(float(out))
    rec["phase"] = "train"
    return out


T.calculate_bpd = bpd_spy
T.train(flow, prior, logger=logger, experiment_name="dropin", exp_output_dir="dropin", data_root="", data_name="MNIST",
        transformations=[], batch_size=B, num_workers=0, optim_name="adam", lr=1e-4, n_epochs=1, print_freq=1,
        save_checkpoint_freq=5, log_param_distribution=False, log_gen_images_per_iter=100, device=flow.device,
        checkpoint_dir=ckpt, result_dir=res, resume_info=None, img_size=S, n_bits=n_bits, temperature=0.7, digits=None,
        fid_kwargs=[], kid_kwargs=[], ssim_psnr_kwargs=None)
final_sd = ({k: v.detach().cpu().clone() for k, v in flow.state_dict().items()},
            {k: v.detach().cpu().clone() for k, v in prior.state_dict().items()})

# sampling as the trainer / evaluation code does it (trainer.py:196-198): prior sample -> Glow.sample
flow.transform = orig_transform
shapes = nf.calculate_output_shapes(L=flow.L, in_channels=flow.in_channel, size=S)
last = prior.sample(shape=(4, *shapes[-1]), temperature=0.0)
img0 = flow.sample([last], temperature=0.0)
img1 = flow.sample([prior.sample(shape=(4, *shapes[-1]), temperature=0.7)], temperature=0.7)
files = sorted(os.listdir(ckpt))
torch.save({"inputs": rec["inputs"], "snapshots": rec["snapshots"], "bpd_inputs": rec["bpd_inputs"], "final": final_sd,
            "last": last.cpu(), "img0": img0.cpu()}, os.path.join(work, "observed.pt"))
print(json.dumps({"losses": rec["losses"], "bpds": rec["bpds"], "n_inputs": len(rec["inputs"]),
                  "n_bpd_inputs": len(rec["bpd_inputs"]), "checkpoints": files,
                  "img1_finite": bool(torch.isfinite(img1).all()), "img1_shape": list(img1.shape),
                  "train_mode_after_sample": bool(flow.training)}))
