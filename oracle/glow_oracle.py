"""CPU oracle for the NFDPM Glow hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
shipped package (``normalizing-flow-with-diffusion-prior-model_b200/``) never does.

It is a *functional* restatement of the reference algorithm (the reference is a tree of
``nn.Module`` objects): every function takes plain tensors / a flat ``state_dict`` with the
reference's key names and runs in fp32 on the CPU with fp64 accumulators, exactly as the
reference's trainers do.  Each function cites the reference lines it follows
(paths relative to the reference root).

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified reference
package in the build container, runs it on seeded weights/inputs and writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file against those
vectors (the reference's own tests hold no golden values — only round-trip properties —
so outputs of the reference run here are the pin).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LOG_2PI = float(np.log(2 * np.pi))  # normalizing_flow/prior.py:23


# --------------------------------------------------------------------------- primitives
def squeeze2x2(x: Tensor) -> Tensor:
    """Space-to-depth, out channel = c*4 + h1*2 + w1 (normalizing_flow/transforms.py:226)."""
    b, c, h, w = x.shape
    v = x.reshape(b, c, h // 2, 2, w // 2, 2)          # b c h h1 w w1
    return v.permute(0, 1, 3, 5, 2, 4).reshape(b, c * 4, h // 2, w // 2)


def unsqueeze2x2(y: Tensor) -> Tensor:
    """Depth-to-space, inverse of :func:`squeeze2x2` (normalizing_flow/transforms.py:238)."""
    b, c4, h, w = y.shape
    v = y.reshape(b, c4 // 4, 2, 2, h, w)               # b c c1 c2 h w
    return v.permute(0, 1, 4, 2, 5, 3).reshape(b, c4 // 4, h * 2, w * 2)


def actnorm_stats(x: Tensor) -> Tuple[Tensor, Tensor]:
    """Data-dependent ActNorm parameters (normalizing_flow/transforms.py:74-78).

    scale = -log(std_unbiased + 1e-6), bias = -mean, both per channel, shape (C,1,1)."""
    scale = -torch.log(x.std(dim=(0, 2, 3)) + 1e-6)
    bias = -x.mean(dim=(0, 2, 3))
    return scale.reshape(-1, 1, 1), bias.reshape(-1, 1, 1)


def actnorm_fwd(x: Tensor, scale: Tensor, bias: Tensor) -> Tuple[Tensor, Tensor]:
    """y = exp(scale)*(x+bias); log-det term = H*W*sum(scale) (transforms.py:80-81)."""
    h, w = x.shape[2:]
    return torch.exp(scale) * (x + bias), h * w * scale.sum()


def actnorm_inv(y: Tensor, scale: Tensor, bias: Tensor) -> Tensor:
    """x = y*exp(-scale) - bias (transforms.py:93)."""
    return y * torch.exp(-scale) - bias


def invconv_fwd(x: Tensor, weight: Tensor) -> Tuple[Tensor, Tensor]:
    """1x1 conv with a dense CxC weight; log-det = H*W*log|det W| computed in fp64 and cast
    to fp32 (transforms.py:130-132)."""
    h, w = x.shape[2:]
    c = weight.shape[0]
    ld = h * w * torch.slogdet(weight.reshape(c, c).double())[1].float()
    return F.conv2d(x, weight), ld


def invconv_inv(y: Tensor, weight: Tensor) -> Tensor:
    """1x1 conv with the fp32 dense inverse (transforms.py:144)."""
    c = weight.shape[0]
    return F.conv2d(y, weight.reshape(c, c).inverse().reshape(c, c, 1, 1))


def zeroconv(x: Tensor, weight: Tensor, bias: Tensor, logs: Tensor) -> Tensor:
    """ZeroConv2d: (conv3x3(x, pad 1) + bias) * exp(3*logs) (normalizing_flow/utils.py:36-44)."""
    return F.conv2d(x, weight, bias, padding=1) * torch.exp(logs * 3.0)


def coupling_net(x_a: Tensor, sd: Dict[str, Tensor], pre: str, init: bool = False) -> Tensor:
    """conv3x3(no bias) -> ActNorm -> ReLU -> conv1x1 -> ActNorm -> ReLU -> ZeroConv3x3
    (normalizing_flow/utils.py:83-89, :64-69).  ``pre`` is the prefix of ``...affcoupling.net.``.
    With ``init`` the inner ActNorms whose ``is_initialized`` flag is 0 are set from the data
    first (transforms.py:74-78 reached through utils.py:69)."""
    h = x_a
    for idx, pad in (("0", 1), ("2", 0)):
        base = f"{pre}{idx}._Conv2dActNorm__"
        h = F.conv2d(h, sd[base + "conv.weight"], None, padding=pad)
        if init and int(sd[base + "actnorm.is_initialized"]) == 0:
            s, b = actnorm_stats(h)
            sd[base + "actnorm.scale"] = s
            sd[base + "actnorm.bias"] = b
            sd[base + "actnorm.is_initialized"] = torch.tensor(1, dtype=torch.uint8)
        h = torch.relu(actnorm_fwd(h, sd[base + "actnorm.scale"], sd[base + "actnorm.bias"])[0])
    return zeroconv(h, sd[pre + "4.weight"], sd[pre + "4.bias"], sd[pre + "4.logs"])


def coupling_fwd(x: Tensor, sd: Dict[str, Tensor], pre: str, init: bool = False) -> Tuple[Tensor, Tensor]:
    """Affine coupling forward (transforms.py:179-184): first half of the net output is the
    log-scale, second half the shift; scale = sigmoid(log_scale + 2); y_b = (x_b + t)*scale;
    per-sample log-det = sum log(scale + 1e-6)."""
    x_a, x_b = x.chunk(2, dim=1)
    log_s, t = coupling_net(x_a, sd, pre, init).chunk(2, dim=1)
    s = torch.sigmoid(log_s + 2.0)
    y = torch.cat([x_a, (x_b + t) * s], dim=1)
    return y, torch.log(s + 1e-6).reshape(x.shape[0], -1).sum(1)


def coupling_inv(y: Tensor, sd: Dict[str, Tensor], pre: str) -> Tensor:
    """Affine coupling inverse (transforms.py:196-200): x_b = y_b/(scale + 1e-6) - t."""
    y_a, y_b = y.chunk(2, dim=1)
    log_s, t = coupling_net(y_a, sd, pre).chunk(2, dim=1)
    s = torch.sigmoid(log_s + 2.0)
    return torch.cat([y_a, y_b / (s + 1e-6) - t], dim=1)


def gaussian_logp(x: Tensor, mean: Tensor, logs: Tensor) -> Tensor:
    """Diagonal Gaussian log-density summed over non-batch dims (normalizing_flow/prior.py:36-37)."""
    v = -0.5 * (LOG_2PI + 2.0 * logs + ((x - mean) ** 2.0) * torch.exp(-2.0 * logs))
    return v.reshape(x.shape[0], -1).sum(1)


def split_prior_params(y: Tensor, sd: Dict[str, Tensor], pre: str) -> Tuple[Tensor, Tensor]:
    """mean, logs = chunk(ZeroConv3x3(y)) or zeros when the prior is not learned
    (transforms.py:266-268)."""
    if (pre + "conv.weight") in sd:
        h = zeroconv(y, sd[pre + "conv.weight"], sd[pre + "conv.bias"], sd[pre + "conv.logs"])
    else:
        h = torch.zeros(y.shape[0], 2 * y.shape[1], *y.shape[2:])
    mean, logs = h.chunk(2, dim=1)
    return mean, logs


# --------------------------------------------------------------------------- step / model
def _step_prefixes(L: int, K: int) -> List[List[str]]:
    """Key prefixes per level in execution order (normalizing_flow/glow.py:163-170)."""
    lv = [[f"blocks.{i}.flows.{j}." for j in range(K)] for i in range(L - 1)]
    lv.append([f"final_flows.{j}." for j in range(K)])
    return lv


def step_fwd(x: Tensor, sd: Dict[str, Tensor], pre: str, ld: Tensor, init: bool = False) -> Tensor:
    """StepFlow forward: ActNorm -> InvConv2d -> AffineCoupling (glow.py:46-48).
    ``ld`` is mutated in place like the reference's ``log_det_jac +=``."""
    if init and int(sd[pre + "actnorm.is_initialized"]) == 0:
        s, b = actnorm_stats(x)
        sd[pre + "actnorm.scale"], sd[pre + "actnorm.bias"] = s, b
        sd[pre + "actnorm.is_initialized"] = torch.tensor(1, dtype=torch.uint8)
    y, d = actnorm_fwd(x, sd[pre + "actnorm.scale"], sd[pre + "actnorm.bias"])
    ld += d
    y, d = invconv_fwd(y, sd[pre + "invconv2d.weight"])
    ld += d
    y, d = coupling_fwd(y, sd, pre + "affcoupling.net.", init)
    ld += d
    return y


def step_inv(y: Tensor, sd: Dict[str, Tensor], pre: str) -> Tensor:
    """StepFlow inverse: coupling^-1 -> invconv^-1 -> actnorm^-1 (glow.py:60-62)."""
    x = coupling_inv(y, sd, pre + "affcoupling.net.")
    x = invconv_inv(x, sd[pre + "invconv2d.weight"])
    return actnorm_inv(x, sd[pre + "actnorm.scale"], sd[pre + "actnorm.bias"])


def glow_transform(sd: Dict[str, Tensor], x: Tensor, L: int, K: int, ld: Tensor,
                   logp: Optional[Tensor], init: bool = False) -> Tuple[List[Tensor], Tensor, Optional[Tensor]]:
    """Glow.transform (glow.py:172-201 with GlowBlock.transform :90-114 and Split.transform
    transforms.py:286-290).  Returns ([z_0..z_{L-1}], ld, logp); ld/logp mutated in place.
    ``logp=None`` disables the split priors (transforms.py:287).  ``init=True`` performs the
    data-dependent ActNorm initialisation on flags that are still 0 and writes the new
    parameters back into ``sd``."""
    levels = _step_prefixes(L, K)
    zs: List[Tensor] = []
    y = x
    for i in range(L - 1):
        y = squeeze2x2(y)
        for pre in levels[i]:
            y = step_fwd(y, sd, pre, ld, init)
        y, z = y.chunk(2, dim=1)
        if logp is not None:
            mean, logs = split_prior_params(y, sd, f"blocks.{i}.split.")
            logp += gaussian_logp(z, mean, logs)
        zs.append(z)
    y = squeeze2x2(y)
    for pre in levels[L - 1]:
        y = step_fwd(y, sd, pre, ld, init)
    zs.append(y)
    return zs, ld, logp


def glow_invert(sd: Dict[str, Tensor], latents: Sequence[Tensor], L: int, K: int,
                temperature: float = 1.0, eps: Optional[Sequence[Tensor]] = None) -> Tensor:
    """Glow.invert (glow.py:203-228, GlowBlock.invert :116-137, Split.invert
    transforms.py:292-309).  ``latents`` holds either all L parts or only the last one; in the
    latter case the split latents are drawn as mean + exp(logs)*temperature*eps
    (prior.py:49-50) with ``eps[i]`` the standard-normal draw for block i (zeros if None)."""
    levels = _step_prefixes(L, K)
    y = latents[-1]
    for pre in reversed(levels[L - 1]):
        y = step_inv(y, sd, pre)
    y = unsqueeze2x2(y)
    for n, i in enumerate(reversed(range(L - 1))):
        idx = -(n + 2)
        z = latents[idx] if len(latents) >= -idx else None   # utils.py:295-300 get_item
        if z is None:
            mean, logs = split_prior_params(y, sd, f"blocks.{i}.split.")
            e = eps[i] if eps is not None else torch.zeros_like(mean)
            z = mean + (torch.exp(logs) * temperature) * e
        y = torch.cat([y, z], dim=1)
        for pre in reversed(levels[i]):
            y = step_inv(y, sd, pre)
        y = unsqueeze2x2(y)
    return y


def gaussian_prior_params(psd: Dict[str, Tensor], shape: Sequence[int]) -> Tuple[Tensor, Tensor]:
    """GaussianPrior: ZeroConv2d(2C,2C) applied to an all-zero map, chunked into mean/logs
    (normalizing_flow/prior.py:79-81).  The conv of zeros is exactly its bias, so this is
    bias*exp(3*logs) per channel — restated literally with the conv to stay faithful."""
    b, c, h, w = shape
    z = torch.zeros(b, 2 * c, h, w)
    key = "_GaussianPrior__conv."
    if (key + "weight") in psd:
        z = zeroconv(z, psd[key + "weight"], psd[key + "bias"], psd[key + "logs"])
    mean, logs = z.chunk(2, dim=1)
    return mean, logs


def gaussian_prior_logp(psd: Dict[str, Tensor], z: Tensor) -> Tensor:
    """GaussianPrior.compute_log_prob (prior.py:70-83)."""
    mean, logs = gaussian_prior_params(psd, z.shape)
    return gaussian_logp(z, mean, logs)


def gaussian_prior_sample(psd: Dict[str, Tensor], shape: Sequence[int], temperature: float, eps: Tensor) -> Tensor:
    """GaussianPrior.sample (prior.py:85-99, :49-50) with the normal draw supplied."""
    mean, logs = gaussian_prior_params(psd, shape)
    return mean + (torch.exp(logs) * temperature) * eps


# --------------------------------------------------------------------------- step glue (a23)
def preprocess_batch(batch: Tensor, n_bits: int, n_bins: float) -> Tensor:
    """[0,1] images -> n_bits quantised, centred (normalizing_flow/utils.py:188-196)."""
    v = batch * 255
    if n_bits < 8:
        v = torch.floor(v / 2 ** (8 - n_bits))
    return v / n_bins - 0.5


def postprocess_batch(batch: Tensor, n_bins: float) -> Tensor:
    """Inverse of the above to uint8 (normalizing_flow/utils.py:210)."""
    return torch.clip(torch.floor((batch + 0.5) * n_bins) * (256.0 / n_bins), 0, 255).to(torch.uint8)


def bpd_loss(log_likelihood: Tensor, n_bins: float, n_pixel: float) -> Tensor:
    """Bits per dimension (normalizing_flow/utils.py:255-256)."""
    return ((np.log(n_bins) * n_pixel - log_likelihood) * (np.log2(np.e) / n_pixel)).mean(dim=0)

# --------------------------------------------------------------------------- data formats either side of the path (§8f)
def dequantize(batch: Tensor, n_bits: int, n_bins: float, noise: Tensor) -> Tensor:
    """preprocess_batch followed by the dequantisation-noise add of the training step
    (normalizing_flow/trainer.py:152-155: ``batch + torch.rand_like(batch) / n_bins``)."""
    return preprocess_batch(batch, n_bits, n_bins) + noise / n_bins


def cat_format(latents: Sequence[Tensor]) -> Tensor:
    """CatFormater.process_latents (diffusion_prior/latent_formaters.py:163-189): latent ``pos`` is squeezed
    ``target - pos`` times (or unsqueezed ``pos - target`` times), target = (L-1)//2; channel concat."""
    target = (len(latents) - 1) // 2
    parts = []
    for pos, t in enumerate(latents):
        deg = target - pos
        for _ in range(abs(deg)):
            t = squeeze2x2(t) if deg > 0 else unsqueeze2x2(t)
        parts.append(t)
    return torch.cat(parts, dim=1)


def cat_unformat(cat: Tensor, latent_dims: Sequence[Sequence[int]]) -> List[Tensor]:
    """CatFormater.postprocess (diffusion_prior/latent_formaters.py:191-236).  The reference walks outwards from the
    target part, un/squeezing the *remaining concatenation* once per hop and cutting one latent off its inner edge; for
    Glow's latent shapes (leading channels = 2(2^t - 1) x trailing channels) that equals cutting the concatenation at
    the parts' channel offsets and un/squeezing every part on its own, which is what is written here (checked against
    the reference's output in tests/golden/formats.npz)."""
    target = (len(latent_dims) - 1) // 2
    out, off = [], 0
    for pos, (C, H, W) in enumerate(latent_dims):
        deg = target - pos
        cnt = int(C) * 4 ** deg if deg >= 0 else int(C) // 4 ** (-deg)
        t = cat[:, off:off + cnt]
        off += cnt
        for _ in range(abs(deg)):
            t = unsqueeze2x2(t) if deg > 0 else squeeze2x2(t)
        out.append(t.contiguous())
    assert off == cat.shape[1]
    return out



def nll_bpd(sd: Dict[str, Tensor], psd: Dict[str, Tensor], x: Tensor, L: int, K: int,
            n_bins: float, n_pixel: float) -> Tensor:
    """The likelihood-evaluation recipe of normalizing_flow/trainer.py:154-161 /:46-52."""
    b = x.shape[0]
    ld = torch.zeros(b, dtype=torch.float64)
    lp = torch.zeros(b, dtype=torch.float64)
    zs, ld, lp = glow_transform(sd, x, L, K, ld, lp)
    lp += gaussian_prior_logp(psd, zs[-1])
    return bpd_loss(ld + lp, n_bins, n_pixel)


def train_grads(sd: Dict[str, Tensor], psd: Dict[str, Tensor], x: Tensor, L: int, K: int, n_bins: float,
                n_pixel: float) -> Tuple[Tensor, Dict[str, Tensor], Dict[str, Tensor]]:
    """Loss and parameter gradients of one training step as normalizing_flow/trainer.py:154-164 computes them
    (fp64 accumulators, transform, prior log-prob, bits-per-dim loss, backward).  Returns
    (loss, {flow key: grad}, {prior key: grad}); parameters without a gradient path are reported as zeros,
    the way torch leaves the all-zero-input GaussianPrior conv weight."""
    sd_g = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in sd.items()}
    psd_g = {k: v.clone().requires_grad_(True) for k, v in psd.items()}
    with torch.enable_grad():
        loss = nll_bpd(sd_g, psd_g, x, L, K, n_bins, n_pixel)
        loss.backward()
    g = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in sd_g.items() if v.dtype.is_floating_point}
    pg = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in psd_g.items()}
    return loss.detach(), g, pg


def grad_signature(g: Tensor, seed: int) -> Tuple[float, float]:
    """(L2 norm, projection on a PCG64 standard-normal vector) of a gradient tensor in fp64: a compact fingerprint
    that pins large gradients in the golden fixtures without storing them."""
    v = g.detach().double().reshape(-1)
    r = torch.from_numpy(np.random.default_rng(seed).standard_normal(v.numel()))
    return float(v.norm()), float((v * r).sum())


def output_shapes(L: int, in_channels: int, size: int) -> List[Tuple[int, int, int]]:
    """Latent shapes (normalizing_flow/utils.py:93-117)."""
    out = []
    for _ in range(L - 1):
        if size % 2 != 0:
            raise ValueError("The input dimension is not divisible by 2!")
        in_channels *= 2
        size //= 2
        out.append((in_channels, size, size))
    out.append((in_channels * 4, size // 2, size // 2))
    return out


# --------------------------------------------------------------------------- seeded weights / inputs
# (the generators live in /synthetic.py so that bench.py's product arm and tools/ get their synthetic data without
# importing the oracle; re-exported here for the tests and the golden generators)
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from synthetic import glow_param_shapes, seeded_input, seeded_state  # noqa: E402,F401


def state_checksum(sd: Dict[str, Tensor]) -> Tuple[float, float]:
    """(sum, abs-sum) in fp64 over all float tensors — used to prove regenerated weights
    equal the ones the golden outputs were made with."""
    s = a = 0.0
    for k in sorted(sd):
        t = sd[k]
        if t.dtype.is_floating_point:
            s += float(t.double().sum())
            a += float(t.double().abs().sum())
    return s, a
