"""Deterministic synthetic weights and inputs (numpy PCG64: the same bytes on any box) for bench.py, tools/, tests/
and the golden-vector generators.  No arithmetic of the flow lives here — only the reference's state_dict LAYOUT
(key, shape, dtype in order; verified against the live reference module by oracle/make_golden.py) and random draws with
the shapes and scales a Glow of that layout uses.  The CPU oracle re-exports these names, so fixtures made through
``oracle.glow_oracle.seeded_state`` and inputs made here are the same bytes."""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor


def _quantise(batch: Tensor, n_bits: int, n_bins: float) -> Tensor:
    """[0,1] images -> n_bits levels centred on 0, the op sequence of normalizing_flow/utils.py:188-196."""
    v = batch * 255
    if n_bits < 8:
        v = torch.floor(v / 2 ** (8 - n_bits))
    return v / n_bins - 0.5


def glow_param_shapes(in_channel: int, L: int, K: int, learn_prior: bool = True,
                      n_features: int = 512) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, dtype) in the reference's state_dict order (verified against the live
    reference module by oracle/make_golden.py)."""
    out: List[Tuple[str, Tuple[int, ...], str]] = []

    def step(pre: str, c: int):
        out.append((pre + "actnorm.scale", (c, 1, 1), "f"))
        out.append((pre + "actnorm.bias", (c, 1, 1), "f"))
        out.append((pre + "actnorm.is_initialized", (), "u8"))
        out.append((pre + "invconv2d.weight", (c, c, 1, 1), "f"))
        n = pre + "affcoupling.net."
        for idx, cin, k in (("0", c // 2, 3), ("2", n_features, 1)):
            out.append((f"{n}{idx}._Conv2dActNorm__conv.weight", (n_features, cin, k, k), "f"))
            out.append((f"{n}{idx}._Conv2dActNorm__actnorm.scale", (n_features, 1, 1), "f"))
            out.append((f"{n}{idx}._Conv2dActNorm__actnorm.bias", (n_features, 1, 1), "f"))
            out.append((f"{n}{idx}._Conv2dActNorm__actnorm.is_initialized", (), "u8"))
        out.append((n + "4.weight", (c, n_features, 3, 3), "f"))
        out.append((n + "4.bias", (c,), "f"))
        out.append((n + "4.logs", (1, c, 1, 1), "f"))

    for i in range(L - 1):
        c = 4 * (2 ** i) * in_channel
        for j in range(K):
            step(f"blocks.{i}.flows.{j}.", c)
        if learn_prior:
            out.append((f"blocks.{i}.split.conv.weight", (c, c // 2, 3, 3), "f"))
            out.append((f"blocks.{i}.split.conv.bias", (c,), "f"))
            out.append((f"blocks.{i}.split.conv.logs", (1, c, 1, 1), "f"))
    c = 2 ** (L + 1) * in_channel
    for j in range(K):
        step(f"final_flows.{j}.", c)
    return out


def seeded_state(in_channel: int, L: int, K: int, seed: int, learn_prior: bool = True,
                 initialized: bool = True, zero_sigma: float = 0.02) -> Tuple[Dict[str, Tensor], Dict[str, Tensor]]:
    """Deterministic, *non-degenerate* weights for parity work (numpy PCG64, so the same
    bytes are regenerated on any box).  ZeroConv2d tensors get N(0, zero_sigma) noise because
    at the reference's zero init the coupling nets output exactly 0 and conv precision would
    be untestable (SURVEY.md §7 hard part 1).  Returns (flow state_dict, GaussianPrior state_dict)."""
    rng = np.random.default_rng(seed)
    sd: Dict[str, Tensor] = {}
    for key, shape, kind in glow_param_shapes(in_channel, L, K, learn_prior):
        if kind == "u8":
            sd[key] = torch.tensor(1 if initialized else 0, dtype=torch.uint8)
            continue
        if key.endswith("invconv2d.weight"):
            c = shape[0]
            q, _ = np.linalg.qr(rng.standard_normal((c, c)))
            w = q + 0.05 * rng.standard_normal((c, c))          # well conditioned, |det| != 1
            arr = w.reshape(shape)
        elif key.endswith("actnorm.scale"):
            arr = 0.1 * rng.standard_normal(shape)
        elif key.endswith("actnorm.bias"):
            arr = 0.1 * rng.standard_normal(shape)
        elif key.endswith("_Conv2dActNorm__conv.weight"):
            fan_in = shape[1] * shape[2] * shape[3]
            arr = rng.standard_normal(shape) / math.sqrt(fan_in)
        elif key.endswith(".logs"):
            arr = 0.1 * rng.standard_normal(shape)
        elif key.endswith(".bias"):
            arr = zero_sigma * rng.standard_normal(shape)
        else:  # ZeroConv weights (coupling net.4 / split.conv)
            fan_in = shape[1] * 9
            arr = zero_sigma * rng.standard_normal(shape) * (16.0 / math.sqrt(fan_in))
        sd[key] = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32))
    cz = 2 ** (L + 1) * in_channel
    psd: Dict[str, Tensor] = {}
    if learn_prior:
        psd["_GaussianPrior__conv.weight"] = torch.from_numpy(
            (0.01 * rng.standard_normal((2 * cz, 2 * cz, 3, 3))).astype(np.float32))
        psd["_GaussianPrior__conv.bias"] = torch.from_numpy((0.1 * rng.standard_normal((2 * cz,))).astype(np.float32))
        psd["_GaussianPrior__conv.logs"] = torch.from_numpy(
            (0.1 * rng.standard_normal((1, 2 * cz, 1, 1))).astype(np.float32))
    return sd, psd


def seeded_input(shape: Sequence[int], seed: int, n_bits: int = 5) -> Tensor:
    """Synthetic dequantised images in [-0.5, 0.5) as the trainer feeds them
    (normalizing_flow/trainer.py:152-155), from numpy PCG64."""
    rng = np.random.default_rng(seed)
    n_bins = 2.0 ** n_bits
    img = torch.from_numpy(rng.random(tuple(shape), dtype=np.float32))
    noise = torch.from_numpy(rng.random(tuple(shape), dtype=np.float32))
    return _quantise(img, n_bits, n_bins) + noise / n_bins
