"""Latent formaters of the diffusion prior (reference diffusion_prior/latent_formaters.py:13-260): the data format
between Glow's latent list and the tensor(s) the diffusion model is trained on / samples.

Same classes, constructors, methods and return values as the reference.  ``CatFormater`` is the one with arithmetic: it
squeezes the shallow latents and unsqueezes the deep ones to the resolution of the middle latent and concatenates along
channels (``process_latents``, :163-189); ``postprocess`` (:191-236) is the inverse.  Here each direction is ONE launch
of ``nfdpm_latent_format`` that moves every part (no einops chain, no ``torch.cat``); under autograd the backward of
either direction is the other direction's launch.  CUDA float32 tensors only — there is no host path.
"""
from __future__ import annotations

import copy
from abc import ABC, abstractmethod
from typing import List

import numpy as np
import torch
import torch.nn as nn

from normalizing_flow import calculate_output_shapes
from normalizing_flow import _engine as E
from normalizing_flow import _native as N


class BaseFormater(nn.Module, ABC):
    """Base class (reference :13-87).  The min/max standardisation of the reference is commented out there, so
    ``standardize_latents`` / ``inv_standardize_latents`` return their argument."""

    def __init__(self, L: int, in_channels: int, size: int):
        super().__init__()
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.latent_dims = np.array(calculate_output_shapes(L=L, in_channels=in_channels, size=size))
        self.mins, self.maxs = None, None

    @abstractmethod
    def process_latents(self, latents: list) -> list:
        ...

    @abstractmethod
    def postprocess(self, latents: list) -> list:
        ...

    def get_num_latent_parts(self) -> int:
        return len(self.latent_dims)

    def standardize_latents(self, latents: list) -> list:
        return latents

    def inv_standardize_latents(self, latents: list) -> list:
        return latents


class IdentityFormater(BaseFormater):
    """Passes the L latents through (reference :90-136)."""

    def __init__(self, L: int, in_channels: int, size: int):
        super().__init__(L, in_channels, size)
        self.postprocessed_latent_shapes = self.latent_dims

    def process_latents(self, latents: list) -> list:
        assert len(latents) == len(self.latent_dims), "IdentityFormater expects L latent tensors from Diffusion prior."
        return self.standardize_latents(latents)

    def postprocess(self, latents: list) -> list:
        return self.inv_standardize_latents(latents)

    def get_input_shapes(self) -> list:
        return self.postprocessed_latent_shapes


def _plan(dims) -> tuple:
    """[(C,H,W)] of the L latents -> (target index, [(degree, channel offset, channel count)], Ct, Ht, Wt)."""
    n = len(dims)
    t = (n - 1) // 2
    Ht, Wt = int(dims[t][1]), int(dims[t][2])
    parts, off = [], 0
    for i, (C, H, W) in enumerate(dims):
        d = t - i                                   # > 0: squeeze d times, < 0: unsqueeze -d times (reference :174-182)
        C, H, W = int(C), int(H), int(W)
        if d >= 0:
            ok, cnt = (H == Ht << d and W == Wt << d), C << (2 * d)
        else:
            ok, cnt = (H << -d == Ht and W << -d == Wt and C % (4 ** -d) == 0), C >> (2 * -d)
        if not ok:
            raise ValueError(f"latent {i} of shape {(C, H, W)} cannot be brought to {Ht}x{Wt} by {abs(d)} "
                             f"{'squeeze' if d > 0 else 'unsqueeze'} steps")
        parts.append((d, off, cnt))
        off += cnt
    return t, parts, off, Ht, Wt


class _ToCat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, *latents):
        _, parts, Ct, Ht, Wt = plan
        B = latents[0].shape[0]
        cat = torch.empty(B, Ct, Ht, Wt, dtype=torch.float32, device=latents[0].device)
        N.latent_format([(t, d, o, c) for t, (d, o, c) in zip(latents, parts)], cat, B, Ct, Ht, Wt, True)
        ctx.plan, ctx.shapes = plan, [t.shape for t in latents]
        return cat

    @staticmethod
    def backward(ctx, g):
        _, parts, Ct, Ht, Wt = ctx.plan
        g = g.contiguous()
        outs = [torch.empty(s, dtype=torch.float32, device=g.device) for s in ctx.shapes]
        N.latent_format([(t, d, o, c) for t, (d, o, c) in zip(outs, parts)], g, g.shape[0], Ct, Ht, Wt, False)
        return (None, *outs)


class _FromCat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, shapes, cat):
        _, parts, Ct, Ht, Wt = plan
        B = cat.shape[0]
        outs = [torch.empty((B,) + tuple(int(v) for v in s), dtype=torch.float32, device=cat.device) for s in shapes]
        N.latent_format([(t, d, o, c) for t, (d, o, c) in zip(outs, parts)], cat, B, Ct, Ht, Wt, False)
        ctx.plan = plan
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        _, parts, Ct, Ht, Wt = ctx.plan
        gs = [g.contiguous() for g in gs]
        B = gs[0].shape[0]
        cat = torch.empty(B, Ct, Ht, Wt, dtype=torch.float32, device=gs[0].device)
        N.latent_format([(t, d, o, c) for t, (d, o, c) in zip(gs, parts)], cat, B, Ct, Ht, Wt, True)
        return None, None, cat


class CatFormater(BaseFormater):
    """Concatenating formater (reference :139-248)."""

    def __init__(self, L: int, in_channels: int, size: int):
        super().__init__(L, in_channels, size)
        # the reference advertises the middle latent's shape with doubled channels (:154-156) — kept as is, although
        # process_latents returns sum(parts) channels
        dim = copy.deepcopy(self.latent_dims[(len(self.latent_dims) - 1) // 2])
        dim[0] *= 2
        self.postprocessed_latent_shapes = [dim]
        self._plan = _plan(self.latent_dims)

    def process_latents(self, latents: list) -> list:
        if len(latents) != len(self.latent_dims):
            raise ValueError(f"CatFormater built for {len(self.latent_dims)} latent parts got {len(latents)}")
        lat = [E.check_input(t, f"latents[{i}]") for i, t in enumerate(latents)]
        for i, (t, d) in enumerate(zip(lat, self.latent_dims)):
            if tuple(t.shape[1:]) != tuple(int(v) for v in d) or t.shape[0] != lat[0].shape[0]:
                raise ValueError(f"latents[{i}] has shape {tuple(t.shape)}, expected (B, {', '.join(str(int(v)) for v in d)})")
        return self.standardize_latents([_ToCat.apply(self._plan, *lat)])

    def postprocess(self, latents: list) -> list:
        assert len(latents) == 1, "CatFormater expects a single latent tensor from Diffusion prior."
        cat = E.check_input(self.inv_standardize_latents(latents)[0], "latents[0]")
        _, _, Ct, Ht, Wt = self._plan
        if tuple(cat.shape[1:]) != (Ct, Ht, Wt):
            raise ValueError(f"latents[0] has shape {tuple(cat.shape)}, expected (B, {Ct}, {Ht}, {Wt})")
        return list(_FromCat.apply(self._plan, [tuple(d) for d in self.latent_dims], cat))

    def get_num_latent_parts(self) -> int:
        return 1

    def get_input_shapes(self) -> list:
        return self.postprocessed_latent_shapes


def get_formater(name: str):
    """'IdentityFormater' | 'CatFormater' -> class; anything else raises ValueError("Invalid formater name") like the
    reference (latent_formaters.py:251-263)."""
    if name == "IdentityFormater":
        return IdentityFormater
    elif name == "CatFormater":
        return CatFormater
    raise ValueError("Invalid formater name")
