"""Fused optimiser step for the training recipe of the reference (normalizing_flow/trainer.py:165-167):

    torch.nn.utils.clip_grad_value_(flow.parameters(), 1)
    torch.nn.utils.clip_grad_norm_(flow.parameters(), 1)
    optimizer.step()                                   # Adam / AdamW from init_optimizer (utils.py:120-137)

With 729 parameter tensors those three lines are ~1300 small launches (torch's foreach norm has no fast path for the
tensor-valued clip coefficient).  ``FusedClipAdam.step()`` performs the same arithmetic in three launches of
libnfdpm_b200 (csrc/optimizer.cu) over a device table of tensor references; gradients of the clip group are clamped and
scaled IN PLACE like the two torch utilities do, so what a caller observes in ``p.grad`` afterwards is unchanged.

A trainer adopts it by replacing the three lines with ``optimizer.step()`` and constructing
``FusedClipAdam(params, lr=..., clip_params=flow.parameters())``.  State is exposed per parameter under torch.optim.Adam's
keys (``step``, ``exp_avg``, ``exp_avg_sq``), so ``state_dict()`` interchanges with a torch Adam checkpoint
(prior.py:102-115 writes ``optimizer.state_dict()``).
"""
from __future__ import annotations

from typing import Iterable, Optional

import numpy as np
import torch

from . import _native as N


class FusedClipAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 decoupled_weight_decay: bool = False, clip_params: Optional[Iterable[torch.Tensor]] = None,
                 clip_value: float = 1.0, max_norm: float = 1.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                        decoupled_weight_decay=decoupled_weight_decay)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FusedClipAdam supports a single parameter group")
        self._params = [p for p in self.param_groups[0]["params"]]
        for p in self._params:
            N.require_cuda(p, "parameter")
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise TypeError("FusedClipAdam needs contiguous float32 parameters")
        clip_ids = {id(p) for p in (clip_params if clip_params is not None else self._params)}
        self._clip = [1 if id(p) in clip_ids else 0 for p in self._params]
        self.clip_value, self.max_norm = float(clip_value), float(max_norm)
        dev = self._params[0].device
        chunk = N.opt_chunk()
        offs, chunks, off = [], [], 0
        for i, p in enumerate(self._params):
            offs.append(off)
            for c in range(0, p.numel(), chunk):
                chunks.append((i, c))
            off += (p.numel() + 63) // 64 * 64
        self._offs = offs
        self._n_chunks = len(chunks)
        self._chunks = torch.tensor(np.asarray(chunks, dtype=np.int32).reshape(-1), device=dev)
        self._exp_avg = torch.zeros(off, dtype=torch.float32, device=dev)
        self._exp_avg_sq = torch.zeros(off, dtype=torch.float32, device=dev)
        self._partial = torch.empty(self._n_chunks, dtype=torch.float32, device=dev)
        self._scal = torch.zeros(3, dtype=torch.float32, device=dev)     # clip coefficient, grad norm, step
        self._refs_host = torch.empty(len(self._params), 4, dtype=torch.int64).pin_memory()
        self._refs_dev = torch.empty(len(self._params), 4, dtype=torch.int64, device=dev)
        self._ref_key = None
        for i, p in enumerate(self._params):
            n = p.numel()
            self.state[p] = {"step": self._scal[2], "exp_avg": self._exp_avg[offs[i]:offs[i] + n].view_as(p),
                             "exp_avg_sq": self._exp_avg_sq[offs[i]:offs[i] + n].view_as(p)}

    # last clip coefficient / total gradient norm of the clip group (device tensors, no sync)
    @property
    def grad_norm(self) -> torch.Tensor:
        return self._scal[1]

    @property
    def clip_coef(self) -> torch.Tensor:
        return self._scal[0]

    def _upload_refs(self) -> None:
        key = tuple(p.grad.data_ptr() if p.grad is not None else 0 for p in self._params)
        if key == self._ref_key and not torch.cuda.is_current_stream_capturing():
            return
        h = self._refs_host.numpy()
        for i, p in enumerate(self._params):
            g = p.grad
            if g is not None and (g.dtype != torch.float32 or not g.is_contiguous() or not g.is_cuda):
                raise TypeError("FusedClipAdam needs contiguous float32 CUDA gradients")
            h[i, 0] = p.data_ptr()
            h[i, 1] = key[i]
            h[i, 2] = self._offs[i]
            h[i, 3] = p.numel() | (self._clip[i] << 32)
        self._refs_dev.copy_(self._refs_host, non_blocking=True)
        self._ref_key = key

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        g = self.param_groups[0]
        self._upload_refs()
        N.fused_clip_adam(self._refs_dev, self._chunks, self._n_chunks, self._exp_avg, self._exp_avg_sq, self._partial,
                          self._scal, self.clip_value, self.max_norm, float(g["lr"]), float(g["betas"][0]),
                          float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]),
                          bool(g["decoupled_weight_decay"]))
        # the parameters were updated behind torch's back: bump their version counters (the packed-weight / LU caches
        # of the flow are keyed on them)
        torch.autograd.graph.increment_version(self._params)
        return loss

    def load_state_dict(self, state_dict):
        """Accepts a torch.optim.Adam / FusedClipAdam state_dict; moments are copied into the flat buffers."""
        st = state_dict["state"]
        ids = state_dict["param_groups"][0]["params"]
        for i, (pid, p) in enumerate(zip(ids, self._params)):
            s = st.get(pid, st.get(str(pid)))
            if s is None:
                continue
            self.state[p]["exp_avg"].copy_(s["exp_avg"])
            self.state[p]["exp_avg_sq"].copy_(s["exp_avg_sq"])
            self._scal[2] = float(s["step"])
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in state_dict["param_groups"][0]:
                self.param_groups[0][k] = state_dict["param_groups"][0][k]
