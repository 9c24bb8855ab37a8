"""Data-parallel training of the Glow flow on the GPUs of one box: one process per GPU, full parameter
replica, the batch sharded by images.  Forward, likelihood evaluation, inverse and sampling need no communication.
Training has ONE exchange step per iteration: the gradient all-reduce (NCCL over NVLink/NVSwitch, fp32, average —
the loss is a batch mean, normalizing_flow/utils.py:256, so equal shards + AVG reproduce the single-GPU gradient),
after which every rank applies the identical clip + Adam update (trainer.py:165-167).

The backward kernels write all gradients of a step into one flat buffer in ``parameters()`` order
(_train.GradSink); levels finish deepest-first, so each level's slice is all-reduced on a communication stream as
soon as its last gradient kernel has been enqueued, overlapping with the backward of the shallower levels.

The reference has no distributed code; a trainer adopts this with three lines:

    dp = GradAllReduce(flow, prior)          # after torch.distributed.init_process_group("nccl")
    dp.broadcast_parameters()                # once, after the data-dependent initialisation on rank 0's first batch
    loss.backward(); dp.finish()             # every step, before clip_grad_* / optimizer.step()

Data-dependent initialisation over the GLOBAL first batch (what a single process with the whole batch computes,
normalizing_flow/utils.py:275-292) instead of rank 0's shard:

    dp.broadcast_parameters()                # equal random init on every rank
    with dp.global_initialization():         # every rank runs the init pass on ITS shard of the first batch;
        nf.data_dependent_nf_initialization(flow, local_loader, device, n_bits, n_bins)   # statistics are combined
"""
from __future__ import annotations

import contextlib
import os
from typing import List, Optional

import torch
import torch.distributed as dist


def shard(n: int, rank: int, world: int) -> slice:
    """Contiguous, equal shard of a global batch of n images (n must be divisible by world: the gradient average
    over ranks equals the global batch mean only for equal shards)."""
    if n % world != 0:
        raise ValueError(f"global batch {n} is not divisible by the world size {world}")
    per = n // world
    return slice(rank * per, (rank + 1) * per)


@torch.no_grad()
def combine_init_stats(scale: torch.Tensor, bias: torch.Tensor, n_local: int, group=None, eps: float = 1e-6) -> None:
    """Replace a freshly initialised ActNorm's ``scale = -log(std + eps)``, ``bias = -mean`` (statistics of THIS rank's
    ``n_local`` samples per channel, reference transforms.py:74-78) by the values of the union of all ranks' samples,
    in place.  The per-rank mean and unbiased variance are recovered from the parameters, gathered (one small
    all-gather of 2C doubles) and merged with the pairwise-variance identity

        M2 = sum_r [(n-1) var_r + n (mean_r - mean)^2],   var = M2 / (world n - 1)

    in float64 on every rank identically, so the replicas stay bit-equal without a broadcast.  Shards must be equal
    (``shard`` enforces it)."""
    world = dist.get_world_size(group)
    if world == 1:
        return
    mean = -bias.detach().double().reshape(-1)
    std = (torch.exp(-scale.detach().double().reshape(-1)) - eps).clamp_min_(0.0)
    mine = torch.stack([mean, std * std])
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    allp = torch.stack(parts)                                   # [world, 2, C]
    g_mean = allp[:, 0].mean(dim=0)
    n = float(n_local)
    m2 = ((n - 1.0) * allp[:, 1] + n * (allp[:, 0] - g_mean) ** 2).sum(dim=0)
    g_std = torch.sqrt(m2 / (world * n - 1.0))
    scale.copy_((-torch.log(g_std + eps)).to(scale.dtype).view_as(scale))
    bias.copy_((-g_mean).to(bias.dtype).view_as(bias))


class GradAllReduce:
    def __init__(self, flow, prior: Optional[torch.nn.Module] = None, group=None):
        if not dist.is_initialized():
            raise RuntimeError("GradAllReduce needs torch.distributed.init_process_group(...) first")
        max_ctas = int(os.environ.get("NFDPM_NCCL_MAX_CTAS", "16"))          # 0: NCCL's default communicator
        if group is None and dist.get_backend() == "nccl" and max_ctas > 0:
            # The gradient all-reduces overlap the backward: every SM an NCCL CTA occupies is one the persistent GEMM grids
            # (one CTA per SM, statically partitioned tiles) have to queue behind.  A communicator with few CTAs trades
            # collective bandwidth (irrelevant: the buckets are overlapped) for less interference.  Measured at N = 2, config-2
            # train step (tools/dp_cta_sweep.sh, profiles/r02_dp_cta_sweep.txt): exposed communication 211 us with NCCL's
            # default, 151 / 126 / 204 / 302 us with at most 16 / 8 / 4 / 2 CTAs; at N = 8: 375 us default, 342 / 276 / 451 us with
            # 32 / 16 / 8.  16 is the default.
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = max_ctas
            opts.config.min_ctas = 1
            group = dist.new_group(backend="nccl", pg_options=opts)
        self.flow, self.prior, self.group = flow, prior, group
        self.world = dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self.cuda = next(flow.parameters()).is_cuda
        self.stream = torch.cuda.Stream() if self.cuda else None
        self.sink = None
        self.covered = 0
        self.launched = 0
        self.last_bucket_bytes = 0
        self.enabled = True            # False: skip the collectives (bench.py: the step without communication, to report
        flow._grad_hook = self         # the exposed communication time as the difference)
        # The GaussianPrior's (few, small) gradients are complete at the very START of the backward (the prior term is the
        # last thing the forward computes): average them on the communication stream as soon as autograd has accumulated
        # them, instead of after the backward, where every launch of that small chain would be exposed.
        self._prior_params = [p for p in prior.parameters() if p.requires_grad] if prior is not None else []
        self._prior_ready = 0
        self._prior_done = False
        for p in self._prior_params:
            p.register_post_accumulate_grad_hook(self._prior_grad_ready)

    def detach(self) -> None:
        if getattr(self.flow, "_grad_hook", None) is self:
            self.flow._grad_hook = None

    # ---- collectives
    def _avg(self, t: torch.Tensor) -> None:
        if self.backend == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
        else:                                   # gloo (CPU tests) has no AVG
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.div_(self.world)

    def broadcast_parameters(self, src: int = 0) -> None:
        """Parameters and buffers (incl. the uint8 ``is_initialized`` flags) of rank ``src`` to every rank: run once
        after the data-dependent ActNorm initialisation (normalizing_flow/utils.py:275-292) so replicas start equal."""
        mods = [self.flow] + ([self.prior] if self.prior is not None else [])
        with torch.no_grad():
            for m in mods:
                # the parameters THEMSELVES (not ``p.data``): an in-place copy under no_grad bumps the version counters that
                # key every parameter-derived cache (LU/fold matrices, packed weights, captured chains); ``p.data.copy_``
                # would leave a rank that already ran a forward on its stale caches
                ts = list(m.parameters()) + list(m.buffers())
                for dt in sorted({t.dtype for t in ts}, key=str):
                    grp = [t for t in ts if t.dtype == dt]
                    flat = torch.cat([t.reshape(-1) for t in grp])
                    dist.broadcast(flat, src=src, group=self.group)
                    off = 0
                    for t in grp:
                        t.copy_(flat[off:off + t.numel()].view_as(t))
                        off += t.numel()
        for m in mods:                                 # host-side caches of the flags / prepared matrices
            for sub in m.modules():
                if hasattr(sub, "_init_known"):
                    sub._init_known = None
            if hasattr(m, "invalidate_caches"):
                m.invalidate_caches()

    @contextlib.contextmanager
    def global_initialization(self):
        """Inside this context every data-dependent ActNorm initialisation (144 of them in an L3/K16 Glow, each depending
        on the layers before it) uses the statistics of all ranks' shards: see ``combine_init_stats``."""
        from . import _engine as E
        prev = E.stats_hook
        E.stats_hook = lambda scale, bias, n: combine_init_stats(scale, bias, n, self.group)
        try:
            yield self
        finally:
            E.stats_hook = prev

    def _avg_prior(self) -> None:
        gs = [p.grad for p in self._prior_params if p.grad is not None]
        if gs:
            flat = torch.cat([g.reshape(-1) for g in gs])
            self._avg(flat)
            off = 0
            for g in gs:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()

    def _prior_grad_ready(self, p) -> None:
        self._prior_ready += 1
        if self._prior_ready < len(self._prior_params) or not self.enabled:
            return
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                self._avg_prior()
            for q in self._prior_params:
                q.grad.record_stream(self.stream)
        else:
            self._avg_prior()
        self._prior_done = True

    # ---- hooks called by _train.GlowTransformFn.backward
    def begin(self, glow, sink) -> None:
        self.sink = sink
        self.covered = 0
        self.launched = 0
        if self.cuda:
            sink.flat.record_stream(self.stream)

    def bucket_done(self, lo: int, hi: int, events=()) -> None:
        """Average the flat-gradient range [lo, hi) over the ranks on the communication stream, once the current stream
        and the streams that recorded ``events`` (the weight-gradient side stream) have produced it.  Buckets arrive in
        backward order — a few StepFlows each (_train.BUCKET_STEPS) — so only the last, small one is not overlapped by
        the rest of the backward."""
        self.covered += hi - lo
        self.launched += 1
        if not self.enabled:
            return
        buf = self.sink.flat[lo:hi]
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream())
            for ev in events:
                self.stream.wait_event(ev)
            with torch.cuda.stream(self.stream):
                self._avg(buf)
        else:
            self._avg(buf)
        self.last_bucket_bytes = 4 * (hi - lo)

    def describe(self) -> str:
        from . import _train as T
        return (f"NCCL AVG on buckets of up to {T.BUCKET_STEPS} StepFlows of the flat gradient buffer ({self.launched} per step, "
                f"last one {self.last_bucket_bytes / 1e6:.1f} MB), launched from inside the backward on a communication "
                f"stream as soon as a bucket's last gradient kernel is enqueued")

    def finish(self) -> None:
        """Join the communication stream and average the (few, small) GaussianPrior gradients.  Call after
        ``loss.backward()`` and before clipping / the optimizer step."""
        if self.cuda and self.enabled:
            torch.cuda.current_stream().wait_stream(self.stream)
        if self.sink is not None and self.covered != self.sink.numel:
            raise RuntimeError(f"gradient all-reduce incomplete: the backward reported {self.covered} of "
                               f"{self.sink.numel} gradient elements")
        self.sink = None
        if self.prior is not None and self.enabled and not self._prior_done:
            self._avg_prior()                   # (gradients that did not come through autograd's accumulation)
        self._prior_ready = 0
        self._prior_done = False
