"""Host logic of the data-parallel path on CPU: world_size-2 ``gloo`` processes exercise the bucketed gradient
all-reduce (StepFlow-group slices of the flat gradient buffer, in backward order), the parameter/buffer broadcast and the
batch sharding.  No kernels run here; the NCCL path is the same code with backend "nccl" (bench.py --gpus N)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import normalizing_flow as nf
        from normalizing_flow._train import GradSink
        torch.manual_seed(100 + rank)                      # replicas start DIFFERENT on purpose
        flow = nf.Glow(1, 3, 2)
        prior = nf.GaussianPrior(16)
        dp = nf.GradAllReduce(flow, prior)
        # --- broadcast: every rank ends with rank 0's parameters and flags
        if rank == 0:
            for m in flow.modules():
                if hasattr(m, "is_initialized"):
                    m.is_initialized.fill_(1)
        v0 = [p._version for p in flow.parameters()]
        flow._pver = list(v0)                              # as if a forward had already keyed its caches on these versions
        dp.broadcast_parameters(src=0)
        assert all(p._version > v for p, v in zip(flow.parameters(), v0)), "broadcast must bump the version counters"
        assert flow._pver is None, "broadcast must invalidate the parameter-derived caches"
        chk = torch.stack([p.detach().double().sum() for p in flow.parameters()]).sum()
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        assert all(float(c) == float(allc[0]) for c in allc), "parameters differ after broadcast"
        assert all(int(m.is_initialized) == 1 for m in flow.modules() if hasattr(m, "is_initialized"))
        # --- bucketed gradient average in the order the backward reports the levels
        params = list(flow.parameters())
        sink = GradSink(params)
        for i, p in enumerate(params):
            sink.get(p).fill_(float(rank + 1) * (i + 1))
        ranges = sink.level_ranges(flow)
        assert len(ranges) == 3 and ranges[0][0] == 0 and ranges[-1][1] == sink.numel
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(2)), "levels must tile the flat buffer"
        # the buckets the backward reports: per level (deepest first) groups of StepFlows, last group first; the first
        # bucket of a level with a Split also holds the Split prior's parameters
        levels = [(blk.flows, blk.split) for blk in flow.blocks] + [(flow.final_flows, None)]
        buckets = []
        for flows, split in reversed(levels):
            last = list((split if split is not None else flows[-1]).parameters())[-1]
            buckets.append(sink.span(next(flows[1].parameters()), last))          # K = 2, one StepFlow per bucket here
            buckets.append(sink.span(next(flows[0].parameters()), list(flows[0].parameters())[-1]))
        assert sorted(buckets) == sorted(set(buckets)) and sum(h - l for l, h in buckets) == sink.numel
        dp.begin(flow, sink)
        for lo, hi in buckets:
            dp.bucket_done(lo, hi)
        for p, v in zip(prior.parameters(), (1.0, 2.0, 3.0)):
            p.grad = torch.full_like(p, v * (rank + 1))
        dp.finish()
        mean_rank = sum(r + 1 for r in range(world)) / world
        for i, p in enumerate(params):
            assert torch.allclose(sink.get(p), torch.full_like(p, mean_rank * (i + 1))), i
        for p, v in zip(prior.parameters(), (1.0, 2.0, 3.0)):
            assert torch.allclose(p.grad, torch.full_like(p, v * mean_rank))
        # --- an incomplete backward is an error, not a silent partial average
        dp.begin(flow, sink)
        dp.bucket_done(*buckets[0])
        try:
            dp.finish()
            raise AssertionError("finish() accepted an incomplete all-reduce")
        except RuntimeError:
            pass
        # --- data-dependent ActNorm initialisation over the GLOBAL batch: every rank initialises from its shard
        # (reference formula, transforms.py:74-78), combine_init_stats merges the shards' statistics
        g = torch.Generator().manual_seed(7)
        xg = torch.randn(8, 5, 6, 6, generator=g) * torch.tensor([0.5, 1.0, 2.0, 3.0, 10.0]).view(1, 5, 1, 1) + 3.0
        xs = xg[nf.shard(8, rank, world)]
        scale = -torch.log(xs.std(dim=(0, 2, 3)) + 1e-6).view(5, 1, 1).clone()
        bias = -xs.mean(dim=(0, 2, 3)).view(5, 1, 1).clone()
        nf.combine_init_stats(scale, bias, xs.shape[0] * 36)
        want_s = -torch.log(xg.double().std(dim=(0, 2, 3)) + 1e-6)
        want_b = -xg.double().mean(dim=(0, 2, 3))
        assert torch.allclose(scale.double().view(-1), want_s, rtol=0, atol=2e-6), (scale.view(-1), want_s)
        assert torch.allclose(bias.double().view(-1), want_b, rtol=0, atol=2e-6)
        both = [torch.zeros(10) for _ in range(world)]
        dist.all_gather(both, torch.cat([scale.view(-1), bias.view(-1)]))
        assert all(torch.equal(b, both[0]) for b in both), "replicas must end bit-identical"
        # the context manager installs / removes the hook the init kernels' wrapper calls
        from normalizing_flow import _engine as E
        with dp.global_initialization():
            assert E.stats_hook is not None
        assert E.stats_hook is None
        # --- sharding
        sl = nf.shard(8, rank, world)
        assert (sl.start, sl.stop) == (rank * 4, rank * 4 + 4)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))


def test_gradient_allreduce_and_broadcast_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}:\n{msg}"


def test_shard_rejects_ragged_batches():
    import normalizing_flow as nf
    with pytest.raises(ValueError):
        nf.shard(10, 0, 4)
