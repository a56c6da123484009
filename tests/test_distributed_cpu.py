"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: weight broadcast, the flat-bucket gradient
all-reduce protocol and its equivalence to a single-process full-batch gradient (checked with the oracle)."""

import os
import socket

import torch
import torch.multiprocessing as mp

from constraints_as_terminations_b200 import dist as cdist
from oracle import ppo_oracle

OBS, ACT = 45, 12


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _flat(tensors):
    return torch.cat([t.reshape(-1) for t in tensors])


def _params(agent):
    return list(agent.critic.parameters()) + list(agent.actor_mean.parameters()) + [agent.actor_logstd]


def _batch(seed, n):
    g = torch.Generator().manual_seed(seed)
    obs, act = torch.randn(n, OBS, generator=g), torch.randn(n, ACT, generator=g)
    logp, adv = torch.randn(n, generator=g) * 0.1 - 15, torch.randn(n, generator=g)
    ret, val = torch.randn(n, generator=g), torch.randn(n, generator=g)
    return obs, act, logp, adv, ret, val


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    r, w, _ = cdist.init_from_env("gloo")
    assert (r, w) == (rank, world) and cdist.world_size() == world and cdist.rank() == rank
    assert cdist.shard_seed(42) == 42 + rank
    # 1. rank 0 seeds the flat parameter vector of everybody
    torch.manual_seed(100 + rank)  # different init per rank on purpose
    agent = ppo_oracle.AgentOracle(OBS, ACT)
    flat = _flat([p.detach() for p in _params(agent)])
    cdist.broadcast_params(flat, src=0)
    off = 0
    with torch.no_grad():
        for p in _params(agent):
            p.copy_(flat[off : off + p.numel()].view_as(p))
            off += p.numel()
    # 2. per-rank shard gradient (mean over the local minibatch), sum-allreduce, scale by 1/world
    shard = _batch(7, 256)
    lo, hi = rank * 128, (rank + 1) * 128
    value_rms = {"mean": torch.tensor(0.1), "var": torch.tensor(1.5)}
    loss, _ = ppo_oracle.ppo_minibatch_loss(agent, value_rms, *(t[lo:hi] for t in shard), norm_adv=False)
    loss.backward()
    bucket = _flat([p.grad for p in _params(agent)])
    scale = cdist.allreduce_grads(bucket)
    out[rank] = (flat.clone(), bucket * scale)
    torch.distributed.destroy_process_group()


def test_flat_bucket_allreduce_equals_full_batch_gradient():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    (flat0, g0), (flat1, g1) = out[0], out[1]
    assert torch.equal(flat0, flat1), "broadcast must leave every rank with rank 0's parameters"
    assert torch.equal(g0, g1), "all ranks must hold the same reduced gradient"
    # single-process reference: gradient of the mean loss over the concatenated (2 x 128) batch
    agent = ppo_oracle.AgentOracle(OBS, ACT)
    off = 0
    with torch.no_grad():
        for p in _params(agent):
            p.copy_(flat0[off : off + p.numel()].view_as(p))
            off += p.numel()
    value_rms = {"mean": torch.tensor(0.1), "var": torch.tensor(1.5)}
    loss, _ = ppo_oracle.ppo_minibatch_loss(agent, value_rms, *_batch(7, 256), norm_adv=False)
    loss.backward()
    want = _flat([p.grad for p in _params(agent)])
    # the entropy term's gradient (-ent_coef per log-std) is batch-size independent, like every mean-reduced term
    torch.testing.assert_close(g0, want, rtol=1e-4, atol=1e-7)


def test_single_process_helpers_are_noops():
    t = torch.arange(4.0)
    assert cdist.world_size() == 1 and cdist.rank() == 0
    cdist.broadcast_params(t)
    assert cdist.allreduce_grads(t) == 1.0 and torch.equal(t, torch.arange(4.0))


def _moments_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    cdist.init_from_env("gloo")
    g = torch.Generator().manual_seed(11)
    data = torch.randn(1000, 5, generator=g) * torch.tensor([1.0, 3.0, 0.1, 10.0, 2.0]) + torch.tensor([0.0, 5.0, -2.0, 100.0, 1e-3])
    shard = data[:300] if rank == 0 else data[300:]  # uneven shards
    mean, var, count = shard.mean(0), shard.var(0, unbiased=False), torch.tensor(float(len(shard)))
    keep = (mean.clone(), var.clone(), count.clone())
    m, v, n = cdist.pooled_moments(mean, var, count)
    assert all(torch.equal(a, b) for a, b in zip(keep, (mean, var, count)))  # inputs untouched
    out[rank] = (m, v, n, data.mean(0), data.var(0, unbiased=False))
    torch.distributed.destroy_process_group()


def test_pooled_moments_are_the_statistics_of_the_union_of_the_shards():
    """dist.pooled_moments (used for the normaliser entries of multi-rank checkpoints): the rank-ordered parallel-moments
    merge of per-rank (mean, population variance, count) equals the moments of all samples, identically on every rank."""
    out = mp.Manager().dict()
    mp.spawn(_moments_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for r in (0, 1):
        m, v, n, want_m, want_v = out[r]
        assert float(n) == 1000.0
        torch.testing.assert_close(m, want_m, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(v, want_v, rtol=1e-5, atol=1e-6)
    assert all(torch.equal(a, b) for a, b in zip(out[0][:3], out[1][:3]))
    # single process: a copy
    m, v, n = cdist.pooled_moments(torch.ones(3), torch.full((3,), 2.0), torch.tensor(5.0))
    assert torch.equal(m, torch.ones(3)) and torch.equal(v, torch.full((3,), 2.0)) and float(n) == 5.0
