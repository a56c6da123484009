"""Multi-GPU gradient exchange on real hardware (needs >= 2 GPUs; skipped otherwise): the one-shot peer-memory
all-reduce + gradient-norm kernel (csrc/peer.cu), the identity "rank-summed CUDA gradients == single-GPU gradient of the
concatenated minibatch", and a 2-rank trainer run (one CUDA graph per epoch on every rank) against the NCCL path."""

import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
needs_two = pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")


def _free_ports(k=1):
    """k distinct free ports (all sockets held open together so the kernel cannot hand the same one out twice)."""
    socks = [socket.socket() for _ in range(k)]
    try:
        for s in socks:
            s.bind(("127.0.0.1", 0))
        return [s.getsockname()[1] for s in socks]
    finally:
        for s in socks:
            s.close()


def _free_port():
    return _free_ports(1)[0]


def _init(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", rank))


def _vec(seed, n):
    return torch.randn(n, generator=torch.Generator().manual_seed(seed))


def _exchange_worker(rank, world, port, out):
    _init(rank, world, port)
    from constraints_as_terminations_b200 import dist as cdist

    dev = torch.device("cuda", rank)
    n = 377241
    ex = cdist.PeerGradExchange(n, dev)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    opt_ws = torch.zeros(8, dtype=torch.int64, device=dev)
    norm = torch.zeros(1, device=dev)
    ok = True
    for epoch in range(4):
        parity = epoch & 1
        ex.arena[parity].copy_(_vec(100 * epoch + rank, n))
        got = ex.reduce(parity, step, opt_ws, max_grad_norm=1.0, grad_norm_out=norm).clone()
        want = sum(_vec(100 * epoch + r, n) for r in range(world))
        torch.cuda.synchronize()
        ok &= torch.equal(got.cpu(), want) if world == 2 else bool(torch.allclose(got.cpu(), want, rtol=1e-6, atol=1e-6))
        ok &= abs(float(norm) - float((want / world).double().norm())) < 1e-3
        ok &= float(ex.arena[parity ^ 1].abs().sum()) == 0.0  # the other arena was zeroed for the next minibatch
        ok &= int(step) == epoch + 1 and int(ex.epoch) == epoch + 1
    ex.check()
    out[rank] = ok
    ex.close()
    torch.distributed.destroy_process_group()


@needs_two
def test_peer_allreduce_sums_in_rank_order_and_yields_the_norm():
    out = mp.Manager().dict()
    mp.spawn(_exchange_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


def _gradient_worker(rank, world, port, out):
    _init(rank, world, port)
    from constraints_as_terminations_b200 import dist as cdist
    from constraints_as_terminations_b200 import ops
    from tests import test_mlp_gpu as T

    dev = torch.device("cuda", rank)
    agent = T.make_agent(seed=1)
    dims = ops.make_dims(T.OBS, T.ACT, precision="tf32")
    layout = ops.mlp_layout(dims)
    params = T.flat_params(agent, layout).to(dev)
    wc = ops.weight_copies(dims, layout, dev)
    ops.cast_weights(dims, params, wc)
    hp = ops.make_hparams(norm_adv=False)  # per-minibatch advantage normalisation is rank-local by design: off for the identity
    M = 4096
    shards = [T._minibatch(agent, M, M, seed=50 + r) for r in range(world)]  # every rank can rebuild every shard

    def grad_of(obs, actions, logp, adv, returns, values, norm_stats, rows, into):
        idx = torch.arange(rows, device=dev)
        ws = ops.mlp_workspace(dims, rows, True, dev)
        loss_acc = torch.zeros(8, device=dev)
        ops.ppo_minibatch_grad(dims, hp, idx, ops.obs_to_operand(dims, obs.to(dev)), actions.to(dev), logp.to(dev), adv.to(dev),
                               returns.to(dev), values.to(dev), norm_stats.to(dev), params, wc, into, loss_acc, ws)  # fmt: skip

    ex = cdist.PeerGradExchange(layout.n_params, dev)
    obs, actions, logp, adv, returns, values, norm_stats, _ = shards[rank]
    grad_of(obs, actions, logp, adv, returns, values, norm_stats, M, ex.arena[0])
    step, opt_ws = torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(8, dtype=torch.int64, device=dev)
    gsum = ex.reduce(0, step, opt_ws).clone()
    ex.check()
    # single-GPU gradient of the concatenated minibatch (mean loss over world * M rows) == mean of the per-shard gradients
    cat = [torch.cat([s[i] for s in shards]) for i in range(6)]
    full = torch.zeros(layout.n_params, device=dev)
    grad_of(*cat, shards[0][6], world * M, full)
    torch.cuda.synchronize()
    rel = float((gsum / world - full).norm() / full.norm())
    out[rank] = rel
    ex.close()
    torch.distributed.destroy_process_group()


@needs_two
def test_rank_summed_gradient_equals_single_gpu_gradient_of_the_concatenated_minibatch():
    out = mp.Manager().dict()
    mp.spawn(_gradient_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    # identical arithmetic per sample; only the order of fp32 summation differs (atomics, split sizes)
    assert max(out.values()) < 2e-5, dict(out)


def _trainer_worker(rank, world, port, peer, out, fused=True):
    # fused: True (one launch, default exchange pattern), False (three launches), "rs" (one launch, reduce-scatter / all-gather)
    os.environ["CATB200_PEER_ALLREDUCE"] = "1" if peer else "0"
    os.environ["CATB200_FUSED_OPT"] = "1" if fused else "0"
    os.environ["CATB200_PEER_RS"] = "1" if fused == "rs" else "0"
    _init(rank, world, port)
    from constraints_as_terminations_b200 import PPOTrainer, solo12_flat_ppo_cfg
    from constraints_as_terminations_b200 import synthetic_env as se

    dev = torch.device("cuda", rank)
    torch.manual_seed(0)
    env = se.SyntheticSolo12Env(512, device=dev, seed=rank, pool=3, episode_length=20, constraints_cfg=se.solo12_constraints_cfg())
    env.load_managers()
    cfg = solo12_flat_ppo_cfg(logger=None, num_steps=8, minibatch_size=1024, updates_epochs=2, num_iterations=4)
    tr = PPOTrainer(env, cfg, device=dev, use_graphs=True)
    assert (tr.peer is not None) == peer
    tr.start()
    for _ in range(3):
        tr.train_iteration()
        tr.losses()
    torch.cuda.synchronize()
    out[(peer, rank) if fused is True else (peer, rank, "split" if not fused else fused)] = tr.agent.parameters_flat().detach().cpu()
    if tr.peer is not None:
        tr.peer.check()
        tr.peer.close()
    torch.distributed.destroy_process_group()


@needs_two
def test_two_rank_trainer_peer_exchange_matches_nccl_and_keeps_ranks_identical():
    out = mp.Manager().dict()
    for peer, port in zip((True, False), _free_ports(2)):
        mp.spawn(_trainer_worker, args=(2, port, peer, out), nprocs=2, join=True)
    assert torch.equal(out[(True, 0)], out[(True, 1)])  # rank-ordered sums: every rank takes bit-identical steps
    assert torch.equal(out[(False, 0)], out[(False, 1)])
    torch.testing.assert_close(out[(True, 0)], out[(False, 0)], rtol=1e-3, atol=3e-4)  # vs NCCL: summation order only


@needs_two
def test_single_launch_peer_optimizer_step_equals_the_three_launch_one():
    """catb200_ppo_minibatch_update_peer (fold + handshake + rank-ordered sum + norm + clip + Adam + operand refresh in ONE
    launch around two local grid barriers) against fold_grads + grad_allreduce_norm + adam_cast: same per-element
    arithmetic.  Two runs of either path already differ in the last bits of the weight gradients (red.global.add order of
    mlp_wgrad_kernel) and Adam's first steps normalise every gradient element by its own magnitude, so near-zero elements
    may move by up to 2 lr per step in either direction: same tolerance as the comparison with NCCL above."""
    out = mp.Manager().dict()
    for fused, port in zip((True, False, "rs"), _free_ports(3)):
        mp.spawn(_trainer_worker, args=(2, port, True, out, fused), nprocs=2, join=True)
    assert torch.equal(out[(True, 0)], out[(True, 1)])
    assert torch.equal(out[(True, 0, "split")], out[(True, 1, "split")])
    torch.testing.assert_close(out[(True, 0)], out[(True, 0, "split")], rtol=1e-3, atol=3e-4)
    # the reduce-scatter / all-gather pattern (default beyond two ranks, forced here): every slice is summed once, in rank
    # order, and copied -> still bit-identical on every rank
    assert torch.equal(out[(True, 0, "rs")], out[(True, 1, "rs")])
    torch.testing.assert_close(out[(True, 0, "rs")], out[(True, 0, "split")], rtol=1e-3, atol=3e-4)
