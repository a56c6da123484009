"""GPU tests of the trainer: replay of one recorded iteration of the REAL reference trainer
(tests/golden/ppo_iter.pt) and an end-to-end run on the synthetic Solo12 env."""

import os
import types

import pytest
import torch

from constraints_as_terminations_b200 import Agent, PPOTrainer, ops, solo12_flat_ppo_cfg
from constraints_as_terminations_b200 import synthetic_env as se

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class ShapeOnlyEnv:
    def __init__(self, n):
        self.num_envs = n
        self.device = torch.device(DEV)
        self.unwrapped = self
        self.single_observation_space = {"policy": types.SimpleNamespace(shape=(se.OBS_DIM,))}
        self.single_action_space = types.SimpleNamespace(shape=(se.ACT_DIM,))


def test_state_dict_keys_match_reference(golden_dir):
    g = torch.load(os.path.join(golden_dir, "ppo_iter.pt"), weights_only=False)
    agent = Agent(ShapeOnlyEnv(4), device=DEV)
    assert list(agent.state_dict().keys()) == list(g["init_state"].keys())
    for k, v in agent.state_dict().items():
        assert v.shape == g["init_state"][k].shape, k
    agent.load_state_dict(g["final_state"])
    for k, v in agent.state_dict().items():
        assert torch.equal(v.cpu(), g["final_state"][k]), k
    # parameters live in one flat vector that the kernels update in place
    assert agent.critic[0].weight.data_ptr() == agent.parameters_flat().data_ptr()
    # the operand-precision compute copies followed the load (default precision tf32: fp32 rounded to 10 mantissa bits)
    w = agent.actor_mean[2].weight
    lay = agent.layout
    assert agent.precision == "tf32"
    got = agent._wc[lay.wc[1][1] : lay.wc[1][1] + w.numel()].view_as(w)
    torch.testing.assert_close(got, w.detach(), rtol=2**-11, atol=0)
    assert int((got.view(torch.int32) & 0x1FFF).abs().max()) == 0


def test_trainer_replays_reference_iteration(golden_dir):
    g = torch.load(os.path.join(golden_dir, "ppo_iter.pt"), weights_only=False)
    cfg_d = g["cfg"]
    n, T = g["num_envs"], cfg_d["num_steps"]
    cfg = solo12_flat_ppo_cfg(logger=None, num_iterations=1, minibatch_size=cfg_d["minibatch_size"], updates_epochs=cfg_d["updates_epochs"])
    tr = PPOTrainer(ShapeOnlyEnv(n), cfg, device=DEV, use_graphs=False)
    tr.agent.load_state_dict(g["init_state"])
    # recorded rollout -> trainer buffers (slot T = bootstrap obs / next_done)
    tr.obs[:T].copy_(g["obs"]); tr.obs[T].copy_(g["next_obs"])
    ops.obs_to_operand(tr.agent.dims, tr.obs.view(-1, se.OBS_DIM), out=tr.obs_op.view(-1, tr.agent.dims.obs_pad))
    tr.actions.copy_(g["actions"]); tr.logprobs.copy_(g["logprobs"]); tr.rewards.copy_(g["rewards"]); tr.values.copy_(g["values"])
    tr.dones[:T].copy_(g["dones"]); tr.dones[T].copy_(g["next_done"])
    tr.true_dones[:T].copy_(g["true_dones"]); tr.true_dones[T].copy_(g["next_true_done"])
    # our own bootstrap value (tf32 tensor-core MLP) vs the reference's fp32 one
    tr.compute_gae(bootstrap=True)
    torch.testing.assert_close(tr.next_value.cpu(), g["next_value"].reshape(-1), rtol=4e-3, atol=4e-3)
    # with the recorded bootstrap value GAE is bit-exact and the value statistics agree to 1e-5
    tr.agent.load_state_dict(g["init_state"])
    tr.next_value.copy_(g["next_value"].reshape(-1))
    tr.compute_gae(bootstrap=False)
    assert torch.equal(tr.advantages.cpu(), g["advantages"]) and torch.equal(tr.returns.cpu(), g["returns"])
    ref = g["rms_after_rollout"]
    torch.testing.assert_close(tr.agent.value_rms.running_mean.cpu(), ref["value_rms.running_mean"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(tr.agent.value_rms.running_var.cpu(), ref["value_rms.running_var"], rtol=1e-5, atol=1e-6)
    assert float(tr.agent.value_rms.count) == float(ref["value_rms.count"])
    st = tr.norm_stats.cpu()
    torch.testing.assert_close((g["values"].reshape(-1) - st[0]) / torch.sqrt(st[1] + 1e-8), g["b_values"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close((g["returns"].reshape(-1) - st[2]) / torch.sqrt(st[3] + 1e-8), g["b_returns"], rtol=1e-4, atol=1e-4)
    # the recorded minibatch permutations -> same sequence of 6 optimizer steps
    tr.set_lr(g["lr"])
    tr.update(perms=[p.to(DEV) for p in g["perms"]])
    losses = tr.losses()
    n_mb = len(g["perms"]) * (n * T // cfg_d["minibatch_size"])
    # tf32 tensor-core numerics vs the reference's fp32 CPU run: losses to 3e-3 (relative or absolute)
    assert losses["mean_pg_loss"] == pytest.approx(g["sum_pg_loss"] / n_mb, rel=3e-3, abs=3e-4)
    assert losses["mean_v_loss"] == pytest.approx(g["sum_v_loss"] / n_mb, rel=3e-3, abs=3e-4)
    assert losses["mean_entropy_loss"] == pytest.approx(g["sum_entropy_loss"] / n_mb, rel=1e-4)
    # parameter update after the 6 Adam steps: same direction and size as the reference's
    init = torch.cat([g["init_state"][k].reshape(-1) for k in g["init_state"] if "_rms." not in k])
    want = torch.cat([g["final_state"][k].reshape(-1) for k in g["final_state"] if "_rms." not in k]) - init
    got = torch.cat([v.detach().cpu().reshape(-1) for k, v in tr.agent.state_dict().items() if "_rms." not in k]) - init
    cos = float(torch.dot(got, want) / (got.norm() * want.norm()))
    assert cos > 0.995, f"update direction cosine {cos:.4f}"
    assert float(got.norm() / want.norm()) == pytest.approx(1.0, abs=0.02)


def _make_trainer(n, T, mb, graphs, seed):
    torch.manual_seed(seed)
    env = se.SyntheticSolo12Env(n, device=DEV, seed=seed, pool=3, episode_length=20, constraints_cfg=se.solo12_constraints_cfg())
    env.load_managers()
    cfg = solo12_flat_ppo_cfg(logger=None, num_steps=T, minibatch_size=mb, updates_epochs=2, num_iterations=4)
    tr = PPOTrainer(env, cfg, device=DEV, use_graphs=graphs)
    tr.start()
    return env, tr


def test_trainer_end_to_end_on_synthetic_env():
    n, T = 512, 8
    env, tr = _make_trainer(n, T, 1024, graphs=True, seed=3)
    p0 = tr.agent.parameters_flat().clone()
    for _ in range(3):
        infos = tr.train_iteration()
        assert len(infos) > 0  # resets happened (episode_length 20) and carried the constraint statistics
        handle = tr.losses_async()  # the non-blocking read-back a logger would use ...
        losses = tr.losses()
        assert tr.read_losses(handle) == losses  # ... returns what the blocking read returns
        assert all(torch.isfinite(torch.tensor(v)) for v in losses.values()), losses
    torch.cuda.synchronize()
    assert float(tr.agent.obs_rms.count) == 1 + n * (1 + 3 * T)
    assert float(tr.agent.value_rms.count) == 1 + 3 * 2 * n * T
    assert int(tr.step_dev) == 3 * 2 * (n * T // 1024)
    assert not torch.equal(p0, tr.agent.parameters_flat())
    assert torch.isfinite(tr.agent.parameters_flat()).all()
    assert float(tr.grads.abs().sum()) == 0.0
    # float dones are probabilities in [0, 1]; hard resets are exactly 1 (cat_env.py:107,121)
    assert float(tr.dones.min()) >= 0.0 and float(tr.dones.max()) <= 1.0
    keys = list(infos[-1].keys())
    assert any(k.startswith("Episode_Constraint_violation/") for k in keys)
    assert tr.kernel_launches() > 0
    # lr annealing reached iteration 3 of 4 (ppo.py:196-199)
    assert float(tr.lr_dev) == pytest.approx(3e-4 * (1 - 2 / 4), rel=1e-6)


def test_graph_and_eager_updates_agree():
    """Eager launches vs CUDA graphs (whole env steps: policy + env device half + append + observation statistics, one
    graph per (slot, state-pointer set); whole epochs incl. the Philox permutation): same parameters, same logged
    statistics.  6 iterations: pool 3 / T 8 -> the (slot, pointer) keys repeat from iteration 5 on, i.e. graphs replay."""
    outs, logs, ngraphs = [], [], []
    for graphs in (False, True):
        env, tr = _make_trainer(256, 8, 512, graphs=graphs, seed=5)
        keys = []
        for _ in range(6):
            infos = tr.train_iteration()
            keys.append(sorted({k for i in infos for k in i}))
            last = {k: float(v) for k, v in infos[-1].items()} if infos else {}
        torch.cuda.synchronize()
        outs.append(tr.agent.parameters_flat().clone())
        logs.append((keys, last))
        ngraphs.append(len(tr._step_graphs))
        assert float(tr.agent.obs_rms.count) == 1 + 256 * (1 + 6 * 8)
    assert ngraphs[0] == 0 and ngraphs[1] == 24  # 8 slots x 3 pointer sets
    assert logs[0][0] == logs[1][0]
    for k, v in logs[0][1].items():
        assert logs[1][1][k] == pytest.approx(v, rel=1e-4, abs=1e-6, nan_ok=True), k
    # same launches either way; the weight gradients are summed with fp32 atomics (red.global.add), whose order varies
    # from run to run: two eager runs differ by up to ~5e-5 per parameter after these 16 Adam steps (measured), and so
    # do an eager and a graph run (48 Adam steps here)
    torch.testing.assert_close(outs[0], outs[1], rtol=1e-3, atol=1e-3)
