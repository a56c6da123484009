"""GPU parity of GAE, value-normalisation statistics, running moments and rollout append (C ABI)."""

import os

import numpy as np
import pytest
import torch

from constraints_as_terminations_b200 import ops
from oracle import ppo_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rollout(T, N, seed):
    g = torch.Generator().manual_seed(seed)
    rewards = torch.rand(T, N, generator=g) * 0.05
    values = torch.randn(T, N, generator=g)
    dones = torch.rand(T + 1, N, generator=g) * (torch.rand(T + 1, N, generator=g) < 0.4)
    dones = torch.where(torch.rand(T + 1, N, generator=g) < 0.03, torch.ones(()), dones)
    true_dones = (torch.rand(T + 1, N, generator=g) < 0.02).float()
    next_value = torch.randn(N, generator=g)
    return rewards, values, dones, true_dones, next_value


def test_gae_matches_reference_golden_bit_exact(golden_dir):
    g = torch.load(os.path.join(golden_dir, "ppo_iter.pt"), weights_only=False)
    dones = torch.cat([g["dones"], g["next_done"][None]]).to(DEV)
    true_dones = torch.cat([g["true_dones"], g["next_true_done"][None]]).to(DEV)
    adv, ret = ops.gae(
        g["rewards"].to(DEV), g["values"].to(DEV), dones, true_dones, g["next_value"].reshape(-1).to(DEV),
        g["cfg"]["gamma"], g["cfg"]["gae_lambda"],
    )  # fmt: skip
    assert torch.equal(adv.cpu(), g["advantages"])
    assert torch.equal(ret.cpu(), g["returns"])


@pytest.mark.parametrize("T,N", [(24, 4096), (24, 1), (5, 63), (1, 130), (33, 257)])
def test_gae_matches_oracle_bit_exact(T, N):
    rewards, values, dones, true_dones, next_value = _rollout(T, N, seed=T * 1000 + N)
    want_adv, want_ret = ppo_oracle.gae(rewards, values, dones[:-1], true_dones[:-1], next_value, dones[-1], true_dones[-1])
    value_rms = torch.tensor([0.0, 1.0, 1.0], device=DEV)
    stats = torch.zeros(4, device=DEV)
    adv, ret = ops.gae(
        rewards.to(DEV), values.to(DEV), dones.to(DEV), true_dones.to(DEV), next_value.to(DEV), 0.99, 0.95,
        value_rms=value_rms, norm_stats=stats,
    )  # fmt: skip
    assert torch.equal(adv.cpu(), want_adv)
    assert torch.equal(ret.cpu(), want_ret)
    # value normalisation statistics (two Chan merges): 1e-5 relative (reduction order differs from torch)
    if T * N > 1:
        s2, values_n, returns_n = ppo_oracle.value_normalisation(ppo_oracle.rms_init(()), values.reshape(-1), want_ret.reshape(-1))
        torch.testing.assert_close(value_rms.cpu(), torch.stack([s2["mean"], s2["var"], s2["count"]]), rtol=1e-5, atol=1e-6)
        st = stats.cpu()
        got_values_n = (values.reshape(-1) - st[0]) / torch.sqrt(st[1] + 1e-8)
        got_returns_n = (want_ret.reshape(-1) - st[2]) / torch.sqrt(st[3] + 1e-8)
        torch.testing.assert_close(got_values_n, values_n, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(got_returns_n, returns_n, rtol=1e-5, atol=1e-5)


def test_gae_full_size_properties():
    T, N = 24, 65536
    rewards, values, dones, true_dones, next_value = (t.to(DEV) for t in _rollout(T, N, seed=3))
    adv, ret = ops.gae(rewards, values, dones, true_dones, next_value, 0.99, 0.95)
    assert torch.equal(ret, adv + values)  # returns = advantages + values, rounded once
    # where the next step is a certain termination (done == 1) the advantage is reward - value
    certain = dones[1:] == 1.0
    assert torch.equal(adv[certain], (rewards - values)[certain])
    # all-terminal rollout: no bootstrapping anywhere
    adv1, _ = ops.gae(rewards, values, torch.ones_like(dones), true_dones, next_value, 0.99, 0.95)
    assert torch.equal(adv1, rewards - values)
    # gamma = 0: one-step advantage
    adv0, _ = ops.gae(rewards, values, dones, true_dones, next_value, 0.0, 0.95)
    assert torch.equal(adv0, rewards + 0.0 - values)


@pytest.mark.parametrize("rows,dim", [(4096, 45), (1, 45), (3, 7), (1000, 1), (98304, 1), (257, 64)])
def test_running_moments_match_oracle(rows, dim):
    g = torch.Generator().manual_seed(rows + dim)
    shape = (dim,) if dim > 1 else ()
    oracle = ppo_oracle.rms_init(shape)
    mean = torch.zeros(dim, device=DEV)
    var = torch.ones(dim, device=DEV)
    count = torch.ones(1, device=DEV)
    ws = ops.Workspace(DEV)
    for step in range(4):
        x = torch.randn(rows, dim, generator=g) * torch.linspace(0.5, 4.0, dim) + (step - 1.5)
        xin = x if dim > 1 else x[:, 0]
        oracle = ppo_oracle.rms_update(oracle, xin)
        want = ppo_oracle.rms_normalize(oracle, xin)
        got = ops.rms_forward(xin.contiguous().to(DEV), mean, var, count, workspace=ws)
        torch.testing.assert_close(mean.cpu().reshape(oracle["mean"].shape), oracle["mean"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(var.cpu().reshape(oracle["var"].shape), oracle["var"], rtol=1e-5, atol=1e-6)
        assert float(count) == float(oracle["count"])
        torch.testing.assert_close(got.cpu(), want, rtol=1e-5, atol=1e-5)
        # given the statistics on the device, the normalisation is correctly rounded IEEE fp32: checked
        # against numpy (torch's vectorised CPU sqrt is not correctly rounded for every input, so torch
        # itself is only matched to 1e-5 above)
        m_np, v_np = mean.cpu().numpy(), var.cpu().numpy()
        d_np = np.sqrt((v_np + np.float32(1e-8)).astype(np.float32)).astype(np.float32)
        want_np = ((x.numpy() - m_np).astype(np.float32) / d_np).astype(np.float32)
        assert np.array_equal(got.cpu().numpy().reshape(rows, dim), want_np)
    # update=False leaves the statistics alone
    before = (mean.clone(), var.clone(), count.clone())
    x = torch.randn(rows, dim, generator=g).to(DEV)
    ops.rms_forward(x if dim > 1 else x[:, 0].contiguous(), mean, var, count, update=False)
    assert torch.equal(mean, before[0]) and torch.equal(var, before[1]) and torch.equal(count, before[2])


def test_obs_rms_golden_sequence(golden_dir):
    """The reference trainer's obs_rms over one rollout (reset obs + 24 steps), ppo.py:187,225."""
    g = torch.load(os.path.join(golden_dir, "ppo_iter.pt"), weights_only=False)
    mean, var, count = torch.zeros(45, device=DEV), torch.ones(45, device=DEV), torch.ones(1, device=DEV)
    ws = ops.Workspace(DEV)
    first = ops.rms_forward(g["env_trace_reset_obs"].to(DEV), mean, var, count, workspace=ws)
    torch.testing.assert_close(first.cpu(), g["obs"][0], rtol=1e-5, atol=1e-5)
    raw = g["env_trace_raw_obs"].to(DEV)
    for t in range(raw.shape[0]):
        normed = ops.rms_forward(raw[t], mean, var, count, workspace=ws)
        want = g["obs"][t + 1] if t + 1 < g["obs"].shape[0] else g["next_obs"]
        torch.testing.assert_close(normed.cpu(), want, rtol=1e-5, atol=1e-5)
    ref = g["rms_after_rollout"]
    torch.testing.assert_close(mean.cpu(), ref["obs_rms.running_mean"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(var.cpu(), ref["obs_rms.running_var"], rtol=1e-5, atol=1e-6)
    assert float(count) == float(ref["obs_rms.count"])


def test_rollout_append():
    n = 1000
    reward, done = torch.rand(n, device=DEV), torch.rand(n, device=DEV)
    time_out = torch.rand(n, device=DEV) < 0.1
    rewards = torch.zeros(3, n, device=DEV)
    dones = torch.zeros(4, n, device=DEV)
    true_dones = torch.zeros(4, n, device=DEV)
    ops.rollout_append(reward, done, time_out, rewards[1], dones[2], true_dones[2])
    assert torch.equal(rewards[1], reward) and torch.equal(dones[2], done) and torch.equal(true_dones[2], time_out.float())
    assert float(rewards[0].abs().sum() + rewards[2].abs().sum() + dones[1].abs().sum()) == 0.0
