"""ORACLE (test infrastructure, never shipped or timed as the product).

CPU restatement, in eager fp32 torch, of the per-update path of the reference's
CleanRL-style trainer (`U/cleanrl/ppo.py`, `U/` = `exts/cat_envs/cat_envs/tasks/utils/`):
running mean/std, the float-dones GAE scan, the actor-critic MLP and the
PPO-clip minibatch loss.  Pinned by `oracle/make_golden.py` against the real
reference module imported in the build container (fixtures in `tests/golden/`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module.
"""

from __future__ import annotations

import math

import torch
import torch.nn as nn


# ------------------------------------------------------------------------------------------------
# RunningMeanStd  (U/cleanrl/ppo.py:12-62)
# ------------------------------------------------------------------------------------------------
def rms_init(shape=()):
    """mean 0, var 1, count 1 (ppo.py:15-17)."""
    return {"mean": torch.zeros(shape), "var": torch.ones(shape), "count": torch.ones(())}


def rms_update(state: dict, x: torch.Tensor) -> dict:
    """Chan parallel-moments merge of the batch into the running stats (ppo.py:27-62)."""
    batch_mean = torch.mean(x, dim=0)
    batch_var = torch.var(x, correction=0, dim=0)
    n = x.shape[0]
    delta = batch_mean - state["mean"]
    tot = state["count"] + n
    new_mean = state["mean"] + delta * n / tot
    m2 = state["var"] * state["count"] + batch_var * n + torch.square(delta) * state["count"] * n / tot
    return {"mean": new_mean, "var": m2 / tot, "count": tot}


def rms_normalize(state: dict, x: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """(x - mean) / sqrt(var + eps) with whatever stats are current (ppo.py:21-25)."""
    return (x - state["mean"]) / torch.sqrt(state["var"] + eps)


# ------------------------------------------------------------------------------------------------
# GAE with float dones  (U/cleanrl/ppo.py:251-277)
# ------------------------------------------------------------------------------------------------
def gae(rewards, values, dones, true_dones, next_value, next_done, next_true_done, gamma=0.99, lam=0.95):
    """Reverse scan; `dones` are termination *probabilities*, `true_dones` are time-outs (floats).

    rewards/values/dones/true_dones: [T, N]; next_value: [1, N] or [N]; next_done/next_true_done: [N].
    Returns (advantages, returns), both [T, N].
    """
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    last = 0
    next_value = next_value.reshape(1, -1)
    for t in reversed(range(T)):
        if t == T - 1:
            nnt = 1.0 - next_done
            tnnt = 1 - next_true_done
            nv = next_value
        else:
            nnt = 1.0 - dones[t + 1]
            tnnt = 1 - true_dones[t + 1]
            nv = values[t + 1]
        delta = rewards[t] + gamma * nv * nnt * tnnt - values[t]
        adv[t] = last = delta + gamma * lam * nnt * tnnt * last
    return adv, adv + values


# ------------------------------------------------------------------------------------------------
# Agent  (U/cleanrl/ppo.py:65-123)
# ------------------------------------------------------------------------------------------------
def _mlp(sizes, last_std):
    layers = []
    for i in range(len(sizes) - 1):
        lin = nn.Linear(sizes[i], sizes[i + 1])
        last = i == len(sizes) - 2
        nn.init.orthogonal_(lin.weight, last_std if last else math.sqrt(2))  # ppo.py:65-68
        nn.init.constant_(lin.bias, 0.0)
        layers.append(lin)
        if not last:
            layers.append(nn.ELU())
    return nn.Sequential(*layers)


class AgentOracle(nn.Module):
    """Separate actor / critic MLPs obs->512->256->128->{A,1}, ELU, state-independent log-std.

    Parameter / buffer names equal the reference's so that `state_dict()`s interchange
    (ppo.py:78-99): critic.{0,2,4,6}, actor_mean.{0,2,4,6}, actor_logstd, obs_rms.*, value_rms.*.
    """

    def __init__(self, obs_dim: int = 45, act_dim: int = 12):
        super().__init__()
        self.critic = _mlp([obs_dim, 512, 256, 128, 1], 1.0)
        self.actor_mean = _mlp([obs_dim, 512, 256, 128, act_dim], 0.01)
        self.actor_logstd = nn.Parameter(torch.zeros(1, act_dim))

    def evaluate(self, obs, action):
        """log-prob sum, entropy sum, value[N,1] for given actions (ppo.py:104-119 with action given)."""
        mean = self.actor_mean(obs)
        logstd = self.actor_logstd.expand_as(mean)
        std = torch.exp(logstd)
        var = std**2
        log_scale = std.log()  # torch.distributions.Normal takes log(exp(logstd)), not logstd itself
        logp = -((action - mean) ** 2) / (2 * var) - log_scale - math.log(math.sqrt(2 * math.pi))
        ent = 0.5 + 0.5 * math.log(2 * math.pi) + log_scale
        return logp.sum(1), ent.sum(1), self.critic(obs)

    def act(self, obs, noise):
        """Sampled action mean + std * noise (what Normal.sample does with eps = noise), its log-prob, value."""
        mean = self.actor_mean(obs)
        std = torch.exp(self.actor_logstd.expand_as(mean))
        action = mean + std * noise
        logp, _, value = self.evaluate(obs, action)
        return action, logp, value


# ------------------------------------------------------------------------------------------------
# PPO-clip minibatch loss  (U/cleanrl/ppo.py:300-344)
# ------------------------------------------------------------------------------------------------
def ppo_minibatch_loss(
    agent: AgentOracle,
    value_rms: dict,
    mb_obs,
    mb_actions,
    mb_logprobs,
    mb_advantages,
    mb_returns_n,
    mb_values_n,
    clip_coef=0.2,
    ent_coef=0.001,
    vf_coef=2.0,
    norm_adv=True,
    clip_vloss=True,
):
    """Returns (loss, dict of scalars).  `*_n` are the value-normalised returns / values (ppo.py:287-288)."""
    newlogprob, entropy, newvalue = agent.evaluate(mb_obs, mb_actions)
    logratio = newlogprob - mb_logprobs
    ratio = logratio.exp()
    with torch.no_grad():
        approx_kl = ((ratio - 1) - logratio).mean()
        clipfrac = ((ratio - 1.0).abs() > clip_coef).float().mean()
    adv = mb_advantages
    if norm_adv:  # ppo.py:315-318, unbiased std
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    pg_loss = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1 - clip_coef, 1 + clip_coef)).mean()
    newvalue = rms_normalize(value_rms, newvalue.view(-1))  # ppo.py:329 (update=False)
    if clip_vloss:  # ppo.py:330-339
        v_unclipped = (newvalue - mb_returns_n) ** 2
        v_clipped = mb_values_n + torch.clamp(newvalue - mb_values_n, -clip_coef, clip_coef)
        v_loss = 0.5 * torch.max(v_unclipped, (v_clipped - mb_returns_n) ** 2).mean()
    else:
        v_loss = 0.5 * ((newvalue - mb_returns_n) ** 2).mean()
    entropy_loss = entropy.mean()
    loss = pg_loss - ent_coef * entropy_loss + v_loss * vf_coef  # ppo.py:344
    return loss, {
        "pg_loss": pg_loss.detach(),
        "v_loss": v_loss.detach(),
        "entropy": entropy_loss.detach(),
        "approx_kl": approx_kl,
        "clipfrac": clipfrac,
    }


def value_normalisation(value_rms: dict, values_flat, returns_flat):
    """The double update of ppo.py:287-288: stats absorb `values`, then `returns`; each is normalised
    with the stats current at that moment.  Returns (new_state, values_n, returns_n)."""
    s1 = rms_update(value_rms, values_flat)
    values_n = rms_normalize(s1, values_flat)
    s2 = rms_update(s1, returns_flat)
    returns_n = rms_normalize(s2, returns_flat)
    return s2, values_n, returns_n
