"""ORACLE (test infrastructure, never shipped or timed as the product).

CPU restatement, in eager fp32 torch, of the two single-`dones` GAE scans the reference's other
front-ends use (`U/` = `exts/cat_envs/cat_envs/tasks/utils/`):

* `skrl_compute_gae`   follows the nested `compute_gae` of the CaT skrl agent, `U/skrl/ppo.py:397-442`
                       (`not_dones = 1 - dones` is the CaT modification, :421).  PINNED: `oracle/make_golden_gae.py`
                       extracts that very function from the reference file, executes it on seeded inputs and
                       stores the outputs in `tests/golden/gae_variants.pt`; `tests/test_oracle_golden.py`
                       replays them through this restatement bit for bit (normalised advantages: 1e-6).
* `rlgames_discount_values`  follows rl_games' `A2CBase.discount_values` (rl_games/common/a2c_common.py),
                       which `CaTA2CAgent.play_steps` calls at `U/rl_games/cat_common.py:96-103` on the float
                       `dones` of `CaTExperienceBuffer` (`U/rl_games/cat_experience.py:27-33`).  rl_games is a
                       third-party dependency that is neither vendored in /root/reference nor version-pinned by it
                       (`setup.py:16-19`) nor installed here: PARITY UNPINNED for this function.  The restatement
                       is the library's published algorithm; it is anchored on the reference's call site (argument
                       order, float dones, `mb_returns = mb_advs + mb_values` :104) and cross-checked against the
                       pinned CleanRL GAE, to which it reduces when no env times out.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs may import this module.
"""

from __future__ import annotations

import torch


def skrl_compute_gae(rewards, dones, values, last_values, discount_factor=0.99, lambda_coefficient=0.95, normalize=True):
    """rewards / dones / values: [T, N] or [T, N, 1]; last_values: [N] or [N, 1].  -> (returns, advantages)."""
    advantage = 0
    advantages = torch.zeros_like(rewards)
    not_dones = 1 - dones  # U/skrl/ppo.py:421
    memory_size = rewards.shape[0]
    for i in reversed(range(memory_size)):  # :425-433
        next_values = values[i + 1] if i < memory_size - 1 else last_values
        advantage = rewards[i] - values[i] + discount_factor * not_dones[i] * (next_values + lambda_coefficient * advantage)
        advantages[i] = advantage
    returns = advantages + values  # :435
    if normalize:  # :437
        advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    return returns, advantages


def rlgames_discount_values(fdones, last_values, mb_fdones, mb_values, mb_rewards, gamma=0.99, tau=0.95):
    """fdones: [N]; last_values: [N, 1]; mb_fdones: [T, N]; mb_values, mb_rewards: [T, N, 1].  -> mb_advs [T, N, 1]."""
    horizon = mb_rewards.shape[0]
    lastgaelam = 0
    mb_advs = torch.zeros_like(mb_rewards)
    for t in reversed(range(horizon)):
        if t == horizon - 1:
            nextnonterminal = 1.0 - fdones
            nextvalues = last_values
        else:
            nextnonterminal = 1.0 - mb_fdones[t + 1]
            nextvalues = mb_values[t + 1]
        nextnonterminal = nextnonterminal.unsqueeze(1)
        delta = mb_rewards[t] + gamma * nextvalues * nextnonterminal - mb_values[t]
        mb_advs[t] = lastgaelam = delta + gamma * tau * nextnonterminal * lastgaelam
    return mb_advs


def sample_inputs(T: int, N: int, seed: int):
    """Seeded rollout-shaped inputs (shared by the golden generator and the GPU parity tests)."""
    g = torch.Generator().manual_seed(seed)
    rewards = torch.rand(T, N, generator=g) * 0.06
    values = torch.randn(T, N, generator=g) * 0.7 + 0.3
    # CaT dones are termination probabilities: mostly 0, some fractional, a few certain terminations
    u = torch.rand(T + 1, N, generator=g)
    dones = torch.where(u < 0.7, torch.zeros(()), torch.where(u < 0.95, torch.rand(T + 1, N, generator=g) * 0.25, torch.ones(())))
    last_values = torch.randn(N, generator=g) * 0.7 + 0.3
    return rewards, values, dones, last_values
