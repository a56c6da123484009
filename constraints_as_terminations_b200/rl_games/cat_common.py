"""GAE of the reference's rl_games agent on float `dones` (reference
`exts/cat_envs/cat_envs/tasks/utils/rl_games/cat_common.py:8-112`, `cat_experience.py:21-33`).

`CaTA2CAgent` only changes rl_games' `A2CAgent` so that `dones` stay float32 termination probabilities
(`init_tensors` :14-35, `play_steps` :37-112) and then calls the stock `discount_values` (:100-102): a python loop
over the horizon with ~8 eager kernels per step.  `discount_values` below is that scan as one kernel launch
(`catb200_gae_float_dones`, variant CATB200_GAE_RLGAMES), with rl_games' argument order and tensor shapes.

rl_games is not installed where this repo is built, so the agent subclass itself cannot be defined here;
`CaTDiscountMixin` carries the one method to override:

    class CaTA2CAgent(CaTDiscountMixin, a2c_continuous.A2CAgent): ...   # plus the reference's float-dones buffers
"""

from __future__ import annotations

import torch

from .. import _lib as L
from .. import ops


def discount_values(
    fdones: torch.Tensor,
    last_extrinsic_values: torch.Tensor,
    mb_fdones: torch.Tensor,
    mb_extrinsic_values: torch.Tensor,
    mb_rewards: torch.Tensor,
    gamma: float,
    tau: float,
) -> torch.Tensor:
    """-> mb_advs, shaped like `mb_rewards` ([horizon, num_actors, 1]).

    fdones [num_actors]: dones after the last step; mb_fdones [horizon, num_actors]: dones observed before each
    step (what `play_steps` stores at :47); all float32 termination probabilities.
    """
    L.require_cuda(mb_rewards, "mb_rewards")
    f = lambda t: t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()  # noqa: E731
    advantages, _ = ops.gae_float_dones(
        L.GAE_RLGAMES, f(mb_rewards), f(mb_extrinsic_values), f(mb_fdones), f(last_extrinsic_values), gamma, gamma * tau,
        last_dones=f(fdones),
    )  # fmt: skip
    return advantages


class CaTDiscountMixin:
    """Mix into an rl_games `A2CBase` subclass: `discount_values` on the fused kernel (uses `self.gamma`, `self.tau`)."""

    def discount_values(self, fdones, last_extrinsic_values, mb_fdones, mb_extrinsic_values, mb_rewards):
        return discount_values(fdones, last_extrinsic_values, mb_fdones, mb_extrinsic_values, mb_rewards, self.gamma, self.tau)
