"""`compute_gae` of the CaT skrl agent (reference `exts/cat_envs/cat_envs/tasks/utils/skrl/ppo.py:397-442`).

In the reference this is a closure inside `PPO._update`: a python loop over the `memory_size` rollout steps
(~8 eager kernels per step) with the CaT change `not_dones = 1 - dones` on the float `terminated` tensor (:421),
followed by a global advantage normalisation (:437).  Here the whole scan, `returns = advantages + values` and the
normalisation are two kernel launches (`catb200_gae_float_dones`, variant CATB200_GAE_SKRL).

To use it from the reference's agent, replace the body of the nested function by a call to this one (the agent class
itself subclasses skrl's `Agent`, which is not installed where this repo is built).
"""

from __future__ import annotations

import torch

from .. import _lib as L
from .. import ops

_workspaces: dict = {}


def compute_gae(
    rewards: torch.Tensor,
    dones: torch.Tensor,
    values: torch.Tensor,
    next_values: torch.Tensor,
    discount_factor: float = 0.99,
    lambda_coefficient: float = 0.95,
    normalize: bool = True,
):
    """-> (returns, advantages), shaped like `rewards` ([memory_size, num_envs, 1] in skrl's memory).

    `dones` is the float `terminated` tensor (termination probabilities), `next_values` the critic's value of the
    state after the last stored step (the reference's closure variable `last_values`, :447-452).
    """
    L.require_cuda(rewards, "rewards")
    f = lambda t: t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()  # noqa: E731
    ws = _workspaces.get(rewards.device)
    if ws is None:
        ws = _workspaces[rewards.device] = ops.Workspace(rewards.device)
    advantages, returns = ops.gae_float_dones(
        L.GAE_SKRL, f(rewards), f(values), f(dones), f(next_values), discount_factor, lambda_coefficient,
        normalize=normalize, workspace=ws,
    )  # fmt: skip
    return returns, advantages
