"""`ConstraintTermCfg` — the per-term config object of the constraint manager.

Boundary type mirrored from the reference
(`exts/cat_envs/cat_envs/tasks/utils/cat/manager_constraint_cfg.py:23-27`): a
manager-term cfg carrying `func`, `params` and the maximum termination
probability `max_p`.  Task files import it as `ConstraintTerm`
(`.../solo12/cat_flat_env_cfg.py:16-18`), so both names are exported.
"""

from __future__ import annotations

from collections.abc import Callable
from dataclasses import MISSING

import torch

from ._isaaclab_compat import ManagerTermBaseCfg, configclass


@configclass
class ConstraintTermCfg(ManagerTermBaseCfg):
    """Configuration of one constraint term: `func(env, **params) -> Tensor[N] | Tensor[N, J]`."""

    func: Callable[..., torch.Tensor] = MISSING
    """Term function. Built-in ones live in :mod:`.constraints` and are fused on the GPU."""

    max_p: float = MISSING
    """Upper bound of the termination probability of this term (mutated by the curriculum)."""


ConstraintTerm = ConstraintTermCfg
