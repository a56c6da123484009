"""PPO hyper-parameter config of the trainer.

Field names, types and defaults mirror the reference's `CleanRlPpoActorCriticCfg`
(`exts/cat_envs/cat_envs/tasks/utils/cleanrl/rl_cfg.py:13-38`) and `solo12_flat_ppo_cfg()` carries the
Solo12 values (`.../solo12/agents/clean_rl_ppo_cfg.py:12-34`), so a cfg object of either code base can drive
either trainer.  The class is assembled from a field table instead of being spelled out.
"""

from __future__ import annotations

import dataclasses
from typing import Any

_REQUIRED = dataclasses.MISSING

# (name, type, default); `_REQUIRED` = must be given, as in the reference (dataclasses.MISSING there too)
_FIELDS: list[tuple[str, Any, Any]] = [
    ("seed", int, 42),
    ("save_interval", int, _REQUIRED),
    # optimisation
    *[(n, float, _REQUIRED) for n in ("learning_rate", "gamma", "gae_lambda", "clip_coef", "ent_coef", "vf_coef", "max_grad_norm")],
    *[(n, int, _REQUIRED) for n in ("num_steps", "num_iterations", "updates_epochs", "minibatch_size")],
    *[(n, bool, _REQUIRED) for n in ("norm_adv", "clip_vloss", "anneal_lr")],
    # bookkeeping
    ("experiment_name", str, _REQUIRED),
    ("logger", Any, "tensorboard"),  # "tensorboard" | "wandb" | None (None = no logging, an extension)
    ("wandb_project", str, _REQUIRED),
    ("load_run", str, _REQUIRED),
    ("load_checkpoint", str, _REQUIRED),
]


def _to_dict(self) -> dict:
    return dataclasses.asdict(self)


CleanRlPpoActorCriticCfg = dataclasses.make_dataclass(
    "CleanRlPpoActorCriticCfg",
    [(n, t) if d is _REQUIRED else (n, t, dataclasses.field(default=d)) for n, t, d in _FIELDS],
    kw_only=True,
    namespace={"to_dict": _to_dict, "__doc__": "PPO hyper-parameters (see module docstring)."},
)

SOLO12_FLAT_VALUES = {
    "save_interval": 50,
    "learning_rate": 3.0e-4, "num_steps": 24, "num_iterations": 2000, "gamma": 0.99, "gae_lambda": 0.95,
    "updates_epochs": 5, "minibatch_size": 16384, "clip_coef": 0.2, "ent_coef": 0.001, "vf_coef": 2.0,
    "max_grad_norm": 1.0, "norm_adv": True, "clip_vloss": True, "anneal_lr": True,
    "experiment_name": "solo12_flat", "logger": "tensorboard", "wandb_project": "solo12_flat",
    "load_run": ".*", "load_checkpoint": "model_.*.pt",
}  # fmt: skip


def solo12_flat_ppo_cfg(**overrides):
    """The Solo12 flat-terrain PPO config (`Solo12FlatPPORunnerCfg` in the reference)."""
    return CleanRlPpoActorCriticCfg(**{**SOLO12_FLAT_VALUES, **overrides})
