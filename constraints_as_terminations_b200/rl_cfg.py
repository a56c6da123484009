"""PPO hyper-parameter config, field-for-field the reference's `CleanRlPpoActorCriticCfg`
(`exts/cat_envs/cat_envs/tasks/utils/cleanrl/rl_cfg.py:13-38`) and the Solo12 values
(`.../solo12/agents/clean_rl_ppo_cfg.py:12-34`)."""

from __future__ import annotations

from dataclasses import MISSING
from typing import Literal

from ._isaaclab_compat import configclass


@configclass
class CleanRlPpoActorCriticCfg:
    seed: int = 42

    save_interval: int = MISSING

    learning_rate: float = MISSING
    num_steps: int = MISSING
    num_iterations: int = MISSING
    gamma: float = MISSING
    gae_lambda: float = MISSING
    updates_epochs: int = MISSING
    minibatch_size: int = MISSING
    clip_coef: float = MISSING
    ent_coef: float = MISSING
    vf_coef: float = MISSING
    max_grad_norm: float = MISSING
    norm_adv: bool = MISSING
    clip_vloss: bool = MISSING
    anneal_lr: bool = MISSING

    experiment_name: str = MISSING
    logger: Literal["tensorboard", "wandb"] | None = "tensorboard"
    wandb_project: str = MISSING

    load_run: str = MISSING
    load_checkpoint: str = MISSING


def solo12_flat_ppo_cfg(**overrides) -> CleanRlPpoActorCriticCfg:
    """`Solo12FlatPPORunnerCfg` of the reference (clean_rl_ppo_cfg.py:12-34)."""
    values = dict(
        save_interval=50, learning_rate=3.0e-4, num_steps=24, num_iterations=2000, gamma=0.99, gae_lambda=0.95,
        updates_epochs=5, minibatch_size=16384, clip_coef=0.2, ent_coef=0.001, vf_coef=2.0, max_grad_norm=1.0,
        norm_adv=True, clip_vloss=True, anneal_lr=True, experiment_name="solo12_flat", logger="tensorboard",
        wandb_project="solo12_flat", load_run=".*", load_checkpoint="model_.*.pt",
    )  # fmt: skip
    values.update(overrides)
    return CleanRlPpoActorCriticCfg(**values)
