"""Command-line options shared by the CleanRL launchers.

Same flag names and override semantics as the reference launcher helpers (`scripts/clean_rl/cli_args.py:11-73`),
table-driven here."""

from __future__ import annotations

import argparse

# flag, kwargs, cfg attribute the flag overrides (None = handled separately)
_OPTIONS = [
    ("--experiment_name", dict(type=str, help="log folder under logs/clean_rl/"), "experiment_name"),
    ("--resume", dict(type=bool, help="resume from a checkpoint"), "resume"),
    ("--load_run", dict(type=str, help="run folder (regex) to load from"), "load_run"),
    ("--checkpoint", dict(type=str, help="checkpoint file (regex) to load"), "load_checkpoint"),
    ("--logger", dict(type=str, choices=("wandb", "tensorboard"), help="scalar logger"), "logger"),
    ("--log_project_name", dict(type=str, help="wandb project name"), None),
]


def add_clean_rl_args(parser: argparse.ArgumentParser) -> None:
    group = parser.add_argument_group("clean_rl", description="CleanRL agent options")
    for flag, kwargs, _ in _OPTIONS:
        group.add_argument(flag, default=None, **kwargs)


def update_clean_rl_cfg(agent_cfg, args_cli: argparse.Namespace):
    """CLI values that were given win over the cfg defaults."""
    seed = getattr(args_cli, "seed", None)
    if seed is not None:
        agent_cfg.seed = seed
    for flag, _, attr in _OPTIONS:
        value = getattr(args_cli, flag.lstrip("-"), None)
        if attr is not None and value is not None:
            setattr(agent_cfg, attr, value)
    if agent_cfg.logger == "wandb" and getattr(args_cli, "log_project_name", None):
        agent_cfg.wandb_project = args_cli.log_project_name
    return agent_cfg


def parse_clean_rl_cfg(task_name: str, args_cli: argparse.Namespace):
    """Default agent cfg of the task (gym registry under Isaac Lab, Solo12 values otherwise) + CLI overrides."""
    try:
        from isaaclab_tasks.utils.parse_cfg import load_cfg_from_registry  # type: ignore

        cfg = load_cfg_from_registry(task_name, "clean_rl_cfg_entry_point")
    except ImportError:
        from constraints_as_terminations_b200 import solo12_flat_ppo_cfg

        cfg = solo12_flat_ppo_cfg()
    return update_clean_rl_cfg(cfg, args_cli)
