"""Shared pieces of the launchers: env construction (Isaac Lab if importable, synthetic Solo12 otherwise),
log-directory layout and checkpoint lookup (reference `scripts/clean_rl/train.py:116-131`, `play.py:79-88`)."""

from __future__ import annotations

import os
import pickle
import re
from datetime import datetime

TASK = "Isaac-Velocity-CaT-Flat-Solo12-v0"


def isaaclab_available() -> bool:
    try:
        import isaaclab  # noqa: F401

        return True
    except ImportError:
        return False


def make_env(task: str, num_envs: int | None, seed: int, device: str):
    """gym.make(task, cfg=env_cfg) under Isaac Lab; the synthetic Solo12 state source otherwise."""
    if isaaclab_available():  # pragma: no cover - needs Isaac Sim
        import gymnasium as gym
        from isaaclab_tasks.utils import parse_env_cfg  # type: ignore

        env_cfg = parse_env_cfg(task, device=device, num_envs=num_envs)
        env_cfg.seed = seed
        return gym.make(task, cfg=env_cfg), env_cfg
    from constraints_as_terminations_b200 import synthetic_env as se

    print(f"[INFO] Isaac Lab not importable: running task '{task}' on the synthetic Solo12 state source (no physics).")
    env = se.SyntheticSolo12Env(num_envs or 4096, device=device, seed=seed, pool=8, constraints_cfg=se.solo12_constraints_cfg())
    env.load_managers()
    print("[INFO] Constraint Manager: ", env.constraint_manager)
    return env, {"task": task, "num_envs": env.num_envs, "seed": seed, "synthetic": True}


def new_log_dir(experiment_name: str) -> str:
    root = os.path.abspath(os.path.join("logs", "clean_rl", experiment_name))
    print(f"[INFO] Logging experiment in directory: {root}")
    return os.path.join(root, datetime.now().strftime("%Y-%m-%d_%H-%M-%S"))


def dump_params(log_dir: str, env_cfg, agent_cfg) -> None:
    """params/{env,agent}.{yaml,pkl} like the reference (train.py:124-127)."""
    os.makedirs(os.path.join(log_dir, "params"), exist_ok=True)
    import yaml

    def as_dict(c):
        return c.to_dict() if hasattr(c, "to_dict") else dict(c) if isinstance(c, dict) else {"repr": repr(c)}

    for name, cfg in (("env", env_cfg), ("agent", agent_cfg)):
        with open(os.path.join(log_dir, "params", f"{name}.yaml"), "w") as f:
            yaml.safe_dump(as_dict(cfg), f, default_flow_style=False)
        with open(os.path.join(log_dir, "params", f"{name}.pkl"), "wb") as f:
            pickle.dump(as_dict(cfg), f)


def get_checkpoint_path(log_root: str, run_dir: str = ".*", checkpoint: str = "model_.*.pt") -> str:
    """Latest run matching `run_dir`, latest checkpoint matching `checkpoint` (isaaclab_tasks' helper semantics)."""
    runs = sorted(d for d in os.listdir(log_root) if os.path.isdir(os.path.join(log_root, d)) and re.match(run_dir, d))
    if not runs:
        raise ValueError(f"No runs present in the directory: '{log_root}' match: '{run_dir}'.")
    run_path = os.path.join(log_root, runs[-1])
    files = [f for f in os.listdir(run_path) if re.match(checkpoint, f)]
    if not files:
        raise ValueError(f"No checkpoints in the directory: '{run_path}' match '{checkpoint}'.")
    files.sort(key=lambda m: f"{m:0>15}")
    return os.path.join(run_path, files[-1])
