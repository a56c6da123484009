"""Env-side MDP terms of the CaT tasks (reference `exts/cat_envs/cat_envs/tasks/utils/mdp/`): the velocity command with
dead zone / Bernoulli resampling / yaw flip and the random push event, as fused kernels on a device-side Philox stream."""

from .commands import UniformVelocityCommandWithDeadzone, UniformVelocityCommandWithDeadzoneCfg, update_velocity_command
from .events import push_by_setting_velocity_with_random_envs, select_pushes
from .observations import ObservationAssembler, ObsTermSpec, solo12_policy_terms

__all__ = [
    "UniformVelocityCommandWithDeadzone",
    "UniformVelocityCommandWithDeadzoneCfg",
    "update_velocity_command",
    "push_by_setting_velocity_with_random_envs",
    "select_pushes",
    "ObservationAssembler",
    "ObsTermSpec",
    "solo12_policy_terms",
]
