"""Velocity command with dead zone (reference `U/mdp/commands.py:19-100`).

`UniformVelocityCommandWithDeadzone._update_command` post-processes `vel_command_b [N,3]` every env step: yaw rate from
the heading error for heading envs, small commands to zero, Bernoulli resampling (p = 0.01 when the command is still,
physics_dt / episode_length_s otherwise) with uniform redraws, and a Bernoulli yaw-rate flip.  The reference does this
with ~25 eager ops, two `torch.bernoulli` draws and two `nonzero()` host syncs; here it is ONE launch
(`catb200_command_update`), one thread per env, random numbers from the device-side Philox stream -- no host sync, so the
call can sit in the CUDA graph of the env step.

The Isaac Lab base class (`isaaclab_tasks...mdp.UniformVelocityCommand`) is third party and not installed here: with Isaac
Lab present the class below subclasses it (the task cfg's `class_type` keeps working); without it the functional form
`update_velocity_command` is what the tests drive.
"""

from __future__ import annotations

import torch

from .. import _lib as L
from .. import ops

try:  # pragma: no cover - needs Isaac Lab
    import isaaclab_tasks.manager_based.locomotion.velocity.mdp as _mdp
    from isaaclab.utils import configclass

    _Base, _BaseCfg = _mdp.UniformVelocityCommand, _mdp.UniformVelocityCommandCfg
    HAVE_ISAACLAB_MDP = True
except Exception:  # noqa: BLE001
    from .._isaaclab_compat import configclass

    _Base, _BaseCfg = object, object
    HAVE_ISAACLAB_MDP = False


def make_command_cfg(ranges, velocity_deadzone, heading_command, heading_control_stiffness, rel_heading_envs, rel_standing_envs,
                     physics_dt, max_episode_length_s) -> L.CommandCfg:  # fmt: skip
    """POD the kernel takes by value; `ranges` has lin_vel_x / lin_vel_y / ang_vel_z (/ heading) pairs like cfg.ranges."""
    c = L.CommandCfg()
    c.lin_vel_x[:] = ranges.lin_vel_x
    c.lin_vel_y[:] = ranges.lin_vel_y
    c.ang_vel_z[:] = ranges.ang_vel_z
    heading = getattr(ranges, "heading", None)
    c.heading[:] = heading if heading is not None else (0.0, 0.0)
    c.velocity_deadzone = velocity_deadzone
    c.heading_control_stiffness = heading_control_stiffness
    c.rel_heading_envs, c.rel_standing_envs = rel_heading_envs, rel_standing_envs
    c.p_step = L.f32(physics_dt / max_episode_length_s)  # commands.py:71-73,81-83 (python double, then fp32)
    c.heading_command = int(bool(heading_command))
    return c


def update_velocity_command(cfg: L.CommandCfg, vel_command_b, heading_target, heading_w, is_heading_env, is_standing_env,
                            rng_state=None, uniforms=None, resampled=None):  # fmt: skip
    """In-place `_update_command` (commands.py:39-93) on `vel_command_b [N,3]`.  Random numbers: `rng_state` (device
    Philox, advanced by 8 N) or `uniforms [N,8]` (tests).  Returns the resample mask (bool [N])."""
    L.require_cuda(vel_command_b, "vel_command_b")
    n = vel_command_b.shape[0]
    if vel_command_b.dtype != torch.float32 or not vel_command_b.is_contiguous() or vel_command_b.shape[1] != 3:
        raise TypeError("vel_command_b must be a contiguous float32 [N, 3] tensor")
    for t, name in ((is_heading_env, "is_heading_env"), (is_standing_env, "is_standing_env")):
        if t is not None and (t.dtype not in (torch.bool, torch.uint8) or not t.is_contiguous() or t.numel() != n):
            raise TypeError(f"{name} must be a contiguous bool tensor [N]")
    if resampled is None:
        resampled = torch.empty(n, dtype=torch.bool, device=vel_command_b.device)
    L.check(
        L.load().catb200_command_update(
            cfg, n, vel_command_b.data_ptr(), L.ptr(heading_target), L.ptr(heading_w), L.ptr(is_heading_env),
            L.ptr(is_standing_env), L.ptr(uniforms), L.ptr(rng_state), resampled.data_ptr(), L.stream(),
        ),
        "command_update",
    )  # fmt: skip
    return resampled


class UniformVelocityCommandWithDeadzone(_Base):
    """Drop-in for the reference class of the same name (commands.py:19-93)."""

    def __init__(self, cfg, env):
        if not HAVE_ISAACLAB_MDP:
            raise ImportError("UniformVelocityCommandWithDeadzone needs Isaac Lab's UniformVelocityCommand; use update_velocity_command() directly")
        super().__init__(cfg, env)  # pragma: no cover
        self.velocity_deadzone = cfg.velocity_deadzone  # pragma: no cover
        self.dt = env.physics_dt  # pragma: no cover
        self.max_episode_length_s = env.max_episode_length_s  # pragma: no cover
        self._kcfg = make_command_cfg(cfg.ranges, cfg.velocity_deadzone, cfg.heading_command, getattr(cfg, "heading_control_stiffness", 0.0),
                                      cfg.rel_heading_envs, cfg.rel_standing_envs, self.dt, self.max_episode_length_s)  # pragma: no cover
        self._rng = ops.make_rng_state(getattr(env.cfg, "seed", 0) or 0, self.device)  # pragma: no cover

    def _update_command(self):  # pragma: no cover - needs Isaac Lab
        heading = self.cfg.heading_command
        mask = update_velocity_command(
            self._kcfg, self.vel_command_b, self.heading_target if heading else None, self.robot.data.heading_w if heading else None,
            self.is_heading_env if heading else None, self.is_standing_env, rng_state=self._rng,
        )  # fmt: skip
        # CommandTerm._resample's bookkeeping for the resampled envs (Isaac Lab: time_left redraw, command_counter += 1)
        lo, hi = self.cfg.resampling_time_range
        self.time_left = torch.where(mask, lo + (hi - lo) * torch.rand_like(self.time_left), self.time_left)
        self.command_counter += mask.to(self.command_counter.dtype)


if HAVE_ISAACLAB_MDP:  # pragma: no cover

    @configclass
    class UniformVelocityCommandWithDeadzoneCfg(_BaseCfg):
        class_type: type = UniformVelocityCommandWithDeadzone
        velocity_deadzone: float = 0.1

else:

    @configclass
    class UniformVelocityCommandWithDeadzoneCfg:
        """Fields of the reference cfg the kernel needs (commands.py:96-100 + the Isaac Lab base cfg)."""

        class_type: type = UniformVelocityCommandWithDeadzone
        velocity_deadzone: float = 0.1
        heading_command: bool = False
        heading_control_stiffness: float = 1.0
        rel_standing_envs: float = 0.0
        rel_heading_envs: float = 1.0
