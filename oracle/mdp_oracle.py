"""ORACLE (test infrastructure): CPU restatement of the env-side MDP terms of the reference, with the random numbers made
explicit so that any generator can drive it: `torch.bernoulli(p)` becomes `u < p` and `uniform_(a, b)` becomes
`a + (b - a) * u` for uniforms u supplied by the caller.

  update_command : UniformVelocityCommandWithDeadzone._update_command, U/mdp/commands.py:39-93, including the body of Isaac
                   Lab's UniformVelocityCommand._resample_command (third party, absent from /root/reference: README pins
                   Isaac Lab 2.1.0; its published implementation draws lin_x, lin_y, ang_z, heading uniformly in the cfg
                   ranges and is_heading / is_standing as `uniform <= rel_*_envs`) -> parity there is UNPINNED, anchored on
                   the reference's call site commands.py:77-78.
  select_pushes  : push_by_setting_velocity_with_random_envs, U/mdp/events.py:59-96.
  assemble_obs   : the policy observation group of S12/cat_flat_env_cfg.py:137-172 as Isaac Lab's ObservationManager
                   computes it (third party: per term gather -> AdditiveUniformNoise `data + rand * (n_max - n_min) + n_min`
                   -> scale -> concatenate; parity UNPINNED there, anchored on the cfg's constants).
Only tests/ may import this module.
"""

from __future__ import annotations

import math

import torch


def wrap_to_pi(angles: torch.Tensor) -> torch.Tensor:
    """isaaclab.utils.math.wrap_to_pi (third party): [-pi, pi], +pi kept for positive odd multiples of pi."""
    wrapped = (angles + math.pi) % (2 * math.pi)
    return torch.where((wrapped == 0) & (angles > 0), torch.tensor(math.pi), wrapped - math.pi)


def update_command(cmd, heading_target, heading_w, is_heading, is_standing, u, *, ranges, deadzone, heading_command, stiffness,
                   rel_heading, rel_standing, physics_dt, max_episode_length_s):  # fmt: skip
    """u: [N, 8] uniforms (0 resample, 1-3 lin_x / lin_y / ang_z, 4 heading, 5 is_heading, 6 is_standing, 7 yaw flip).
    Returns new (cmd, heading_target, is_heading, is_standing, resampled)."""
    cmd, heading_target = cmd.clone(), heading_target.clone()
    is_heading, is_standing = is_heading.clone(), is_standing.clone()
    if heading_command:  # commands.py:46-58
        ids = is_heading.nonzero(as_tuple=False).flatten()
        err = wrap_to_pi(heading_target[ids] - heading_w[ids])
        cmd[ids, 2] = torch.clip(stiffness * err, min=ranges["ang_vel_z"][0], max=ranges["ang_vel_z"][1])
    cmd *= torch.any(torch.abs(cmd[:, :3]) > deadzone, dim=1).unsqueeze(1)  # :60-65
    no_vel = (torch.norm(cmd[:, :3], dim=1) < deadzone).float()  # :68-70
    p_step = torch.tensor(physics_dt / max_episode_length_s, dtype=torch.float32)
    p = 0.01 * no_vel + p_step * (1 - no_vel)  # :71-73
    res = u[:, 0] < p  # torch.bernoulli(p) :74-76
    ids = res.nonzero(as_tuple=False).flatten()
    if len(ids) > 0:  # :77-78 -> UniformVelocityCommand._resample_command
        for col, key in ((0, "lin_vel_x"), (1, "lin_vel_y"), (2, "ang_vel_z")):
            lo, hi = (torch.tensor(v, dtype=torch.float32) for v in ranges[key])
            cmd[ids, col] = lo + (hi - lo) * u[ids, 1 + col]
        if heading_command:
            lo, hi = (torch.tensor(v, dtype=torch.float32) for v in ranges["heading"])
            heading_target[ids] = lo + (hi - lo) * u[ids, 4]
            is_heading[ids] = u[ids, 5] <= rel_heading
        is_standing[ids] = u[ids, 6] <= rel_standing
    flip = (u[:, 7] < p_step).float()  # torch.bernoulli(full_like(., p_ang_vel)) :81-93
    cmd[:, 2] *= 1 - 2 * flip
    return cmd, heading_target, is_heading, is_standing, res


def select_pushes(root_vel_w, u, *, physics_dt, max_episode_length_s, velocity_range):
    """u: [N, 7] uniforms (0 Bernoulli, 1-6 the six velocity components).  Returns (pushed mask, new root velocities)."""
    p_push = torch.tensor(physics_dt / (max_episode_length_s * 2), dtype=torch.float32)  # events.py:67-69
    pushed = u[:, 0] < p_push  # :73-77
    keys = ["x", "y", "z", "roll", "pitch", "yaw"]
    ranges = torch.tensor([velocity_range.get(k, (0.0, 0.0)) for k in keys], dtype=torch.float32)  # :86-90
    vel = root_vel_w.clone()
    sample = ranges[:, 0] + (ranges[:, 1] - ranges[:, 0]) * u[:, 1:7]  # sample_uniform :91-93
    vel[pushed] = sample[pushed]
    return pushed, vel


def assemble_obs(sources, terms, u):
    """sources: list of [N, D] tensors; terms: list of dicts {ids, noise (lo, hi) | None, scale}; u: [N, n_cols] uniforms."""
    cols, c = [], 0
    for src, t in zip(sources, terms):
        ids = list(range(src.shape[1])) if t["ids"] is None else list(t["ids"])
        data = src[:, ids]
        if t["noise"] is not None:
            lo, hi = t["noise"]
            data = data + u[:, c : c + len(ids)] * (hi - lo) + lo  # AdditiveUniformNoiseCfg / uniform_noise
        scale = t["scale"]
        data = data * (torch.tensor(scale, dtype=torch.float32) if isinstance(scale, (tuple, list)) else scale)
        cols.append(data)
        c += len(ids)
    return torch.cat(cols, dim=1)
