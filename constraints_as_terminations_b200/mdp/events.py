"""Random push event (reference `U/mdp/events.py:59-96`, `push_by_setting_velocity_with_random_envs`).

Every env step each env is pushed with probability p = physics_dt / (2 * max_episode_length_s): its root velocity is
replaced by a uniform draw in `velocity_range`.  The reference draws a Bernoulli vector, `nonzero()`s it (host sync),
gathers, samples and scatters; here ONE launch (`catb200_push_select`) writes the mask and the new velocities for all
envs (unpushed envs keep theirs), so the write into the simulator needs no index list and no sync.
"""

from __future__ import annotations

import torch

from .. import _lib as L

_KEYS = ("x", "y", "z", "roll", "pitch", "yaw")


def select_pushes(root_vel_w: torch.Tensor, p_push: float, velocity_range: dict, rng_state=None, uniforms=None):
    """-> (pushed bool [N], velocities [N,6]: uniform draws for the pushed envs, `root_vel_w` unchanged elsewhere)."""
    L.require_cuda(root_vel_w, "root_vel_w")
    n = root_vel_w.shape[0]
    ranges = torch.tensor([velocity_range.get(k, (0.0, 0.0)) for k in _KEYS], dtype=torch.float32)  # events.py:86-90
    lo, hi = ranges[:, 0].contiguous().to(root_vel_w.device), ranges[:, 1].contiguous().to(root_vel_w.device)
    vel = root_vel_w.to(torch.float32).contiguous().clone()
    pushed = torch.empty(n, dtype=torch.bool, device=root_vel_w.device)
    L.check(
        L.load().catb200_push_select(n, L.f32(p_push), lo.data_ptr(), hi.data_ptr(), vel.data_ptr(), L.ptr(uniforms), L.ptr(rng_state),
                                     pushed.data_ptr(), L.stream()),
        "push_select",
    )  # fmt: skip
    return pushed, vel


def push_by_setting_velocity_with_random_envs(env, env_ids, velocity_range, asset_cfg=None, rng_state=None):
    """Event term with the reference's signature (events.py:59-64); `env_ids` is ignored there too (all envs draw)."""
    name = asset_cfg.name if asset_cfg is not None else "robot"
    asset = env.scene[name]
    p_push = env.physics_dt / (env.max_episode_length_s * 2)  # events.py:67-69
    if rng_state is None:
        rng_state = getattr(env, "_catb200_push_rng", None)
        if rng_state is None:
            from .. import ops

            rng_state = env._catb200_push_rng = ops.make_rng_state(int(getattr(getattr(env, "cfg", None), "seed", 0) or 0) + 17, env.device)
    pushed, vel = select_pushes(asset.data.root_vel_w, p_push, velocity_range, rng_state=rng_state)
    asset.write_root_velocity_to_sim(vel)  # all envs: the unpushed ones receive the velocity they already have
    return pushed
