"""ORACLE (test infrastructure, never shipped or timed as the product).

CPU restatement, in eager fp32 torch ops, of the per-step CaT path of
Gepetto/constraints-as-terminations.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs may import this module.

Every function cites the reference lines it follows (paths relative to the
reference root, `U/` = `exts/cat_envs/cat_envs/tasks/utils/`).  The restatement
is pinned against the reference itself: `oracle/make_golden.py` imports the
real reference modules (under import stubs for Isaac Lab) in the build
container, runs them on seeded synthetic Solo12 state and commits inputs +
outputs under `tests/golden/`; `tests/test_oracle_golden.py` replays them here
bit-for-bit.  The reference ships no tests or golden vectors of its own
(SURVEY.md §4), so those generated fixtures are the pin.
"""

from __future__ import annotations

import torch


# ------------------------------------------------------------------------------------------------
# term functions  (U/cat/constraints.py)
# ------------------------------------------------------------------------------------------------
def _robot(env, asset_cfg):
    return env.scene[asset_cfg.name].data


def _history_force_peak(env, asset_cfg):
    """max over history of ||F|| per selected body -> [N, B]  (constraints.py:102-107,151-158,207-209)."""
    f = env.scene[asset_cfg.name].data.net_forces_w_history
    return torch.max(torch.norm(f[:, :, asset_cfg.body_ids], dim=-1), dim=1)[0]


def _command(env):
    return env.command_manager.get_command("base_velocity")


def term_joint_position(env, limit, asset_cfg):  # constraints.py:23-31
    return torch.abs(_robot(env, asset_cfg).joint_pos[:, asset_cfg.joint_ids]) - limit


def term_joint_position_when_moving_forward(env, limit, velocity_deadzone, asset_cfg):  # constraints.py:34-54
    d = _robot(env, asset_cfg)
    c = torch.abs(d.joint_pos[:, asset_cfg.joint_ids] - d.default_joint_pos[:, asset_cfg.joint_ids]) - limit
    gate = (torch.abs(_command(env)[:, 1]) < velocity_deadzone).float().unsqueeze(1)
    return c * gate


def term_joint_torque(env, limit, asset_cfg):  # constraints.py:57-65
    return torch.abs(_robot(env, asset_cfg).applied_torque[:, asset_cfg.joint_ids]) - limit


def term_joint_velocity(env, limit, asset_cfg):  # constraints.py:68-75
    return torch.abs(_robot(env, asset_cfg).joint_vel[:, asset_cfg.joint_ids]) - limit


def term_joint_acceleration(env, limit, asset_cfg):  # constraints.py:78-85
    return torch.abs(_robot(env, asset_cfg).joint_acc[:, asset_cfg.joint_ids]) - limit


def term_upsidedown(env, limit, asset_cfg):  # constraints.py:88-94  (bool)
    return _robot(env, asset_cfg).projected_gravity_b[:, 2] > limit


def term_contact(env, asset_cfg):  # constraints.py:97-110  (bool)
    return torch.any(_history_force_peak(env, asset_cfg) > 1.0, dim=1)


def term_base_orientation(env, limit, asset_cfg):  # constraints.py:113-119
    return torch.norm(_robot(env, asset_cfg).projected_gravity_b[:, :2], dim=1) - limit


def term_air_time(env, limit, velocity_deadzone, asset_cfg):  # constraints.py:122-141
    sensor = env.scene[asset_cfg.name]
    touchdown = sensor.compute_first_contact(env.step_dt)[:, asset_cfg.body_ids]
    last_air = sensor.data.last_air_time[:, asset_cfg.body_ids]
    moving = (torch.norm(_command(env)[:, :3], dim=1) > velocity_deadzone).float().unsqueeze(1)
    return (limit - last_air) * touchdown.float() * moving


def term_n_foot_contact(env, number_of_desired_feet, min_command_value, asset_cfg):  # constraints.py:144-168
    n_in_contact = (_history_force_peak(env, asset_cfg) > 1.0).sum(1)
    miss = torch.abs(n_in_contact - number_of_desired_feet)
    moving = (torch.norm(_command(env)[:, :3], dim=1) > min_command_value).float()
    return miss * moving


def term_joint_range(env, limit, asset_cfg):  # constraints.py:171-181
    d = _robot(env, asset_cfg)
    return torch.abs(d.joint_pos[:, asset_cfg.joint_ids] - d.default_joint_pos[:, asset_cfg.joint_ids]) - limit


def term_action_rate(env, limit, asset_cfg):  # constraints.py:184-198
    a = env.action_manager._action[:, asset_cfg.joint_ids]
    ap = env.action_manager._prev_action[:, asset_cfg.joint_ids]
    return torch.abs(a - ap) / env.step_dt - limit


def term_foot_contact_force(env, limit, asset_cfg):  # constraints.py:201-211
    return _history_force_peak(env, asset_cfg) - limit


def term_min_base_height(env, limit, asset_cfg):  # constraints.py:214-220
    return limit - env.scene[asset_cfg.name].data.root_pos_w[:, 2]


def term_no_move(env, velocity_deadzone, joint_vel_limit, asset_cfg):  # constraints.py:223-235
    qd = _robot(env, asset_cfg).joint_vel[:, asset_cfg.joint_ids]
    still = (torch.norm(_command(env)[:, :3], dim=1) < velocity_deadzone).float().unsqueeze(1)
    return (torch.abs(qd) - joint_vel_limit) * still


TERM_ORACLES = {
    "joint_position": term_joint_position,
    "joint_position_when_moving_forward": term_joint_position_when_moving_forward,
    "joint_torque": term_joint_torque,
    "joint_velocity": term_joint_velocity,
    "joint_acceleration": term_joint_acceleration,
    "upsidedown": term_upsidedown,
    "contact": term_contact,
    "base_orientation": term_base_orientation,
    "air_time": term_air_time,
    "n_foot_contact": term_n_foot_contact,
    "joint_range": term_joint_range,
    "action_rate": term_action_rate,
    "foot_contact_force": term_foot_contact_force,
    "min_base_height": term_min_base_height,
    "no_move": term_no_move,
}


# ------------------------------------------------------------------------------------------------
# CaT probability engine  (U/cat/constraint_manager.py:22-116)
# ------------------------------------------------------------------------------------------------
class CatOracle:
    """State = one Polyak running max row per term; everything else is recomputed each step."""

    def __init__(self, tau: float = 0.95, min_p: float = 0.0):
        self.tau = tau
        self.min_p = min_p
        self.running_max: dict[str, torch.Tensor] = {}
        self.probs: dict[str, torch.Tensor] = {}
        self.raw: dict[str, torch.Tensor] = {}

    def add(self, name: str, c: torch.Tensor, max_p: float) -> torch.Tensor:
        # constraint_manager.py:45-49 - bool/int -> float, [N] -> [N,1]
        if not torch.is_floating_point(c):
            c = c.float()
        if c.ndim == 1:
            c = c.unsqueeze(1)
        self.raw[name] = c
        # :55 - column max over all envs, clamped from below
        col_max = c.max(dim=0, keepdim=True)[0].clamp(min=1e-6)
        # :58-61 - first call assigns, later calls Polyak-average (two roundings, python-double (1-tau))
        if name in self.running_max:
            self.running_max[name] = self.running_max[name] * self.tau + (1.0 - self.tau) * col_max
        else:
            self.running_max[name] = col_max
        # :64-72 - probability only where the constraint is violated; the `mask.any()` guard is a no-op
        rm = self.running_max[name].expand_as(c)
        p_violating = self.min_p + torch.clamp(c / rm, 0.0, 1.0) * (max_p - self.min_p)
        probs = torch.where(c > 0.0, p_violating, torch.zeros_like(c))
        self.probs[name] = probs
        return probs

    def combined(self) -> torch.Tensor:  # :78-82
        if not self.probs:
            return torch.tensor([])
        return torch.cat(list(self.probs.values()), dim=1).max(1).values


class ManagerOracle:
    """`ConstraintManager.compute / reset` (U/cat/constraint_manager.py:190-229) over a list of terms.

    `terms` is a list of `(name, fn, params, max_p_getter)`; `max_p_getter()` is read on every
    compute like the reference re-reads `term_cfg.max_p` (:217).
    """

    def __init__(self, env, terms, tau: float = 0.95, min_p: float = 0.0):
        self.env = env
        self.terms = terms
        self.cat = CatOracle(tau, min_p)
        n = env.num_envs
        self.episode_sums = {t[0]: torch.zeros(n) for t in terms}
        self.mean_values = {t[0]: torch.zeros(n) for t in terms}

    def compute(self) -> torch.Tensor:
        for name, fn, params, max_p in self.terms:  # :216-217
            self.cat.add(name, fn(self.env, **params), max_p() if callable(max_p) else max_p)
        cstr_prob = self.cat.combined()  # :220
        for name, *_ in self.terms:  # :223-227
            term_max = self.cat.probs[name].max(1).values
            self.episode_sums[name] += term_max.gt(0.0).float()
            self.mean_values[name] += term_max
        return cstr_prob

    def reset(self, env_ids=None) -> dict[str, torch.Tensor]:  # :190-211
        ids = slice(None) if env_ids is None else env_ids
        out = {}
        for name in self.episode_sums:
            length = self.env.episode_length_buf[ids]
            out[f"Episode_Constraint_violation/{name}"] = (self.episode_sums[name][ids] / length).mean() * 100
            out[f"Episode_Constraint_probability/{name}"] = (self.mean_values[name][ids] / length).mean()
            self.episode_sums[name][ids] = 0.0
            self.mean_values[name][ids] = 0.0
        return out


def step_epilogue(raw_reward: torch.Tensor, cstr_prob: torch.Tensor, reset_buf: torch.Tensor):
    """Reward scaling and float dones of `CaTEnv.step` (U/cat/cat_env.py:100-107,118-121)."""
    reward = torch.clip(raw_reward * (1.0 - cstr_prob), min=0.0, max=None)
    dones = cstr_prob.clone()
    ids = reset_buf.nonzero(as_tuple=False).squeeze(-1)
    if len(ids) > 0:
        dones[ids] = 1.0
    return reward, dones


def terms_from_cfg(cfg_items, resolve_scene=None):
    """Build the `(name, oracle_fn, params, max_p_getter)` list from a ConstraintsCfg-like mapping.

    The term function is looked up by `func.__name__`, so the same task cfg object drives the product
    (which dispatches on its own `constraints.*` functions) and this oracle.
    """
    items = cfg_items.items() if isinstance(cfg_items, dict) else cfg_items.__dict__.items()
    out = []
    for name, term in items:
        if term is None:
            continue
        fn = TERM_ORACLES[term.func.__name__]
        if resolve_scene is not None:
            for v in term.params.values():
                if hasattr(v, "resolve"):
                    v.resolve(resolve_scene)
        out.append((name, fn, term.params, (lambda t=term: t.max_p)))
    return out
