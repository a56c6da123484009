"""Constraint term functions of the CaT manager, backed by the fused sm_100a term kernel.

Same names, signatures and return shapes / dtypes as the reference term library
(`exts/cat_envs/cat_envs/tasks/utils/cat/constraints.py:23-235`), so a task cfg such as the Solo12
one (`.../solo12/cat_flat_env_cfg.py:259-355`) works unchanged.  The difference is *how* they run:

* inside `ConstraintManager.compute()` the functions are never called.  Each one carries a
  `fused_spec` describing its arithmetic as a `catb200_op` over named source tensors, and the manager
  evaluates all terms of all envs in one kernel straight from the simulator's state tensors.
* called directly (`joint_torque(env, limit=..., asset_cfg=...)`), a function evaluates just its own
  columns through the same kernel (`catb200_cat_eval_terms`) and returns a tensor like the reference.

Anything that is not one of these 15 functions (a user's own python term) still works: the manager
calls it and feeds the returned tensor to the kernel as a generic column block.
"""

from __future__ import annotations

from collections.abc import Callable
from dataclasses import dataclass, field

import inspect

import torch

from . import _lib as L


@dataclass
class SourceRef:
    """A tensor the term reads, identified by a stable key so that terms sharing it stage it once."""

    key: str
    fetch: Callable[[object], torch.Tensor]
    bodies: int = 0  # contact history only: number of bodies B of the [N, H, B, 3] tensor


@dataclass
class TermSpec:
    op: int
    src0: SourceRef
    ids: object = None  # list[int] | slice | None -> resolved against the source row length
    src1: SourceRef | None = None
    src2: SourceRef | None = None
    p0: float = 0.0
    p1: float = 0.0
    p2: float = 0.0
    single_column: bool = False  # op reduces over its ids to one column
    returns_bool: bool = False
    squeeze: bool = False  # reference returns [N] rather than [N, 1]
    extra: dict = field(default_factory=dict)


def _ids_list(ids, length: int) -> list[int]:
    if ids is None or (isinstance(ids, slice) and ids == slice(None)):
        return list(range(length))
    if isinstance(ids, slice):
        return list(range(length))[ids]
    if isinstance(ids, torch.Tensor):
        return [int(i) for i in ids.tolist()]
    return [int(i) for i in ids]


# ---- source tensors (attribute paths of the Isaac Lab env) -----------------------
def _robot_src(asset_cfg, attr: str) -> SourceRef:
    name = asset_cfg.name
    return SourceRef(f"{name}.data.{attr}", lambda env: getattr(env.scene[name].data, attr))


def _command_src() -> SourceRef:
    return SourceRef("command.base_velocity", lambda env: env.command_manager.get_command("base_velocity"))


def _forces_src(asset_cfg) -> SourceRef:
    name = asset_cfg.name
    return SourceRef(f"{name}.data.net_forces_w_history", lambda env: env.scene[name].data.net_forces_w_history, bodies=-1)


# ---- stand-alone evaluation of one term --------------------------------------------------------------
def _standalone(spec: TermSpec, env) -> torch.Tensor:
    from .constraint_manager import build_plan  # local import: the manager imports this module

    plan, tensors, n_cols = build_plan(env, [("term", spec, 0)])
    num_envs = tensors[0].shape[0]
    out = torch.empty((num_envs, n_cols), dtype=torch.float32, device=tensors[0].device)
    L.check(L.load().catb200_cat_eval_terms(plan, num_envs, out.data_ptr(), L.stream()), "cat_eval_terms")
    if spec.squeeze:
        out = out[:, 0]
    if spec.returns_bool:
        out = out > 0.5
    return out


def _term(spec_fn):
    """Decorator: turn a `fused_spec` builder into a reference-compatible term function."""

    def func(env, **params) -> torch.Tensor:
        return _standalone(spec_fn(env, **params), env)

    func.__name__ = spec_fn.__name__
    func.__qualname__ = spec_fn.__qualname__
    func.__doc__ = spec_fn.__doc__
    # Isaac Lab's ManagerBase._resolve_common_term_cfg checks `inspect.signature(func)` against the keys of
    # `term_cfg.params`: expose the reference's own signature (env, limit, asset_cfg, ...) -> Tensor, not (env, **params)
    func.__signature__ = inspect.signature(spec_fn).replace(return_annotation=torch.Tensor)
    func.fused_spec = spec_fn
    return func


# ---- the 15 term functions -----------------------------------------------------------------------------
@_term
def joint_position(env, limit: float, asset_cfg) -> TermSpec:
    """|q_j| - limit per selected joint (reference constraints.py:23-31)."""
    return TermSpec(L.OP_ABS_MINUS, _robot_src(asset_cfg, "joint_pos"), asset_cfg.joint_ids, p0=limit)


@_term
def joint_position_when_moving_forward(env, limit: float, velocity_deadzone: float, asset_cfg) -> TermSpec:
    """(|q_j - q0_j| - limit) * [|cmd_y| < deadzone] (reference constraints.py:34-54)."""
    return TermSpec(
        L.OP_ABSDIFF_MINUS_GATE_Y,
        _robot_src(asset_cfg, "joint_pos"),
        asset_cfg.joint_ids,
        src1=_robot_src(asset_cfg, "default_joint_pos"),
        src2=_command_src(),
        p0=limit,
        p1=velocity_deadzone,
    )


@_term
def joint_torque(env, limit: float, asset_cfg) -> TermSpec:
    """|tau_j| - limit (reference constraints.py:57-65)."""
    return TermSpec(L.OP_ABS_MINUS, _robot_src(asset_cfg, "applied_torque"), asset_cfg.joint_ids, p0=limit)


@_term
def joint_velocity(env, limit: float, asset_cfg) -> TermSpec:
    """|qd_j| - limit (reference constraints.py:68-75)."""
    return TermSpec(L.OP_ABS_MINUS, _robot_src(asset_cfg, "joint_vel"), asset_cfg.joint_ids, p0=limit)


@_term
def joint_acceleration(env, limit: float, asset_cfg) -> TermSpec:
    """|qdd_j| - limit (reference constraints.py:78-85)."""
    return TermSpec(L.OP_ABS_MINUS, _robot_src(asset_cfg, "joint_acc"), asset_cfg.joint_ids, p0=limit)


@_term
def upsidedown(env, limit: float, asset_cfg) -> TermSpec:
    """projected_gravity_b.z > limit, a bool per env (reference constraints.py:88-94)."""
    return TermSpec(
        L.OP_COMPONENT_GT, _robot_src(asset_cfg, "projected_gravity_b"), [2], p0=limit, returns_bool=True, squeeze=True
    )


@_term
def contact(env, asset_cfg) -> TermSpec:
    """any selected body whose contact force peaked above 1 N over the history (reference constraints.py:97-110)."""
    return TermSpec(
        L.OP_CONTACT_ANY,
        _forces_src(asset_cfg),
        asset_cfg.body_ids,
        p0=1.0,
        single_column=True,
        returns_bool=True,
        squeeze=True,
    )


@_term
def base_orientation(env, limit: float, asset_cfg) -> TermSpec:
    """||projected_gravity_b.xy|| - limit (reference constraints.py:113-119)."""
    return TermSpec(
        L.OP_NORM2_MINUS, _robot_src(asset_cfg, "projected_gravity_b"), [0, 1], p0=limit, single_column=True, squeeze=True
    )


@_term
def air_time(env, limit: float, velocity_deadzone: float, asset_cfg) -> TermSpec:
    """(limit - last_air_time) * first_contact * [||cmd|| > deadzone] per foot (reference constraints.py:122-141)."""
    name = asset_cfg.name
    return TermSpec(
        L.OP_AIR_TIME,
        SourceRef(f"{name}.data.last_air_time", lambda e: e.scene[name].data.last_air_time),
        asset_cfg.body_ids,
        # Isaac Lab's ContactSensor.compute_first_contact is third-party arithmetic: it stays an input
        src1=SourceRef(f"{name}.first_contact", lambda e: e.scene[name].compute_first_contact(e.step_dt)),
        src2=_command_src(),
        p0=limit,
        p1=velocity_deadzone,
    )


@_term
def n_foot_contact(env, number_of_desired_feet: int, min_command_value: float, asset_cfg) -> TermSpec:
    """|#feet in contact - desired| * [||cmd|| > min_command] (reference constraints.py:144-168)."""
    return TermSpec(
        L.OP_N_CONTACT,
        _forces_src(asset_cfg),
        asset_cfg.body_ids,
        src2=_command_src(),
        p0=float(number_of_desired_feet),
        p1=min_command_value,
        p2=1.0,
        single_column=True,
        squeeze=True,
    )


@_term
def joint_range(env, limit: float, asset_cfg) -> TermSpec:
    """|q_j - q0_j| - limit (reference constraints.py:171-181)."""
    return TermSpec(
        L.OP_ABSDIFF_MINUS,
        _robot_src(asset_cfg, "joint_pos"),
        asset_cfg.joint_ids,
        src1=_robot_src(asset_cfg, "default_joint_pos"),
        p0=limit,
    )


@_term
def action_rate(env, limit: float, asset_cfg) -> TermSpec:
    """|a - a_prev| / step_dt - limit (reference constraints.py:184-198)."""
    return TermSpec(
        L.OP_ACTION_RATE,
        SourceRef("action_manager._action", lambda e: e.action_manager._action),
        asset_cfg.joint_ids,
        src1=SourceRef("action_manager._prev_action", lambda e: e.action_manager._prev_action),
        p0=limit,
        p1=env.step_dt,
    )


@_term
def foot_contact_force(env, limit: float, asset_cfg) -> TermSpec:
    """max over history of ||F_b|| - limit per selected body (reference constraints.py:201-211)."""
    return TermSpec(L.OP_FORCE_PEAK_MINUS, _forces_src(asset_cfg), asset_cfg.body_ids, p0=limit)


@_term
def min_base_height(env, limit: float, asset_cfg) -> TermSpec:
    """limit - root height (reference constraints.py:214-220)."""
    return TermSpec(L.OP_LIMIT_MINUS, _robot_src(asset_cfg, "root_pos_w"), [2], p0=limit, squeeze=True)


@_term
def no_move(env, velocity_deadzone: float, joint_vel_limit: float, asset_cfg) -> TermSpec:
    """(|qd_j| - limit) * [||cmd|| < deadzone] (reference constraints.py:223-235)."""
    return TermSpec(
        L.OP_ABS_MINUS_GATE_STILL,
        _robot_src(asset_cfg, "joint_vel"),
        asset_cfg.joint_ids,
        src2=_command_src(),
        p0=joint_vel_limit,
        p1=velocity_deadzone,
    )


BUILTIN_TERMS = (
    joint_position,
    joint_position_when_moving_forward,
    joint_torque,
    joint_velocity,
    joint_acceleration,
    upsidedown,
    contact,
    base_orientation,
    air_time,
    n_foot_contact,
    joint_range,
    action_rate,
    foot_contact_force,
    min_base_height,
    no_move,
)
