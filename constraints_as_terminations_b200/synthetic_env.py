"""Synthetic Solo12 state source with the attribute paths of the Isaac Lab env.

Isaac Sim / Isaac Lab are not installed where this repo is built and measured,
so the trainer-side hot path is driven by this stand-in.  It exposes *exactly*
the tensors the reference's term functions, manager, env step and trainer read
(reference `constraints.py:23-235`, `constraint_manager.py:128,142,196`,
`cat_env.py:92-147`, `cleanrl/ppo.py:158-161,186,215-226`) with the same names,
shapes and dtypes, filled from seeded distributions instead of PhysX.

It is a test / bench fixture, not part of the accelerated path: the only work it
does per step is swapping pre-generated state tensors in and calling the
constraint manager + the fused reward/dones epilogue, as `CaTEnv.step` does.
"""

from __future__ import annotations

import math
from types import SimpleNamespace

import torch

from .curriculums import modify_constraint_p

# URDF link order of solo12_mpi.urdf (reference assets/.../solo12_mpi.urdf:14-705); any fixed
# convention is valid because the manager only ever sees resolved integer ids.
BODY_NAMES = ["base_link"] + [
    f"{leg}_{part}" for leg in ("FL", "FR", "HL", "HR") for part in ("SHOULDER", "UPPER_LEG", "LOWER_LEG", "FOOT")
]
# joint order of the action cfg (reference cat_flat_env_cfg.py:116-129)
JOINT_NAMES = [f"{leg}_{j}" for leg in ("FL", "FR", "HR", "HL") for j in ("HAA", "HFE", "KFE")]
DEFAULT_JOINT_POS = [0.05, 0.4, -0.8, -0.05, 0.4, -0.8, -0.05, -0.4, 0.8, 0.05, -0.4, 0.8]
NUM_BODIES = len(BODY_NAMES)
NUM_JOINTS = len(JOINT_NAMES)
HISTORY = 3
OBS_DIM = 45
ACT_DIM = 12

STATE_FIELDS = (
    "joint_pos",
    "joint_vel",
    "joint_acc",
    "applied_torque",
    "projected_gravity_b",
    "root_pos_w",
    "net_forces_w_history",
    "last_air_time",
    "first_contact",
    "command",
    "action",
    "prev_action",
    "obs",
    "raw_reward",
)


def sample_state(num_envs: int, generator: torch.Generator, adversarial: bool = False) -> dict[str, torch.Tensor]:
    """One synthetic Solo12 state (CPU tensors), distributions per SURVEY.md §8(d) cfg 1."""
    g = generator
    n = num_envs

    def randn(*shape):
        return torch.randn(*shape, generator=g, dtype=torch.float32)

    def rand(*shape):
        return torch.rand(*shape, generator=g, dtype=torch.float32)

    q0 = torch.tensor(DEFAULT_JOINT_POS, dtype=torch.float32)
    s = {}
    s["joint_pos"] = q0 + 0.5 * randn(n, NUM_JOINTS)
    s["joint_vel"] = 8.0 * randn(n, NUM_JOINTS)
    s["joint_acc"] = 400.0 * randn(n, NUM_JOINTS)
    s["applied_torque"] = 2.0 * randn(n, NUM_JOINTS)
    grav = randn(n, 3) * 0.35 + torch.tensor([0.0, 0.0, -1.0])
    s["projected_gravity_b"] = grav / grav.norm(dim=1, keepdim=True)
    s["root_pos_w"] = torch.cat([randn(n, 2), 0.25 + 0.08 * randn(n, 1)], dim=1)
    forces = 20.0 * randn(n, HISTORY, NUM_BODIES, 3)
    forces = forces * (rand(n, HISTORY, NUM_BODIES, 1) < 0.3).float()
    s["net_forces_w_history"] = forces
    s["last_air_time"] = 0.5 * rand(n, NUM_BODIES)
    s["first_contact"] = rand(n, NUM_BODIES) < 0.1
    lo = torch.tensor([-0.3, -0.7, -0.78])
    hi = torch.tensor([1.0, 0.7, 0.78])
    cmd = lo + (hi - lo) * rand(n, 3)
    cmd = cmd * (rand(n, 1) >= 0.02).float()  # 2 % standing-still commands
    s["command"] = cmd
    s["action"] = randn(n, ACT_DIM)
    s["prev_action"] = randn(n, ACT_DIM)
    s["obs"] = randn(n, OBS_DIM) * torch.linspace(0.2, 3.0, OBS_DIM) + torch.linspace(-1.0, 1.0, OBS_DIM)
    s["raw_reward"] = 0.02 * rand(n) * 3.0
    if adversarial and n >= 8:
        # rows that exercise the clamp / exact-zero / all-negative paths of CaT.add
        s["joint_vel"][0] = 16.0  # |qd| - 16 == 0 exactly: not a violation
        s["joint_vel"][1] = 0.0
        s["applied_torque"][:, 3] = 0.5  # a column nobody violates (running max decays to the 1e-6 clamp)
        s["joint_acc"][2] = 1.0e6  # huge outlier drives the column max
        s["command"][3] = 0.0
        s["net_forces_w_history"][4] = 0.0
        s["projected_gravity_b"][5] = torch.tensor([0.0, 0.0, 1.0])  # upside down
        s["first_contact"][6] = True
        s["last_air_time"][6] = 0.0
    return s


class _Articulation:
    def __init__(self):
        self.data = SimpleNamespace()
        self.joint_names = list(JOINT_NAMES)
        self.body_names = list(BODY_NAMES)


class _ContactSensor:
    def __init__(self):
        self.data = SimpleNamespace()
        self.body_names = list(BODY_NAMES)
        self._first_contact = None

    def compute_first_contact(self, dt: float, abs_tol: float = 1.0e-8) -> torch.Tensor:
        # Isaac Lab derives this from current_contact_time; here it is part of the synthetic state.
        return self._first_contact


class _Scene(dict):
    pass


class _CommandManager:
    def __init__(self):
        self._commands = {}

    def get_command(self, name: str) -> torch.Tensor:
        return self._commands[name]


class _Space:
    def __init__(self, shape):
        self.shape = tuple(shape)


class SyntheticSolo12Env:
    """Fake `CaTEnv` for Isaac-Velocity-CaT-Flat-Solo12-v0 (trainer-side only, no physics)."""

    def __init__(
        self,
        num_envs: int,
        device: str | torch.device = "cpu",
        seed: int = 0,
        pool: int = 4,
        episode_length: int = 500,
        constraints_cfg=None,
        curriculum: bool = True,
        adversarial: bool = False,
        stochastic_terminations: bool = False,
    ):
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        self.step_dt = 0.02
        self.max_episode_length = int(episode_length)
        self.common_step_counter = 0
        self.cfg = SimpleNamespace(constraints=constraints_cfg)
        self.extras: dict = {}
        self._curriculum_on = curriculum
        # optional mode: `dones` becomes a hard 0 / 1 mask sampled from the constraint probability (device Philox);
        # default off = the reference's behaviour (float probability, soft discount in GAE)
        self.stochastic_terminations = bool(stochastic_terminations)
        self._term_rng = None
        self.fuse_reset = True  # gather the reset statistics inside the constraint step (False: separate reset launch)

        self.scene = _Scene(robot=_Articulation(), contact_forces=_ContactSensor())
        self.command_manager = _CommandManager()
        self.action_manager = SimpleNamespace(_action=None, _prev_action=None)
        self.single_observation_space = {"policy": _Space((OBS_DIM,))}
        self.single_action_space = _Space((ACT_DIM,))
        self.unwrapped = self

        gen = torch.Generator().manual_seed(seed)
        self._pool = []
        for _ in range(max(1, pool)):
            cpu_state = sample_state(self.num_envs, gen, adversarial=adversarial)
            self._pool.append({k: v.to(self.device) for k, v in cpu_state.items()})
        self._cursor = 0
        robot = self.scene["robot"]
        robot.data.default_joint_pos = (
            torch.tensor(DEFAULT_JOINT_POS, dtype=torch.float32).repeat(self.num_envs, 1).to(self.device)
        )
        # staggered episode phases so that resets trickle in like in a real run, without a host sync
        phase = torch.randint(0, self.max_episode_length, (self.num_envs,), generator=gen)
        self._phase_np = phase.numpy().copy()  # host mirror of episode_length_buf: reset schedule without a sync
        self.episode_length_buf = phase.to(self.device, dtype=torch.long)
        self.reset_buf = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        self.reset_time_outs = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        self.reward_buf = torch.zeros(self.num_envs, dtype=torch.float32, device=self.device)
        self.obs_buf = {"policy": None}
        self.constraint_manager = None
        self.load_state(self._pool[0])

    # -- state plumbing -------------------------------------------------------------------------
    def load_state(self, state: dict[str, torch.Tensor]) -> None:
        """Point the Isaac-Lab-style attribute paths at the tensors of `state` (no copies)."""
        robot, sensor = self.scene["robot"], self.scene["contact_forces"]
        robot.data.joint_pos = state["joint_pos"]
        robot.data.joint_vel = state["joint_vel"]
        robot.data.joint_acc = state["joint_acc"]
        robot.data.applied_torque = state["applied_torque"]
        robot.data.projected_gravity_b = state["projected_gravity_b"]
        robot.data.root_pos_w = state["root_pos_w"]
        sensor.data.net_forces_w_history = state["net_forces_w_history"]
        sensor.data.last_air_time = state["last_air_time"]
        sensor._first_contact = state["first_contact"]
        self.command_manager._commands["base_velocity"] = state["command"]
        self.action_manager._action = state["action"]
        self.action_manager._prev_action = state["prev_action"]
        self._raw_reward = state["raw_reward"]
        self.obs_buf = {"policy": state["obs"]}
        self._state = state

    def load_managers(self, manager_cls=None, **manager_kwargs):
        """Mirror of `CaTEnv.load_managers` (reference cat_env.py:18-40)."""
        if manager_cls is None:
            from .constraint_manager import ConstraintManager as manager_cls
        if getattr(self.cfg, "constraints", None) is not None:
            self.constraint_manager = manager_cls(self.cfg.constraints, self, **manager_kwargs)
        return self.constraint_manager

    # -- gym-style API used by the trainer --------------------------------------------------------
    def reset(self):
        self.load_state(self._pool[0])
        return self.obs_buf, {}

    def _advance(self):
        self._cursor = (self._cursor + 1) % len(self._pool)
        self.load_state(self._pool[self._cursor])

    def step(self, action: torch.Tensor):
        """Trainer-visible part of `CaTEnv.step` (reference cat_env.py:92-147) on synthetic state."""
        resetting = self.step_host()
        self.step_device(resetting)
        self.step_finish_host()
        return self.obs_buf, self.reward_buf, self._dones, self.reset_time_outs, self.extras

    # The step in two halves, so that a trainer can replay the device half from a CUDA graph (one launch per env step):
    #   step_host()   : everything the host does -- state pointers, host-side counters, the time-out schedule (a host
    #                   mirror of the episode lengths: no device read), the curriculum and the extras["log"] dict;
    #   step_device() : every kernel -- counters, time-outs, the fused constraint step (+ reset statistics), the zeroing
    #                   of the reset envs' episode lengths.  With `uniform=True` it does the same launches whether or not
    #                   anybody resets this step (what a graph records once must be valid for every replay);
    #   step_finish_host() : the curriculum of the envs that reset (reference: inside _reset_idx, i.e. after this step's
    #                   constraint computation -- the new max_p takes effect from the next step on).
    def pointer_key(self):
        """Identifies the set of state tensors `step_device()` reads after the last `step_host()` (a CUDA graph of the
        device half is valid for exactly one such set)."""
        return self._cursor

    def step_host(self) -> bool:
        self._advance()
        self.common_step_counter += 1
        self._phase_np += 1
        due = self._phase_np >= self.max_episode_length
        resetting = bool(due.any())
        mgr = self.constraint_manager
        if resetting:
            self.extras["log"] = dict()
            self._phase_np[due] = 0
        if mgr is not None:
            mgr.refresh_params()
        self._resetting = resetting
        return resetting

    def step_finish_host(self) -> None:
        if self._resetting and self.constraint_manager is not None:
            self._curriculum()

    def step_device(self, resetting: bool | None = None, uniform: bool = False, fused_out=None):
        resetting = self._resetting if resetting is None else resetting
        self.episode_length_buf += 1
        # time-outs (Isaac Lab's termination manager in the real env)
        torch.ge(self.episode_length_buf, self.max_episode_length, out=self.reset_time_outs)
        self.reset_buf = self.reset_time_outs
        mgr = self.constraint_manager
        fuse = (resetting or uniform) and self.fuse_reset
        if mgr is not None:
            # when some env resets this step, its episode statistics are gathered by the same two launches
            self.reward_buf, self._dones = mgr.compute_step(self._raw_reward, self.reset_buf, fuse_reset=fuse, fused_out=fused_out)
        else:
            self.reward_buf = self._raw_reward
            self._dones = self.reset_buf.float()
        if self.stochastic_terminations and mgr is not None:
            if self._term_rng is None:
                from . import ops

                self._term_rng = ops.make_rng_state(0x7E57, self.device)
            self.termination_mask = mgr.sample_terminations(self._term_rng, probs=self._dones, with_ids=False)
            self._dones = self.termination_mask.float()
        if resetting or uniform:
            if mgr is not None and resetting:
                # (assigned, not .update()d: the fused statistics stay one packed device vector until somebody reads them)
                self.extras["log"] = mgr.fused_reset_stats(fused_out) if fuse else mgr.reset_masked(self.reset_buf)
            self.episode_length_buf.masked_fill_(self.reset_buf, 0)

    def _curriculum(self):
        mgr = self.constraint_manager
        if self._curriculum_on:
            for name in mgr.active_terms:
                if name in SOLO12_CURRICULUM_TERMS:
                    modify_constraint_p(self, None, name, num_steps=24 * 1000, init_max_p=0.25)

    def _reset_masked(self, mask: torch.Tensor):
        """CaT-relevant part of `CaTEnv._reset_idx` (cat_env.py:149-200), envs selected by a device mask."""
        mgr = self.constraint_manager
        self.extras["log"] = dict()
        if mgr is not None:
            self._curriculum()
            # (assigned, not .update()d: the fused statistics stay one packed device vector until somebody reads them)
            self.extras["log"] = mgr.fused_reset_stats() if self.fuse_reset else mgr.reset_masked(mask)
        self.episode_length_buf.masked_fill_(mask, 0)

    def _reset_idx(self, env_ids: torch.Tensor):
        """Same with explicit env ids, the way Isaac Lab calls it (cat_env.py:149-200)."""
        mgr = self.constraint_manager
        self.extras["log"] = dict()
        if mgr is not None:
            self._curriculum()
            self.extras["log"].update(mgr.reset(env_ids))
        self.episode_length_buf[env_ids] = 0


SOLO12_CURRICULUM_TERMS = (
    "joint_torque",
    "joint_velocity",
    "joint_acceleration",
    "action_rate",
    "hip_position",
    "base_orientation",
    "air_time",
    "two_foot_contact",
)


def solo12_constraints_cfg(stress: bool = False, constraints_module=None, term_cls=None, scene_entity_cls=None):
    """The 13-term `ConstraintsCfg` of the Solo12 flat task (reference cat_flat_env_cfg.py:259-355).

    `stress=True` appends the three extra terms of the 16-term / 93-column sweep (SURVEY.md §8d cfg 5).
    Returned as an ordered dict, which the manager accepts like a configclass instance
    (reference constraint_manager.py:242).  The three optional arguments let the golden-vector
    generator build the very same cfg out of the reference's own term functions and cfg class.
    """
    from ._isaaclab_compat import SceneEntityCfg

    if constraints_module is None:
        from . import constraints
    else:
        constraints = constraints_module
    if term_cls is None:
        from .manager_constraint_cfg import ConstraintTermCfg as Term
    else:
        Term = term_cls
    if scene_entity_cls is not None:
        SceneEntityCfg = scene_entity_cls

    all_joints = [".*_HAA", ".*_HFE", ".*_KFE"]
    cfg = {
        "joint_torque": Term(
            func=constraints.joint_torque,
            max_p=0.25,
            params={"limit": 3.0, "asset_cfg": SceneEntityCfg("robot", joint_names=all_joints)},
        ),
        "joint_velocity": Term(
            func=constraints.joint_velocity,
            max_p=0.25,
            params={"limit": 16.0, "asset_cfg": SceneEntityCfg("robot", joint_names=all_joints)},
        ),
        "joint_acceleration": Term(
            func=constraints.joint_acceleration,
            max_p=0.25,
            params={"limit": 800.0, "asset_cfg": SceneEntityCfg("robot", joint_names=all_joints)},
        ),
        "action_rate": Term(
            func=constraints.action_rate,
            max_p=0.25,
            params={"limit": 80.0, "asset_cfg": SceneEntityCfg("robot", joint_names=all_joints)},
        ),
        "contact": Term(
            func=constraints.contact,
            max_p=1.0,
            params={"asset_cfg": SceneEntityCfg("contact_forces", body_names=["base_link", ".*_UPPER_LEG"])},
        ),
        "foot_contact_force": Term(
            func=constraints.foot_contact_force,
            max_p=1.0,
            params={"limit": 50.0, "asset_cfg": SceneEntityCfg("contact_forces", body_names=".*_FOOT")},
        ),
        "front_hfe_position": Term(
            func=constraints.joint_position,
            max_p=1.0,
            params={"limit": 1.3, "asset_cfg": SceneEntityCfg("robot", joint_names=["FL_HFE", "FR_HFE"])},
        ),
        "upsidedown": Term(
            func=constraints.upsidedown, max_p=1.0, params={"limit": 0.0, "asset_cfg": SceneEntityCfg("robot")}
        ),
        "hip_position": Term(
            func=constraints.joint_position_when_moving_forward,
            max_p=0.25,
            params={
                "limit": 0.2,
                "velocity_deadzone": 0.1,
                "asset_cfg": SceneEntityCfg("robot", joint_names=[".*_HAA"]),
            },
        ),
        "base_orientation": Term(
            func=constraints.base_orientation,
            max_p=0.25,
            params={"limit": 0.1, "asset_cfg": SceneEntityCfg("robot")},
        ),
        "air_time": Term(
            func=constraints.air_time,
            max_p=0.25,
            params={
                "limit": 0.25,
                "velocity_deadzone": 0.1,
                "asset_cfg": SceneEntityCfg("contact_forces", body_names=".*_FOOT"),
            },
        ),
        "no_move": Term(
            func=constraints.no_move,
            max_p=0.1,
            params={
                "velocity_deadzone": 0.1,
                "joint_vel_limit": 4.0,
                "asset_cfg": SceneEntityCfg("robot", joint_names=all_joints),
            },
        ),
        "two_foot_contact": Term(
            func=constraints.n_foot_contact,
            max_p=0.25,
            params={
                "number_of_desired_feet": 2,
                "min_command_value": 0.5,
                "asset_cfg": SceneEntityCfg("contact_forces", body_names=".*_FOOT"),
            },
        ),
    }
    if stress:
        cfg["joint_range"] = Term(
            func=constraints.joint_range,
            max_p=0.25,
            params={"limit": 1.0, "asset_cfg": SceneEntityCfg("robot", joint_names=all_joints)},
        )
        cfg["min_base_height"] = Term(
            func=constraints.min_base_height,
            max_p=1.0,
            params={"limit": 0.15, "asset_cfg": SceneEntityCfg("robot")},
        )
        cfg["rear_hfe_position"] = Term(
            func=constraints.joint_position,
            max_p=1.0,
            params={"limit": 1.3, "asset_cfg": SceneEntityCfg("robot", joint_names=["HL_HFE", "HR_HFE"])},
        )
    return cfg


def curriculum_max_p(common_step_counter: int, num_steps: int = 24 * 1000, init_max_p: float = 0.25) -> float:
    """Closed form of `modify_constraint_p` (reference curriculums.py:28-35), for test schedules."""
    progress = min(common_step_counter / num_steps, 1.0)
    return 1 / (20 + progress * (1 / init_max_p - 20))


__all__ = [
    "SyntheticSolo12Env",
    "sample_state",
    "solo12_constraints_cfg",
    "curriculum_max_p",
    "BODY_NAMES",
    "JOINT_NAMES",
    "STATE_FIELDS",
    "OBS_DIM",
    "ACT_DIM",
]

assert math.isclose(curriculum_max_p(0), 0.05)
