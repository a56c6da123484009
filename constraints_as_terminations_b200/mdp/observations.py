"""Policy observation assembly (reference task cfg `S12/cat_flat_env_cfg.py:137-172`).

The Solo12 policy observation is six Isaac Lab `ObservationTermCfg`s -- base angular velocity, velocity command, projected
gravity, joint positions, joint velocities (12 joints in a fixed order), last action -- each a column selection of a
state tensor with additive uniform noise and a scale, concatenated to 45 numbers.  Isaac Lab's ObservationManager runs a
dozen eager ops per term (gather, `rand_like`, mul, add, scale, `cat`); `ObservationAssembler` does all terms for all
envs in ONE launch (`catb200_obs_assemble`), noise from the device Philox stream, so the call can sit inside the CUDA
graph of the env step.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Sequence

import torch

from .. import _lib as L
from .. import ops


@dataclass
class ObsTermSpec:
    """One observation term: `fetch(env) -> [N, D]` float32 state tensor, the columns taken from it (None = all),
    uniform noise bounds (None = no noise) and a scale (scalar or one value per column)."""

    name: str
    fetch: Callable
    ids: Sequence[int] | None = None
    noise: tuple[float, float] | None = None
    scale: float | Sequence[float] = 1.0
    _cols: int = field(default=0, init=False)


def solo12_policy_terms(joint_ids: Sequence[int] | None = None) -> list[ObsTermSpec]:
    """The six terms of `ObservationsCfg.PolicyCfg` (cat_flat_env_cfg.py:145-170), noise enabled (`enable_corruption`).
    `joint_ids`: indices of FL_HAA, FL_HFE, FL_KFE, FR_*, HR_*, HL_* in the articulation's joint order (`preserve_order`)."""
    robot = lambda env: env.scene["robot"].data  # noqa: E731
    return [
        ObsTermSpec("base_ang_vel", lambda env: robot(env).root_ang_vel_b, noise=(-0.001, 0.001), scale=0.25),
        ObsTermSpec("velocity_commands", lambda env: env.command_manager.get_command("base_velocity"), scale=(2.0, 2.0, 0.25)),
        ObsTermSpec("projected_gravity", lambda env: robot(env).projected_gravity_b, noise=(-0.05, 0.05), scale=0.1),
        ObsTermSpec("joint_pos", lambda env: robot(env).joint_pos, ids=joint_ids, noise=(-0.01, 0.01), scale=1.0),
        ObsTermSpec("joint_vel", lambda env: robot(env).joint_vel, ids=joint_ids, noise=(-0.2, 0.2), scale=0.05),
        ObsTermSpec("actions", lambda env: env.action_manager._action, scale=1.0),
    ]


class ObservationAssembler:
    def __init__(self, terms: Sequence[ObsTermSpec], device, seed: int = 0):
        if not 0 < len(terms) <= 8:
            raise ValueError("between 1 and 8 observation terms")
        self.terms = list(terms)
        self.device = torch.device(device)
        self.rng_state = ops.make_rng_state(seed, self.device)
        self._live = []

    def _plan(self, env) -> tuple[L.ObsPlan, int]:
        plan = L.ObsPlan()
        plan.n_terms = len(self.terms)
        self._live = []
        total = 0
        for t, spec in zip(plan.terms, self.terms):
            src = spec.fetch(env)
            L.require_cuda(src, spec.name)
            if src.dtype != torch.float32 or src.ndim != 2 or src.stride(1) != 1:
                raise TypeError(f"observation source '{spec.name}' must be a float32 [N, D] tensor with contiguous rows")
            ids = list(range(src.shape[1])) if spec.ids is None else [int(i) for i in spec.ids]
            if not 0 < len(ids) <= 32 or max(ids) > 255 or max(ids) >= src.shape[1] or min(ids) < 0:
                raise ValueError(f"observation term '{spec.name}': bad column selection {ids}")
            t.src, t.row_stride, t.n_cols = src.data_ptr(), src.stride(0), len(ids)
            lo, hi = spec.noise if spec.noise is not None else (0.0, 0.0)
            t.n_min, t.n_max, t.noise_span = lo, hi, hi - lo  # the difference in double, rounded once (python semantics)
            scale = [float(spec.scale)] * len(ids) if not isinstance(spec.scale, (tuple, list)) else [float(x) for x in spec.scale]
            if len(scale) != len(ids):
                raise ValueError(f"observation term '{spec.name}': {len(scale)} scales for {len(ids)} columns")
            for k, (i, sc) in enumerate(zip(ids, scale)):
                t.ids[k], t.scale[k] = i, sc
            total += len(ids)
            self._live.append(src)
        return plan, total

    def assemble(self, env, out: torch.Tensor | None = None, uniforms: torch.Tensor | None = None) -> torch.Tensor:
        """-> obs [N, n_cols]; `uniforms` [N, n_cols] replaces the Philox draws (tests)."""
        plan, n_cols = self._plan(env)
        n = self._live[0].shape[0]
        if out is None:
            out = torch.empty((n, n_cols), dtype=torch.float32, device=self.device)
        if out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != n * n_cols:
            raise TypeError("out must be a contiguous float32 [N, n_cols] tensor")
        L.check(
            L.load().catb200_obs_assemble(plan, n, out.data_ptr(), L.ptr(uniforms), self.rng_state.data_ptr(), L.stream()),
            "obs_assemble",
        )
        return out
