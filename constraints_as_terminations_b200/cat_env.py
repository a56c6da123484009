"""`CaTEnv`: Isaac Lab manager-based RL env with constraints-as-terminations.

Counterpart of the reference's `exts/cat_envs/cat_envs/tasks/utils/cat/cat_env.py` (`CaTEnv` :17-200): same
constructor / `step()` contract -- `(obs_dict, reward[N], dones[N] float32 in [0,1], time_outs[N] bool, extras)`
where `dones` is the constraint termination *probability* (1.0 for envs Isaac Lab hard-resets) -- on top of the
fused constraint kernels:

* `load_managers()` adds `self.constraint_manager = ConstraintManager(cfg.constraints, self)` (reference :34-40);
* `step()` follows Isaac Lab's `ManagerBasedRLEnv.step` order and inserts ONE fused call,
  `constraint_manager.compute_step(raw_reward, reset_buf, fuse_reset=True)`, for the reference's lines :100-107,:118-121
  (constraint probability, `clip(reward * (1 - p), 0)`, `dones = p`, `dones[reset] = 1`) and for the
  `constraint_manager.reset(reset ids)` of :181;
* `_reset_idx()` gathers the per-term episode statistics of the envs being reset *before* Isaac Lab zeroes their
  episode lengths and merges them into `extras["log"]` (reference :178-182), delegating everything else to
  `ManagerBasedRLEnv._reset_idx` instead of re-implementing it.

Isaac Lab is not importable where this repository is built and tested; `tests/test_cat_env_gpu.py` executes this
module (constructor, load_managers, step, _reset_idx) on a test double of `ManagerBasedRLEnv` backed by the synthetic
Solo12 state, against the oracle of the reference's step; `SyntheticSolo12Env` mirrors the same three hooks for the
trainer-side runs.  With Isaac Lab present it subclasses the real `ManagerBasedRLEnv`.
"""

from __future__ import annotations

from collections.abc import Sequence

import torch

from .constraint_manager import ConstraintManager

try:
    from isaaclab.envs.manager_based_rl_env import ManagerBasedRLEnv

    HAVE_ISAACLAB_ENV = True
except Exception:  # noqa: BLE001
    ManagerBasedRLEnv = object
    HAVE_ISAACLAB_ENV = False


class CaTEnv(ManagerBasedRLEnv):
    """Manager-based RL env whose `terminated` output is the CaT termination probability."""

    def __init__(self, *args, **kwargs):
        if not HAVE_ISAACLAB_ENV:
            raise ImportError(
                "CaTEnv needs Isaac Lab (isaaclab.envs.ManagerBasedRLEnv); without it use "
                "constraints_as_terminations_b200.synthetic_env.SyntheticSolo12Env for trainer-side runs"
            )
        super().__init__(*args, **kwargs)

    # -- managers ---------------------------------------------------------------------------------
    def load_managers(self):
        super().load_managers()
        if hasattr(self.cfg, "constraints"):
            self.constraint_manager = ConstraintManager(self.cfg.constraints, self)
            print("[INFO] Constraint Manager: ", self.constraint_manager)

    # -- stepping ---------------------------------------------------------------------------------
    def _simulate(self):
        """Decimated physics stepping, as in ManagerBasedRLEnv.step."""
        rendering = self.sim.has_gui() or self.sim.has_rtx_sensors()
        for _ in range(self.cfg.decimation):
            self._sim_step_counter += 1
            self.action_manager.apply_action()
            self.scene.write_data_to_sim()
            self.sim.step(render=False)
            if rendering and self._sim_step_counter % self.cfg.sim.render_interval == 0:
                self.sim.render()
            self.scene.update(dt=self.physics_dt)

    def step(self, action: torch.Tensor):
        self.action_manager.process_action(action.to(self.device))
        self.recorder_manager.record_pre_step()
        self._simulate()

        self.episode_length_buf += 1
        self.common_step_counter += 1
        self.reset_buf = self.termination_manager.compute()
        self.reset_terminated = self.termination_manager.terminated
        self.reset_time_outs = self.termination_manager.time_outs

        raw_reward = self.reward_manager.compute(dt=self.step_dt)
        if hasattr(self, "constraint_manager"):
            # constraint probability, constrained reward, float dones (incl. dones[reset] = 1) and the episode
            # statistics of the envs flagged in reset_buf (which _reset_idx below would otherwise gather with a
            # second call) in one fused call; clones because the manager owns and reuses its output buffers
            reward, dones = self.constraint_manager.compute_step(raw_reward, self.reset_buf, fuse_reset=True)
            self.reward_buf, dones = reward.clone(), dones.clone()
            self._fused_reset_pending = True
        else:
            self.reward_buf = raw_reward
            dones = self.reset_buf.to(torch.float32)

        if len(self.recorder_manager.active_terms) > 0:
            self.obs_buf = self.observation_manager.compute()
            self.recorder_manager.record_post_step()

        reset_env_ids = self.reset_buf.nonzero(as_tuple=False).squeeze(-1)
        if len(reset_env_ids) == 0:
            self._fused_reset_pending = False
        if len(reset_env_ids) > 0:
            self.recorder_manager.record_pre_reset(reset_env_ids)
            self._reset_idx(reset_env_ids)
            self.scene.write_data_to_sim()
            if self.sim.has_rtx_sensors() and self.cfg.rerender_on_reset:
                self.sim.render()
            self.recorder_manager.record_post_reset(reset_env_ids)

        self.command_manager.compute(dt=self.step_dt)
        if "interval" in self.event_manager.available_modes:
            self.event_manager.apply(mode="interval", dt=self.step_dt)
        self.obs_buf = self.observation_manager.compute()
        return self.obs_buf, self.reward_buf, dones, self.reset_time_outs, self.extras

    def _reset_idx(self, env_ids: Sequence[int]):
        info = None
        if hasattr(self, "constraint_manager"):
            if getattr(self, "_fused_reset_pending", False):
                # called from step() for exactly the envs flagged in reset_buf: their statistics were gathered (and
                # cleared) by compute_step(..., fuse_reset=True), with the episode lengths as they were then
                self._fused_reset_pending = False
                info = self.constraint_manager.fused_reset_stats()
                for term_cfg in self.constraint_manager._class_term_cfgs:
                    term_cfg.func.reset(env_ids=env_ids)
            else:  # reset() of the whole env, or any other caller
                info = self.constraint_manager.reset(env_ids)  # needs the episode lengths Isaac Lab is about to zero
        super()._reset_idx(env_ids)
        if info is not None:
            self.extras["log"].update(info)
