"""Test double for `isaaclab.envs.manager_based_rl_env.ManagerBasedRLEnv`: the members `CaTEnv` touches
(U/cat/cat_env.py:18-200 of the reference; constraints_as_terminations_b200/cat_env.py here), backed by the synthetic
Solo12 state source instead of Isaac Sim.  Installing it in `sys.modules` lets `cat_env.py` -- which subclasses the
real class when Isaac Lab is importable -- execute under test: constructor, load_managers, step, _reset_idx."""

from __future__ import annotations

import sys
import types

import torch

from constraints_as_terminations_b200 import synthetic_env as se


class _Recorder:
    active_terms: list = []

    def record_pre_step(self): pass
    def record_post_step(self): pass
    def record_pre_reset(self, env_ids): pass
    def record_post_reset(self, env_ids): pass


class _Sim:
    def __init__(self):
        self.steps = 0

    def has_gui(self): return False
    def has_rtx_sensors(self): return False
    def render(self): pass

    def step(self, render=False):
        self.steps += 1


class ManagerBasedRLEnv(se.SyntheticSolo12Env):
    """cfg: SimpleNamespace(num_envs, device, seed, pool, episode_length, decimation, constraints, sim.render_interval)."""

    def __init__(self, cfg, render_mode=None, **kwargs):
        super().__init__(cfg.num_envs, device=cfg.device, seed=cfg.seed, pool=cfg.pool, episode_length=cfg.episode_length,
                         constraints_cfg=getattr(cfg, "constraints", None), curriculum=False)  # fmt: skip
        self.cfg = cfg
        self.sim = _Sim()
        self.physics_dt = self.step_dt / cfg.decimation
        self._sim_step_counter = 0
        self.recorder_manager = _Recorder()
        env = self

        class _Scene(se._Scene):
            def write_data_to_sim(self): pass

            def update(self, dt):
                pass

        scene = _Scene(self.scene)
        self.scene = scene
        self.action_manager.process_action = lambda a: setattr(env, "_last_action", a)
        self.action_manager.apply_action = lambda: None
        self.termination_manager = types.SimpleNamespace(
            compute=lambda: env._terminations(), terminated=None, time_outs=None,
        )  # fmt: skip
        self.reward_manager = types.SimpleNamespace(compute=lambda dt: env._raw_reward)
        self.observation_manager = types.SimpleNamespace(compute=lambda: env.obs_buf)
        self.event_manager = types.SimpleNamespace(available_modes=[], apply=lambda **kw: None)
        self.command_manager.compute = lambda dt: None
        self.extras = {}
        self.load_managers()

    def _terminations(self):
        time_outs = self.episode_length_buf >= self.max_episode_length
        self.termination_manager.terminated = torch.zeros_like(time_outs)
        self.termination_manager.time_outs = time_outs
        return time_outs

    # ---- what Isaac Lab's base class provides and CaTEnv extends -----------------------------------------------
    def load_managers(self):  # the base managers already exist in this double
        pass

    def _reset_idx(self, env_ids):
        self.extras["log"] = dict()
        self.episode_length_buf[env_ids] = 0

    def step(self, action):  # never reached: CaTEnv overrides it
        raise NotImplementedError


def install():
    """Register the double as `isaaclab.envs.manager_based_rl_env` and (re)import cat_env against it."""
    import importlib

    for name in ("isaaclab", "isaaclab.envs"):
        sys.modules.setdefault(name, types.ModuleType(name))
    mod = types.ModuleType("isaaclab.envs.manager_based_rl_env")
    mod.ManagerBasedRLEnv = ManagerBasedRLEnv
    sys.modules["isaaclab.envs.manager_based_rl_env"] = mod
    from constraints_as_terminations_b200 import cat_env

    return importlib.reload(cat_env)


def uninstall():
    import importlib

    for name in ("isaaclab.envs.manager_based_rl_env", "isaaclab.envs", "isaaclab"):
        sys.modules.pop(name, None)
    from constraints_as_terminations_b200 import cat_env

    importlib.reload(cat_env)
