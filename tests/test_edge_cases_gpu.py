"""Edge cases of the manager on the GPU: empty selections, strided / unaligned sources (the non-bulk staging
path of the eval kernel), cfg swaps and in-place param edits, integer max_p."""

import math

import pytest
import torch

from constraints_as_terminations_b200 import ConstraintManager, ConstraintTermCfg, constraints
from constraints_as_terminations_b200 import synthetic_env as se
from constraints_as_terminations_b200._isaaclab_compat import SceneEntityCfg
from oracle import cat_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pair(n, seed=0, stress=False):
    state = se.sample_state(n, torch.Generator().manual_seed(seed))
    cpu_env = se.SyntheticSolo12Env(n, device="cpu", pool=1)
    gpu_env = se.SyntheticSolo12Env(n, device=DEV, pool=1)
    cpu_env.load_state(state)
    gpu_env.load_state({k: v.to(DEV) for k, v in state.items()})
    cfg_cpu, cfg_gpu = se.solo12_constraints_cfg(stress=stress), se.solo12_constraints_cfg(stress=stress)
    oracle = cat_oracle.ManagerOracle(cpu_env, cat_oracle.terms_from_cfg(cfg_cpu, resolve_scene=cpu_env.scene))
    mgr = ConstraintManager(cfg_gpu, gpu_env)
    return cpu_env, gpu_env, cfg_cpu, oracle, mgr


def test_reset_with_empty_selection_gives_nan_like_torch():
    cpu_env, gpu_env, _, oracle, mgr = _pair(64)
    mgr.compute()
    oracle.compute()
    before = mgr._stats.clone()
    got = mgr.reset(torch.empty(0, dtype=torch.long, device=DEV))
    want = oracle.reset(torch.empty(0, dtype=torch.long))
    for k in want:
        assert math.isnan(float(want[k])) and math.isnan(float(got[k]))
    assert torch.equal(mgr._stats, before)  # nothing was cleared
    none_sel = mgr.reset_masked(torch.zeros(64, dtype=torch.bool, device=DEV))
    assert all(math.isnan(float(v)) for v in none_sel.values())


def test_strided_and_unaligned_sources_take_the_fallback_staging_path():
    n = 96
    cpu_env, gpu_env, _, oracle, mgr = _pair(n, seed=3, stress=True)
    robot = gpu_env.scene["robot"].data
    # strided view: [N, 12] columns of a wider [N, 20] buffer (row stride 20 != row length 12)
    wide = torch.zeros(n, 20, device=DEV)
    wide[:, 4:16] = robot.joint_vel
    robot.joint_vel = wide[:, 4:16]
    assert not robot.joint_vel.is_contiguous()
    # contiguous but only 4-byte aligned: a view starting one float into a larger buffer
    flat = torch.zeros(n * 12 + 1, device=DEV)
    flat[1:] = robot.applied_torque.reshape(-1)
    robot.applied_torque = flat[1:].view(n, 12)
    assert robot.applied_torque.data_ptr() % 16 != 0
    for _ in range(2):
        assert torch.equal(mgr.compute().cpu(), oracle.compute())
    assert torch.equal(mgr.cat.get_raw_constraints().cpu(), torch.cat(list(oracle.cat.raw.values()), dim=1))


def test_cfg_swap_and_inplace_param_edit_are_picked_up():
    cpu_env, gpu_env, cfg_cpu, oracle, mgr = _pair(128, seed=5)
    assert torch.equal(mgr.compute().cpu(), oracle.compute())
    # in-place edit of a scalar param (the reference re-reads params every step, constraint_manager.py:217)
    mgr.get_term_cfg("joint_velocity").params["limit"] = 4.0
    cfg_cpu["joint_velocity"].params["limit"] = 4.0
    assert torch.equal(mgr.compute().cpu(), oracle.compute())
    # set_term_cfg with a brand-new cfg object (different function, integer max_p)
    new = ConstraintTermCfg(func=constraints.joint_velocity, max_p=1, params={"limit": 2.0, "asset_cfg": SceneEntityCfg("robot", joint_names=[".*"])})
    new.params["asset_cfg"].resolve(gpu_env.scene)
    mgr.set_term_cfg("joint_acceleration", new)
    for i, (name, fn, params, max_p) in enumerate(oracle.terms):
        if name == "joint_acceleration":
            p = dict(new.params)
            oracle.terms[i] = (name, cat_oracle.term_joint_velocity, p, 1)
    assert torch.equal(mgr.compute().cpu(), oracle.compute())
    assert torch.equal(mgr.cat.get_running_maxes().cpu(), torch.cat(list(oracle.cat.running_max.values()), dim=1))


def test_terms_that_nobody_violates_decay_to_the_clamp():
    """Column max <= 0 every step: running max follows the 1e-6 clamp path (constraint_manager.py:55)."""
    n = 40
    env = se.SyntheticSolo12Env(n, device=DEV, pool=1)
    env.scene["robot"].data.joint_vel = torch.zeros(n, 12, device=DEV)
    cfg = {"v": ConstraintTermCfg(func=constraints.joint_velocity, max_p=0.5, params={"limit": 1.0, "asset_cfg": SceneEntityCfg("robot", joint_names=[".*"])})}
    mgr = ConstraintManager(cfg, env)
    orc = cat_oracle.CatOracle()
    for _ in range(3):
        p = mgr.compute()
        orc.add("v", torch.zeros(n, 12).abs() - 1.0, 0.5)
        assert float(p.abs().sum()) == 0.0
        assert torch.equal(mgr.cat.get_running_maxes().cpu(), orc.running_max["v"])
    assert float(mgr._episode_sums["v"].sum()) == 0.0


def _random_cfg(rng, constraints_module=None):
    """A random ConstraintsCfg: random subset / order of the 15 term functions with random ids and scalars."""
    import random

    c = constraints
    J, B = se.JOINT_NAMES, se.BODY_NAMES

    def joints():
        k = rng.randint(1, 12)
        return rng.sample(J, k)

    def bodies():
        k = rng.randint(1, 6)
        return rng.sample(B, k)

    makers = {
        "joint_position": lambda: (c.joint_position, {"limit": rng.uniform(0.2, 1.5), "asset_cfg": SceneEntityCfg("robot", joint_names=joints())}),
        "joint_position_when_moving_forward": lambda: (c.joint_position_when_moving_forward, {"limit": rng.uniform(0.05, 0.5), "velocity_deadzone": rng.uniform(0.05, 0.4), "asset_cfg": SceneEntityCfg("robot", joint_names=joints())}),
        "joint_torque": lambda: (c.joint_torque, {"limit": rng.uniform(1.0, 4.0), "asset_cfg": SceneEntityCfg("robot", joint_names=joints())}),
        "joint_velocity": lambda: (c.joint_velocity, {"limit": rng.uniform(4.0, 20.0), "asset_cfg": SceneEntityCfg("robot", joint_names=joints())}),
        "joint_acceleration": lambda: (c.joint_acceleration, {"limit": rng.uniform(200.0, 900.0), "asset_cfg": SceneEntityCfg("robot", joint_names=joints())}),
        "upsidedown": lambda: (c.upsidedown, {"limit": rng.uniform(-0.9, 0.2), "asset_cfg": SceneEntityCfg("robot")}),
        "contact": lambda: (c.contact, {"asset_cfg": SceneEntityCfg("contact_forces", body_names=bodies())}),
        "base_orientation": lambda: (c.base_orientation, {"limit": rng.uniform(0.05, 0.5), "asset_cfg": SceneEntityCfg("robot")}),
        "air_time": lambda: (c.air_time, {"limit": rng.uniform(0.1, 0.4), "velocity_deadzone": rng.uniform(0.05, 0.5), "asset_cfg": SceneEntityCfg("contact_forces", body_names=bodies())}),
        "n_foot_contact": lambda: (c.n_foot_contact, {"number_of_desired_feet": rng.randint(0, 4), "min_command_value": rng.uniform(0.1, 0.8), "asset_cfg": SceneEntityCfg("contact_forces", body_names=bodies())}),
        "joint_range": lambda: (c.joint_range, {"limit": rng.uniform(0.2, 1.2), "asset_cfg": SceneEntityCfg("robot", joint_names=joints())}),
        "action_rate": lambda: (c.action_rate, {"limit": rng.uniform(20.0, 100.0), "asset_cfg": SceneEntityCfg("robot", joint_names=joints())}),
        "foot_contact_force": lambda: (c.foot_contact_force, {"limit": rng.uniform(5.0, 60.0), "asset_cfg": SceneEntityCfg("contact_forces", body_names=bodies())}),
        "min_base_height": lambda: (c.min_base_height, {"limit": rng.uniform(0.1, 0.35), "asset_cfg": SceneEntityCfg("robot")}),
        "no_move": lambda: (c.no_move, {"velocity_deadzone": rng.uniform(0.05, 0.6), "joint_vel_limit": rng.uniform(1.0, 8.0), "asset_cfg": SceneEntityCfg("robot", joint_names=joints())}),
    }
    names = rng.sample(sorted(makers), rng.randint(1, 15))
    # the same function may appear twice under different names
    names += [rng.choice(sorted(makers)) for _ in range(rng.randint(0, 3))]
    cfg = {}
    for i, name in enumerate(names):
        func, params = makers[name]()
        cfg[f"t{i}_{name}"] = ConstraintTermCfg(func=func, max_p=rng.choice([1.0, 0.25, 0.1, rng.uniform(0.01, 1.0)]), params=params)
    return cfg


@pytest.mark.parametrize("trial", range(12))
def test_random_term_configurations_match_oracle_bit_exact(trial):
    import copy
    import random

    rng = random.Random(1000 + trial)
    n = rng.choice([1, 7, 32, 33, 100, 257, 1024, 2500])
    tau, min_p = rng.choice([(0.95, 0.0), (0.9, 0.01), (0.5, 0.0)])
    cfg = _random_cfg(rng)
    cpu_env = se.SyntheticSolo12Env(n, device="cpu", pool=1)
    gpu_env = se.SyntheticSolo12Env(n, device=DEV, pool=1)
    cfg_cpu = copy.deepcopy(cfg)
    oracle = cat_oracle.ManagerOracle(cpu_env, cat_oracle.terms_from_cfg(cfg_cpu, resolve_scene=cpu_env.scene), tau=tau, min_p=min_p)
    mgr = ConstraintManager(cfg, gpu_env, tau=tau, min_p=min_p)
    gen = torch.Generator().manual_seed(trial)
    for step in range(3):
        state = se.sample_state(n, gen, adversarial=(step == 1))
        cpu_env.load_state(state)
        gpu_env.load_state({k: v.to(DEV) for k, v in state.items()})
        want = oracle.compute()
        got = mgr.compute()
        assert torch.equal(got.cpu(), want), f"trial {trial} step {step}: cstr_prob"
        assert torch.equal(mgr.cat.get_running_maxes().cpu(), torch.cat(list(oracle.cat.running_max.values()), dim=1))
    assert torch.equal(mgr.cat.get_raw_constraints().cpu(), torch.cat(list(oracle.cat.raw.values()), dim=1))
    assert torch.equal(mgr._stats[0, : len(cfg)].cpu(), torch.stack(list(oracle.episode_sums.values())))
    assert torch.equal(mgr._stats[1, : len(cfg)].cpu(), torch.stack(list(oracle.mean_values.values())))
