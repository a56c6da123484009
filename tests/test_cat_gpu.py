"""GPU parity of the per-step CaT path (through the C ABI) against the golden fixtures of the real
reference and against the CPU oracle.  fp32 results are expected bit-identical; the only toleranced
outputs are the reset means (torch's reduction order is implementation defined)."""

import os

import pytest
import torch

from constraints_as_terminations_b200 import CaT, ConstraintManager, ConstraintTermCfg, constraints
from constraints_as_terminations_b200 import synthetic_env as se
from constraints_as_terminations_b200._isaaclab_compat import SceneEntityCfg
from oracle import cat_oracle
from tests.helpers import replay_cat_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("fixture", ["cat_solo12.pt", "cat_stress.pt"])
def test_manager_matches_reference_golden(golden_dir, fixture):
    gold = torch.load(os.path.join(golden_dir, fixture), weights_only=False)
    n = gold["num_envs"]
    env = se.SyntheticSolo12Env(n, device=DEV, seed=gold["seed"], pool=1, adversarial=True)
    mgr = ConstraintManager(se.solo12_constraints_cfg(stress=gold["stress"]), env)
    env.constraint_manager = mgr
    assert mgr.active_terms == gold["names"]

    def step_fn(step, state, rec):
        reward, dones = mgr.compute_step(state["raw_reward"], rec["reset_buf"])
        out = {
            "cstr_prob": mgr._cstr_prob_buf.clone(),
            "running_max": mgr.cat.get_running_maxes().squeeze(0).clone(),
            "reward": reward.clone(),
            "dones": dones.clone(),
        }
        if "raw" in rec:
            out["raw"] = mgr.cat.get_raw_constraints()
            out["probs"] = torch.cat(list(mgr.cat.probs.values()), dim=1)
        return out

    def set_max_p(values):
        for name, v in zip(mgr.active_terms, values):
            cfg = mgr.get_term_cfg(name)
            cfg.max_p = v
            mgr.set_term_cfg(name, cfg)

    replay_cat_golden(gold, env, step_fn, mgr.reset, set_max_p, exact=True, device=DEV)
    torch.cuda.synchronize()
    assert torch.equal(torch.stack([mgr._episode_sums[k] for k in mgr.active_terms]).cpu(), gold["episode_sums"])
    assert torch.equal(torch.stack([mgr._cstr_mean_values[k] for k in mgr.active_terms]).cpu(), gold["mean_values"])


@pytest.mark.parametrize("num_envs", [1, 31, 33, 1000, 4096, 20011])  # 20011: 626 tiles = up to 3 per persistent CTA, the last one ragged
def test_manager_matches_oracle(num_envs):
    steps = 5
    cpu_env = se.SyntheticSolo12Env(num_envs, device="cpu", seed=11, pool=1)
    gpu_env = se.SyntheticSolo12Env(num_envs, device=DEV, seed=11, pool=1)
    cfg_cpu, cfg_gpu = se.solo12_constraints_cfg(stress=True), se.solo12_constraints_cfg(stress=True)
    oracle = cat_oracle.ManagerOracle(cpu_env, cat_oracle.terms_from_cfg(cfg_cpu, resolve_scene=cpu_env.scene))
    mgr = ConstraintManager(cfg_gpu, gpu_env)
    gen = torch.Generator().manual_seed(5)
    for step in range(steps):
        state = se.sample_state(num_envs, gen, adversarial=step == 2)
        cpu_env.load_state(state)
        gpu_env.load_state({k: v.to(DEV) for k, v in state.items()})
        for c in (cfg_cpu, mgr):  # curriculum-like change of max_p between steps
            term = c["joint_velocity"] if isinstance(c, dict) else c.get_term_cfg("joint_velocity")
            term.max_p = 0.05 + 0.04 * step
        want = oracle.compute()
        got = mgr.compute()
        assert torch.equal(got.cpu(), want), f"step {step}: cstr_prob differs"
        want_rm = torch.cat(list(oracle.cat.running_max.values()), dim=1)
        assert torch.equal(mgr.cat.get_running_maxes().cpu(), want_rm)
    assert torch.equal(mgr._stats[0].cpu(), torch.stack(list(oracle.episode_sums.values())))
    assert torch.equal(mgr._stats[1].cpu(), torch.stack(list(oracle.mean_values.values())))
    # per-term dict views and debug matrices
    assert torch.equal(mgr.cat.probs["air_time"].cpu(), oracle.cat.probs["air_time"])
    assert torch.equal(mgr.cat.raw_constraints["two_foot_contact"].cpu(), oracle.cat.raw["two_foot_contact"])
    assert mgr.cat.max_p["joint_velocity"].shape == (12,)
    assert mgr.cat.get_names() == mgr.active_terms and len(mgr.cat.get_vals()) == len(mgr.active_terms)
    # reset by ids == reset by mask == oracle (means: 1e-5 relative)
    cpu_env.episode_length_buf[:] = torch.arange(1, num_envs + 1)
    gpu_env.episode_length_buf[:] = torch.arange(1, num_envs + 1, device=DEV)
    ids = torch.arange(0, num_envs, 3)
    twin = torch.zeros(num_envs, dtype=torch.bool, device=DEV)
    twin[ids.to(DEV)] = True
    stats_before = mgr._stats.clone()
    by_mask = mgr.reset_masked(twin)
    mgr._stats.copy_(stats_before)
    by_ids = mgr.reset(ids.to(DEV))
    want = oracle.reset(ids)
    for k in want:
        torch.testing.assert_close(by_ids[k].cpu(), want[k], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(by_mask[k].cpu(), want[k], rtol=1e-5, atol=1e-7)
    assert torch.equal(mgr._stats[0].cpu(), torch.stack(list(oracle.episode_sums.values())))
    all_out = mgr.reset()
    want_all = oracle.reset(None)
    for k in want_all:  # rows already zeroed for `ids`: 0/len means are exact zeros there
        torch.testing.assert_close(all_out[k].cpu(), want_all[k], rtol=1e-5, atol=1e-7)
    assert float(mgr._stats.abs().sum()) == 0.0


def test_standalone_term_functions_match_oracle():
    n = 777
    state = se.sample_state(n, torch.Generator().manual_seed(2), adversarial=True)
    cpu_env = se.SyntheticSolo12Env(n, device="cpu", pool=1)
    gpu_env = se.SyntheticSolo12Env(n, device=DEV, pool=1)
    cpu_env.load_state(state)
    gpu_env.load_state({k: v.to(DEV) for k, v in state.items()})
    cfg = se.solo12_constraints_cfg(stress=True)
    for name, term in cfg.items():
        for v in term.params.values():
            if hasattr(v, "resolve"):
                v.resolve(cpu_env.scene)
        want = cat_oracle.TERM_ORACLES[term.func.__name__](cpu_env, **term.params)
        got = term.func(gpu_env, **term.params)
        assert got.shape == want.shape and got.dtype == want.dtype, f"{name}: {got.shape}/{got.dtype} vs {want.shape}/{want.dtype}"
        assert torch.equal(got.cpu(), want), f"{name}: values differ"


def test_python_terms_go_through_the_generic_path():
    n = 500
    state = se.sample_state(n, torch.Generator().manual_seed(4))
    env = se.SyntheticSolo12Env(n, device=DEV, pool=1)
    env.load_state({k: v.to(DEV) for k, v in state.items()})

    def wide(env, scale):  # 45 columns: wider than one 32-id block
        return env.obs_buf["policy"] * scale - 1.0

    def flag(env):  # bool [N]
        return env.scene["robot"].data.root_pos_w[:, 2] < 0.2

    def cpu_term(env):  # lives on the wrong device, int dtype
        return (torch.arange(env.num_envs) % 7 - 3).to(torch.int64)

    cfg = {
        "wide": ConstraintTermCfg(func=wide, max_p=0.3, params={"scale": 0.5}),
        "torque": ConstraintTermCfg(func=constraints.joint_torque, max_p=0.25, params={"limit": 3.0, "asset_cfg": SceneEntityCfg("robot", joint_names=[".*"])}),
        "flag": ConstraintTermCfg(func=flag, max_p=1.0, params={}),
        "cpu_term": ConstraintTermCfg(func=cpu_term, max_p=0.5, params={}),
    }
    mgr = ConstraintManager(cfg, env)
    oracle = cat_oracle.CatOracle()
    for step in range(3):
        got = mgr.compute()
        oracle.add("wide", wide(env, 0.5).cpu(), 0.3)
        oracle.add("torque", (env.scene["robot"].data.applied_torque.abs() - 3.0).cpu(), 0.25)
        oracle.add("flag", flag(env).cpu(), 1.0)
        oracle.add("cpu_term", cpu_term(env), 0.5)
        assert torch.equal(got.cpu(), oracle.combined()), f"step {step}"
    assert mgr.cat.probs["wide"].shape == (n, 45)
    assert torch.equal(mgr.cat.probs["wide"].cpu(), oracle.probs["wide"])
    assert torch.equal(mgr.cat.probs["cpu_term"].cpu(), oracle.probs["cpu_term"])


def test_standalone_cat_matches_oracle():
    n = 300
    gen = torch.Generator().manual_seed(9)
    cat, oracle = CaT(tau=0.9, min_p=0.01), cat_oracle.CatOracle(tau=0.9, min_p=0.01)
    for step in range(4):
        a = torch.randn(n, 7, generator=gen)
        b = torch.randn(n, generator=gen) > 0.5
        cat.add("a", a.to(DEV), 0.2)
        cat.add("b", b.to(DEV), 1.0)
        oracle.add("a", a, 0.2)
        oracle.add("b", b, 1.0)
        assert torch.equal(cat.get_probs().cpu(), oracle.combined())
        assert torch.equal(cat.get_running_maxes().cpu(), torch.cat(list(oracle.running_max.values()), dim=1))
        assert torch.equal(cat.probs["a"].cpu(), oracle.probs["a"])
    assert cat.get_names() == ["a", "b"]
    assert cat.get_max_p().shape == (8,)
    cat.reset()
    assert cat.get_probs().numel() == 0 and len(cat.running_maxes) == 2


def test_full_size_properties():
    """BASELINE sizes (65536 envs, 16 terms): properties that need no oracle run."""
    n = 65536
    env = se.SyntheticSolo12Env(n, device=DEV, seed=1, pool=2)
    mgr = ConstraintManager(se.solo12_constraints_cfg(stress=True), env)
    env.constraint_manager = mgr
    reset = torch.zeros(n, dtype=torch.bool, device=DEV)
    reset[::97] = True
    reward, dones = mgr.compute_step(env._raw_reward, reset)
    p = mgr._cstr_prob_buf
    assert p.shape == (n,) and float(p.min()) >= 0.0 and float(p.max()) <= 1.0
    assert torch.equal(dones[reset], torch.ones_like(dones[reset]))
    assert torch.equal(dones[~reset], p[~reset])
    assert torch.equal(reward, torch.clip(env._raw_reward * (1 - p), min=0))
    probs = torch.cat(list(mgr.cat.probs.values()), dim=1)
    raw = mgr.cat.get_raw_constraints()
    assert torch.equal(probs.max(1).values, p)
    assert torch.equal(probs > 0, raw > 0)  # violation mask bit-exact
    # first call assigns the column max (clamped): running max == max(colmax, 1e-6)
    assert torch.equal(mgr.cat.get_running_maxes().squeeze(0), raw.max(0).values.clamp(min=1e-6))
    # statistics after one step: indicator and value of the per-term max
    for name in mgr.active_terms:
        tmax = mgr.cat.probs[name].max(1).values
        assert torch.equal(mgr._episode_sums[name], (tmax > 0).float())
        assert torch.equal(mgr._cstr_mean_values[name], tmax)
    # second step, same state: Polyak update rm' = 0.95*rm + (1-0.95)*colmax and idempotent raw values
    rm0 = mgr.cat.get_running_maxes().clone()
    mgr.compute()
    colmax = raw.max(0).values.clamp(min=1e-6)
    assert torch.equal(mgr.cat.get_running_maxes().squeeze(0), rm0.squeeze(0) * 0.95 + (1.0 - 0.95) * colmax)
    assert torch.equal(mgr.cat.get_raw_constraints(), raw)


@pytest.mark.parametrize("num_envs", [33, 1000, 4096])
def test_fused_step_reset_equals_step_then_reset(num_envs):
    """compute_step(fuse_reset=True) == compute_step() followed by reset(ids of reset_buf) (the order of CaTEnv.step,
    reference cat_env.py:100 then :181) and == the oracle: cstr_prob / reward / dones / statistics bit-exact, the
    episode means to 1e-5 (double-precision sums vs torch's fp32 reduction)."""
    steps = 4
    envs = [se.SyntheticSolo12Env(num_envs, device=DEV, seed=3, pool=1) for _ in range(2)]
    cpu_env = se.SyntheticSolo12Env(num_envs, device="cpu", seed=3, pool=1)
    mgrs = [ConstraintManager(se.solo12_constraints_cfg(), e) for e in envs]
    cfg_cpu = se.solo12_constraints_cfg()
    oracle = cat_oracle.ManagerOracle(cpu_env, cat_oracle.terms_from_cfg(cfg_cpu, resolve_scene=cpu_env.scene))
    gen = torch.Generator().manual_seed(17)
    for step in range(steps):
        state = se.sample_state(num_envs, gen, adversarial=step == 1)
        reset_cpu = torch.rand(num_envs, generator=gen) < (0.0 if step == 2 else 0.2)  # step 2: nobody resets
        ep_len = torch.randint(1, 400, (num_envs,), generator=gen)
        cpu_env.load_state(state)
        cpu_env.episode_length_buf[:] = ep_len
        for e in envs:
            e.load_state({k: v.to(DEV) for k, v in state.items()})
            e.episode_length_buf[:] = ep_len.to(DEV)
        reset = reset_cpu.to(DEV)
        raw_reward = envs[0]._raw_reward
        # A: fused
        rew_a, dones_a = mgrs[0].compute_step(raw_reward, reset, fuse_reset=True)
        out_a = mgrs[0].fused_reset_stats()
        # B: two calls
        rew_b, dones_b = mgrs[1].compute_step(envs[1]._raw_reward, reset)
        out_b = mgrs[1].reset(reset.nonzero().squeeze(-1))
        # oracle
        want_p = oracle.compute()
        want_rew, want_dones = cat_oracle.step_epilogue(cpu_env._raw_reward, want_p, reset_cpu)
        want_out = oracle.reset(reset_cpu.nonzero().squeeze(-1))
        assert torch.equal(rew_a, rew_b) and torch.equal(dones_a, dones_b)
        assert torch.equal(rew_a.cpu(), want_rew) and torch.equal(dones_a.cpu(), want_dones)
        assert torch.equal(mgrs[0]._stats, mgrs[1]._stats), f"step {step}: statistics after reset differ"
        assert torch.equal(mgrs[0]._stats[0].cpu(), torch.stack(list(oracle.episode_sums.values())))
        assert list(out_a.keys()) == list(out_b.keys()) == list(want_out.keys())
        for k in want_out:
            torch.testing.assert_close(out_a[k].cpu(), want_out[k], rtol=1e-5, atol=1e-7, equal_nan=True)
            torch.testing.assert_close(out_a[k], out_b[k], rtol=1e-6, atol=1e-8, equal_nan=True)
    with pytest.raises(ValueError):
        mgrs[0].compute_step(raw_reward, None, fuse_reset=True)


def test_optional_stochastic_termination_mask_is_the_documented_philox_draw():
    """north_star's "sampling the stochastic termination mask ... bit-exact for the termination-mask indices": optional
    mode (default off -- the reference keeps `dones` a probability).  mask[i] = (u_i < dones[i]) with u from Philox4x32-10;
    probabilities are bit-exact against the oracle, so the index list is too."""
    from constraints_as_terminations_b200 import ops
    from oracle import philox_oracle

    n, seed = 5000, 31337
    cpu_env = se.SyntheticSolo12Env(n, device="cpu", seed=2, pool=1)
    gpu_env = se.SyntheticSolo12Env(n, device=DEV, seed=2, pool=1)
    oracle = cat_oracle.ManagerOracle(cpu_env, cat_oracle.terms_from_cfg(se.solo12_constraints_cfg(), resolve_scene=cpu_env.scene))
    mgr = ConstraintManager(se.solo12_constraints_cfg(), gpu_env)
    rng = ops.make_rng_state(seed, DEV)
    reset = torch.zeros(n, dtype=torch.bool)
    reset[::97] = True
    offset = 0
    for _ in range(3):
        want_p = oracle.compute()
        _, want_dones = cat_oracle.step_epilogue(cpu_env._raw_reward, want_p, reset)
        mgr.compute_step(gpu_env._raw_reward, reset.to(DEV))
        mask, ids, count = mgr.sample_terminations(rng)
        want_mask = philox_oracle.bernoulli_mask(want_dones.numpy(), seed, offset)
        offset += n
        assert torch.equal(mask.cpu(), torch.from_numpy(want_mask))
        assert torch.equal(ids[: int(count)].cpu(), torch.from_numpy(want_mask).nonzero().flatten())
        assert mask.cpu()[reset].all()  # hard resets (dones == 1) always terminate
    assert rng.cpu().tolist() == [seed, 3 * n]
    # the env-level switch: hard 0 / 1 dones, every reset env included
    env = se.SyntheticSolo12Env(512, device=DEV, seed=0, pool=2, episode_length=5, constraints_cfg=se.solo12_constraints_cfg(), stochastic_terminations=True)
    env.load_managers()
    for _ in range(6):
        _, _, dones, time_outs, _ = env.step(torch.zeros(512, se.ACT_DIM, device=DEV))
        assert set(dones.unique().tolist()) <= {0.0, 1.0} and bool(dones[time_outs].eq(1).all())
