"""`CaTEnv` (constraints_as_terminations_b200/cat_env.py) executed for real on top of a test double of Isaac Lab's
ManagerBasedRLEnv (tests/isaaclab_stub.py): constructor -> load_managers -> step -> _reset_idx, compared step by step with
the CPU oracle of the reference's `CaTEnv.step` CaT block (U/cat/cat_env.py:98-121,181-182)."""

import types

import pytest
import torch

from constraints_as_terminations_b200 import synthetic_env as se
from oracle import cat_oracle
from tests import isaaclab_stub

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_cat_env_step_and_reset_match_the_oracle():
    cat_env = isaaclab_stub.install()
    try:
        n, steps = 300, 40
        cfg = types.SimpleNamespace(num_envs=n, device=DEV, seed=4, pool=3, episode_length=12, decimation=4,
                                    constraints=se.solo12_constraints_cfg(), sim=types.SimpleNamespace(render_interval=4), rerender_on_reset=False)  # fmt: skip
        env = cat_env.CaTEnv(cfg)
        assert cat_env.HAVE_ISAACLAB_ENV and hasattr(env, "constraint_manager")
        assert len(env.constraint_manager.active_terms) == 13
        # CPU twin: same synthetic states, the oracle manager, the reference's step lines
        cpu = se.SyntheticSolo12Env(n, device="cpu", seed=4, pool=3, episode_length=12, curriculum=False)
        oracle = cat_oracle.ManagerOracle(cpu, cat_oracle.terms_from_cfg(se.solo12_constraints_cfg(), resolve_scene=cpu.scene))
        resets_seen = 0
        for step in range(steps):
            # the double does not move the state by itself: advance both twins the same way before the step
            env._advance()
            cpu._advance()
            obs, reward, dones, time_outs, extras = env.step(torch.zeros(n, se.ACT_DIM, device=DEV))
            assert env.sim.steps == (step + 1) * cfg.decimation
            cpu.episode_length_buf += 1
            reset = cpu.episode_length_buf >= cpu.max_episode_length
            cstr = oracle.compute()
            want_reward, want_dones = cat_oracle.step_epilogue(cpu._raw_reward, cstr, reset)
            assert torch.equal(reward.cpu(), want_reward), f"step {step}: reward"
            assert torch.equal(dones.cpu(), want_dones), f"step {step}: dones"
            assert torch.equal(time_outs.cpu(), reset), f"step {step}: time outs"
            assert dones.dtype == torch.float32 and obs["policy"].shape == (n, se.OBS_DIM)
            if bool(reset.any()):
                resets_seen += 1
                ids = reset.nonzero().flatten()
                want_log = oracle.reset(ids)
                cpu.episode_length_buf[ids] = 0
                log = extras["log"]
                assert set(want_log) <= set(log)
                for k, v in want_log.items():
                    torch.testing.assert_close(torch.as_tensor(log[k]).detach().cpu().reshape(()), v.reshape(()), rtol=1e-5, atol=1e-7, msg=lambda m, k=k: f"step {step} {k}: {m}")
                assert int(env.episode_length_buf[ids.to(DEV)].abs().sum()) == 0
        assert resets_seen >= 3
        # a reset from outside step() (env.reset() in Isaac Lab) takes the non-fused path; envs whose episode length is 0
        # make the reference's mean NaN (0 / 0, constraint_manager.py:196-204) and ours alike
        ids = torch.arange(0, n, 7, device=DEV)
        want = oracle.reset(ids.cpu())
        env._reset_idx(ids)
        for k, v in want.items():
            torch.testing.assert_close(torch.as_tensor(env.extras["log"][k]).detach().cpu().reshape(()), v.reshape(()), rtol=1e-5, atol=1e-7, equal_nan=True)
    finally:
        isaaclab_stub.uninstall()
