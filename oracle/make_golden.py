"""ORACLE tooling (build container only): generate tests/golden/ from the REAL reference.

Run:  python oracle/make_golden.py            (needs /root/reference; writes tests/golden/*.pt)

What it records
  cat_solo12.pt   the reference `ConstraintManager` (+ `CaT`, the 13 reference term functions and
                  `modify_constraint_p`) driven for 24 steps on seeded synthetic Solo12 state at N=256
                  with the max_p curriculum active and adversarial rows; two reset events.
  cat_stress.pt   the 16-term / 93-column stress layout, 6 steps, N=96 (ragged vs the 32-env tile).
  ppo_iter.pt     one full iteration of the reference `PPO()` trainer on a seeded fake env at N=64:
                  the rollout buffers, GAE outputs, value normalisation, RunningMeanStd states, the
                  minibatch permutations, the summed losses and the agent's state_dict before/after.
                  PPO() is a monolithic function, so its locals are captured with a line tracer.

Inputs are regenerated from seeds by `sample_state` (torch CPU generator), so fixtures hold outputs plus
a checksum of the inputs.  The reference itself is executed from /root/reference; nothing is copied.
"""

from __future__ import annotations

import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from constraints_as_terminations_b200 import synthetic_env as se  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def state_checksum(state: dict) -> float:
    return float(sum(v.double().sum().item() for v in state.values()))


def run_reference_cat(num_envs: int, steps: int, seed: int, stress: bool, reset_at: dict[int, list[int] | None]):
    ref = ref_loader.load_reference_cat()
    from constraints_as_terminations_b200._isaaclab_compat import SceneEntityCfg

    cfg = se.solo12_constraints_cfg(
        stress=stress,
        constraints_module=ref["constraints"],
        term_cls=ref["manager_constraint_cfg"].ConstraintTermCfg,
        scene_entity_cls=SceneEntityCfg,
    )
    env = se.SyntheticSolo12Env(num_envs, device="cpu", seed=seed, pool=1, adversarial=True)
    mgr = ref["constraint_manager"].ConstraintManager(cfg, env)
    mgr.cat._device = torch.device("cpu")
    env.constraint_manager = mgr
    gen = torch.Generator().manual_seed(seed + 1000)
    out = {"steps": [], "num_envs": num_envs, "seed": seed, "stress": stress, "names": list(mgr.active_terms)}
    for step in range(steps):
        state = se.sample_state(num_envs, gen, adversarial=(step % 3 == 0))
        env.load_state(state)
        env.episode_length_buf += 1
        env.common_step_counter = step * 400
        for name in mgr.active_terms:  # the curriculum mutates max_p between steps (curriculums.py:21-41)
            if name in se.SOLO12_CURRICULUM_TERMS:
                ref["curriculums"].modify_constraint_p(env, None, name, num_steps=24 * 1000, init_max_p=0.25)
        cstr_prob = mgr.compute()
        reset_buf = torch.rand(num_envs, generator=gen) < 0.05
        reward = torch.clip(state["raw_reward"] * (1.0 - cstr_prob), min=0.0, max=None)  # cat_env.py:102-106
        dones = cstr_prob.clone()
        ids = reset_buf.nonzero(as_tuple=False).squeeze(-1)
        if len(ids) > 0:
            dones[ids] = 1.0
        rec = {
            "checksum": state_checksum(state),
            "max_p": [float(mgr.get_term_cfg(n).max_p) for n in mgr.active_terms],
            "cstr_prob": cstr_prob.clone(),
            "running_max": mgr.cat.get_running_maxes().clone().squeeze(0),
            "reset_buf": reset_buf.clone(),
            "reward": reward,
            "dones": dones,
        }
        if step in (0, 1, steps - 1):
            rec["raw"] = mgr.cat.get_raw_constraints().clone()
            rec["probs"] = torch.cat(list(mgr.cat.probs.values()), dim=1).clone()
        if step in reset_at:
            sel = reset_at[step]
            env_ids = None if sel is None else torch.tensor(sel, dtype=torch.long)
            rec["reset_ids"] = sel
            rec["reset_out"] = {k: v.clone() for k, v in mgr.reset(env_ids).items()}
            if env_ids is None:
                env.episode_length_buf[:] = 0
            else:
                env.episode_length_buf[env_ids] = 0
            env.episode_length_buf.clamp_(min=0)
        out["steps"].append(rec)
    out["episode_sums"] = torch.stack([mgr._episode_sums[n] for n in mgr.active_terms]).clone()
    out["mean_values"] = torch.stack([mgr._cstr_mean_values[n] for n in mgr.active_terms]).clone()
    out["manager_str"] = str(mgr)
    return out


class _FakeTrainEnv:
    """Seeded stand-in for the gym env the reference trainer drives (ppo.py:158-161,186,215-226)."""

    def __init__(self, num_envs, seed):
        self.num_envs = num_envs
        self.unwrapped = self
        self.single_observation_space = {"policy": types.SimpleNamespace(shape=(se.OBS_DIM,))}
        self.single_action_space = types.SimpleNamespace(shape=(se.ACT_DIM,))
        self.gen = torch.Generator().manual_seed(seed)
        self.trace = []

    def _obs(self):
        scale = torch.linspace(0.2, 3.0, se.OBS_DIM)
        return {"policy": torch.randn(self.num_envs, se.OBS_DIM, generator=self.gen) * scale + 0.3}

    def reset(self):
        obs = self._obs()
        self.trace.append(("reset", obs["policy"].clone()))
        return obs, {}

    def step(self, action):
        obs = self._obs()
        reward = torch.rand(self.num_envs, generator=self.gen) * 0.05
        dones = torch.rand(self.num_envs, generator=self.gen) * (torch.rand(self.num_envs, generator=self.gen) < 0.3)
        hard = torch.rand(self.num_envs, generator=self.gen) < 0.03
        dones = torch.where(hard, torch.ones_like(dones), dones)
        timeouts = torch.rand(self.num_envs, generator=self.gen) < 0.02
        self.trace.append(("step", obs["policy"].clone(), reward.clone(), dones.clone(), timeouts.clone()))
        return obs, reward, dones, timeouts, {"log": {}}


def run_reference_ppo(num_envs=64, seed=7, minibatch=512, epochs=2, run_path="/tmp/catb200_golden_run"):
    ppo = ref_loader.load_reference_ppo()
    src_lines = open(ppo.__file__).read().splitlines()

    def line_of(text, nth=0):
        hits = [i + 1 for i, ln in enumerate(src_lines) if text in ln]
        return hits[nth]

    at_init = line_of("obs = torch.zeros(")  # first statement after the optimizer exists
    at_update = line_of("sum_pg_loss = sum_entropy_loss")  # after GAE and both value_rms calls
    at_mb = line_of("mb_inds = b_inds[start:end]")
    at_end = line_of("num_updates = UPDATES_EPOCHS")

    cfg = types.SimpleNamespace(
        logger="tensorboard", learning_rate=3.0e-4, num_steps=24, num_iterations=1, gamma=0.99, gae_lambda=0.95,
        updates_epochs=epochs, minibatch_size=minibatch, clip_coef=0.2, ent_coef=0.001, vf_coef=2.0,
        max_grad_norm=1.0, norm_adv=True, clip_vloss=True, anneal_lr=True, save_interval=1000,
    )  # fmt: skip
    cap = {"perms": []}

    def clone_sd(agent):
        return {k: v.detach().clone() for k, v in agent.state_dict().items()}

    def tracer(frame, event, arg):
        if frame.f_code.co_name != "PPO":
            return None

        def local(frame, event, arg):
            if event != "line":
                return local
            lo = frame.f_locals
            if frame.f_lineno == at_init and "init_state" not in cap:
                cap["init_state"] = clone_sd(lo["agent"])
            elif frame.f_lineno == at_update:
                for k in ("obs", "actions", "logprobs", "rewards", "dones", "true_dones", "values", "advantages",
                          "returns", "next_value", "next_done", "next_true_done", "b_values", "b_returns", "next_obs"):  # fmt: skip
                    cap[k] = lo[k].detach().clone()
                cap["rms_after_rollout"] = {k: v for k, v in clone_sd(lo["agent"]).items() if "_rms." in k}
            elif frame.f_lineno == at_mb and lo["start"] == 0:
                cap["perms"].append(lo["b_inds"].clone())
            elif frame.f_lineno == at_end:
                cap["final_state"] = clone_sd(lo["agent"])
                for k in ("sum_pg_loss", "sum_entropy_loss", "sum_v_loss", "sum_surrogate_loss"):
                    cap[k] = float(lo[k])
                cap["clipfracs"] = list(lo["clipfracs"])
                cap["lr"] = lo["optimizer"].param_groups[0]["lr"]
            return local

        return local

    torch.manual_seed(seed)
    env = _FakeTrainEnv(num_envs, seed)
    sys.settrace(tracer)
    try:
        ppo.PPO(env, cfg, run_path)
    finally:
        sys.settrace(None)
    cap["env_trace_reset_obs"] = env.trace[0][1]
    cap["env_trace_raw_obs"] = torch.stack([t[1] for t in env.trace[1:]])
    cap["env_trace_timeouts"] = torch.stack([t[4] for t in env.trace[1:]])
    cap["cfg"] = vars(cfg)
    cap["num_envs"] = num_envs
    cap["seed"] = seed
    return cap


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(1)  # fixed reduction order for the fixtures
    cat = run_reference_cat(256, 24, seed=0, stress=False, reset_at={9: [3, 17, 200, 255], 20: None})
    torch.save(cat, os.path.join(GOLDEN, "cat_solo12.pt"))
    stress = run_reference_cat(96, 6, seed=3, stress=True, reset_at={4: [0, 95]})
    torch.save(stress, os.path.join(GOLDEN, "cat_stress.pt"))
    ppo = run_reference_ppo()
    torch.save(ppo, os.path.join(GOLDEN, "ppo_iter.pt"))
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
