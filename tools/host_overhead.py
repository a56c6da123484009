"""Host-side cost (perf_counter, no syncs) of the pieces of one synthetic env step + manager call."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from constraints_as_terminations_b200 import _lib as L

env, tr = bench.make_trainer(4096, torch.device("cuda", 0), seed=0)
for _ in range(2):
    tr.train_iteration()
torch.cuda.synchronize()
mgr = env.constraint_manager
acc = {}
def t(name, fn, reps=300):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    acc[name] = (time.perf_counter() - t0) / reps * 1e6
    torch.cuda.synchronize()

t("advance(load_state)", env._advance)
t("ep_len += 1", lambda: env.episode_length_buf.__iadd__(1))
t("reset_time_outs = ep_len >= max", lambda: env.episode_length_buf >= env.max_episode_length)
t("numpy phase", lambda: ((env._phase_np.__iadd__(0)) >= 500).any())
t("mgr._param_values", mgr._param_values)
t("mgr._built.refresh", lambda: mgr._built.refresh(env))
t("mgr._refresh_max_p", mgr._refresh_max_p)
t("mgr._ensure_plan", mgr._ensure_plan)
rb = env.reset_buf
t("mgr.compute_step", lambda: mgr.compute_step(env._raw_reward, rb))
t("mgr.compute", mgr.compute)
t("curriculum x8", env._curriculum)
t("mgr.reset_masked", lambda: mgr.reset_masked(rb))
t("masked_fill_", lambda: env.episode_length_buf.masked_fill_(rb, 0))
t("env.step (all)", lambda: env.step(tr.actions[0]))
t("L.stream()", L.stream)
t("policy (graph replay)", lambda: tr._policy(0))
t("rollout_append", lambda: __import__("constraints_as_terminations_b200").ops.rollout_append(env.reward_buf, tr.dones[1], env.reset_time_outs, tr.rewards[0], tr.dones[1], tr.true_dones[1], validate=False))
t("ingest_obs", lambda: tr._ingest_obs(env.obs_buf["policy"], 1))
for k, v in acc.items():
    print(f"{k:36s} {v:8.1f} us")
