"""ORACLE tooling: run the UNMODIFIED reference trainer (`PPO()` of U/cleanrl/ppo.py) with the UNMODIFIED reference
`ConstraintManager` / `CaT` / term functions / curriculum on the synthetic Solo12 state source, and time it.

This is the reference arm of bench.py (`--impl reference`: on the host cores) and its `gpu_eager_baseline` leg (the
same reference code with CUDA tensors on the B200, TF32 matmuls enabled as scripts/clean_rl/train.py:86-87 does).
Only tests/, __graft_entry__.smoke() and bench.py may import it; the product path never does.

The one piece of the reference that cannot be imported is `CaTEnv` (it subclasses Isaac Lab's ManagerBasedRLEnv, which
is not installed); `RefCaTEnv.step` below restates its CaT block (U/cat/cat_env.py:98-121,181-182) line for line around
the synthetic state, calling the reference manager's own `compute()` / `reset()`.
"""

from __future__ import annotations

import contextlib
import os
import shutil
import sys
import tempfile
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from constraints_as_terminations_b200 import synthetic_env as se  # noqa: E402
from oracle import ref_loader  # noqa: E402


def make_ref_env(num_envs, device, seed, pool=8, stress=False, episode_length=500):
    """Synthetic Solo12 state source + the reference's own constraint manager, stepped like CaTEnv.step."""
    ref = ref_loader.load_reference_cat()
    from constraints_as_terminations_b200._isaaclab_compat import SceneEntityCfg

    cfg = se.solo12_constraints_cfg(
        stress=stress, constraints_module=ref["constraints"], term_cls=ref["manager_constraint_cfg"].ConstraintTermCfg,
        scene_entity_cls=SceneEntityCfg,
    )  # fmt: skip
    modify_constraint_p = ref["curriculums"].modify_constraint_p

    class RefCaTEnv(se.SyntheticSolo12Env):
        def load_managers(self):  # cat_env.py:38-40
            self.constraint_manager = ref["constraint_manager"].ConstraintManager(self.cfg.constraints, self)
            return self.constraint_manager

        def step(self, action):
            self._advance()
            self.episode_length_buf += 1
            self.common_step_counter += 1
            self.reset_time_outs = self.episode_length_buf >= self.max_episode_length
            self.reset_buf = self.reset_time_outs
            cstr_prob = self.constraint_manager.compute()  # cat_env.py:100
            self.reward_buf = torch.clip(self._raw_reward * (1.0 - cstr_prob), min=0.0, max=None)  # :102-106
            dones = cstr_prob.clone()  # :107
            reset_env_ids = self.reset_buf.nonzero(as_tuple=False).squeeze(-1)  # :118
            if len(reset_env_ids) > 0:  # :120 (a device->host sync, as in the reference)
                dones[reset_env_ids] = 1.0  # :121
                self.extras["log"] = dict()
                for name in self.constraint_manager.active_terms:  # CurriculumCfg of the task (cat_flat_env_cfg.py:384-)
                    if name in se.SOLO12_CURRICULUM_TERMS:
                        modify_constraint_p(self, reset_env_ids, name, num_steps=24 * 1000, init_max_p=0.25)
                self.extras["log"].update(self.constraint_manager.reset(reset_env_ids))  # :181-182
                self.episode_length_buf[reset_env_ids] = 0
            return self.obs_buf, self.reward_buf, dones, self.reset_time_outs, self.extras

    env = RefCaTEnv(num_envs, device=device, seed=seed, pool=pool, constraints_cfg=cfg, episode_length=episode_length)
    env.load_managers()
    return env


@contextlib.contextmanager
def _device_choice(use_cuda: bool):
    """The reference picks its device with `torch.cuda.is_available()` (ppo.py:163, constraint_manager.py:32): answer
    for it, for the duration of the run, without touching its code."""
    orig = torch.cuda.is_available
    torch.cuda.is_available = lambda: use_cuda
    try:
        yield
    finally:
        torch.cuda.is_available = orig


def reference_ppo_cfg(num_iterations: int):
    """S12/agents/clean_rl_ppo_cfg.py:12-34 (restated by the package's solo12_flat_ppo_cfg; pinned against the reference
    file by tests/test_reference_cfg.py)."""
    from constraints_as_terminations_b200 import solo12_flat_ppo_cfg

    cfg = solo12_flat_ppo_cfg(logger="tensorboard", num_iterations=num_iterations)
    cfg.save_interval = 10**9
    return cfg


def run_reference_trainer(num_envs: int, steps: int, warmup: int, device: str = "cpu", seed: int = 0, tf32: bool = True) -> dict:
    """`PPO(envs, cfg, run_path)` for warmup + steps iterations; returns the wall time of the last `steps` iterations
    (from the first env step of iteration warmup + 1 to the return of PPO(), device-synchronised on both ends)."""
    ppo = ref_loader.load_reference_ppo()
    dev = torch.device(device)
    use_cuda = dev.type == "cuda"
    if use_cuda:  # scripts/clean_rl/train.py:86-87
        torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
        torch.backends.cudnn.allow_tf32 = bool(tf32)
    cfg = reference_ppo_cfg(warmup + steps)
    T = int(cfg.num_steps)
    run_path = tempfile.mkdtemp(prefix="cat_ref_run_")
    marks = {}
    with _device_choice(use_cuda):
        env = make_ref_env(num_envs, dev, seed)
        inner_step = env.step
        count = {"n": 0}

        def timed_step(action):
            if count["n"] == warmup * T:
                if use_cuda:
                    torch.cuda.synchronize()
                marks["t0"] = time.perf_counter()
            count["n"] += 1
            return inner_step(action)

        env.step = timed_step
        torch.manual_seed(seed)
        stdout = sys.stdout
        sys.stdout = open(os.devnull, "w")  # PPO() prints per save; keep the bench's single JSON line clean
        try:
            ppo.PPO(env, cfg, run_path)
        finally:
            sys.stdout.close()
            sys.stdout = stdout
        if use_cuda:
            torch.cuda.synchronize()
        marks["t1"] = time.perf_counter()
    shutil.rmtree(run_path, ignore_errors=True)
    assert count["n"] == (warmup + steps) * T, "the reference trainer did not take the expected number of env steps"
    seconds = marks["t1"] - marks["t0"]
    return {
        "seconds": seconds, "seconds_per_iteration": seconds / steps, "env_steps_per_sec": num_envs * T * steps / seconds,
        "iterations": steps, "warmup": warmup, "num_envs": num_envs, "device": str(dev),
        "reference_root": "oracle/_ref/ref_hotpath.tar.gz" if ref_loader.REFERENCE_ROOT.startswith(tempfile.gettempdir()) else ref_loader.REFERENCE_ROOT,
    }  # fmt: skip


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    print(run_reference_trainer(n, steps=2, warmup=1, device=sys.argv[2] if len(sys.argv) > 2 else "cpu"))
