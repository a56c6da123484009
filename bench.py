#!/usr/bin/env python
"""Benchmark of the CaT PPO hot path (BASELINE.json metric: env-steps/sec, Solo12 CaT-Flat @4096 envs/GPU).

    python bench.py --gpus 1 --steps 20 --warmup 5                     # this repo's CUDA path (1 GPU)
    torchrun --nproc-per-node N ... bench.py --gpus N --steps K --warmup W   # one rank per GPU, NCCL
    python bench.py --impl reference --steps K --warmup W              # reference arm: CPU oracle port

A "step" is one full PPO iteration on the trainer side of Isaac-Velocity-CaT-Flat-Solo12-v0 with the
reference's hyper-parameters: 24 env steps x num_envs (policy forward + sampling, the 13-term constraint
manager with reward / dones epilogue, rollout append, running observation normalisation), then GAE, value
normalisation and 5 epochs x 6 minibatches of PPO-clip update (forward, backward, clip, Adam) -- nothing
skipped.  Isaac Sim is not installed here, so physics is replaced by a synthetic Solo12 state source with
the same tensors (`constraints_as_terminations_b200/synthetic_env.py`); the number is trainer-side
env-steps/s and says so in `data`.

Prints ONE JSON line (rank 0).  `value` has the env state already resident in HBM; `e2e` feeds every env
step's state from pinned host memory (H2D inside the timed region) and reads the losses back (D2H).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "40", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )  # fmt: skip
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        mx = next((float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()), None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------------------
def make_trainer(num_envs, device, seed, host_fed=False, pool=8, graphs=True, distributed=True):
    from constraints_as_terminations_b200 import PPOTrainer, solo12_flat_ppo_cfg
    from constraints_as_terminations_b200 import synthetic_env as se

    cls = HostFedEnv if host_fed else se.SyntheticSolo12Env
    env = cls(num_envs, device=device, seed=seed, pool=pool, constraints_cfg=se.solo12_constraints_cfg())
    env.load_managers()
    cfg = solo12_flat_ppo_cfg(logger=None)
    torch.manual_seed(cfg.seed + seed)
    trainer = PPOTrainer(env, cfg, device=device, use_graphs=graphs, distributed=distributed)
    trainer.start()
    return env, trainer


def _host_fed_env_cls():
    from constraints_as_terminations_b200 import synthetic_env as se

    class HostFed(se.SyntheticSolo12Env):
        """Same env, but every step's state arrives from pinned host memory (the e2e leg).  The states of one rollout
        (24 env steps) sit back to back in one pinned buffer and move with ONE cudaMemcpyAsync per iteration on a copy
        stream into one of two device staging sets: while the kernels of iteration i read set i & 1, the copy for
        iteration i + 1 fills the other set.  Every step's inputs therefore cross PCIe inside the timed region, but the
        per-step host cost is a pointer swap.  The synthetic state does not depend on the actions, which is what makes
        reading ahead legitimate here; with a real simulator the state is produced on the device and there is no such
        copy at all."""

        STEPS = 24  # env steps per staging set = one rollout

        def __init__(self, num_envs, device, seed, pool, constraints_cfg):
            super().__init__(num_envs, device=device, seed=seed, pool=pool, constraints_cfg=constraints_cfg)
            # packed layout of one state: the state tensors are views into the packed buffers
            layout, off = {}, 0
            for k, v in self._pool[0].items():
                layout[k] = (off, v.numel() * v.element_size(), v.dtype, tuple(v.shape))
                off += (layout[k][1] + 255) // 256 * 256
            self._packed_bytes = off

            def views(buf):
                return {k: buf[o : o + n].view(dt).view(shape) for k, (o, n, dt, shape) in layout.items()}

            steps = self.STEPS
            self._host = torch.empty((steps, off), dtype=torch.uint8).pin_memory()  # state of rollout step k at row k
            for k in range(steps):
                st = self._pool[(k + 1) % len(self._pool)]  # step k of a rollout follows the pool like the resident env
                for name, dst in views(self._host[k]).items():
                    dst.copy_(st[name].cpu())
            first = {name: v.clone() for name, v in self._pool[0].items()}  # the state reset() exposes
            self._staging_buf = [torch.empty((steps, off), dtype=torch.uint8, device=device) for _ in range(2)]
            self._staging = [[views(b[k]) for k in range(steps)] for b in self._staging_buf]
            self._pool = None  # nothing stays resident on the device except the two staging sets
            self.h2d_bytes = sum(n for _, n, _, _ in layout.values())
            self._copy_stream = torch.cuda.Stream(device=device)
            self._ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._consumed = [torch.cuda.Event(), torch.cuda.Event()]
            self._count = 0  # env steps taken so far
            self.load_state(first)
            self._prefetch(0)

        def _prefetch(self, buf):
            """Enqueue the H2D copy of one rollout's states into staging set `buf` on the copy stream."""
            main = torch.cuda.current_stream()
            self._consumed[buf].record(main)  # the copy must not overwrite data that enqueued kernels still read
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(self._consumed[buf])
                self._staging_buf[buf].copy_(self._host, non_blocking=True)
                self._ready[buf].record(self._copy_stream)

        def reset(self):
            return self.obs_buf, {}

        def _advance(self):
            it, k = divmod(self._count, self.STEPS)
            buf = it & 1
            if k == 0:  # first step of a rollout: its states were requested one iteration ago; request the next ones
                torch.cuda.current_stream().wait_event(self._ready[buf])
                self._prefetch(buf ^ 1)
            self.load_state(self._staging[buf][k])
            self._count += 1

    return HostFed


HostFedEnv = None


def timed_iterations(trainer, steps, warmup, world, device, read_losses):
    for _ in range(warmup):
        trainer.train_iteration()
        if read_losses:
            trainer.losses()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    launches0 = trainer.kernel_launches()
    sampler = ClockSampler(device.index or 0)
    sampler.start()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    for _ in range(steps):
        trainer.train_iteration()
        if read_losses:
            trainer.losses()  # device -> host read of the step's result
    end.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(end)
    if world > 1:
        torch.distributed.barrier()
        t = torch.tensor([ms], device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t)
    clocks = sampler.stop()
    return ms, trainer.kernel_launches() - launches0, clocks


def kernel_profile(trainer, iters=2, record=True):
    """Device time of every kernel over `iters` full iterations of the timed workload (CUPTI timestamps via
    torch.profiler, CUDA graphs included): name -> {us, launches, share}.  Iterations contain the gradient
    all-reduce, so under torchrun EVERY rank must call this; only ranks with record=True profile."""
    from torch.profiler import ProfilerActivity, profile

    torch.cuda.synchronize()
    if not record:
        for _ in range(iters):
            trainer.train_iteration()
        torch.cuda.synchronize()
        return {}, 0.0
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(iters):
            trainer.train_iteration()
        torch.cuda.synchronize()
    rows = {}
    for ev in prof.key_averages():
        t = getattr(ev, "device_time_total", None) or getattr(ev, "cuda_time_total", 0.0)
        if t <= 0:
            continue
        name = ev.key.split("(")[0].replace("void ", "").replace("catb200::", "")
        rows[name] = {"us": t / iters, "launches": ev.count / iters}
    total = sum(r["us"] for r in rows.values())
    for r in rows.values():
        r["share"] = r["us"] / total
    return dict(sorted(rows.items(), key=lambda kv: -kv[1]["us"])), total


# tensor-core flops per sample of the three tcgen05 GEMM modes (both nets, obs padded to 64): SURVEY.md §8d
_MAC_FWD = 2 * (64 * 512 + 512 * 256 + 256 * 128)   # also the weight-gradient GEMMs
_MAC_DGRAD = 2 * (256 * 128 + 512 * 256)


def _traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures."""
    path = os.path.join(ROOT, "profiles", "r1b_traffic.json")
    return json.load(open(path)) if os.path.isfile(path) else {}


def _gemm_flops_by_kernel(trainer):
    """Tensor-core flops per iteration attributed to each tcgen05 kernel name.  Forward / dgrad launches with more
    than 2 x 148 output tiles run the persistent kernel, the others (rollout-sized launches, the 256 -> 128 layer)
    the one-tile-per-CTA kernel (tc_gemm.cu: use_persistent)."""
    n, T = trainer.num_envs, trainer.T
    opt_rows = trainer.batch_size * int(trainer.cfg.updates_epochs)  # minibatch rows per iteration
    mb = trainer.minibatch_size
    layers = [(64, 512), (512, 256), (256, 128)]  # (K, N) of the hidden layers, both nets batched per launch
    flops = {}

    def add(kind, rows_per_launch, total_rows, k, n_out):
        tiles = 2 * ((rows_per_launch + 127) // 128) * (n_out // 128)
        name = f"tc_gemm_persist_kernel<{kind}, 128>" if tiles > 2 * 148 else f"tc_gemm_kernel<{kind}, 128>"
        flops[name] = flops.get(name, 0.0) + 2.0 * 2 * k * n_out * total_rows

    for k, n_out in layers:
        add(0, mb, opt_rows, k, n_out)          # update forward
        add(0, n, n * (T + 1), k, n_out)        # rollout policy + bootstrap value
    for k, n_out in layers[1:]:
        add(1, mb, opt_rows, n_out, k)          # dgrad: dZ_l [M, N_l] x W_l -> [M, K_l]
    return flops


def dominant_kernel_roofline(prof, total_us, trainer, peaks):
    """Roofline entry of the kernel with the largest share of the iteration."""
    flops = _gemm_flops_by_kernel(trainer)
    if not prof:  # no CUPTI records (e.g. the process itself runs under ncu, which owns the profiling interface)
        return {"kernel": None, "note": "kernel shares unavailable: CUPTI produced no records in this process"}
    name, row = next(iter(prof.items()))
    out = {"kernel": name, "share_of_step": row["share"], "us_per_step": row["us"], "launches_per_step": row["launches"]}
    if name in flops:
        peak = peaks.get("bf16_tflops_sustained") or peaks["bf16_tflops"]
        ach = flops[name] / (row["us"] * 1e-6) / 1e12
        out.update({"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "flops_per_step": flops[name], "traffic": _traffic().get(name),
                    "peak_source": peaks["source"] + " (sustained bf16: the kernel runs inside a long step)"})
    return out


def kernel_rooflines(device, num_envs, peaks):
    """Per-kernel achieved bandwidth / throughput, timed alone with CUDA events on the launch stream,
    L2 flushed between timed launches (a 256 MiB write)."""
    from constraints_as_terminations_b200 import ops
    from constraints_as_terminations_b200 import synthetic_env as se

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    out = {}

    def time_kernel(fn, reps=20):
        for _ in range(3):
            fn()
        times = []
        for _ in range(reps):
            flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            times.append(s.elapsed_time(e) * 1e-3)
        times.sort()
        return sum(times[: max(1, len(times) // 2)]) / max(1, len(times) // 2)  # mean of the faster half

    T = 24
    for n in sorted({num_envs, 65536, 1 << 20}):
        # GAE: 24*T*N + 12*N algorithmic bytes
        rewards, values = torch.rand(T, n, device=device), torch.randn(T, n, device=device)
        dones, tdones = torch.rand(T + 1, n, device=device), torch.zeros(T + 1, n, device=device)
        nv = torch.randn(n, device=device)
        adv, ret = torch.empty_like(rewards), torch.empty_like(rewards)
        sec = time_kernel(lambda: ops.gae(rewards, values, dones, tdones, nv, 0.99, 0.95, advantages=adv, returns=ret))
        nbytes = 24 * T * n + 12 * n
        out[f"gae@{n}"] = {"bound": "hbm", "achieved": nbytes / sec / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": nbytes / sec / 1e9 / peaks["hbm_gbs"], "us": sec * 1e6, "bytes": nbytes}
        del rewards, values, dones, tdones, adv, ret
        # fused constraint step (cat_eval + cat_apply): 940 B/env-step algorithmic
        env = se.SyntheticSolo12Env(n, device=device, seed=1, pool=1, constraints_cfg=se.solo12_constraints_cfg())
        mgr = env.load_managers()
        reset = torch.zeros(n, dtype=torch.bool, device=device)
        sec = time_kernel(lambda: mgr.compute_step(env._raw_reward, reset))
        nbytes = 940 * n
        out[f"cat_step@{n}"] = {"bound": "hbm", "achieved": nbytes / sec / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": nbytes / sec / 1e9 / peaks["hbm_gbs"], "us": sec * 1e6, "bytes": nbytes, "launches": 2}
        del env, mgr
    # one PPO optimizer step on a 16384-row minibatch (gather, 3 fwd + 2 dgrad + 3 wgrad tcgen05 GEMMs, head/loss,
    # reduce, clip + Adam + weight cast): 2.2525 MFLOP/sample of tensor work
    env, tr = make_trainer(num_envs, device, seed=0, graphs=False, distributed=False)  # rank-local probe
    tr.train_iteration()
    mb = tr.minibatch_size
    perm = torch.randperm(tr.batch_size, device=device)
    sec = time_kernel(lambda: tr._minibatch(perm[:mb]), reps=10)
    flops = 2.2525e6 * mb
    out[f"ppo_minibatch_step@{mb}"] = {"bound": "tensor", "achieved": flops / sec / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": flops / sec / 1e12 / peaks["bf16_tflops"], "us": sec * 1e6, "flops": flops, "launches": 14}
    del env, tr
    # DRAM traffic per launch from the committed `ncu --set full` captures (profiles/), where available
    tpath = os.path.join(ROOT, "profiles", "r1b_traffic.json")
    if os.path.isfile(tpath):
        for k, v in json.load(open(tpath)).items():
            if k in out:
                out[k]["traffic"] = v
    return out


def run_ours(args):
    global HostFedEnv
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)
    import __graft_entry__

    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        torch.distributed.barrier()
    HostFedEnv = _host_fed_env_cls()
    peaks = load_peaks()
    N, T = args.envs, 24

    env, trainer = make_trainer(N, device, seed=rank, graphs=not args.no_graphs)
    ms, launches, clocks = timed_iterations(trainer, args.steps, args.warmup, world, device, read_losses=False)
    value = N * T * world * args.steps / (ms * 1e-3)
    losses = trainer.losses()
    prof, prof_total = kernel_profile(trainer, record=(rank == 0))
    dominant = dominant_kernel_roofline(prof, prof_total, trainer, peaks) if rank == 0 else None
    working_set = sum(t.numel() * t.element_size() for t in (trainer.obs, trainer.obs_op, trainer.actions, trainer.rewards, trainer.dones, trainer.values, trainer.advantages, trainer.returns, trainer.train_ws))
    del env, trainer
    torch.cuda.empty_cache()

    # ---- e2e: host-resident env state, H2D each env step, D2H of the losses each iteration
    env, trainer = make_trainer(N, device, seed=rank, host_fed=True, graphs=not args.no_graphs)
    e_steps = max(3, args.steps // 2)
    e_ms, _, _ = timed_iterations(trainer, e_steps, max(3, args.warmup // 2), world, device, read_losses=True)
    e2e = {
        "value": N * T * world * e_steps / (e_ms * 1e-3),
        "unit": UNIT,
        "h2d_bytes_per_step": env.h2d_bytes * T,
        "d2h_bytes_per_step": 32,
        "ms_per_step": e_ms / e_steps,
        "feed": "per iteration one cudaMemcpyAsync of the rollout's 24 packed env states from pinned host memory on a copy "
                "stream into one of two device staging sets (read one iteration ahead), and one blocking read of the losses",
    }
    del env, trainer
    torch.cuda.empty_cache()

    line = None
    if rank == 0:
        roof = kernel_rooflines(device, N, peaks)
        cpu = cpu_baseline(N, sample_steps=4, sample_minibatches=2)
        main = dominant
        main["note"] = main.get("note") or ("dominant kernel of the step by device time (torch.profiler/CUPTI over 2 iterations of the timed workload); "
                        "HBM-bound kernels (GAE, CaT) timed alone with CUDA events are in `rooflines`")
        gae_main = roof[f"gae@{N}"]
        gae_main["note"] = f"{gae_main['bytes']/1e6:.2f} MB per launch: launch-latency bound at {N} envs; see the 65536 / 1M-env entries"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32 (CaT/GAE/moments/loss/Adam), bf16 operands + fp32 accumulate (hidden-layer GEMMs)",
            "data": "synthetic Solo12 state (no Isaac Sim physics): trainer-side env-steps/s",
            "config": {
                "workload": "Isaac-Velocity-CaT-Flat-Solo12-v0 trainer side, 4096 envs/GPU, CleanRL PPO cfg (T=24, 5 epochs x 6 minibatches of 16384)",
                "envs_per_gpu": N, "num_steps": T, "constraint_terms": 13, "constraint_columns": 78,
                "minibatch": 16384, "epochs": 5, "parallelism": f"dp{world} (one env shard per GPU, 1 gradient allreduce per optimizer step)",
                "l2": f"per-step working set {working_set/1e6:.0f} MB > 126 MB L2 (inputs larger than L2)",
                "cuda_graphs": not args.no_graphs,
            },
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": main,
            "rooflines": roof,
            "kernel_shares": {k: {"share": round(v["share"], 4), "us_per_step": round(v["us"], 1), "launches_per_step": v["launches"]} for k, v in list(prof.items())[:16]},
            "cpu_baseline": cpu,
            "losses": losses,
        }  # fmt: skip
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, timed on the host cores
# ----------------------------------------------------------------------------------------------------------
class CpuReferencePath:
    """The reference's CPU code path restated by oracle/ (the reference itself is Python on Isaac Lab and
    cannot travel to the GPU box): per env step the 13-term ConstraintManager + reward/dones, Agent
    sampling, obs RunningMeanStd; per update GAE, value normalisation and PPO-clip minibatches with Adam."""

    def __init__(self, num_envs, seed=0):
        from constraints_as_terminations_b200 import synthetic_env as se
        from oracle import cat_oracle, ppo_oracle

        self.se, self.cat_oracle, self.po = se, cat_oracle, ppo_oracle
        torch.manual_seed(seed)
        self.N, self.T = num_envs, 24
        self.env = se.SyntheticSolo12Env(num_envs, device="cpu", seed=seed, pool=4)
        cfg = se.solo12_constraints_cfg()
        self.mgr = cat_oracle.ManagerOracle(self.env, cat_oracle.terms_from_cfg(cfg, resolve_scene=self.env.scene))
        self.agent = ppo_oracle.AgentOracle(se.OBS_DIM, se.ACT_DIM)
        params = list(self.agent.critic.parameters()) + list(self.agent.actor_mean.parameters()) + [self.agent.actor_logstd]
        self.params = params
        self.opt = torch.optim.Adam(params, lr=3e-4, eps=1e-5)
        self.obs_rms = ppo_oracle.rms_init((se.OBS_DIM,))
        self.value_rms = ppo_oracle.rms_init(())
        N, T = self.N, self.T
        self.obs = torch.zeros(T, N, se.OBS_DIM)
        self.actions = torch.zeros(T, N, se.ACT_DIM)
        self.logprobs, self.rewards, self.values = torch.zeros(T, N), torch.zeros(T, N), torch.zeros(T, N)
        self.dones, self.true_dones = torch.zeros(T + 1, N), torch.zeros(T + 1, N)
        self.next_obs = self._norm(self.env.obs_buf["policy"])

    def _norm(self, raw):
        self.obs_rms = self.po.rms_update(self.obs_rms, raw)
        return self.po.rms_normalize(self.obs_rms, raw)

    def env_step(self, t):
        env = self.env
        self.obs[t] = self.next_obs
        with torch.no_grad():
            action, logp, value = self.agent.act(self.next_obs, torch.randn(self.N, self.se.ACT_DIM))
        self.actions[t], self.logprobs[t], self.values[t] = action, logp, value.flatten()
        env._advance()
        env.episode_length_buf += 1
        reset = env.episode_length_buf >= env.max_episode_length
        cstr = self.mgr.compute()
        reward, dones = self.cat_oracle.step_epilogue(env._raw_reward, cstr, reset)
        env.episode_length_buf[reset] = 0
        self.rewards[t], self.dones[t + 1], self.true_dones[t + 1] = reward, dones, reset.float()
        self.next_obs = self._norm(env.obs_buf["policy"])

    def gae(self):
        with torch.no_grad():
            nv = self.agent.critic(self.next_obs).reshape(1, -1)
            adv, ret = self.po.gae(self.rewards, self.values, self.dones[:-1], self.true_dones[:-1], nv, self.dones[-1], self.true_dones[-1])
            self.value_rms, self.b_values, self.b_returns = self.po.value_normalisation(self.value_rms, self.values.reshape(-1), ret.reshape(-1))
        self.b_adv = adv.reshape(-1)

    def minibatch(self, idx):
        B = self.N * self.T
        loss, _ = self.po.ppo_minibatch_loss(
            self.agent, self.value_rms, self.obs.reshape(B, -1)[idx], self.actions.reshape(B, -1)[idx], self.logprobs.reshape(-1)[idx],
            self.b_adv[idx], self.b_returns[idx], self.b_values[idx],
        )  # fmt: skip
        self.opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.params, 1.0)
        self.opt.step()


def pick_threads(num_envs):
    """torch's CPU ops on these small tensors do not scale to every core of a big host (128 threads make
    one env step ~100x slower than 8-16 do), so give the CPU arm the thread count that is fastest for it:
    time one env step + one minibatch at a few counts up to all cores and keep the best."""
    cores = os.cpu_count() or 1
    candidates = sorted({min(cores, c) for c in (8, 16, 32, 64, cores)})
    ref = CpuReferencePath(num_envs)
    mb = min(16384, num_envs * 24)
    idx = torch.randperm(num_envs * 24)[:mb]
    ref.env_step(0)
    ref.gae()
    best, best_t = candidates[0], float("inf")
    for c in candidates:
        torch.set_num_threads(c)
        ref.env_step(1)
        t0 = time.perf_counter()
        ref.env_step(2)
        t_step = time.perf_counter() - t0
        ref.minibatch(idx)
        t0 = time.perf_counter()
        ref.minibatch(idx)
        t_mb = time.perf_counter() - t0
        est = 24 * t_step + 30 * t_mb
        if est < best_t:
            best, best_t = c, est
    torch.set_num_threads(best)
    return best, cores


def cpu_baseline(num_envs, sample_steps=4, sample_minibatches=2):
    """Bounded sample of the same workload on the host cores; extrapolated to a full iteration."""
    threads, cores = pick_threads(num_envs)
    ref = CpuReferencePath(num_envs)
    ref.env_step(0)  # warm-up
    t0 = time.perf_counter()
    for t in range(sample_steps):
        ref.env_step(t)
    t_step = (time.perf_counter() - t0) / sample_steps
    t0 = time.perf_counter()
    ref.gae()
    t_gae = time.perf_counter() - t0
    perm = torch.randperm(num_envs * 24)
    mb = min(16384, num_envs * 24)
    ref.minibatch(perm[:mb])  # warm-up
    t0 = time.perf_counter()
    for i in range(sample_minibatches):
        ref.minibatch(perm[i * mb : (i + 1) * mb] if (i + 1) * mb <= perm.numel() else perm[:mb])
    t_mb = (time.perf_counter() - t0) / sample_minibatches
    n_mb = 5 * max(1, num_envs * 24 // mb)
    t_iter = 24 * t_step + t_gae + n_mb * t_mb
    return {
        "value": num_envs * 24 / t_iter, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": f"{sample_steps} env steps ({t_step*1e3:.1f} ms each) + 1 GAE ({t_gae*1e3:.1f} ms) + {sample_minibatches} minibatches of {mb} ({t_mb*1e3:.1f} ms each), extrapolated to 24 steps + GAE + {n_mb} minibatches = {t_iter:.2f} s per iteration; torch CPU oracle port, {threads} threads (fastest of the counts tried on a {cores}-core host)",
    }  # fmt: skip


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port) on this box's host cores.  Each step is a
    bounded sample of one iteration: the full 24-step rollout + GAE + one of the five epochs (6 minibatches);
    the other four epochs repeat identical work and are accounted for by scaling that epoch's time."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, T = args.envs, 24
    threads, cores = pick_threads(N)
    ref = CpuReferencePath(N)
    mb = min(16384, N * T)
    n_mb_epoch = max(1, N * T // mb)

    def one_step():
        t0 = time.perf_counter()
        for t in range(T):
            ref.env_step(t)
        ref.gae()
        t1 = time.perf_counter()
        perm = torch.randperm(N * T)
        for i in range(n_mb_epoch):
            ref.minibatch(perm[i * mb : (i + 1) * mb])
        t2 = time.perf_counter()
        return (t1 - t0) + 5 * (t2 - t1)

    for _ in range(args.warmup):
        one_step()
    total = sum(one_step() for _ in range(args.steps))
    value = N * T * args.steps / total
    sample = f"per step: full 24-step rollout + GAE + 1 of 5 epochs ({n_mb_epoch} minibatches of {mb}) measured, epoch time x5; torch CPU oracle port, {threads} threads (fastest of the counts tried on a {cores}-core host)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic Solo12 state (no Isaac Sim physics): trainer-side env-steps/s",
        "config": {"workload": "Isaac-Velocity-CaT-Flat-Solo12-v0 trainer side, 4096 envs, CleanRL PPO cfg (T=24, 5 epochs x 6 minibatches of 16384)", "envs_per_gpu": N},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))  # fmt: skip


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU")
    ap.add_argument("--no-graphs", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
