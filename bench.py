#!/usr/bin/env python
"""Benchmark of the CaT PPO hot path (BASELINE.json metric: env-steps/sec, Solo12 CaT-Flat @4096 envs/GPU).

    python bench.py --gpus 1 --steps 20 --warmup 5                     # this repo's CUDA path (1 GPU)
    torchrun --nproc-per-node N ... bench.py --gpus N --steps K --warmup W   # one rank per GPU, NCCL
    python bench.py --impl reference --steps K --warmup W              # reference arm: the unmodified reference on the host cores

A "step" is one full PPO iteration on the trainer side of Isaac-Velocity-CaT-Flat-Solo12-v0 with the
reference's hyper-parameters: 24 env steps x num_envs (policy forward + sampling, the 13-term constraint
manager with reward / dones epilogue, rollout append, running observation normalisation), then GAE, value
normalisation and 5 epochs x 6 minibatches of PPO-clip update (forward, backward, clip, Adam) -- nothing
skipped.  Isaac Sim is not installed here, so physics is replaced by a synthetic Solo12 state source with
the same tensors (`constraints_as_terminations_b200/synthetic_env.py`); the number is trainer-side
env-steps/s and says so in `data`.

Prints ONE JSON line (rank 0).  `value` has the env state already resident in HBM; `e2e` feeds every env
step's state from pinned host memory (H2D inside the timed region) and reads the losses back (D2H).  Beside them:
`roofline` (dominant kernel of the step), `rooflines` (HBM-bound kernels timed alone, incl. 65 536 / 1 M envs),
`sweep` (BASELINE.json configs 3-5: 16 384 / 65 536 envs per GPU and the 16-term stress layout), `cpu_baseline` (the
unmodified reference on the host cores, bounded sample) and `gpu_eager_baseline` (the unmodified reference's torch-eager
code with CUDA tensors on this B200, TF32 on: the practically relevant "before").
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


def workload_config(envs_per_gpu, world):
    """`config` of the JSON line: identical for both arms (same keys, same values) for the same command line."""
    T, mb = 24, min(16384, envs_per_gpu * 24)
    return {
        "workload": f"Isaac-Velocity-CaT-Flat-Solo12-v0 trainer side, {envs_per_gpu} envs/GPU, CleanRL PPO cfg (T=24, 5 epochs x {envs_per_gpu * T // mb} minibatches of {mb})",
        "envs_per_gpu": envs_per_gpu, "num_steps": T, "constraint_terms": 13, "constraint_columns": 78,
        "minibatch": mb, "epochs": 5,
        "parallelism": f"dp{world} (one env shard per GPU, 1 gradient allreduce per optimizer step)",
        "l2": "per-step working set (rollout buffers + minibatch activations, > 300 MB) exceeds the 126 MB L2: inputs larger than L2",
    }  # fmt: skip


DATA = "synthetic Solo12 state (no Isaac Sim physics): trainer-side env-steps/s"


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
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "10", "-i", str(self.index)],
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
PRECISION = None  # --precision: tf32 (default) | bf16


def make_trainer(num_envs, device, seed, host_fed=False, pool=8, graphs=True, distributed=True):
    from constraints_as_terminations_b200 import PPOTrainer, solo12_flat_ppo_cfg
    from constraints_as_terminations_b200 import synthetic_env as se

    cls = HostFedEnv if host_fed else se.SyntheticSolo12Env
    env = cls(num_envs, device=device, seed=seed, pool=pool, constraints_cfg=se.solo12_constraints_cfg())
    env.load_managers()
    cfg = solo12_flat_ppo_cfg(logger=None)
    torch.manual_seed(cfg.seed + seed)
    trainer = PPOTrainer(env, cfg, device=device, use_graphs=graphs, distributed=distributed, precision=PRECISION)
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

        def pointer_key(self):  # which staging set / step the current state tensors belong to
            it, k = divmod(self._count - 1, self.STEPS)
            return (it & 1, k)

    return HostFed


HostFedEnv = None


def timed_iterations(trainer, steps, warmup, world, device, read_losses, sample_clocks=True):
    for _ in range(warmup):
        trainer.train_iteration()
        if read_losses:
            trainer.losses()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    launches0 = trainer.kernel_launches()
    sampler = ClockSampler(device.index or 0) if sample_clocks else None
    if sampler:
        sampler.start()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    pending = None
    for _ in range(steps):
        trainer.train_iteration()
        if read_losses:
            # device -> host read of every step's result: the copy (32 bytes into pinned memory) is enqueued right behind
            # the step and the host reads it while it submits the next step, as an asynchronous logger does -- the GPU
            # queue never drains; the last step's result is read before the closing event
            handle = trainer.losses_async()
            if pending is not None:
                trainer.read_losses(pending)
            pending = handle
    if pending is not None:
        trainer.read_losses(pending)
    end.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(end)
    if world > 1:
        torch.distributed.barrier()
        t = torch.tensor([ms], device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t)
    clocks = sampler.stop() if sampler else None
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


def _traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures of the
    CURRENT kernels (profiles/r2_traffic.json names the ncu CSV each value came from).  ncu flushes the caches before
    every replay, so these are cold-cache figures: an upper bound on what the same launch moves inside the step."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.isfile(path):
        return {}
    return {k: v for k, v in json.load(open(path)).items() if isinstance(v, dict)}


def _gemm_kernel_name(mode, p, rows, n_out, k_pad):
    """Kernel a forward (mode 0) / dgrad (mode 1) launch is dispatched to -- the default rules of csrc/tc_gemm.cu
    (`launch_gemm_any` / `use_tile256`), spelled the way CUPTI prints the instantiation."""
    bn = 256 if n_out % 256 == 0 else 128
    ctas256 = 2 * ((rows + 255) // 256) * (n_out // bn)
    tile256 = os.environ.get("CATB200_TILE256", "")
    if tile256 != "0" and ctas256 >= 96 and (tile256 == "1" or (mode == 0 and p == 1 and k_pad >= 256)):
        return f"mlp_gemm256_kernel<{mode}, {p}, {bn}>"
    stages = os.environ.get("CATB200_GEMM_STAGES", "4")
    return f"mlp_gemm_kernel<{mode}, {p}, {stages if stages in ('3', '5') else '4'}>"


def _gemm_flops_by_kernel(trainer):
    """ALGORITHMIC tensor-core flops per iteration attributed to each tcgen05 kernel instantiation (csrc/tc_gemm.cu):
    layer 0 counted at K = 45 (0.7508 MFLOP/sample forward), not at the K = 64 the padded operand rows
    carry; both nets (critic + actor) run in every launch.  mlp_gemm_kernel<0, P, S> / mlp_gemm256_kernel<0, P, BN> =
    forward, <1, ...> = dgrad, mlp_wgrad_kernel<P, BN> = weight gradients (BN = 64 for the 45 -> 512 layer)."""
    n, T = trainer.num_envs, trainer.T
    mb, n_opt = trainer.minibatch_size, (trainer.batch_size // trainer.minibatch_size) * int(trainer.cfg.updates_epochs)
    p = 1 if trainer.agent.precision == "tf32" else 0
    layers = [(45, 64, 512), (512, 512, 256), (256, 256, 128)]  # (K, K padded, N) of the hidden layers
    flops = {}

    def add(name, rows, launches, k, n_out):
        flops[name] = flops.get(name, 0.0) + 2.0 * 2 * k * n_out * rows * launches

    for k, k_pad, n_out in layers:
        add(_gemm_kernel_name(0, p, mb, n_out, k_pad), mb, n_opt, k, n_out)   # update forward
        add(_gemm_kernel_name(0, p, n, n_out, k_pad), n, T + 1, k, n_out)     # rollout policy + bootstrap value
    for k, k_pad, n_out in layers[1:]:
        add(_gemm_kernel_name(1, p, mb, k, n_out), mb, n_opt, n_out, k)       # dgrad: dZ_l [M, N_l] x W_l -> [M, K_l]
    for k, k_pad, n_out in layers:
        add(f"mlp_wgrad_kernel<{p}, {64 if k_pad < 128 else 128}>", mb, n_opt, k, n_out)
    return flops


def _gemm_bytes_by_kernel(trainer):
    """ALGORITHMIC HBM bytes per iteration of the same kernels: every activation tensor a launch consumes
    is read once and every one it produces is written once, both nets, at the operand size of the precision (fp32 storage
    for tf32 operands, 2 bytes for bf16); the weights (1.5 MB, L2 resident) are left out.  forward l: read H_{l-1} (X once
    for both nets), write H_l; dgrad l: read dZ_l and H_{l-1}, write dZ_{l-1}; wgrad l: read dZ_l and H_{l-1}."""
    n, T = trainer.num_envs, trainer.T
    mb, n_opt = trainer.minibatch_size, (trainer.batch_size // trainer.minibatch_size) * int(trainer.cfg.updates_epochs)
    tf32 = trainer.agent.precision == "tf32"
    p, es = (1, 4) if tf32 else (0, 2)
    layers = [(45, 64, 512), (512, 512, 256), (256, 256, 128)]
    nbytes = {}

    def add(name, v):
        nbytes[name] = nbytes.get(name, 0.0) + v

    for li, (k, k_pad, n_out) in enumerate(layers):
        a_in = k_pad * (1 if li == 0 else 2)  # the observation rows are shared by the two nets
        for rows, launches in ((mb, n_opt), (n, T + 1)):
            add(_gemm_kernel_name(0, p, rows, n_out, k_pad), es * rows * launches * (a_in + 2 * n_out))
    for k, k_pad, n_out in layers[1:]:
        add(_gemm_kernel_name(1, p, mb, k, n_out), es * mb * n_opt * 2 * (n_out + k + k))
    for li, (k, k_pad, n_out) in enumerate(layers):
        add(f"mlp_wgrad_kernel<{p}, {64 if k_pad < 128 else 128}>", es * mb * n_opt * (2 * n_out + k_pad * (1 if li == 0 else 2)))
    return nbytes


def dominant_kernel_roofline(prof, total_us, trainer, peaks):
    """Roofline entry of the kernel with the largest share of the iteration.  A GEMM launch has two roofs: the tensor
    pipe (algorithmic flops / peak) and HBM (algorithmic bytes of its activation operands / measured copy bandwidth);
    `bound` names the one that leaves less headroom, both fractions are printed."""
    flops = _gemm_flops_by_kernel(trainer)
    nbytes = _gemm_bytes_by_kernel(trainer)
    if not prof:  # no CUPTI records (e.g. the process itself runs under ncu, which owns the profiling interface)
        return {"kernel": None, "note": "kernel shares unavailable: CUPTI produced no records in this process"}
    name, row = next(iter(prof.items()))
    out = {"kernel": name, "share_of_step": row["share"], "us_per_step": row["us"], "launches_per_step": row["launches"]}
    if name in flops:
        bf16_peak = peaks.get("bf16_tflops_sustained") or peaks["bf16_tflops"]
        tf32 = trainer.agent.precision == "tf32"
        # kind::tf32 issues half the MACs per tcgen05.mma of kind::f16 (K = 8 vs 16 per instruction at the same
        # dispatch rate; nominal 1.1 vs 2.25 PFLOP/s dense, B200_PROFILING.md): peak = half the MEASURED bf16 rate
        t_peak = bf16_peak * (0.5 if tf32 else 1.0)
        t_ach = flops[name] / (row["us"] * 1e-6) / 1e12
        h_ach = nbytes[name] / (row["us"] * 1e-6) / 1e9
        tensor = {"achieved": t_ach, "peak": t_peak, "unit": "TFLOP/s", "frac": t_ach / t_peak, "flops_per_step": flops[name],
                  "peak_source": peaks["source"] + (" (sustained bf16 cuBLAS rate x 0.5: tf32 operands; the kernel runs inside a long step)" if tf32 else " (sustained bf16: the kernel runs inside a long step)")}
        hbm = {"achieved": h_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": h_ach / peaks["hbm_gbs"], "bytes_per_step": nbytes[name],
               "peak_source": peaks["source"] + " (copy bandwidth)"}
        bound = "hbm" if hbm["frac"] >= tensor["frac"] else "tensor"
        pick = hbm if bound == "hbm" else tensor
        out.update({"bound": bound, "achieved": pick["achieved"], "peak": pick["peak"], "unit": pick["unit"], "frac": pick["frac"],
                    "tensor_roof": tensor, "hbm_roof": hbm,
                    "frac_of_measured_bf16_peak": t_ach / bf16_peak,
                    "traffic": (_traffic().get(name) or {}).get("bytes_per_launch"),
                    "traffic_source": (_traffic().get(name) or {}).get("source"),
                    "algorithmic_bytes_per_launch": nbytes[name] / max(row["launches"], 1.0),
                    "peak_source": pick["peak_source"]})
    return out


def kernel_rooflines(device, num_envs, peaks):
    """Per-kernel achieved bandwidth / throughput, timed alone with CUDA events on the launch stream,
    L2 flushed between timed launches (a 256 MiB write)."""
    from constraints_as_terminations_b200 import _lib as _L
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
        before = _L.launch_count()
        mgr.compute_step(env._raw_reward, reset)
        cat_launches = _L.launch_count() - before  # 1: single-launch kernel (grid barrier, tiles stay in shared memory); 2: eval + apply
        nbytes = 940 * n
        out[f"cat_step@{n}"] = {"bound": "hbm", "achieved": nbytes / sec / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": nbytes / sec / 1e9 / peaks["hbm_gbs"], "us": sec * 1e6, "bytes": nbytes, "launches": cat_launches}
        del env, mgr
    # a6: running observation moments + normalise (+ operand copy) + rollout append: 428 B/env-step
    from constraints_as_terminations_b200 import RunningMeanStd
    from constraints_as_terminations_b200 import _lib as L

    roof_a6 = {}
    for n in sorted({num_envs, 65536, 1 << 20}):
        rms = RunningMeanStd(shape=(se.OBS_DIM,)).to(device)
        raw = torch.randn(n, se.OBS_DIM, device=device)
        norm, out_op = torch.empty_like(raw), torch.empty(n, 64, dtype=L.operand_dtype(L.PREC_NAMES[PRECISION or L.default_precision()]), device=device)
        rew, done, tout = torch.rand(n, device=device), torch.rand(n, device=device), torch.zeros(n, dtype=torch.bool, device=device)
        r_t, d_t, td_t = torch.empty(n, device=device), torch.empty(n, device=device), torch.empty(n, device=device)

        def a6():
            ops.rollout_append(rew, done, tout, r_t, d_t, td_t)
            rms(raw, update=True, out=norm, out_op=out_op)

        sec = time_kernel(a6)
        nbytes = 428 * n
        roof_a6[f"obs_norm_append@{n}"] = {"bound": "hbm", "achieved": nbytes / sec / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": nbytes / sec / 1e9 / peaks["hbm_gbs"], "us": sec * 1e6, "bytes": nbytes, "launches": 3}
        del rms, raw, out_op
    # one PPO optimizer step on a 16384-row minibatch (gather, 3 fwd + 2 dgrad + 3 wgrad tcgen05 GEMMs, head/loss, fold,
    # grad norm, Adam + operand-copy refresh): 2.2525 MFLOP/sample of tensor work
    env, tr = make_trainer(num_envs, device, seed=0, graphs=False, distributed=False)  # rank-local probe
    tr.train_iteration()
    mb = tr.minibatch_size
    perm = torch.randperm(tr.batch_size, device=device)
    before = L.launch_count()
    tr._minibatch(perm[:mb])
    launches = L.launch_count() - before
    sec = time_kernel(lambda: tr._minibatch(perm[:mb]), reps=10)
    flops = 2.2525e6 * mb
    tf32 = tr.agent.precision == "tf32"
    peak = peaks["bf16_tflops"] * (0.5 if tf32 else 1.0)
    # the same step against HBM: every activation tensor read / written once per consuming / producing launch
    es = 4 if tf32 else 2
    per_row = (2 * 64                                                     # gather: operand rows in + out
               + (64 + 2 * 512) + (2 * 512 + 2 * 256) + (2 * 256 + 2 * 128)  # forward
               + 4 * 128                                                   # heads: H3 in, dZ3 out (both nets)
               + 2 * (128 + 256 + 256) + 2 * (256 + 512 + 512)             # dgrad
               + (2 * 128 + 2 * 256) + (2 * 256 + 2 * 512) + (2 * 512 + 64))  # wgrad
    hbytes = float(es * per_row * mb)
    t_frac, h_frac = flops / sec / 1e12 / peak, hbytes / sec / 1e9 / peaks["hbm_gbs"]
    entry = {"bound": "hbm" if h_frac >= t_frac else "tensor", "us": sec * 1e6, "flops": flops, "bytes": hbytes, "launches": launches,
             "tensor_roof": {"achieved": flops / sec / 1e12, "peak": peak, "unit": "TFLOP/s", "frac": t_frac,
                             "peak_source": "measured burst bf16 cuBLAS rate" + (" x 0.5 (tf32 operands)" if tf32 else "")},
             "hbm_roof": {"achieved": hbytes / sec / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": h_frac}}
    entry.update({k: entry[entry["bound"] + "_roof"][k] for k in ("achieved", "peak", "unit", "frac")})
    out[f"ppo_minibatch_step@{mb}"] = entry
    del env, tr
    out.update(roof_a6)
    # DRAM traffic per launch from the committed `ncu --set full` captures of the current kernels, where available
    for k, v in _traffic().items():
        if k in out:
            out[k]["traffic"] = v.get("bytes_per_launch")
            out[k]["traffic_source"] = v.get("source")
    return out


def sweep_point(num_envs, device, rank, world, stress, graphs, steps=4, warmup=3):
    """One point of the north_star size sweep (BASELINE.json configs 3-5): full PPO iterations at `num_envs` envs per GPU."""
    from constraints_as_terminations_b200 import PPOTrainer, solo12_flat_ppo_cfg
    from constraints_as_terminations_b200 import synthetic_env as se

    env = se.SyntheticSolo12Env(num_envs, device=device, seed=rank, pool=2, constraints_cfg=se.solo12_constraints_cfg(stress=stress))
    env.load_managers()
    cfg = solo12_flat_ppo_cfg(logger=None)
    tr = PPOTrainer(env, cfg, device=device, use_graphs=graphs, precision=PRECISION)
    tr.start()
    ms, _, _ = timed_iterations(tr, steps, warmup, world, device, read_losses=False, sample_clocks=False)
    n_terms = len(env.constraint_manager.active_terms)
    del env, tr
    torch.cuda.empty_cache()
    return {
        "envs_per_gpu": num_envs, "constraint_terms": n_terms, "constraint_columns": 93 if stress else 78, "n_gpus": world,
        "value": num_envs * 24 * world * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
    }  # fmt: skip


def run_ours(args):
    global HostFedEnv
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa_cpus = []
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)
        from constraints_as_terminations_b200 import dist as cdist

        # one rank per GPU: keep this rank's host threads and its pinned staging buffers on the GPU's own NUMA node
        numa_cpus = cdist.bind_to_gpu_numa_node(local_rank)
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
    precision = trainer.agent.precision
    prof, prof_total = kernel_profile(trainer, record=(rank == 0))
    dominant = dominant_kernel_roofline(prof, prof_total, trainer, peaks) if rank == 0 else None
    gemm_flops, gemm_bytes = _gemm_flops_by_kernel(trainer), _gemm_bytes_by_kernel(trainer)
    del env, trainer
    torch.cuda.empty_cache()

    # ---- e2e: host-resident env state, H2D each env step, D2H of the losses each iteration
    env, trainer = make_trainer(N, device, seed=rank, host_fed=True, graphs=not args.no_graphs)
    e_steps = max(3, args.steps // 2)
    # warm-up: the env-step graphs are captured per (rollout slot, staging set) from the third iteration on, and the two
    # staging sets alternate per iteration -> four iterations until every graph exists (a capture is host work, not the path)
    e_ms, _, _ = timed_iterations(trainer, e_steps, max(5, args.warmup), world, device, read_losses=True, sample_clocks=False)
    e2e = {
        "value": N * T * world * e_steps / (e_ms * 1e-3),
        "unit": UNIT,
        "h2d_bytes_per_step": env.h2d_bytes * T,
        "d2h_bytes_per_step": 32,
        "ms_per_step": e_ms / e_steps, "steps": e_steps, "warmup": max(5, args.warmup),
        "feed": "per iteration one cudaMemcpyAsync of the rollout's 24 packed env states from pinned host memory on a copy "
                "stream into one of two device staging sets (read one iteration ahead); the losses of every iteration are copied to pinned host memory right behind it and read by the host while it submits the next iteration (the last one before the closing event)",
    }
    del env, trainer
    torch.cuda.empty_cache()

    # ---- north_star size sweep (every rank takes part: the iterations contain the gradient all-reduce)
    sweep = None
    if not args.no_sweep:
        sweep = {
            "note": "full PPO iterations (24 env steps + GAE + 5 epochs of 16384-row minibatches), same precision / graphs as the headline; BASELINE.json configs 3-5",
            "points": [
                sweep_point(16384, device, rank, world, stress=False, graphs=not args.no_graphs),
                sweep_point(65536, device, rank, world, stress=False, graphs=not args.no_graphs),
                sweep_point(65536, device, rank, world, stress=True, graphs=not args.no_graphs),
            ],
        }  # fmt: skip

    # a second precision beside the headline (bf16 operands when the headline is tf32, and vice versa), same workload
    other = None
    if not args.no_sweep:
        global PRECISION
        keep, PRECISION = PRECISION, ("bf16" if precision == "tf32" else "tf32")
        env, trainer = make_trainer(N, device, seed=rank, graphs=not args.no_graphs)
        o_ms, _, _ = timed_iterations(trainer, max(3, args.steps // 2), 3, world, device, read_losses=False, sample_clocks=False)
        other = {"precision": PRECISION, "value": N * T * world * max(3, args.steps // 2) / (o_ms * 1e-3), "unit": UNIT, "ms_per_step": o_ms / max(3, args.steps // 2)}
        PRECISION = keep
        del env, trainer
        torch.cuda.empty_cache()

    if world > 1:  # the remaining legs are rank-local: let the other ranks go
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    roof = kernel_rooflines(device, N, peaks)
    cpu = gpu_eager = None
    if world == 1 and not args.no_baselines:
        gpu_eager = gpu_eager_baseline(N, device)
        cpu = cpu_baseline(N)
    main = dominant
    main["note"] = main.get("note") or ("dominant kernel of the step by device time (torch.profiler/CUPTI over 2 iterations of the timed workload); "
                    "HBM-bound kernels (GAE, CaT, obs normalisation) timed alone with CUDA events are in `rooflines`")
    gemm = "tf32 operands (fp32 storage rounded to tf32, tcgen05 kind::tf32) + fp32 accumulate" if precision == "tf32" else "bf16 operands + fp32 accumulate"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": f"fp32 (CaT/GAE/moments/heads/loss/Adam), {gemm} (hidden-layer GEMMs: the reference's GPU numerics are TF32 matmuls)",
        "data": DATA,
        "config": workload_config(N, world),
        "precision": precision,
        "cuda_graphs": not args.no_graphs,
        "host_binding": (f"rank 0 bound to {len(numa_cpus)} CPUs of its GPU's NUMA node" if numa_cpus else "none"),
        "e2e": e2e,
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": main,
        "rooflines": roof,
        "kernel_shares": {k: {"share": round(v["share"], 4), "us_per_step": round(v["us"], 1), "launches_per_step": v["launches"],
                              **({"tflops": round(gemm_flops[k] / v["us"] / 1e6, 1), "algorithmic_gbs": round(gemm_bytes[k] / v["us"] / 1e3, 1)} if k in gemm_flops else {})} for k, v in list(prof.items())[:18]},
        "sweep": sweep,
        "other_precision": other,
        "cpu_baseline": cpu,
        "gpu_eager_baseline": gpu_eager,
        "losses": losses,
    }  # fmt: skip
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
    time one env step + one minibatch of the oracle port (the same op mix) at a few counts up to all cores and keep
    the best."""
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


def reference_kind():
    """"reference": the unmodified reference (from /root/reference, or the archive oracle/build_ref.py packed) runs;
    "port": it is not reachable on this box and the oracle port stands in."""
    from oracle import ref_loader

    return "reference" if ref_loader.reference_available() else "port"


def run_cpu_reference(num_envs, steps, warmup):
    """Full PPO iterations of the reference on the host cores: warmup untimed, then exactly `steps` timed.
    -> (seconds for the `steps` iterations, kind, threads, cores, description)"""
    threads, cores = pick_threads(min(num_envs, 4096))
    kind = reference_kind()
    if kind == "reference":
        from oracle import ref_runner

        r = ref_runner.run_reference_trainer(num_envs, steps, warmup, device="cpu")
        what = (f"the UNMODIFIED reference (ConstraintManager / CaT / term functions / curriculum / PPO() of {r['reference_root']}) "
                f"driven through its own API on the synthetic Solo12 state, torch CPU, {threads} threads (fastest of the counts tried on a {cores}-core host)")
        return r["seconds"], kind, threads, cores, what
    ref = CpuReferencePath(num_envs)
    mb = min(16384, num_envs * 24)

    def one_iteration():
        for t in range(24):
            ref.env_step(t)
        ref.gae()
        for _ in range(5):
            perm = torch.randperm(num_envs * 24)
            for i in range(max(1, num_envs * 24 // mb)):
                ref.minibatch(perm[i * mb : (i + 1) * mb])

    for _ in range(warmup):
        one_iteration()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_iteration()
    what = f"oracle PORT of the reference path (the reference tree is not reachable on this box), torch CPU, {threads} threads (fastest of the counts tried on a {cores}-core host)"
    return time.perf_counter() - t0, kind, threads, cores, what


def cpu_baseline(num_envs):
    """Bounded sample inside the GPU arm's run: 1 warm-up + 3 full PPO iterations of the reference on the host cores
    (every iteration complete: 24 env steps, GAE, 5 epochs x 6 minibatches -- nothing extrapolated)."""
    from oracle import ref_runner

    steps, warmup = 3, 1
    with ref_runner._device_choice(False):  # the reference picks its device with torch.cuda.is_available(): host cores here
        seconds, kind, threads, cores, what = run_cpu_reference(num_envs, steps, warmup)
    return {
        "value": num_envs * 24 * steps / seconds, "unit": UNIT, "cores": threads, "kind": kind,
        "sample": f"{steps} full PPO iterations after {warmup} warm-up ({seconds / steps:.2f} s each) of the headline workload; {what}",
    }  # fmt: skip


def gpu_eager_baseline(num_envs, device):
    """The practically relevant "before" (SURVEY.md §8d last row): the unmodified reference's own
    torch-eager code with CUDA tensors on this B200, TF32 matmuls on as scripts/clean_rl/train.py:86-87 sets them."""
    if reference_kind() != "reference":
        return {"unavailable": "reference tree not reachable on this box (oracle/_ref/ref_hotpath.tar.gz missing)"}
    from oracle import ref_runner

    steps, warmup = 5, 2
    r = ref_runner.run_reference_trainer(num_envs, steps, warmup, device=str(device), tf32=True)
    torch.cuda.empty_cache()
    return {
        "value": r["env_steps_per_sec"], "unit": UNIT, "ms_per_step": r["seconds_per_iteration"] * 1e3, "steps": steps, "warmup": warmup,
        "kind": "reference", "numerics": "fp32 storage, TF32 matmuls (torch.backends.cuda.matmul.allow_tf32 = True)",
        "what": f"reference PPO() + ConstraintManager from {r['reference_root']}, eager torch on {device}, same synthetic Solo12 state and hyper-parameters as the headline; wall clock with device syncs",
    }  # fmt: skip


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores, same workload as the
    GPU arm's command line: `--gpus N` shards of `--envs` envs each (one process, N x envs envs), full PPO iterations,
    W untimed then exactly K timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, int(os.environ.get("WORLD_SIZE", str(args.gpus))))
    N, T = args.envs, 24
    total_envs = N * world
    seconds, kind, threads, cores, what = run_cpu_reference(total_envs, args.steps, args.warmup)
    value = total_envs * T * args.steps / seconds
    sample = f"{args.steps} full PPO iterations after {args.warmup} warm-up on {total_envs} envs ({world} shard(s) of {N}): {what}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": seconds / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": DATA,
        "config": workload_config(N, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
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
    ap.add_argument("--precision", default=None, choices=["tf32", "bf16"], help="operand precision of the hidden-layer GEMMs (default tf32)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the 16384 / 65536-env sweep and the second-precision leg")
    ap.add_argument("--no-baselines", action="store_true", help="skip cpu_baseline and gpu_eager_baseline")
    args = ap.parse_args()
    global PRECISION
    PRECISION = args.precision
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""  # the reference arm runs on the host cores (CUDA is initialised lazily)
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
