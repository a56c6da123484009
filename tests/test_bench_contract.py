"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the driver's keys, the GPU arm
refuses to run without a device (no CPU fallback), and the flops attribution of the roofline object adds up."""

import json
import os
import subprocess
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_json_line():
    proc = _run("--impl", "reference", "--envs", "64", "--steps", "1", "--warmup", "0")
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    sys.path.insert(0, ROOT)
    import bench

    assert d["config"] == bench.workload_config(64, 1)  # the same dict the GPU arm prints for the same command line
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, cwd=ROOT, env=env, timeout=120)
    assert proc.returncode == 0 and proc.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        import pytest

        pytest.skip("CUDA present")
    proc = _run("--steps", "1", "--warmup", "3")
    assert proc.returncode != 0 and "no CPU fallback" in (proc.stderr + proc.stdout)


def test_gemm_flops_attribution_adds_up():
    sys.path.insert(0, ROOT)
    import bench

    tr = SimpleNamespace(num_envs=4096, T=24, batch_size=4096 * 24, minibatch_size=16384, cfg=SimpleNamespace(updates_epochs=5),
                         agent=SimpleNamespace(precision="tf32"))
    flops = bench._gemm_flops_by_kernel(tr)
    opt_rows, roll_rows = 5 * 4096 * 24, 4096 * 25
    mac_fwd = 2 * (45 * 512 + 512 * 256 + 256 * 128)   # both nets, layer 0 at its true K = 45
    mac_dgrad = 2 * (256 * 128 + 512 * 256)
    # the names are the instantiations CUPTI reports (csrc/tc_gemm.cu dispatch): 128 x 128 persistent tiles, 256-row tiles for
    # the minibatch-sized forward launches with K >= 256, weight gradients by column-tile width
    fwd = {"mlp_gemm_kernel<0, 1, 4>", "mlp_gemm256_kernel<0, 1, 256>", "mlp_gemm256_kernel<0, 1, 128>"}
    assert set(flops) == fwd | {"mlp_gemm_kernel<1, 1, 4>", "mlp_wgrad_kernel<1, 64>", "mlp_wgrad_kernel<1, 128>"}
    assert sum(flops[k] for k in fwd) == 2.0 * mac_fwd * (opt_rows + roll_rows)
    assert flops["mlp_gemm256_kernel<0, 1, 256>"] == 2.0 * 2 * 512 * 256 * opt_rows   # rollout-sized launches stay on 128-row tiles
    assert flops["mlp_gemm_kernel<1, 1, 4>"] == 2.0 * mac_dgrad * opt_rows
    assert flops["mlp_wgrad_kernel<1, 64>"] + flops["mlp_wgrad_kernel<1, 128>"] == 2.0 * mac_fwd * opt_rows
    # 750 848 FLOP/sample forward = the hidden layers on the tensor cores + the fp32 heads (128 -> 12, 128 -> 1)
    assert 2.0 * mac_fwd + 2 * (128 * 12 + 128) == 750848


def test_dominant_kernel_roofline_reports_both_roofs_and_names_the_binding_one():
    """A GEMM launch has a tensor roof (algorithmic flops) and an HBM roof (algorithmic activation bytes): bench.py prints
    both and `bound` is the one that leaves less headroom.  Numbers of the round-2 line: the tf32 dgrad kernel, 60 launches
    in 1413 us -> 0.33 of the tensor roof, 0.82 of the HBM roof."""
    sys.path.insert(0, ROOT)
    import bench

    tr = SimpleNamespace(num_envs=4096, T=24, batch_size=4096 * 24, minibatch_size=16384, cfg=SimpleNamespace(updates_epochs=5),
                         agent=SimpleNamespace(precision="tf32"))
    peaks = {"hbm_gbs": 6547.2, "bf16_tflops": 1682.4, "bf16_tflops_sustained": 1396.2, "source": "measured"}
    prof = {"mlp_gemm_kernel<1, 1, 4>": {"us": 1413.0, "launches": 60.0, "share": 0.18}}
    r = bench.dominant_kernel_roofline(prof, 7880.0, tr, peaks)
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] == 6547.2
    # dZ_l + H_{l-1} in, dZ_{l-1} out, both nets, fp32, for the 256- and the 512-wide layer, 30 minibatches of 16384 rows
    assert r["hbm_roof"]["bytes_per_step"] == 4 * 16384 * 30 * 2 * ((128 + 256 + 256) + (256 + 512 + 512))
    assert abs(r["hbm_roof"]["frac"] - 0.816) < 2e-3 and abs(r["tensor_roof"]["frac"] - 0.3266) < 2e-3
    assert r["frac"] == r["hbm_roof"]["frac"] and r["tensor_roof"]["peak"] == 1396.2 / 2
    # bf16 operands: half the bytes, twice the tensor peak
    tr.agent.precision = "bf16"
    r16 = bench.dominant_kernel_roofline({"mlp_gemm_kernel<1, 0, 4>": {"us": 900.0, "launches": 60.0, "share": 0.2}}, 6000.0, tr, peaks)
    assert r16["hbm_roof"]["bytes_per_step"] * 2 == r["hbm_roof"]["bytes_per_step"] and r16["tensor_roof"]["peak"] == 1396.2
    # a kernel without a flops attribution (e.g. the CaT step) keeps its share but claims no roof
    other = bench.dominant_kernel_roofline({"cat_eval_kernel<2>": {"us": 500.0, "launches": 24.0, "share": 0.3}}, 1600.0, tr, peaks)
    assert "frac" not in other and other["kernel"] == "cat_eval_kernel<2>"
