"""Launcher parity: train -> checkpoint with the reference's state_dict keys -> play / export."""

import glob
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_train_then_play(tmp_path):
    env = dict(os.environ, PYTHONPATH=ROOT)
    train = [sys.executable, os.path.join(ROOT, "scripts/clean_rl/train.py"), "--headless", "--num_envs", "256", "--num_iterations", "4",
             "--save_interval", "2", "--minibatch_size", "1024", "--experiment_name", "launcher_test", "--seed", "3"]  # fmt: skip
    out = subprocess.run(train, cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    runs = glob.glob(str(tmp_path / "logs/clean_rl/launcher_test/*"))
    assert len(runs) == 1
    assert os.path.isfile(os.path.join(runs[0], "params", "agent.yaml")) and os.path.isfile(os.path.join(runs[0], "params", "env.pkl"))
    ckpts = sorted(glob.glob(os.path.join(runs[0], "model_*.pt")))
    assert [os.path.basename(c) for c in ckpts] == ["model_1.pt", "model_3.pt"]  # (iteration + 1) % save_interval == 0
    sd = torch.load(ckpts[-1], map_location="cpu")
    assert "critic.0.weight" in sd and "actor_mean.6.bias" in sd and "actor_logstd" in sd and "obs_rms.running_mean" in sd
    assert float(sd["obs_rms.count"]) == 1 + 256 * (1 + 24 * 3)  # saved after iteration 3
    assert glob.glob(os.path.join(runs[0], "events.out.tfevents.*")), "tensorboard scalars missing"
    play = [sys.executable, os.path.join(ROOT, "scripts/clean_rl/play.py"), "--headless", "--num_envs", "32", "--video_length", "5",
            "--experiment_name", "launcher_test"]  # fmt: skip
    out = subprocess.run(play, cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert os.path.isfile(os.path.join(runs[0], "exported", "model.pt"))
    line = [ln for ln in out.stdout.splitlines() if "deterministic action" in ln][0]
    assert float(line.split(":")[-1]) < 5e-2  # bf16 tensor-core path vs fp32 exported module
