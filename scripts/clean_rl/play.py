"""Play a trained checkpoint and export the policy: counterpart of the reference's `scripts/clean_rl/play.py`
(lines 69-150): newest `model_*.pt` of the newest run under logs/clean_rl/<experiment>/, `Agent.load_state_dict`,
export to `exported/model.pt` (TorchScript) and `exported/model.onnx` (when the `onnx` package is present),
then a rollout with the stochastic policy."""

from __future__ import annotations

import argparse
import os
import sys
from pathlib import Path

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import cli_args  # noqa: E402
import common  # noqa: E402


def main():
    parser = argparse.ArgumentParser(description="Play a checkpoint of an RL agent from CleanRL.")
    parser.add_argument("--video", action="store_true", default=False)
    parser.add_argument("--video_length", type=int, default=200, help="Length of the rollout (in steps).")
    parser.add_argument("--num_envs", type=int, default=None)
    parser.add_argument("--task", type=str, default=common.TASK)
    parser.add_argument("--seed", type=int, default=None)
    parser.add_argument("--headless", action="store_true", default=False)
    parser.add_argument("--device", type=str, default="cuda:0")
    cli_args.add_clean_rl_args(parser)
    args_cli, _ = parser.parse_known_args()

    import torch

    from constraints_as_terminations_b200 import Agent

    agent_cfg = cli_args.parse_clean_rl_cfg(args_cli.task, args_cli)
    log_root = os.path.abspath(os.path.join("logs", "clean_rl", agent_cfg.experiment_name))
    print(f"[INFO] Loading experiment from directory: {log_root}")
    resume_path = common.get_checkpoint_path(log_root, agent_cfg.load_run, agent_cfg.load_checkpoint)
    print(f"[INFO] Loading model: {resume_path}")
    log_dir = os.path.dirname(resume_path)

    env, _ = common.make_env(args_cli.task, args_cli.num_envs or 64, agent_cfg.seed, args_cli.device)
    actor = Agent(env).to(torch.device(args_cli.device))
    actor.load_state_dict(torch.load(resume_path, map_location=args_cli.device))
    actor.eval()
    obs = env.reset()[0]["policy"]

    exported = os.path.join(log_dir, "exported")
    Path(exported).mkdir(parents=True, exist_ok=True)
    export = actor.export_module().to(args_cli.device)  # plain torch module over the same weights
    dummy = torch.randn(1, obs.shape[-1], device=args_cli.device)
    pt_path = os.path.join(exported, "model.pt")
    torch.jit.trace(export, dummy).save(pt_path)
    print(f"[INFO] Exported .pt model to {pt_path}")
    try:
        import onnx  # noqa: F401

        onnx_path = os.path.join(exported, "model.onnx")
        torch.onnx.export(export, dummy, onnx_path, export_params=True, opset_version=16, do_constant_folding=True,
                          input_names=["input"], output_names=["output"])  # fmt: skip
        print(f"[INFO] Exported ONNX model to {onnx_path}")
    except ImportError:
        print("[INFO] `onnx` is not installed: skipped the ONNX export")
    # the exported graph and the kernel path agree on the deterministic action
    with torch.no_grad():
        a_kernel = actor(obs)
        a_export = export(obs)
    print(f"[INFO] max |kernel - exported| deterministic action: {(a_kernel - a_export).abs().max().item():.3e}")

    for _ in range(args_cli.video_length):
        with torch.no_grad():
            actions, _, _, _ = actor.get_action_and_value(actor.obs_rms(obs, update=False))
        next_obs, rewards, next_done, timeouts, info = env.step(actions)
        obs = next_obs["policy"]
    torch.cuda.synchronize()
    print(f"[INFO] Rolled out {args_cli.video_length} steps; last mean reward {rewards.mean().item():.4f}, "
          f"mean termination probability {next_done.mean().item():.4f}")


if __name__ == "__main__":
    main()
