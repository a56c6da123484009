"""Train a CaT PPO agent: same flags, log layout and checkpoints as the reference's
`scripts/clean_rl/train.py` (lines 20-50, 92-148), on top of the catb200 kernels.

    python scripts/clean_rl/train.py --task=Isaac-Velocity-CaT-Flat-Solo12-v0 --headless [--num_envs N --seed S --num_iterations K]

With Isaac Lab installed the env is `gym.make(task)` (launch through Isaac Lab's python so the app is up);
without it the synthetic Solo12 state source stands in (trainer-side only)."""

from __future__ import annotations

import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import cli_args  # noqa: E402
import common  # noqa: E402


def main():
    parser = argparse.ArgumentParser(description="Train an RL agent with CleanRL (B200-native CaT hot path).")
    parser.add_argument("--video", action="store_true", default=False, help="Record videos during training.")
    parser.add_argument("--video_length", type=int, default=200, help="Length of the recorded video (in steps).")
    parser.add_argument("--video_interval", type=int, default=2000, help="Interval between video recordings (in steps).")
    parser.add_argument("--num_envs", type=int, default=None, help="Number of environments to simulate.")
    parser.add_argument("--task", type=str, default=common.TASK, help="Name of the task.")
    parser.add_argument("--seed", type=int, default=None, help="Seed used for the environment")
    parser.add_argument("--num_iterations", type=int, default=None, help="RL Policy training iterations.")
    parser.add_argument("--headless", action="store_true", default=False, help="Accepted for CLI parity (no GUI here).")
    parser.add_argument("--device", type=str, default=None, help="cuda device, default cuda:LOCAL_RANK")
    # extensions (not in the reference CLI): handy for short runs
    parser.add_argument("--save_interval", type=int, default=None, help="Checkpoint every N iterations.")
    parser.add_argument("--minibatch_size", type=int, default=None, help="PPO minibatch size.")
    cli_args.add_clean_rl_args(parser)
    args_cli, _ = parser.parse_known_args()

    import torch

    from constraints_as_terminations_b200 import PPO
    from constraints_as_terminations_b200 import dist as cdist

    torch.backends.cuda.matmul.allow_tf32 = True  # reference train.py:86-89 (no torch matmul on our hot path)
    torch.backends.cudnn.allow_tf32 = True
    rank, world, local_rank = cdist.init_from_env()
    device = args_cli.device or f"cuda:{local_rank}"
    agent_cfg = cli_args.parse_clean_rl_cfg(args_cli.task, args_cli)
    if args_cli.num_iterations is not None:
        agent_cfg.num_iterations = args_cli.num_iterations
    if args_cli.save_interval is not None:
        agent_cfg.save_interval = args_cli.save_interval
    if args_cli.minibatch_size is not None:
        agent_cfg.minibatch_size = args_cli.minibatch_size
    seed = cdist.shard_seed(agent_cfg.seed)
    torch.manual_seed(seed)
    env, env_cfg = common.make_env(args_cli.task, args_cli.num_envs, seed, device)
    log_dir = common.new_log_dir(agent_cfg.experiment_name)
    if rank == 0:
        common.dump_params(log_dir, env_cfg, agent_cfg)
    PPO(env, agent_cfg, log_dir)
    if hasattr(env, "close"):
        env.close()


if __name__ == "__main__":
    main()
