"""ms per PPO iteration of the timed bench workload (no e2e / rooflines / CPU legs): quick A/B of kernel variants."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
env, tr = bench.make_trainer(n, dev, seed=0)
ms, launches, clocks = bench.timed_iterations(tr, steps, 5, 1, dev, read_losses=False)
print({"envs": n, "ms_per_iter": round(ms / steps, 4), "env_steps_per_s": round(n * 24 * steps / (ms * 1e-3)), "sm_mhz": clocks.get("sm_mhz"),
       "env": {k: v for k, v in os.environ.items() if k.startswith("CATB200_")}})
