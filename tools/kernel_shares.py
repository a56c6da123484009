"""Device time per kernel (CUPTI) over 2 iterations of the bench workload: quick A/B of kernel variants."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
env, tr = bench.make_trainer(n, dev, seed=0)
for _ in range(5):
    tr.train_iteration()
prof, total = bench.kernel_profile(tr)
print("total us/iter", round(total, 1), {k: v for k, v in os.environ.items() if k.startswith("CATB200_")})
for k, v in list(prof.items())[:14]:
    print(f"{v['us']:9.1f} us {v['launches']:6.1f}x {v['us'] / v['launches']:7.2f} us/launch  {k[:60]}")
