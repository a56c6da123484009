"""Per-launch device times (CUPTI) of ONE minibatch inside the replayed epoch graph, in launch order: which layer's
GEMM costs what when the caches are in their steady state (ncu's per-launch times are cold-cache and serialised).

    python tools/mb_timeline.py [envs] [minibatch index within the last epoch]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from torch.profiler import ProfilerActivity, profile

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
env, tr = bench.make_trainer(n, dev, seed=0)
for _ in range(5):
    tr.train_iteration()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.update()
    torch.cuda.synchronize()
evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy" not in e.name and "Memset" not in e.name),
             key=lambda e: e.time_range.start)
names = [e.name.split("(")[0].replace("void ", "").replace("catb200::", "") for e in evs]
# one minibatch = from one gather_kernel to the next
starts = [i for i, nm in enumerate(names) if nm.startswith("gather_kernel")]
print("launches in update():", len(evs), "minibatches:", len(starts), {k: v for k, v in os.environ.items() if k.startswith("CATB200_")})
for which in (len(starts) // 2, len(starts) - 2):
    a, b = starts[which], starts[which + 1]
    t0 = evs[a].time_range.start
    print(f"-- minibatch {which}: {evs[b].time_range.start - t0:.1f} us start to start")
    for e, nm in zip(evs[a:b], names[a:b]):
        print(f"  +{e.time_range.start - t0:7.1f} us  {e.time_range.end - e.time_range.start:6.1f} us  {nm[:60]}")
