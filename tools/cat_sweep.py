"""Time the fused constraint step (cat_eval + cat_apply) and GAE alone at several env counts: CUDA events on the
launch stream, L2 flushed before every timed launch, plus the per-kernel split from CUPTI.

    CATB200_EVAL_WARPS=4 python tools/cat_sweep.py 4096 16384 65536 262144 1048576
"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from constraints_as_terminations_b200 import ops
from constraints_as_terminations_b200 import synthetic_env as se
from torch.profiler import ProfilerActivity, profile

dev = torch.device("cuda", 0)
sizes = [int(a) for a in sys.argv[1:]] or [4096, 16384, 65536, 262144, 1048576]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


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
        times.append(s.elapsed_time(e) * 1e3)
    times.sort()
    half = max(1, len(times) // 2)
    return sum(times[:half]) / half


out = {"eval_warps": os.environ.get("CATB200_EVAL_WARPS", "auto")}
for n in sizes:
    env = se.SyntheticSolo12Env(n, device=dev, seed=1, pool=1, constraints_cfg=se.solo12_constraints_cfg())
    mgr = env.load_managers()
    reset = torch.zeros(n, dtype=torch.bool, device=dev)
    step = lambda: mgr.compute_step(env._raw_reward, reset)
    us = time_kernel(step)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            flush.fill_(1)
            step()
        torch.cuda.synchronize()
    split = {}
    for ev in prof.key_averages():
        if "cat_" in ev.key:
            t = getattr(ev, "device_time_total", None) or getattr(ev, "cuda_time_total", 0.0)
            split[ev.key.split("(")[0].replace("void ", "").replace("catb200::", "")] = round(t / ev.count, 2)
    T = 24
    rewards, values = torch.rand(T, n, device=dev), torch.randn(T, n, device=dev)
    dones, tdones = torch.rand(T + 1, n, device=dev), torch.zeros(T + 1, n, device=dev)
    nv = torch.randn(n, device=dev)
    adv, ret = torch.empty_like(rewards), torch.empty_like(rewards)
    gae_us = time_kernel(lambda: ops.gae(rewards, values, dones, tdones, nv, 0.99, 0.95, advantages=adv, returns=ret))
    out[str(n)] = {
        "cat_step_us": round(us, 2), "cat_GBps": round(940 * n / us / 1e3, 1), "split_us": split,
        "gae_us": round(gae_us, 2), "gae_GBps": round((24 * T * n + 12 * n) / gae_us / 1e3, 1),
    }
    del env, mgr, rewards, values, dones, tdones, adv, ret
print(json.dumps(out))
