"""Host time of one rollout (24 env steps) with an empty GPU queue, resident vs host-fed env, plus a cProfile of it."""
import cProfile, os, pstats, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
bench.HostFedEnv = bench._host_fed_env_cls()
for host_fed in (False, True):
    env, tr = bench.make_trainer(4096, dev, seed=0, host_fed=host_fed)
    for _ in range(4):
        tr.train_iteration()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        tr.iteration += 1
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tr.collect_rollout()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        tr.compute_gae(); tr.update(); tr.finish_iteration()
        ts.append((t1 - t0, t2 - t1))
    print("host_fed", host_fed, "rollout host ms", [round(a * 1e3, 2) for a, _ in ts], "gpu drain ms", [round(b * 1e3, 2) for _, b in ts])
    if host_fed or "--both" in sys.argv:
        pr = cProfile.Profile()
        torch.cuda.synchronize()
        pr.enable()
        for _ in range(3):
            tr.iteration += 1
            tr.collect_rollout()
            torch.cuda.synchronize()
        pr.disable()
        st = pstats.Stats(pr)
        st.sort_stats("tottime").print_stats(22)
    del env, tr
