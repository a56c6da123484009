"""Wall / device time of the phases of one PPO iteration (rollout, GAE, update) with syncs in between."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    graphs = "--no-graphs" not in sys.argv
    dev = torch.device("cuda", 0)
    env, tr = bench.make_trainer(n, dev, seed=0, graphs=graphs)
    for _ in range(3):
        tr.train_iteration()
    torch.cuda.synchronize()
    acc = {"rollout": 0.0, "gae": 0.0, "update": 0.0, "finish": 0.0, "policy": 0.0, "env": 0.0, "post": 0.0}
    iters = 5
    for _ in range(iters):
        tr.iteration += 1
        t0 = time.perf_counter(); tr.collect_rollout(); torch.cuda.synchronize(); t1 = time.perf_counter()
        tr.compute_gae(); torch.cuda.synchronize(); t2 = time.perf_counter()
        tr.update(); torch.cuda.synchronize(); t3 = time.perf_counter()
        tr.finish_iteration(); torch.cuda.synchronize(); t4 = time.perf_counter()
        acc["rollout"] += t1 - t0; acc["gae"] += t2 - t1; acc["update"] += t3 - t2; acc["finish"] += t4 - t3
    # host-side cost of the rollout pieces (no syncs: pure python + launch overhead)
    for t in range(tr.T):
        a = time.perf_counter()
        tr.agent.get_action_and_value(tr.obs_op[t], out=(tr.actions[t], tr.logprobs[t], tr.values[t]))
        b = time.perf_counter()
        out = env.step(tr.actions[t])
        c = time.perf_counter()
        acc["policy"] += b - a; acc["env"] += c - b
    torch.cuda.synchronize()
    print({k: round(v / iters * 1e3, 3) for k, v in acc.items() if k in ("rollout", "gae", "update", "finish")}, "ms per iteration")
    print({k: round(v / tr.T * 1e6, 1) for k, v in acc.items() if k in ("policy", "env")}, "us host time per env step")

main()
