import cProfile, os, pstats, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
env, tr = bench.make_trainer(4096, torch.device("cuda", 0), seed=0)
for _ in range(3):
    tr.train_iteration()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    tr.iteration += 1
    tr.collect_rollout()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
