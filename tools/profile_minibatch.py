"""One PPO minibatch step (13 launches) between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/<name> python tools/profile_minibatch.py [envs] [tf32|bf16]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    bench.PRECISION = sys.argv[2] if len(sys.argv) > 2 else None
    dev = torch.device("cuda", 0)
    env, tr = bench.make_trainer(n, dev, seed=0, graphs=False, distributed=False)
    tr.train_iteration()
    mb = tr.minibatch_size
    perm = torch.randperm(tr.batch_size, device=dev)
    tr._minibatch(perm[:mb])
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush.fill_(1)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    tr._minibatch(perm[mb : 2 * mb])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()

main()
