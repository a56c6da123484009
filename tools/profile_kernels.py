"""Launch the HBM-bound kernels a few times at a given env count, for `ncu -k regex:...` captures."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from constraints_as_terminations_b200 import ops
from constraints_as_terminations_b200 import synthetic_env as se

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = "cuda:0"
env = se.SyntheticSolo12Env(n, device=dev, seed=1, pool=2, constraints_cfg=se.solo12_constraints_cfg())
mgr = env.load_managers()
reset = torch.zeros(n, dtype=torch.bool, device=dev)
T = 24
rewards, values = torch.rand(T, n, device=dev), torch.randn(T, n, device=dev)
dones, tdones = torch.rand(T + 1, n, device=dev), torch.zeros(T + 1, n, device=dev)
nv = torch.randn(n, device=dev)
adv, ret = torch.empty_like(rewards), torch.empty_like(rewards)
obs = torch.randn(n, 45, device=dev)
mean, var, count = torch.zeros(45, device=dev), torch.ones(45, device=dev), torch.ones(1, device=dev)
out, out16 = torch.empty_like(obs), torch.empty(n, 64, dtype=torch.bfloat16, device=dev)
for i in range(4):
    env._advance()
    mgr.compute_step(env._raw_reward, reset)
    ops.gae(rewards, values, dones, tdones, nv, 0.99, 0.95, advantages=adv, returns=ret)
    ops.rms_forward(obs, mean, var, count, out=out, out16=out16)
torch.cuda.synchronize()
print("done")
