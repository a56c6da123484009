"""Per-tensor relative error of the CUDA minibatch gradient against the fp32 / rounding-emulating oracle, and the
run-to-run spread of two identical trainers (fp32 atomics order).  Diagnostic: python tools/grad_errors.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import test_mlp_gpu as T
from tests import test_trainer_gpu as TT
from constraints_as_terminations_b200 import ops
from oracle import ppo_oracle

DEV = "cuda:0"
for prec in ("tf32", "bf16"):
    for B, M in ((24576, 16384), (3000, 1000), (512, 512), (6000, 5000)):
        agent = T.make_agent(seed=1)
        dims, layout, params, w16 = T.device_agent(agent, prec)
        obs, actions, logp, adv, returns, values, norm_stats, idx = T._minibatch(agent, B, M, seed=B + M)
        obs16 = ops.obs_to_operand(dims, obs.to(DEV))
        ws = ops.mlp_workspace(dims, M, True, DEV)
        hp = ops.make_hparams()
        outs = []
        for rep in range(2):
            grads = torch.zeros(layout.n_params, device=DEV)
            loss_acc = torch.zeros(8, device=DEV)
            ops.ppo_minibatch_grad(dims, hp, idx.to(DEV), obs16, actions.to(DEV), logp.to(DEV), adv.to(DEV), returns.to(DEV),
                                   values.to(DEV), norm_stats.to(DEV), params, w16, grads, loss_acc, ws)
            torch.cuda.synchronize()
            outs.append(grads.cpu())
        got = outs[0]
        print(f"{prec} B={B} M={M}: repeat-call spread {float((outs[0]-outs[1]).norm()/outs[0].norm()):.2e}")
        val_n = (values - norm_stats[0]) / torch.sqrt(norm_stats[1] + 1e-8)
        ret_n = (returns - norm_stats[2]) / torch.sqrt(norm_stats[3] + 1e-8)
        value_rms = {"mean": norm_stats[2], "var": norm_stats[3]}
        a = ppo_oracle.AgentOracle(T.OBS, T.ACT)
        a.load_state_dict(agent.state_dict())
        loss, info = ppo_oracle.ppo_minibatch_loss(a, value_rms, obs[idx], actions[idx], logp[idx], adv[idx], ret_n[idx], val_n[idx])
        loss.backward()
        want = T.flat_grads(a, layout)
        print(f"   whole {float((got-want).norm()/want.norm()):.2e}")
        offs = sorted([*layout.w[0], *layout.b[0], *layout.w[1], *layout.b[1], layout.logstd, layout.n_params])
        for z in range(2):
            row = []
            for l in range(4):
                for kind, off in (("w", layout.w[z][l]), ("b", layout.b[z][l])):
                    nxt = [o for o in offs if o > off][0]
                    row.append(f"{kind}{l} {float((got[off:nxt]-want[off:nxt]).norm()/(want[off:nxt].norm()+1e-12)):.1e}")
            print(f"   net{z}: " + "  ".join(row))
# run-to-run spread of the trainer (eager / eager and eager / graph)
for modes in ((False, False), (False, True)):
    outs = []
    for graphs in modes:
        env, tr = TT._make_trainer(256, 8, 512, graphs=graphs, seed=5)
        for _ in range(2):
            tr.train_iteration()
        torch.cuda.synchronize()
        outs.append(tr.agent.parameters_flat().clone())
    d = (outs[0] - outs[1]).abs()
    print(f"trainer graphs={modes}: max |dp| {float(d.max()):.3e}, rel {float(d.norm()/outs[0].norm()):.3e}")
