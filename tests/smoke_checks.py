"""One small invocation of the hot path on the GPU, checked against the CPU oracle (used by
`__graft_entry__.smoke()`)."""

from __future__ import annotations

import torch

from constraints_as_terminations_b200 import ConstraintManager, ops
from constraints_as_terminations_b200 import synthetic_env as se
from oracle import cat_oracle, ppo_oracle


def run(device: str = "cuda:0") -> None:
    n, T = 512, 24
    # ---- per-step path: 13 Solo12 terms, 3 steps ----
    cpu_env = se.SyntheticSolo12Env(n, device="cpu", seed=0, pool=1)
    gpu_env = se.SyntheticSolo12Env(n, device=device, seed=0, pool=1)
    cfg_cpu, cfg_gpu = se.solo12_constraints_cfg(), se.solo12_constraints_cfg()
    oracle = cat_oracle.ManagerOracle(cpu_env, cat_oracle.terms_from_cfg(cfg_cpu, resolve_scene=cpu_env.scene))
    mgr = ConstraintManager(cfg_gpu, gpu_env)
    gen = torch.Generator().manual_seed(1)
    for _ in range(3):
        state = se.sample_state(n, gen)
        cpu_env.load_state(state)
        gpu_env.load_state({k: v.to(device) for k, v in state.items()})
        want = oracle.compute()
        reset = torch.zeros(n, dtype=torch.bool)
        reset[::50] = True
        want_reward, want_dones = cat_oracle.step_epilogue(state["raw_reward"], want, reset)
        reward, dones = mgr.compute_step(gpu_env._raw_reward, reset.to(device))
        assert torch.equal(mgr._cstr_prob_buf.cpu(), want), "cstr_prob mismatch"
        assert torch.equal(reward.cpu(), want_reward) and torch.equal(dones.cpu(), want_dones)
    # ---- per-update path: GAE + moments ----
    g = torch.Generator().manual_seed(2)
    rewards, values = torch.rand(T, n, generator=g), torch.randn(T, n, generator=g)
    dones = torch.rand(T + 1, n, generator=g) * (torch.rand(T + 1, n, generator=g) < 0.3)
    true_dones = (torch.rand(T + 1, n, generator=g) < 0.02).float()
    next_value = torch.randn(n, generator=g)
    want_adv, want_ret = ppo_oracle.gae(rewards, values, dones[:-1], true_dones[:-1], next_value, dones[-1], true_dones[-1])
    adv, ret = ops.gae(*(t.to(device) for t in (rewards, values, dones, true_dones, next_value)), 0.99, 0.95)
    assert torch.equal(adv.cpu(), want_adv) and torch.equal(ret.cpu(), want_ret), "GAE mismatch"
    x = torch.randn(n, 45, generator=g) * 3 + 1
    st = ppo_oracle.rms_update(ppo_oracle.rms_init((45,)), x)
    mean, var, count = torch.zeros(45, device=device), torch.ones(45, device=device), torch.ones(1, device=device)
    got = ops.rms_forward(x.to(device), mean, var, count)
    torch.testing.assert_close(got.cpu(), ppo_oracle.rms_normalize(st, x), rtol=1e-5, atol=1e-5)
    # ---- single-`dones` GAE of the skrl front-end (bit-exact) ----
    from constraints_as_terminations_b200.skrl import compute_gae
    from oracle import gae_variants_oracle as go

    r, v, d, lv = go.sample_inputs(T, n, 5)
    want_ret, want_adv = go.skrl_compute_gae(r, d[:T], v, lv, normalize=False)
    ret2, adv2 = compute_gae(r.to(device), d[:T].to(device), v.to(device), lv.to(device), normalize=False)
    assert torch.equal(ret2.cpu(), want_ret) and torch.equal(adv2.cpu(), want_adv), "skrl GAE mismatch"
    # ---- rollout policy + one PPO minibatch (tcgen05 GEMMs, head / loss, clip + Adam) vs the fp32 oracle ----
    torch.manual_seed(0)
    agent = ppo_oracle.AgentOracle(se.OBS_DIM, se.ACT_DIM)
    dims = ops.make_dims(se.OBS_DIM, se.ACT_DIM)
    layout = ops.mlp_layout(dims)
    flat = torch.cat([p.detach().reshape(-1) for p in list(agent.critic.parameters()) + list(agent.actor_mean.parameters()) + [agent.actor_logstd]])
    params = flat.to(device)
    w16 = ops.weight_copies(dims, layout, device)
    ops.cast_weights(dims, params, w16)
    rows = 2048
    obs = torch.randn(rows, se.OBS_DIM, generator=g)
    obs16 = ops.obs_to_operand(dims, obs.to(device))
    value = torch.empty(rows, device=device)
    mean_out = torch.empty(rows, se.ACT_DIM, device=device)
    ops.mlp_act(dims, obs16, params, w16, ops.mlp_workspace(dims, rows, False, device), value=value, mean_out=mean_out)
    with torch.no_grad():
        want_v = agent.critic(obs).flatten()
        want_m = agent.actor_mean(obs)
    # tf32 operands / fp32 accumulation (the default precision) vs fp32: 4e-3 on O(1) outputs
    torch.testing.assert_close(value.cpu(), want_v, rtol=4e-3, atol=4e-3)
    torch.testing.assert_close(mean_out.cpu(), want_m, rtol=4e-3, atol=4e-3)
    # device-side Philox: the minibatch permutation is a permutation and matches the CPU restatement bit for bit
    from oracle import philox_oracle

    rng = ops.make_rng_state(1234, device)
    perm = ops.random_permutation(24 * n, rng)
    assert torch.equal(perm.cpu(), torch.from_numpy(philox_oracle.random_permutation(24 * n, 1234, 0))), "permutation mismatch"
    assert torch.equal(perm.sort().values.cpu(), torch.arange(24 * n))
    # a short training run end to end (graphs off: two iterations only): finite losses, parameters move
    from constraints_as_terminations_b200 import PPOTrainer, solo12_flat_ppo_cfg

    env = se.SyntheticSolo12Env(256, device=device, seed=3, pool=2, constraints_cfg=se.solo12_constraints_cfg())
    env.load_managers()
    trainer = PPOTrainer(env, solo12_flat_ppo_cfg(logger=None), device=device, use_graphs=False, distributed=False)
    trainer.start()
    before = trainer.agent.parameters_flat().clone()
    for _ in range(2):
        trainer.train_iteration()
    losses = trainer.losses()
    assert all(torch.isfinite(torch.tensor(x)) for x in losses.values()), losses
    assert float((trainer.agent.parameters_flat() - before).abs().max()) > 0.0
    # the same with CUDA graphs: whole env steps (policy + fused constraint step + append + observation statistics) and
    # whole epochs (Philox permutation + 6 minibatches) replay as single launches from iteration 2 on
    env = se.SyntheticSolo12Env(256, device=device, seed=3, pool=2, constraints_cfg=se.solo12_constraints_cfg())
    env.load_managers()
    trainer = PPOTrainer(env, solo12_flat_ppo_cfg(logger=None), device=device, use_graphs=True, distributed=False)
    trainer.start()
    for _ in range(4):
        trainer.train_iteration()
    losses = trainer.losses()
    assert all(torch.isfinite(torch.tensor(x)) for x in losses.values()), losses
    assert len(trainer._step_graphs) > 0 and len(trainer._epoch_graphs) > 0
    torch.cuda.synchronize()
