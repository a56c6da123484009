"""Training equivalence: the CUDA trainer (tf32 tensor-core MLP) against the fp32 oracle trainer over 20 PPO iterations.

Both trainers see the same synthetic Solo12 states (the env does not react to actions), the same Normal.sample() noise
and the same minibatch permutations: the CUDA side draws them from Philox on the device, the oracle side recomputes
the very same numbers with oracle/philox_oracle.py (permutations bit-exact, normals to 1e-6).  What differs is the
arithmetic of the hidden-layer GEMMs (tf32 operands, fp32 accumulation vs fp32 on the CPU) and reduction orders, so the
two runs drift apart slowly; the test bounds that drift on the quantities the reference logs (U/cleanrl/ppo.py:356-366):
policy loss, value loss, entropy, approx KL, and on the parameters themselves.  The reference's own GPU runs carry the
same kind of difference against its CPU runs (TF32 matmuls, scripts/clean_rl/train.py:86-87).

PPO training is sensitive to perturbations (clip / max branches flip, Adam normalises tiny gradients), so the band is
not guessed: the fp32 oracle trainer is run again from weights perturbed by 2^-11 relative -- the resolution of a tf32
operand -- and the drift of the CUDA trainer from the oracle must stay within 2x the drift between those two fp32 runs.
"""

import os

import numpy as np
import pytest
import torch

from constraints_as_terminations_b200 import PPOTrainer, solo12_flat_ppo_cfg
from constraints_as_terminations_b200 import synthetic_env as se
from oracle import cat_oracle, philox_oracle, ppo_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
N, T, MB, EPOCHS, ITERS = 256, 24, 1024, 5, 20
ENV_SEED = 11


class OracleTrainer:
    """fp32 CPU restatement of the reference loop (ppo.py:186-354) fed with externally supplied noise / permutations."""

    def __init__(self, init_state, rng_seed, lr, num_iterations):
        self.env = se.SyntheticSolo12Env(N, device="cpu", seed=ENV_SEED, pool=3, episode_length=60)
        self.mgr = cat_oracle.ManagerOracle(self.env, cat_oracle.terms_from_cfg(se.solo12_constraints_cfg(), resolve_scene=self.env.scene))
        self.agent = ppo_oracle.AgentOracle(se.OBS_DIM, se.ACT_DIM)
        self.agent.load_state_dict({k: v.clone() for k, v in init_state.items() if "_rms." not in k})
        self.params = list(self.agent.critic.parameters()) + list(self.agent.actor_mean.parameters()) + [self.agent.actor_logstd]
        self.opt = torch.optim.Adam(self.params, lr=lr, eps=1e-5)
        self.lr, self.num_iterations = lr, num_iterations
        self.obs_rms, self.value_rms = ppo_oracle.rms_init((se.OBS_DIM,)), ppo_oracle.rms_init(())
        self.seed, self.offset = rng_seed, 0
        self.obs = torch.zeros(T, N, se.OBS_DIM)
        self.actions = torch.zeros(T, N, se.ACT_DIM)
        self.logprobs, self.rewards, self.values = torch.zeros(T, N), torch.zeros(T, N), torch.zeros(T, N)
        self.dones, self.true_dones = torch.zeros(T + 1, N), torch.zeros(T + 1, N)
        self.next_obs = self._norm(self.env.obs_buf["policy"])
        self.iteration = 0

    def _norm(self, raw):
        self.obs_rms = ppo_oracle.rms_update(self.obs_rms, raw)
        return ppo_oracle.rms_normalize(self.obs_rms, raw)

    def _noise(self):
        z = philox_oracle.normal(self.seed, self.offset, N * se.ACT_DIM)
        self.offset += N * se.ACT_DIM
        return torch.from_numpy(z).reshape(N, se.ACT_DIM)

    def iterate(self):
        self.iteration += 1
        for g in self.opt.param_groups:  # ppo.py:196-199
            g["lr"] = (1.0 - (self.iteration - 1.0) / self.num_iterations) * self.lr
        env = self.env
        for t in range(T):
            self.obs[t] = self.next_obs
            with torch.no_grad():
                action, logp, value = self.agent.act(self.next_obs, self._noise())
            self.actions[t], self.logprobs[t], self.values[t] = action, logp, value.flatten()
            env._advance()
            env.episode_length_buf += 1
            env.common_step_counter += 1
            reset = env.episode_length_buf >= env.max_episode_length
            cstr = self.mgr.compute()
            reward, dones = cat_oracle.step_epilogue(env._raw_reward, cstr, reset)
            if bool(reset.any()):  # curriculum off on both sides: max_p stays at the cfg values
                self.mgr.reset(reset.nonzero().flatten())
                env.episode_length_buf[reset] = 0
            self.rewards[t], self.dones[t + 1], self.true_dones[t + 1] = reward, dones, reset.float()
            self.next_obs = self._norm(env.obs_buf["policy"])
        with torch.no_grad():
            nv = self.agent.critic(self.next_obs).reshape(1, -1)
            adv, ret = ppo_oracle.gae(self.rewards, self.values, self.dones[:-1], self.true_dones[:-1], nv, self.dones[-1], self.true_dones[-1])
            self.value_rms, b_values, b_returns = ppo_oracle.value_normalisation(self.value_rms, self.values.reshape(-1), ret.reshape(-1))
        b_adv = adv.reshape(-1)
        B = N * T
        sums = {"pg_loss": 0.0, "v_loss": 0.0, "entropy": 0.0, "approx_kl": 0.0, "clipfrac": 0.0}
        n_mb = 0
        for _ in range(EPOCHS):
            perm = torch.from_numpy(philox_oracle.random_permutation(B, self.seed, self.offset))
            self.offset += 2
            for s in range(0, B, MB):
                idx = perm[s : s + MB]
                loss, info = ppo_oracle.ppo_minibatch_loss(
                    self.agent, self.value_rms, self.obs.reshape(B, -1)[idx], self.actions.reshape(B, -1)[idx],
                    self.logprobs.reshape(-1)[idx], b_adv[idx], b_returns[idx], b_values[idx],
                )  # fmt: skip
                self.opt.zero_grad()
                loss.backward()
                torch.nn.utils.clip_grad_norm_(self.params, 1.0)
                self.opt.step()
                for k in sums:
                    sums[k] += float(info[k])
                n_mb += 1
        self.dones[0], self.true_dones[0] = self.dones[T].clone(), self.true_dones[T].clone()
        return {k: v / n_mb for k, v in sums.items()}

    def flat(self):
        return torch.cat([p.detach().reshape(-1) for p in self.params])


def _run(precision):
    torch.manual_seed(0)
    env = se.SyntheticSolo12Env(N, device=DEV, seed=ENV_SEED, pool=3, episode_length=60, constraints_cfg=se.solo12_constraints_cfg(), curriculum=False)
    env.load_managers()
    cfg = solo12_flat_ppo_cfg(logger=None, num_steps=T, minibatch_size=MB, updates_epochs=EPOCHS, num_iterations=ITERS)
    tr = PPOTrainer(env, cfg, device=DEV, use_graphs=True, seed=5, precision=precision)
    init = {k: v.detach().cpu().clone() for k, v in tr.agent.state_dict().items()}
    seed = int(tr.rng_state[0])
    ora = OracleTrainer(init, seed, cfg.learning_rate, ITERS)
    tr.start()
    rows = []
    for _ in range(ITERS):
        tr.train_iteration()
        g = tr.losses()
        o = ora.iterate()
        rows.append((g, o))
    assert int(tr.rng_state[1]) == ora.offset, "both sides consumed the same random numbers"
    got = tr.agent.parameters_flat().detach().cpu()
    want = ora.flat()
    p0 = torch.cat([init[k].reshape(-1) for k in init if "_rms." not in k])
    return rows, got, want, p0, init, seed, cfg.learning_rate


def _table(rows, got, want, p0, name):
    pg = np.array([[g["mean_pg_loss"], o["pg_loss"]] for g, o in rows])
    vl = np.array([[g["mean_v_loss"], o["v_loss"]] for g, o in rows])
    kl = np.array([[g["approx_kl"], o["approx_kl"]] for g, o in rows])
    en = np.array([[g["mean_entropy_loss"], o["entropy"]] for g, o in rows])
    cf = np.array([[g["clipfrac"], o["clipfrac"]] for g, o in rows])
    lines = [f"{name} CUDA trainer vs fp32 oracle trainer, {N} envs x {T} steps, {EPOCHS} epochs x {N * T // MB} minibatches of {MB}",
             "iter   pg(gpu)   pg(cpu)    v(gpu)    v(cpu)   kl(gpu)   kl(cpu)  clipfrac gpu/cpu   entropy gpu/cpu"]  # fmt: skip
    for i in range(ITERS):
        lines.append(f"{i:3d} {pg[i,0]:9.5f} {pg[i,1]:9.5f} {vl[i,0]:9.5f} {vl[i,1]:9.5f} {kl[i,0]:9.6f} {kl[i,1]:9.6f} {cf[i,0]:.4f}/{cf[i,1]:.4f} {en[i,0]:.5f}/{en[i,1]:.5f}")
    upd_g, upd_o = got - p0, want - p0
    cos = float(torch.dot(upd_g, upd_o) / (upd_g.norm() * upd_o.norm()))
    rel = float((got - want).norm() / upd_o.norm())
    lines.append(f"parameter update after {ITERS} iterations ({ITERS * EPOCHS * (N * T // MB)} Adam steps): cosine {cos:.5f}, |gpu - cpu| / |update| = {rel:.4f}")
    print("\n" + "\n".join(lines))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):  # kept as evidence (copied to profiles/)
        open(os.path.join(out_dir, f"equivalence_{name}.txt"), "w").write("\n".join(lines) + "\n")
    return pg, vl, kl, en, cf, cos, rel


def _natural_drift(init, seed, lr):
    """Per-metric max |difference| over the run between the fp32 oracle trainer and itself started from weights
    perturbed by 2^-11 relative (two perturbation signs; the larger drift counts)."""
    def run(pert):
        g = torch.Generator().manual_seed(123)
        sd = {k: (v * (1 + pert * torch.randn(v.shape, generator=g)) if v.is_floating_point() and "_rms." not in k else v) for k, v in init.items()}
        o = OracleTrainer(sd, seed, lr, ITERS)
        return [o.iterate() for _ in range(ITERS)], o.flat()

    base, base_p = run(0.0)
    drift = {k: 0.0 for k in ("pg_loss", "v_loss", "entropy", "approx_kl", "clipfrac")}
    prel = 0.0
    for pert in (2.0**-11, -(2.0**-11)):
        rows, p = run(pert)
        for k in drift:
            drift[k] = max(drift[k], max(abs(x[k] - y[k]) for x, y in zip(rows, base)))
        prel = max(prel, float((p - base_p).norm()))
    return drift, prel


def test_tf32_trainer_tracks_the_fp32_oracle_trainer():
    rows, got, want, p0, init, seed, lr = _run("tf32")
    pg, vl, kl, en, cf, cos, rel = _table(rows, got, want, p0, "tf32")
    # iteration 1 (same parameters on both sides, only the GEMM arithmetic differs): logged scalars agree to 1e-4
    assert abs(pg[0, 0] - pg[0, 1]) < 1e-4 and abs(vl[0, 0] - vl[0, 1]) < 1e-4 * max(1.0, vl[0, 1]) and abs(kl[0, 0] - kl[0, 1]) < 1e-5
    # whole run: within 2x the drift between two fp32 runs that differ by a tf32-sized perturbation of the weights
    nat, nat_p = _natural_drift(init, seed, lr)
    gpu = {"pg_loss": np.abs(pg[:, 0] - pg[:, 1]).max(), "v_loss": np.abs(vl[:, 0] - vl[:, 1]).max(), "entropy": np.abs(en[:, 0] - en[:, 1]).max(),
           "approx_kl": np.abs(kl[:, 0] - kl[:, 1]).max(), "clipfrac": np.abs(cf[:, 0] - cf[:, 1]).max()}  # fmt: skip
    gpu_p = float((got - want).norm())
    lines = ["max |difference| over the run:  metric   CUDA tf32 vs fp32 oracle   fp32 oracle vs fp32 oracle perturbed by 2^-11"]
    for k in nat:
        lines.append(f"   {k:10s} {gpu[k]:.3e}   {nat[k]:.3e}")
    lines.append(f"   |params|   {gpu_p:.3e}   {nat_p:.3e}")
    print("\n".join(lines))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        open(os.path.join(out_dir, "equivalence_tf32.txt"), "a").write("\n".join(lines) + "\n")
    for k in nat:
        assert gpu[k] <= 2.0 * nat[k] + 1e-4, f"{k}: CUDA-vs-oracle drift {gpu[k]:.3e} exceeds 2x the fp32 perturbation drift {nat[k]:.3e}"
    assert gpu_p <= 2.0 * nat_p
    assert cos > 0.99 and rel < 0.15


def test_bf16_trainer_drift_is_recorded():
    """The bf16 operand mode next to it, for the record (no band claimed beyond the first iteration and the direction)."""
    pg, vl, kl, en, cf, cos, rel = _table(*_run("bf16")[:4], "bf16")
    assert abs(pg[0, 0] - pg[0, 1]) < 2e-3 and abs(vl[0, 0] - vl[0, 1]) < 2e-2 * max(1.0, vl[0, 1])
    assert cos > 0.9
