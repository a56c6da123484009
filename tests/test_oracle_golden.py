"""Pin the CPU oracle (oracle/*.py) to outputs of the REAL reference (tests/golden/, made by
oracle/make_golden.py from /root/reference).  CPU only; bit-exact unless a tolerance is stated."""

import os

import pytest
import torch

from constraints_as_terminations_b200 import synthetic_env as se
from oracle import cat_oracle, ppo_oracle
from tests.helpers import replay_cat_golden


def _load(golden_dir, name):
    path = os.path.join(golden_dir, name)
    if not os.path.isfile(path):
        pytest.skip(f"{name} missing")
    return torch.load(path, weights_only=False)


@pytest.mark.parametrize("fixture", ["cat_solo12.pt", "cat_stress.pt"])
def test_cat_oracle_matches_reference_bit_exact(golden_dir, fixture):
    gold = _load(golden_dir, fixture)
    n = gold["num_envs"]
    env = se.SyntheticSolo12Env(n, device="cpu", seed=gold["seed"], pool=1, adversarial=True)
    cfg = se.solo12_constraints_cfg(stress=gold["stress"])
    terms = cat_oracle.terms_from_cfg(cfg, resolve_scene=env.scene)
    assert [t[0] for t in terms] == gold["names"]
    mgr = cat_oracle.ManagerOracle(env, terms)

    def step_fn(step, state, rec):
        cstr_prob = mgr.compute()
        reward, dones = cat_oracle.step_epilogue(state["raw_reward"], cstr_prob, rec["reset_buf"])
        out = {
            "cstr_prob": cstr_prob,
            "running_max": torch.cat(list(mgr.cat.running_max.values()), dim=1).squeeze(0),
            "reward": reward,
            "dones": dones,
        }
        if "raw" in rec:
            out["raw"] = torch.cat(list(mgr.cat.raw.values()), dim=1)
            out["probs"] = torch.cat(list(mgr.cat.probs.values()), dim=1)
        return out

    def reset_fn(env_ids):
        return mgr.reset(env_ids)

    def set_max_p(values):
        for (name, _, _, _), term_cfg, v in zip(terms, cfg.values(), values):
            term_cfg.max_p = v

    replay_cat_golden(gold, env, step_fn, reset_fn, set_max_p, exact=True)
    assert torch.equal(torch.stack(list(mgr.episode_sums.values())), gold["episode_sums"])
    assert torch.equal(torch.stack(list(mgr.mean_values.values())), gold["mean_values"])


def test_gae_oracle_matches_reference_bit_exact(golden_dir):
    g = _load(golden_dir, "ppo_iter.pt")
    adv, ret = ppo_oracle.gae(
        g["rewards"], g["values"], g["dones"], g["true_dones"], g["next_value"], g["next_done"], g["next_true_done"],
        g["cfg"]["gamma"], g["cfg"]["gae_lambda"],
    )  # fmt: skip
    assert torch.equal(adv, g["advantages"])
    assert torch.equal(ret, g["returns"])


def test_running_moments_oracle_matches_reference(golden_dir):
    g = _load(golden_dir, "ppo_iter.pt")
    torch.set_num_threads(1)
    # obs_rms: reset obs, then every step's raw obs (ppo.py:187,225); stored obs are the normalised ones
    st = ppo_oracle.rms_init((se.OBS_DIM,))
    st = ppo_oracle.rms_update(st, g["env_trace_reset_obs"])
    first = ppo_oracle.rms_normalize(st, g["env_trace_reset_obs"])
    assert torch.equal(first, g["obs"][0])
    for t in range(g["env_trace_raw_obs"].shape[0]):
        st = ppo_oracle.rms_update(st, g["env_trace_raw_obs"][t])
        normed = ppo_oracle.rms_normalize(st, g["env_trace_raw_obs"][t])
        if t + 1 < g["obs"].shape[0]:
            assert torch.equal(normed, g["obs"][t + 1])
        else:
            assert torch.equal(normed, g["next_obs"])
    ref = g["rms_after_rollout"]
    assert torch.equal(st["mean"], ref["obs_rms.running_mean"])
    assert torch.equal(st["var"], ref["obs_rms.running_var"])
    assert torch.equal(st["count"], ref["obs_rms.count"])
    # value_rms double update (ppo.py:287-288)
    s2, values_n, returns_n = ppo_oracle.value_normalisation(
        ppo_oracle.rms_init(()), g["values"].reshape(-1), g["returns"].reshape(-1)
    )
    assert torch.equal(values_n, g["b_values"])
    assert torch.equal(returns_n, g["b_returns"])
    assert torch.equal(s2["mean"], ref["value_rms.running_mean"])
    assert torch.equal(s2["count"], ref["value_rms.count"])


def test_ppo_update_oracle_matches_reference(golden_dir):
    """Replay the minibatch updates of one reference iteration with the oracle loss + torch Adam."""
    g = _load(golden_dir, "ppo_iter.pt")
    torch.set_num_threads(1)
    cfg = g["cfg"]
    agent = ppo_oracle.AgentOracle(se.OBS_DIM, se.ACT_DIM)
    init = {k: v for k, v in g["init_state"].items() if "_rms." not in k}
    agent.load_state_dict(init)
    # Adam over parameters in the reference's registration order (critic, actor_mean, actor_logstd)
    params = list(agent.critic.parameters()) + list(agent.actor_mean.parameters()) + [agent.actor_logstd]
    opt = torch.optim.Adam(params, lr=g["lr"], eps=1e-5)
    ref_rms = g["rms_after_rollout"]
    value_rms = {"mean": ref_rms["value_rms.running_mean"], "var": ref_rms["value_rms.running_var"]}
    b_obs = g["obs"].reshape(-1, se.OBS_DIM)
    b_act = g["actions"].reshape(-1, se.ACT_DIM)
    b_logp, b_adv = g["logprobs"].reshape(-1), g["advantages"].reshape(-1)
    sums = {"pg_loss": 0.0, "v_loss": 0.0, "entropy": 0.0}
    mbs = cfg["minibatch_size"]
    for perm in g["perms"]:
        for start in range(0, perm.numel(), mbs):
            idx = perm[start : start + mbs]
            loss, info = ppo_oracle.ppo_minibatch_loss(
                agent, value_rms, b_obs[idx], b_act[idx], b_logp[idx], b_adv[idx], g["b_returns"][idx],
                g["b_values"][idx], cfg["clip_coef"], cfg["ent_coef"], cfg["vf_coef"],
            )  # fmt: skip
            for k in sums:
                sums[k] += float(info[k])
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, cfg["max_grad_norm"])
            opt.step()
    assert sums["pg_loss"] == pytest.approx(g["sum_pg_loss"], rel=1e-5, abs=1e-7)
    assert sums["v_loss"] == pytest.approx(g["sum_v_loss"], rel=1e-5)
    assert sums["entropy"] == pytest.approx(g["sum_entropy_loss"], rel=1e-6)
    final = g["final_state"]
    for k, v in agent.state_dict().items():
        torch.testing.assert_close(v, final[k], rtol=1e-5, atol=1e-7, msg=lambda m, k=k: f"{k}: {m}")


def test_rollout_action_oracle_matches_reference(golden_dir):
    """log-prob / value of the stored rollout actions under the initial weights (ppo.py:208-212)."""
    g = _load(golden_dir, "ppo_iter.pt")
    torch.set_num_threads(1)
    agent = ppo_oracle.AgentOracle(se.OBS_DIM, se.ACT_DIM)
    agent.load_state_dict({k: v for k, v in g["init_state"].items() if "_rms." not in k})
    with torch.no_grad():
        logp, _, value = agent.evaluate(g["obs"][0], g["actions"][0])
    torch.testing.assert_close(logp, g["logprobs"][0], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(value.flatten(), g["values"][0], rtol=1e-6, atol=1e-6)


def test_gae_variant_oracles_match_golden(golden_dir):
    """skrl compute_gae restatement vs outputs of the reference's own function (bit-exact raw, 1e-6 normalised);
    rl_games discount_values restatement: regression vs the stored vectors and cross-check against the pinned
    CleanRL GAE, to which it reduces when nothing times out."""
    from oracle import gae_variants_oracle as go

    golden = _load(golden_dir, "gae_variants.pt")
    assert golden["skrl_pinned"] and not golden["rlgames_pinned"]
    for case in golden["cases"]:
        T, N, seed = case["T"], case["N"], case["seed"]
        rewards, values, dones, last_values = go.sample_inputs(T, N, seed)
        ret, adv = go.skrl_compute_gae(rewards, dones[:T], values, last_values)
        assert torch.equal(ret, case["skrl_returns"])
        if T * N > 1:
            torch.testing.assert_close(adv, case["skrl_advantages"], rtol=1e-6, atol=1e-6)
        rg = go.rlgames_discount_values(dones[T], last_values.unsqueeze(-1), dones[:T], values.unsqueeze(-1), rewards.unsqueeze(-1))
        assert torch.equal(rg.squeeze(-1), case["rlgames_advs"])
        zeros = torch.zeros(T, N)
        adv_c, _ = ppo_oracle.gae(rewards, values, dones[:T], zeros, last_values, dones[T], torch.zeros(N))
        assert torch.equal(rg.squeeze(-1), adv_c)


def test_skrl_oracle_matches_reference_source():
    """When /root/reference is present (build container), run the reference's own compute_gae next to the oracle."""
    from oracle import gae_variants_oracle as go
    from oracle import make_golden_gae, ref_loader

    if not os.path.isfile(make_golden_gae.SKRL_PPO):
        pytest.skip("reference tree absent (GPU box)")
    assert ref_loader.reference_available()
    ref_gae = make_golden_gae.load_reference_skrl_compute_gae()
    rewards, values, dones, last_values = go.sample_inputs(24, 96, 11)
    args = (rewards.unsqueeze(-1), dones[:24].unsqueeze(-1), values.unsqueeze(-1), last_values.unsqueeze(-1))
    ret, adv = ref_gae(*args, 0.99, 0.95)
    o_ret, o_adv = go.skrl_compute_gae(*args)
    assert torch.equal(ret, o_ret) and torch.equal(adv, o_adv)
