"""Shared replay logic: drive an implementation (oracle on CPU, CUDA product on the GPU) through the
exact step protocol `oracle/make_golden.py` used with the real reference, comparing every output."""

from __future__ import annotations

import torch

from constraints_as_terminations_b200 import synthetic_env as se


def state_checksum(state: dict) -> float:
    return float(sum(v.double().sum().item() for v in state.values()))


def assert_same(actual: torch.Tensor, expected: torch.Tensor, exact: bool, what: str, rtol=1e-5, atol=1e-7):
    actual = actual.detach().cpu()
    if exact:
        if not torch.equal(actual, expected):
            diff = (actual.double() - expected.double()).abs()
            raise AssertionError(f"{what}: not bit-identical, max |diff| {diff.max().item():.3e} at {diff.argmax().item()}")
    else:
        torch.testing.assert_close(actual, expected, rtol=rtol, atol=atol, msg=lambda m: f"{what}: {m}")


def replay_cat_golden(gold, env, step_fn, reset_fn, set_max_p, exact: bool, device="cpu", exact_keys=None):
    """`step_fn(step, state, rec) -> dict` of outputs named like the fixture; `reset_fn(env_ids) -> dict`."""
    n = gold["num_envs"]
    gen = torch.Generator().manual_seed(gold["seed"] + 1000)
    exact_keys = exact_keys or set()
    for step, rec in enumerate(gold["steps"]):
        state = se.sample_state(n, gen, adversarial=(step % 3 == 0))
        assert state_checksum(state) == rec["checksum"], "synthetic inputs differ from the fixture's"
        _ = torch.rand(n, generator=gen)  # the reset_buf draw of the generator (kept in rec)
        dev_state = {k: v.to(device) for k, v in state.items()}
        env.load_state(dev_state)
        env.episode_length_buf += 1
        env.common_step_counter = step * 400
        set_max_p(rec["max_p"])
        rec_dev = dict(rec)
        rec_dev["reset_buf"] = rec["reset_buf"].to(device)
        out = step_fn(step, dev_state, rec_dev)
        for key, val in out.items():
            assert_same(val, rec[key], exact or key in exact_keys, f"step {step} {key}")
        # bit-exact "mask" outputs whatever the float tolerance: violation mask and hard-done indices
        if "raw" in out and "probs" in out:
            assert torch.equal((out["probs"].cpu() > 0), (rec["raw"] > 0)), f"step {step}: violation mask"
        assert torch.equal((out["dones"].cpu() >= 1).nonzero(), (rec["dones"] >= 1).nonzero()), f"step {step}: done ids"
        if "reset_out" in rec:
            sel = rec["reset_ids"]
            env_ids = None if sel is None else torch.tensor(sel, dtype=torch.long, device=device)
            got = reset_fn(env_ids)
            assert list(got.keys()) == list(rec["reset_out"].keys())
            for k, v in got.items():
                # means over envs: torch's reduction order is implementation defined -> 1e-5 relative
                torch.testing.assert_close(v.detach().cpu(), rec["reset_out"][k], rtol=1e-5, atol=1e-7, msg=lambda m, k=k: f"reset {k}: {m}")
            if env_ids is None:
                env.episode_length_buf[:] = 0
            else:
                env.episode_length_buf[env_ids] = 0
