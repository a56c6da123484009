"""ORACLE tooling (build container only): golden vectors for the single-`dones` GAE variants.

Run:  python oracle/make_golden_gae.py        (needs /root/reference; writes tests/golden/gae_variants.pt)

skrl: the reference keeps `compute_gae` as a function nested inside `PPO._update`
(`exts/cat_envs/cat_envs/tasks/utils/skrl/ppo.py:397-442`); skrl itself is not installed, so the module
cannot be imported.  The function's own source is cut out of the reference file with `ast`, compiled as is
and executed here (its free variable `last_values` is supplied as a global, exactly what the enclosing
method binds at :447-452).  Nothing is copied into the repo: only the outputs are stored.

rl_games: `discount_values` lives in the rl_games package (third party, absent) -> no reference execution
possible; the fixture stores the oracle restatement's outputs for regression only and says so (`pinned: False`).
"""

from __future__ import annotations

import ast
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import gae_variants_oracle as go  # noqa: E402
from oracle import ref_loader  # noqa: E402

SKRL_PPO = os.path.join(ref_loader.REFERENCE_ROOT, "exts/cat_envs/cat_envs/tasks/utils/skrl/ppo.py")


def load_reference_skrl_compute_gae():
    """-> callable(rewards, dones, values, last_values, discount_factor, lambda_coefficient) running the reference code."""
    source = open(SKRL_PPO).read()
    tree = ast.parse(source)
    found = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "compute_gae"]
    assert len(found) == 1, "expected exactly one compute_gae in the reference skrl agent"
    module = ast.Module(body=[found[0]], type_ignores=[])
    ns = {"torch": torch}
    exec(compile(module, SKRL_PPO, "exec"), ns)
    fn = ns["compute_gae"]

    def call(rewards, dones, values, last_values, discount_factor, lambda_coefficient):
        ns["last_values"] = last_values  # the closure variable of PPO._update (:447-452)
        return fn(rewards=rewards, dones=dones, values=values, next_values=last_values,
                  discount_factor=discount_factor, lambda_coefficient=lambda_coefficient)

    return call


def main():
    ref_gae = load_reference_skrl_compute_gae()
    cases = []
    for T, N, seed in ((24, 64, 0), (24, 257, 1), (5, 33, 2), (1, 8, 3)):
        rewards, values, dones, last_values = go.sample_inputs(T, N, seed)
        # skrl memory tensors are [T, N, 1]
        ret, adv = ref_gae(rewards.unsqueeze(-1), dones[:T].unsqueeze(-1), values.unsqueeze(-1), last_values.unsqueeze(-1), 0.99, 0.95)
        o_ret, o_adv = go.skrl_compute_gae(rewards.unsqueeze(-1), dones[:T].unsqueeze(-1), values.unsqueeze(-1), last_values.unsqueeze(-1))
        assert torch.equal(ret, o_ret) and torch.equal(adv, o_adv), "oracle restatement differs from the reference"
        rg = go.rlgames_discount_values(dones[T], last_values.unsqueeze(-1), dones[:T], values.unsqueeze(-1), rewards.unsqueeze(-1))
        cases.append({
            "T": T, "N": N, "seed": seed,
            "skrl_returns": ret.squeeze(-1).clone(), "skrl_advantages": adv.squeeze(-1).clone(),
            "rlgames_advs": rg.squeeze(-1).clone(),
        })
    out = {"cases": cases, "skrl_pinned": True, "rlgames_pinned": False,
           "note": "skrl_* produced by executing the reference's own compute_gae; rlgames_* by the oracle restatement (rl_games absent)"}
    path = os.path.join(ROOT, "tests", "golden", "gae_variants.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
