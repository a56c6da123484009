"""GPU parity of the single-`dones` GAE kernels through the C ABI.

skrl variant: against golden vectors produced by executing the reference's own `compute_gae`
(`oracle/make_golden_gae.py`) -- returns and raw advantages bit-exact, normalised advantages 1e-5.
rl_games variant: against the oracle restatement of rl_games' published `discount_values` (third-party code that is
not part of /root/reference: parity unpinned there), bit-exact.
"""

import os

import pytest
import torch

from oracle import gae_variants_oracle as go

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "gae_variants.pt")


def _skrl(T, N, seed, normalize):
    from constraints_as_terminations_b200.skrl import compute_gae

    rewards, values, dones, last_values = go.sample_inputs(T, N, seed)
    args = [t.unsqueeze(-1).to(DEV) for t in (rewards, dones[:T], values, last_values)]
    ret, adv = compute_gae(*args, discount_factor=0.99, lambda_coefficient=0.95, normalize=normalize)
    assert ret.shape == (T, N, 1) and adv.shape == (T, N, 1)
    return ret.squeeze(-1).cpu(), adv.squeeze(-1).cpu()


def test_skrl_gae_matches_reference_golden():
    golden = torch.load(GOLDEN)
    assert golden["skrl_pinned"]
    for case in golden["cases"]:
        T, N, seed = case["T"], case["N"], case["seed"]
        ret, adv = _skrl(T, N, seed, normalize=True)
        assert torch.equal(ret, case["skrl_returns"]), f"returns differ T={T} N={N}"
        if T * N > 1:
            # the reference normalises with torch's fp32 mean / std (reduction order implementation defined);
            # the kernel reduces in double: 1e-5 relative (north_star tolerance), written here
            torch.testing.assert_close(adv, case["skrl_advantages"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("T,N", [(24, 4096), (24, 1000), (3, 31), (1, 1), (24, 200_000)])
def test_skrl_gae_raw_bit_exact_vs_oracle(T, N):
    rewards, values, dones, last_values = go.sample_inputs(T, N, seed=T * 7 + N)
    o_ret, o_adv = go.skrl_compute_gae(rewards, dones[:T], values, last_values, normalize=False)
    ret, adv = _skrl(T, N, T * 7 + N, normalize=False)
    assert torch.equal(ret, o_ret)
    assert torch.equal(adv, o_adv)
    # and the normalised variant: statistics within 1e-5 of torch's
    if T * N > 1:
        _, adv_n = _skrl(T, N, T * 7 + N, normalize=True)
        _, o_adv_n = go.skrl_compute_gae(rewards, dones[:T], values, last_values, normalize=True)
        torch.testing.assert_close(adv_n, o_adv_n, rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("T,N", [(24, 4096), (16, 777), (2, 33), (1, 5), (24, 200_000)])
def test_rlgames_discount_values_bit_exact_vs_oracle(T, N):
    from constraints_as_terminations_b200.rl_games import CaTDiscountMixin, discount_values

    rewards, values, dones, last_values = go.sample_inputs(T, N, seed=T * 13 + N)
    o_advs = go.rlgames_discount_values(dones[T], last_values.unsqueeze(-1), dones[:T], values.unsqueeze(-1), rewards.unsqueeze(-1))
    args = (dones[T].to(DEV), last_values.unsqueeze(-1).to(DEV), dones[:T].to(DEV), values.unsqueeze(-1).to(DEV), rewards.unsqueeze(-1).to(DEV))
    advs = discount_values(*args, gamma=0.99, tau=0.95)
    assert advs.shape == (T, N, 1)
    assert torch.equal(advs.cpu(), o_advs)

    class Agent(CaTDiscountMixin):
        gamma, tau = 0.99, 0.95

    assert torch.equal(Agent().discount_values(*args).cpu(), o_advs)


def test_rlgames_golden_regression():
    from constraints_as_terminations_b200.rl_games import discount_values

    golden = torch.load(GOLDEN)
    for case in golden["cases"]:
        T, N, seed = case["T"], case["N"], case["seed"]
        rewards, values, dones, last_values = go.sample_inputs(T, N, seed)
        advs = discount_values(dones[T].to(DEV), last_values.unsqueeze(-1).to(DEV), dones[:T].to(DEV), values.unsqueeze(-1).to(DEV),
                               rewards.unsqueeze(-1).to(DEV), gamma=0.99, tau=0.95)
        assert torch.equal(advs.squeeze(-1).cpu(), case["rlgames_advs"])


def test_float_dones_gae_rejects_bad_arguments():
    from constraints_as_terminations_b200 import _lib as L
    from constraints_as_terminations_b200 import ops

    r = torch.zeros(4, 8, device=DEV)
    with pytest.raises(ValueError):
        ops.gae_float_dones(L.GAE_RLGAMES, r, r, r, torch.zeros(8, device=DEV), 0.99, 0.9)  # no last_dones
    with pytest.raises(RuntimeError):
        ops.gae_float_dones(7, r, r, r, torch.zeros(8, device=DEV), 0.99, 0.9)  # unknown variant
    with pytest.raises(RuntimeError):
        ops.gae_float_dones(L.GAE_SKRL, r.cpu(), r.cpu(), r.cpu(), torch.zeros(8), 0.99, 0.9)
