"""GPU parity of the actor-critic MLP kernels, the PPO minibatch gradient and the Adam step.

The hidden-layer GEMMs run on tcgen05 in one of two operand precisions, both with fp32 accumulation:
  tf32 (default) : the reference's own GPU numerics (torch.backends.cuda.matmul.allow_tf32,
                   scripts/clean_rl/train.py:86-87): fp32 storage, operands rounded to a 10-bit mantissa
  bf16           : half the operand bytes, 8-bit mantissa
Each is compared
  (a) tightly against the fp32 oracle with the operand rounding emulated at the points the kernels round, and
  (b) against the plain fp32 oracle with the tolerance that precision allows (stated in TOL / at each assert)."""

import math

import pytest
import torch

from constraints_as_terminations_b200 import ops
from oracle import ppo_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
OBS, ACT = 45, 12


def flat_params(agent: ppo_oracle.AgentOracle, layout) -> torch.Tensor:
    """Oracle parameters -> the flat fp32 vector in libcatb200's layout (== reference parameter order)."""
    flat = torch.zeros(layout.n_params)
    for z, net in enumerate((agent.critic, agent.actor_mean)):
        for l, idx in enumerate((0, 2, 4, 6)):
            w, b = net[idx].weight.detach(), net[idx].bias.detach()
            flat[layout.w[z][l] : layout.w[z][l] + w.numel()] = w.reshape(-1)
            flat[layout.b[z][l] : layout.b[z][l] + b.numel()] = b
    flat[layout.logstd : layout.logstd + ACT] = agent.actor_logstd.detach().reshape(-1)
    return flat


def flat_grads(agent, layout) -> torch.Tensor:
    flat = torch.zeros(layout.n_params)
    for z, net in enumerate((agent.critic, agent.actor_mean)):
        for l, idx in enumerate((0, 2, 4, 6)):
            w, b = net[idx].weight.grad, net[idx].bias.grad
            flat[layout.w[z][l] : layout.w[z][l] + w.numel()] = w.reshape(-1)
            flat[layout.b[z][l] : layout.b[z][l] + b.numel()] = b
    flat[layout.logstd : layout.logstd + ACT] = agent.actor_logstd.grad.reshape(-1)
    return flat


def flat_slices(layout) -> dict:
    """name -> (begin, end) of every parameter tensor inside the flat vector."""
    offs = sorted([*layout.w[0], *layout.b[0], *layout.w[1], *layout.b[1], layout.logstd, layout.n_params])
    out = {}
    for z in range(2):
        for l in range(4):
            for kind, off in (("w", layout.w[z][l]), ("b", layout.b[z][l])):
                out[f"net{z}.layer{l}.{kind}"] = (off, next(o for o in offs if o > off))
    out["logstd"] = (layout.logstd, layout.logstd + ACT)
    return out


def test_layout_follows_reference_parameter_order():
    agent = ppo_oracle.AgentOracle(OBS, ACT)
    layout = ops.mlp_layout(ops.make_dims(OBS, ACT))
    # reference registration order: critic.*, actor_mean.*, actor_logstd (ppo.py:78-97) -> agent.parameters()
    want = torch.cat([p.detach().reshape(-1) for p in list(agent.critic.parameters()) + list(agent.actor_mean.parameters()) + [agent.actor_logstd]])
    assert layout.n_params == want.numel() == 377241
    assert torch.equal(flat_params(agent, layout), want)


def tf32_round(x: torch.Tensor) -> torch.Tensor:
    """cvt.rna.tf32.f32: round the magnitude to a 10-bit mantissa, ties away from zero."""
    bits = x.detach().contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    bits = (bits + 0x1000) & 0xFFFFE000
    bits = torch.where(bits >= 2**31, bits - 2**32, bits)
    return bits.to(torch.int32).view(torch.float32).view_as(x)


def make_round(prec):  # operand rounding with fp32 storage and a straight-through gradient
    if prec == "bf16":
        return lambda x: x + (x.to(torch.bfloat16).float() - x).detach()
    return lambda x: x + (tf32_round(x) - x).detach()


def emulated_forward(agent, obs, prec):
    """fp32 oracle with the operand rounding where the kernels round: inputs, hidden weights, hidden activations."""
    rnd = make_round(prec)
    outs = []
    for net in (agent.critic, agent.actor_mean):
        h = rnd(obs)
        for idx in (0, 2, 4):
            h = rnd(torch.nn.functional.elu(h @ rnd(net[idx].weight).T + net[idx].bias))
        outs.append(h @ net[6].weight.T + net[6].bias)
    return outs[1], outs[0]  # action mean, value


# tolerances per precision.  tf32: products of two 11-bit-significand operands, fp32 accumulation -> relative error
# ~2^-11 per product, far less after summation; bf16: 2^-8.
TOL = {
    "tf32": dict(fwd_emul=2e-3, fwd_mean_abs=1e-4, fwd_fp32=4e-3, loss_emul=5e-4, loss_fp32=2e-3, grad_emul=2e-3, grad_fp32=5e-3,
                 w_mult=3.0, b_mult=4.0, logstd_mult=4.0),
    "bf16": dict(fwd_emul=1e-2, fwd_mean_abs=5e-4, fwd_fp32=3e-2, loss_emul=2e-3, loss_fp32=2e-2, grad_emul=1.5e-2, grad_fp32=6e-2,
                 w_mult=3.5, b_mult=5.0, logstd_mult=8.0),
}
PRECS = ["tf32", "bf16"]


def make_agent(seed=0, scale_heads=True):
    torch.manual_seed(seed)
    agent = ppo_oracle.AgentOracle(OBS, ACT)
    with torch.no_grad():
        agent.actor_logstd.copy_(torch.linspace(-0.5, 0.3, ACT).reshape(1, ACT))
        for net in (agent.critic, agent.actor_mean):
            for idx in (0, 2, 4, 6):
                net[idx].bias.normal_(0, 0.1)
        if scale_heads:
            agent.actor_mean[6].weight.mul_(30.0)  # std=0.01 init makes the mean ~0: give it some signal
    return agent


def device_agent(agent, prec="tf32"):
    dims = ops.make_dims(OBS, ACT, precision=prec)
    layout = ops.mlp_layout(dims)
    params = flat_params(agent, layout).to(DEV)
    w16 = ops.weight_copies(dims, layout, DEV)
    ops.cast_weights(dims, params, w16)
    return dims, layout, params, w16


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("rows", [4096, 1000, 1, 129])
def test_rollout_forward_matches_oracle(rows, prec):
    tol = TOL[prec]
    agent = make_agent()
    dims, layout, params, w16 = device_agent(agent, prec)
    g = torch.Generator().manual_seed(rows)
    obs = torch.randn(rows, OBS, generator=g)
    noise = torch.randn(rows, ACT, generator=g)
    obs16 = ops.obs_to_operand(dims, obs.to(DEV))
    want_op = obs.to(torch.bfloat16).float() if prec == "bf16" else tf32_round(obs)
    assert torch.equal(obs16[:, :OBS].float().cpu(), want_op) and float(obs16[:, OBS:].abs().sum()) == 0
    ws = ops.mlp_workspace(dims, rows, False, DEV)
    action, logprob, value, mean = (torch.empty(rows, ACT, device=DEV), torch.empty(rows, device=DEV), torch.empty(rows, device=DEV), torch.empty(rows, ACT, device=DEV))
    ops.mlp_act(dims, obs16, params, w16, ws, noise=noise.to(DEV), action=action, logprob=logprob, value=value, mean_out=mean)
    with torch.no_grad():
        e_mean, e_value = emulated_forward(agent, obs, prec)
        f_action, f_logp, f_value = agent.act(obs, noise)
    # (a) emulated-rounding oracle: the accumulation order differs, which now and then flips the operand
    # rounding of a hidden activation (one flip = 2^-8 (bf16) / 2^-11 (tf32) relative on that activation), and a far
    # smaller error on average
    torch.testing.assert_close(mean.cpu(), e_mean, rtol=tol["fwd_emul"], atol=tol["fwd_emul"])
    torch.testing.assert_close(value.cpu(), e_value.flatten(), rtol=tol["fwd_emul"], atol=tol["fwd_emul"])
    assert float((value.cpu() - e_value.flatten()).abs().mean()) < tol["fwd_mean_abs"]
    # action / log-prob are exact functions of (mean, noise): check them against the device mean in fp32
    std = agent.actor_logstd.detach().exp()
    want_action = mean.cpu() + std * noise
    torch.testing.assert_close(action.cpu(), want_action, rtol=1e-6, atol=1e-6)
    want_logp = (-((action.cpu() - mean.cpu()) ** 2) / (2 * std**2) - agent.actor_logstd.detach() - math.log(math.sqrt(2 * math.pi))).sum(1)
    torch.testing.assert_close(logprob.cpu(), want_logp, rtol=1e-5, atol=1e-5)
    # (b) plain fp32 oracle: operand rounding -> 3e-2 (bf16) / 4e-3 (tf32) on O(1) outputs
    torch.testing.assert_close(value.cpu(), f_value.flatten(), rtol=tol["fwd_fp32"], atol=tol["fwd_fp32"])
    torch.testing.assert_close(action.cpu(), f_action, rtol=tol["fwd_fp32"], atol=tol["fwd_fp32"])
    # evaluating given actions (ppo.py:110) reproduces the log-prob
    lp2 = torch.empty(rows, device=DEV)
    ops.mlp_act(dims, obs16, params, w16, ws, action_in=action, logprob=lp2)
    torch.testing.assert_close(lp2, logprob, rtol=1e-6, atol=1e-6)
    # deterministic action == mean
    det = torch.empty(rows, ACT, device=DEV)
    ops.mlp_act(dims, obs16, params, w16, ws, action=det)
    assert torch.equal(det, mean)


def _minibatch(agent, B, M, seed, margin=None):
    """Synthetic rollout pool of B samples and M minibatch indices.  The PPO loss is piecewise: every sample sits on one
    side of the ratio clip (0.8 / 1.2), of the value clamp (+-0.2) and of the max() between clipped and unclipped
    value loss, and its gradient jumps when it changes sides.  With `margin`, the minibatch is drawn only from samples
    whose fp32-oracle quantities are at least `margin` away from every such boundary, so that an implementation whose
    outputs differ from the oracle's by less than the margin takes the same branches and the gradient comparison
    measures arithmetic, not branch flips."""
    g = torch.Generator().manual_seed(seed)
    obs = torch.randn(B, OBS, generator=g)
    noise = torch.randn(B, ACT, generator=g)
    with torch.no_grad():
        actions, logp, values = agent.act(obs.to(torch.bfloat16).float(), noise)
        logp = logp + 0.3 * torch.randn(B, generator=g)  # old policy differs: exercises both clip branches
    adv = torch.randn(B, generator=g) * 2.0 + 0.5
    values = values.flatten() + 0.2 * torch.randn(B, generator=g)
    returns = values + adv
    norm_stats = torch.tensor([0.1, 1.3, 0.15, 1.7])
    order = torch.randperm(B, generator=g)
    if margin is not None:
        with torch.no_grad():
            newlogp, _, v = agent.evaluate(obs, actions)
        ratio = (newlogp - logp).exp()
        nv = (v.flatten() - norm_stats[2]) / torch.sqrt(norm_stats[3] + 1e-8)
        val_n = (values - norm_stats[0]) / torch.sqrt(norm_stats[1] + 1e-8)
        ret_n = (returns - norm_stats[2]) / torch.sqrt(norm_stats[3] + 1e-8)
        diff = nv - val_n
        e_u, e_c = nv - ret_n, val_n + diff.clamp(-0.2, 0.2) - ret_n
        safe = ((ratio - 0.8).abs() > margin) & ((ratio - 1.2).abs() > margin) & ((diff.abs() - 0.2).abs() > margin)
        safe &= (diff.abs() < 0.2) | ((e_u.abs() - e_c.abs()).abs() > margin)
        order = order[safe[order]]
        assert order.numel() >= M, f"only {order.numel()} of {B} samples are {margin} away from every branch boundary"
    return obs, actions, logp, adv, returns, values, norm_stats, order[:M]


@pytest.mark.parametrize("prec,margin", [("tf32", 0.01), ("tf32", None), ("bf16", None)])
# 5000: ragged last 128-row tile
@pytest.mark.parametrize("B,M", [(24 * 1024, 16384), (3000, 1000), (768, 512), (7000, 5000)])
def test_minibatch_gradient_matches_oracle(B, M, prec, margin):
    """tf32 / margin 0.01: every sample of the minibatch takes the same PPO branches as in the fp32 oracle -> the
    gradient agrees with fp32 to 5e-3 (rel. Frobenius), what tf32 operand rounding allows.  Without the margin a few
    samples per thousand flip a clip branch (their log-prob / value differs by ~1e-3 from fp32) and each flip moves the
    gradient by that sample's whole contribution: the unrestricted comparisons use the looser bf16 bounds."""
    T = TOL[prec] if margin is not None or prec == "bf16" else {**TOL["bf16"], "loss_emul": TOL["tf32"]["loss_emul"], "loss_fp32": TOL["tf32"]["loss_fp32"]}
    agent = make_agent(seed=1)
    dims, layout, params, w16 = device_agent(agent, prec)
    obs, actions, logp, adv, returns, values, norm_stats, idx = _minibatch(agent, B, M, seed=B + M, margin=margin)
    obs16 = ops.obs_to_operand(dims, obs.to(DEV))
    grads = torch.zeros(layout.n_params, device=DEV)
    loss_acc = torch.zeros(8, device=DEV)
    ws = ops.mlp_workspace(dims, M, True, DEV)
    hp = ops.make_hparams()
    ops.ppo_minibatch_grad(dims, hp, idx.to(DEV), obs16, actions.to(DEV), logp.to(DEV), adv.to(DEV), returns.to(DEV),
                           values.to(DEV), norm_stats.to(DEV), params, w16, grads, loss_acc, ws)  # fmt: skip
    torch.cuda.synchronize()
    # oracle: value-normalised returns / values as the reference feeds them (ppo.py:287-288)
    val_n = (values - norm_stats[0]) / torch.sqrt(norm_stats[1] + 1e-8)
    ret_n = (returns - norm_stats[2]) / torch.sqrt(norm_stats[3] + 1e-8)
    value_rms = {"mean": norm_stats[2], "var": norm_stats[3]}

    class Emulated(ppo_oracle.AgentOracle):
        def evaluate(self, o, a):
            mean, value = emulated_forward(self, o, prec)
            std = torch.exp(self.actor_logstd.expand_as(mean))
            lp = -((a - mean) ** 2) / (2 * std**2) - std.log() - math.log(math.sqrt(2 * math.pi))
            ent = 0.5 + 0.5 * math.log(2 * math.pi) + std.log()
            return lp.sum(1), ent.sum(1), value

    results = {}
    for name, cls in (("fp32", ppo_oracle.AgentOracle), ("emulated", Emulated)):
        a = cls(OBS, ACT)
        a.load_state_dict(agent.state_dict())
        loss, info = ppo_oracle.ppo_minibatch_loss(a, value_rms, obs[idx], actions[idx], logp[idx], adv[idx], ret_n[idx], val_n[idx])
        loss.backward()
        results[name] = (float(loss), info, flat_grads(a, layout))
    got = grads.cpu()
    acc = loss_acc.cpu()
    for name, (loss, info, want) in results.items():
        tol = T["loss_emul"] if name == "emulated" else T["loss_fp32"]
        assert acc[7] == 1.0
        assert float(acc[0]) == pytest.approx(float(info["pg_loss"]), rel=tol, abs=tol)
        assert float(acc[1]) == pytest.approx(float(info["v_loss"]), rel=tol, abs=tol)
        assert float(acc[2]) == pytest.approx(float(info["entropy"]), rel=1e-5)
        assert float(acc[3]) == pytest.approx(float(info["approx_kl"]), rel=tol, abs=tol)
        assert float(acc[4]) == pytest.approx(float(info["clipfrac"]), abs=5e-3)
        assert float(acc[6]) == pytest.approx(loss, rel=tol, abs=tol)
        # whole gradient and every parameter tensor: relative Frobenius error
        gtol = T["grad_emul"] if name == "emulated" else T["grad_fp32"]
        rel = float((got - want).norm() / want.norm())
        assert rel < gtol, f"{name}: relative gradient error {rel:.3e}"
        for z in range(2):
            for l in range(4):
                for kind, off in (("w", layout.w[z][l]), ("b", layout.b[z][l])):
                    nxt = sorted(o for o in [*layout.w[0], *layout.b[0], *layout.w[1], *layout.b[1], layout.logstd, layout.n_params] if o > off)[0]
                    a_, b_ = got[off:nxt], want[off:nxt]
                    r = float((a_ - b_).norm() / (b_.norm() + 1e-12))
                    # bias gradients are column sums of mixed-sign per-sample terms (cancellation, like the log-std
                    # gradient below): 4 x instead of 2.5 x the whole-gradient bound (measured 4.8e-2 on the actor's
                    # first-layer bias at M = 5000, identically on the persistent and the one-tile GEMM paths)
                    bound = (T["b_mult"] if kind == "b" else T["w_mult"]) * gtol
                    assert r < bound, f"{name}: net {z} layer {l} {kind}: relative error {r:.3e}"
        ls = slice(layout.logstd, layout.logstd + ACT)
        # log-std gradient: a sum of mixed-sign per-sample terms (cancellation) -> judged norm-wise
        r = float((got[ls] - want[ls]).norm() / want[ls].norm())
        assert r < T["logstd_mult"] * gtol, f"{name}: log-std gradient relative error {r:.3e}"


@pytest.mark.parametrize("var,prec", [("CATB200_TILE256", "tf32"), ("CATB200_PAIRS", "tf32"), ("CATB200_PAIRS", "bf16"), ("CATB200_WGRAD256", "tf32"), ("CATB200_ZIGZAG", "tf32"), ("CATB200_L2_HINTS", "tf32"), ("CATB200_L2_HINTS", "bf16")])
def test_tile_variants_match_each_other(var, prec):
    """The GEMM tile variants -- CTA pairs (cta_group::2, 256 x 256 tiles; CATB200_PAIRS, opt-in: measured no faster) the single-CTA 256-row tiles (CATB200_TILE256; default: long-K forward launches only), the 256-column weight-gradient tiles
    (CATB200_WGRAD256, opt-in), the alternating row order (CATB200_ZIGZAG, opt-in) and the L2 eviction-priority hints on the TMA
    traffic (CATB200_L2_HINTS) -- against the plain 128 x 128 tiles: same
    gradient on a ragged 16500-row minibatch (the last pair tile has one CTA partly and one wholly beyond M).  The
    switches are read once per process -> subprocesses."""
    import os
    import subprocess
    import sys

    code = (
        "import torch, sys; sys.path.insert(0, %r)\n"
        "from tests import test_mlp_gpu as T\n"
        "from constraints_as_terminations_b200 import ops\n"
        "agent = T.make_agent(seed=1); dims, layout, params, wc = T.device_agent(agent, %r)\n"
        "obs, actions, logp, adv, returns, values, ns, idx = T._minibatch(agent, 20000, 16500, seed=5)\n"
        "g = torch.zeros(layout.n_params, device='cuda:0'); la = torch.zeros(8, device='cuda:0')\n"
        "ops.ppo_minibatch_grad(dims, ops.make_hparams(), idx.cuda(), ops.obs_to_operand(dims, obs.cuda()), actions.cuda(), logp.cuda(), adv.cuda(),"
        " returns.cuda(), values.cuda(), ns.cuda(), params, wc, g, la, ops.mlp_workspace(dims, 16500, True, 'cuda:0'))\n"
        "torch.save(g.cpu(), sys.argv[1])\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), prec)
    outs = []
    for flag in ("0", "1"):
        path = f"/tmp/catb200_{var}_{prec}_{flag}.pt"
        env = dict(os.environ, CATB200_PAIRS="0", CATB200_TILE256="0", CATB200_WGRAD256="0", CATB200_ZIGZAG="0", CATB200_L2_HINTS="0")
        env[var] = flag
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=120)
        outs.append(torch.load(path))
    rel = float((outs[0] - outs[1]).norm() / outs[0].norm())
    assert rel < 1e-5, rel  # same products, same fp32 accumulation per output element; only the atomics order of wgrad differs


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("B,M", [(20000, 16500), (700, 333)])
def test_head_kernels_match_each_other(B, M, prec):
    """head_mma_kernel (mma.sync tf32 fragments, 16 samples per warp, the default) against head_kernel<true> (one warp per
    sample, fp32 FMAs; CATB200_HEAD=warp): same losses and the same gradient on ragged minibatches.  The forward products
    of the mma kernel use hi + lo tf32 weight terms (fp32-grade), its backward products plain tf32 operands: the head
    weight gradients and everything downstream of dZ3 agree to tf32 rounding of dL/dmean (2^-11 relative per term).
    The switch is read once per process -> subprocesses."""
    import os
    import subprocess
    import sys

    code = (
        "import torch, sys; sys.path.insert(0, %r)\n"
        "from tests import test_mlp_gpu as T\n"
        "from constraints_as_terminations_b200 import ops\n"
        "agent = T.make_agent(seed=3); dims, layout, params, wc = T.device_agent(agent, %r)\n"
        "obs, actions, logp, adv, returns, values, ns, idx = T._minibatch(agent, %d, %d, seed=7)\n"
        "g = torch.zeros(layout.n_params, device='cuda:0'); la = torch.zeros(8, device='cuda:0')\n"
        "ops.ppo_minibatch_grad(dims, ops.make_hparams(), idx.cuda(), ops.obs_to_operand(dims, obs.cuda()), actions.cuda(), logp.cuda(), adv.cuda(),"
        " returns.cuda(), values.cuda(), ns.cuda(), params, wc, g, la, ops.mlp_workspace(dims, %d, True, 'cuda:0'))\n"
        "torch.save((g.cpu(), la.cpu()), sys.argv[1])\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), prec, B, M, M)
    outs = []
    for variant in ("warp", "mma"):
        path = f"/tmp/catb200_head_{variant}_{prec}_{M}.pt"
        subprocess.run([sys.executable, "-c", code, path], check=True, env=dict(os.environ, CATB200_HEAD=variant), timeout=120)
        outs.append(torch.load(path))
    (g0, la0), (g1, la1) = outs
    assert torch.allclose(la0, la1, rtol=2e-5, atol=1e-6), (la0, la1)   # losses, KL, clip fraction: fp32-grade forward
    layout = device_agent(make_agent(seed=3), prec)[1]
    for name, (lo, hi) in flat_slices(layout).items():
        a, b = g0[lo:hi], g1[lo:hi]
        rel = float((a - b).norm() / a.norm().clamp_min(1e-20))
        # bf16: dZ3 is stored with an 8-bit mantissa, so a 2^-11 difference in dL/dmean flips roundings of 2^-9
        assert rel < (2e-3 if prec == "tf32" else 1.5e-2), f"{name}: {rel:.3e}"


@pytest.mark.parametrize("prec", PRECS)
def test_adam_step_matches_torch(prec):
    agent = make_agent(seed=2)
    dims, layout, params, w16 = device_agent(agent, prec)
    rnd = (lambda x: x.to(torch.bfloat16)) if prec == "bf16" else tf32_round
    n = layout.n_params
    g = torch.Generator().manual_seed(0)
    ref_p = torch.nn.Parameter(params.cpu().clone())
    opt = torch.optim.Adam([ref_p], lr=3e-4, eps=1e-5)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    lr = torch.tensor(3e-4, device=DEV)
    step = torch.zeros(1, dtype=torch.int32, device=DEV)
    opt_ws = torch.zeros(8, dtype=torch.int64, device=DEV)
    norm_out = torch.zeros(1, device=DEV)
    for it in range(4):
        grad = torch.randn(n, generator=g) * (0.001 if it == 1 else 0.05)  # it==1: norm below the clip threshold
        grads = grad.to(DEV).clone()
        ops.adam_step(dims, params, grads, m, v, w16, lr, step, opt_ws, max_grad_norm=1.0, eps=1e-5, grad_norm_out=norm_out)
        ref_p.grad = grad.clone()
        total = torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
        opt.step()
        assert float(norm_out) == pytest.approx(float(total), rel=1e-5)
        assert float(grads.abs().sum()) == 0.0  # gradient buffer zeroed for the next minibatch
        torch.testing.assert_close(params.cpu(), ref_p.detach(), rtol=1e-5, atol=1e-7)
    assert int(step) == 4
    # operand-precision compute copies refreshed: W and W^T of a hidden layer
    w2 = params[layout.w[1][1] : layout.w[1][1] + 256 * 512].view(256, 512).cpu()
    got_w = w16[layout.wc[1][1] : layout.wc[1][1] + 256 * 512].view(256, 512).cpu()
    got_wt = w16[layout.wtc[1][1] : layout.wtc[1][1] + 256 * 512].view(512, 256).cpu()
    assert torch.equal(got_w, rnd(w2)) and torch.equal(got_wt, rnd(w2.T.contiguous()))
    w1 = params[layout.w[0][0] : layout.w[0][0] + 512 * 45].view(512, 45).cpu()
    got_w1 = w16[layout.wc[0][0] : layout.wc[0][0] + 512 * 64].view(512, 64).cpu()
    assert torch.equal(got_w1[:, :45], rnd(w1.contiguous())) and float(got_w1[:, 45:].abs().sum()) == 0.0
    # grad_scale (1/world after a sum-allreduce) is equivalent to scaling the gradient
    p2, m2, v2 = params.clone(), m.clone(), v.clone()
    step2 = step.clone()
    grad = torch.randn(n, generator=g).to(DEV) * 0.05
    ops.adam_step(dims, params, (grad * 4).clone(), m, v, w16, lr, step, opt_ws, grad_scale=0.25)
    ops.adam_step(dims, p2, grad.clone(), m2, v2, w16, lr, step2, opt_ws, grad_scale=1.0)
    torch.testing.assert_close(params, p2, rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("prec", PRECS)
def test_fused_minibatch_update_matches_grad_then_adam(prec):
    """catb200_ppo_minibatch_update (gradient + ONE fold / norm / clip / Adam / operand-copy launch) against the two
    separate calls it replaces, three optimizer steps: same parameters, moments, operand copies, step counter, gradient
    norm; gradient and scratch left clean.  Only the atomics order of the weight-gradient sums differs between two runs of
    either path; the two states are re-synchronised after every step so that a flipped operand rounding (one bf16 / tf32
    ulp of a weight) cannot compound."""
    agent = make_agent(seed=4)
    obs, actions, logp, adv, returns, values, ns, idx = _minibatch(agent, 9000, 5000, seed=11)
    hp = ops.make_hparams()
    states = []
    for _ in range(2):
        dims, layout, params, wc = device_agent(agent, prec)
        n = layout.n_params
        states.append(dict(
            dims=dims, params=params, wc=wc, grads=torch.zeros(n, device=DEV), m=torch.zeros(n, device=DEV), v=torch.zeros(n, device=DEV),
            la=torch.zeros(8, device=DEV), lr=torch.tensor(3e-4, device=DEV), step=torch.zeros(1, dtype=torch.int32, device=DEV),
            opt_ws=torch.zeros(8, dtype=torch.int64, device=DEV), norm=torch.zeros(1, device=DEV), ws=ops.mlp_workspace(dims, 5000, True, DEV),
        ))  # fmt: skip
    obs_op = ops.obs_to_operand(states[0]["dims"], obs.to(DEV))
    data = [t.to(DEV) for t in (actions, logp, adv, returns, values, ns)]
    A, B = states
    for k in range(3):
        mb = idx.to(DEV).roll(37 * k)
        ops.ppo_minibatch_grad(A["dims"], hp, mb, obs_op, *data, A["params"], A["wc"], A["grads"], A["la"], A["ws"])
        ops.adam_step(A["dims"], A["params"], A["grads"], A["m"], A["v"], A["wc"], A["lr"], A["step"], A["opt_ws"],
                      max_grad_norm=1.0, eps=1e-5, grad_norm_out=A["norm"])
        ops.ppo_minibatch_update(B["dims"], hp, mb, obs_op, *data, B["params"], B["wc"], B["grads"], B["la"], B["ws"], B["m"], B["v"],
                                 B["lr"], B["step"], B["opt_ws"], max_grad_norm=1.0, eps=1e-5, grad_norm_out=B["norm"])
        torch.cuda.synchronize()
        for S in (A, B):
            assert float(S["grads"].abs().sum()) == 0.0
            assert int(S["step"]) == k + 1
            assert int(S["opt_ws"].view(torch.int32)[0]) == 0  # ticket back at zero
            assert float(S["opt_ws"].view(torch.float64)[4]) == 0.0  # sum of squares reset
        assert float(B["norm"]) == pytest.approx(float(A["norm"]), rel=1e-5)
        torch.testing.assert_close(B["la"], A["la"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(B["m"], A["m"], rtol=1e-4, atol=1e-7)
        torch.testing.assert_close(B["v"], A["v"], rtol=1e-4, atol=1e-12)
        # Adam's first steps move every weight by ~lr whatever the gradient's size: parameters are compared to a
        # fraction of one step, operand copies to one unit in the last place of their precision
        assert float((B["params"] - A["params"]).abs().max()) < 0.05 * 3e-4
        assert float((B["wc"].float() - A["wc"].float()).abs().max()) < (2e-2 if prec == "bf16" else 2e-3)
        for key in ("params", "m", "v", "wc"):
            B[key].copy_(A[key])
