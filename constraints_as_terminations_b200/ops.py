"""Thin functional wrappers over the C ABI (`include/catb200.h`) for the trainer-side kernels.

Each function checks devices / dtypes / contiguity, hands raw pointers to libcatb200 on the current
CUDA stream and returns torch tensors.  No arithmetic happens in python and there is no fallback:
non-CUDA inputs raise.
"""

from __future__ import annotations

import torch

from . import _lib as L


def _f32c(t: torch.Tensor, what: str) -> torch.Tensor:
    L.require_cuda(t, what)
    if t.dtype != torch.float32:
        raise TypeError(f"{what} must be float32, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{what} must be contiguous")
    return t


class Workspace:
    """Zero-initialised device scratch, grown on demand, that kernels leave clean between calls."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.buf = None

    def get(self, nbytes: int) -> torch.Tensor:
        if self.buf is None or self.buf.numel() * 8 < nbytes:
            self.buf = L.zeros_workspace(nbytes, self.device)
        return self.buf


# --------------------------------------------------------------------------------------------------
# RunningMeanStd.forward  (reference cleanrl/ppo.py:21-62)
# --------------------------------------------------------------------------------------------------
def rms_forward(
    x: torch.Tensor,
    mean: torch.Tensor,
    var: torch.Tensor,
    count: torch.Tensor,
    eps: float = 1e-8,
    update: bool = True,
    out: torch.Tensor | None = None,
    normalize: bool = True,
    workspace: Workspace | None = None,
    out_op: torch.Tensor | None = None,
    validate: bool = True,
) -> torch.Tensor | None:
    """x: [rows, dim] or [rows] (dim = 1).  Updates (mean, var, count) in place when `update`, then
    returns (x - mean) / sqrt(var + eps) with the updated statistics (written to `out` if given).
    `out_op` (bf16 or fp32 [rows, pad]) additionally receives the zero-padded tensor-core operand copy the MLP
    reads (fp32 = rounded to tf32)."""
    dim = 1 if x.ndim == 1 else x.shape[-1]
    rows = x.numel() // dim
    if validate:
        _f32c(x, "x")
        for t, name in ((mean, "mean"), (var, "var"), (count, "count")):
            _f32c(t, name)
        if mean.numel() != dim or var.numel() != dim or count.numel() != 1:
            raise ValueError("running statistics do not match the feature dimension of x")
    if normalize:
        if out is None:
            out = torch.empty_like(x)
        if validate:
            _f32c(out, "out")
            if out.numel() != x.numel():
                raise ValueError("out must have as many elements as x")
            if out_op is not None and (out_op.dtype not in (torch.bfloat16, torch.float32) or not out_op.is_contiguous() or out_op.numel() % rows):
                raise ValueError("out_op must be a contiguous bf16 / fp32 tensor [rows, pad]")
    lib = L.load()
    ws_ptr, ws_bytes = None, 0
    if update:
        need = lib.catb200_rms_workspace_bytes(dim)
        ws = (workspace or Workspace(x.device)).get(need)
        ws_ptr, ws_bytes = ws.data_ptr(), ws.numel() * 8
    L.check(
        lib.catb200_rms_forward(
            x.data_ptr(), rows, dim, mean.data_ptr(), var.data_ptr(), count.data_ptr(), eps, int(update),
            out.data_ptr() if normalize else None, L.ptr(out_op), (out_op.numel() // rows) if out_op is not None else 0,
            L.PREC_TF32 if out_op is not None and out_op.dtype == torch.float32 else L.PREC_BF16, ws_ptr, ws_bytes, L.stream(),
        ),
        "rms_forward",
    )  # fmt: skip
    return out if normalize else None


# --------------------------------------------------------------------------------------------------
# rollout append  (reference cleanrl/ppo.py:203-205,215-216)
# --------------------------------------------------------------------------------------------------
def rollout_append(reward, done, time_out, rewards_t, dones_t1, true_dones_t1, validate: bool = True) -> None:
    """rewards[t] = reward; dones[t+1] = done; true_dones[t+1] = float(time_out)."""
    n = reward.numel()
    if time_out.dtype == torch.bool:
        time_out = time_out.view(torch.uint8)
    if not validate:
        L.check(
            L.load().catb200_rollout_append(
                reward.data_ptr(), done.data_ptr(), time_out.data_ptr(), n, rewards_t.data_ptr(), dones_t1.data_ptr(),
                true_dones_t1.data_ptr(), L.stream(),
            ),
            "rollout_append",
        )  # fmt: skip
        return
    for t, name in ((reward, "reward"), (done, "done"), (rewards_t, "rewards[t]"), (dones_t1, "dones[t+1]"), (true_dones_t1, "true_dones[t+1]")):  # fmt: skip
        _f32c(t, name)
        if t.numel() != n:
            raise ValueError(f"{name} has {t.numel()} elements, expected {n}")
    L.require_cuda(time_out, "time_out")
    if time_out.dtype == torch.bool:
        time_out = time_out.view(torch.uint8)
    if time_out.dtype != torch.uint8 or not time_out.is_contiguous() or time_out.numel() != n:
        raise TypeError("time_out must be a contiguous bool / uint8 tensor with one entry per env")
    L.check(
        L.load().catb200_rollout_append(
            reward.data_ptr(), done.data_ptr(), time_out.data_ptr(), n, rewards_t.data_ptr(), dones_t1.data_ptr(),
            true_dones_t1.data_ptr(), L.stream(),
        ),
        "rollout_append",
    )  # fmt: skip


# --------------------------------------------------------------------------------------------------
# GAE (+ value normalisation statistics)  (reference cleanrl/ppo.py:251-277,287-288)
# --------------------------------------------------------------------------------------------------
def gae(
    rewards: torch.Tensor,
    values: torch.Tensor,
    dones: torch.Tensor,
    true_dones: torch.Tensor,
    next_value: torch.Tensor,
    gamma: float,
    gae_lambda: float,
    advantages: torch.Tensor | None = None,
    returns: torch.Tensor | None = None,
    value_rms: torch.Tensor | None = None,
    norm_stats: torch.Tensor | None = None,
    workspace: Workspace | None = None,
):
    """rewards, values: [T, N]; dones, true_dones: [T+1, N] (slot T = next_done / next_true_done);
    next_value: [N].  Returns (advantages, returns).  With `value_rms` (3 floats mean,var,count) the two
    RunningMeanStd updates of ppo.py:287-288 are fused in and `norm_stats` (4 floats) receives the
    statistics values / returns are to be normalised with."""
    T, N = rewards.shape
    _f32c(rewards, "rewards"), _f32c(values, "values"), _f32c(dones, "dones"), _f32c(true_dones, "true_dones")
    _f32c(next_value, "next_value")
    if values.shape != (T, N) or dones.shape != (T + 1, N) or true_dones.shape != (T + 1, N) or next_value.numel() != N:
        raise ValueError("gae: shapes must be rewards/values [T,N], dones/true_dones [T+1,N], next_value [N]")
    advantages = torch.empty_like(rewards) if advantages is None else _f32c(advantages, "advantages")
    returns = torch.empty_like(rewards) if returns is None else _f32c(returns, "returns")
    lib = L.load()
    ws_ptr, ws_bytes = None, 0
    if value_rms is not None:
        _f32c(value_rms, "value_rms")
        if norm_stats is None:
            norm_stats = torch.empty(4, dtype=torch.float32, device=rewards.device)
        _f32c(norm_stats, "norm_stats")
        ws = (workspace or Workspace(rewards.device)).get(lib.catb200_gae_workspace_bytes())
        ws_ptr, ws_bytes = ws.data_ptr(), ws.numel() * 8
    L.check(
        lib.catb200_gae(
            rewards.data_ptr(), values.data_ptr(), dones.data_ptr(), true_dones.data_ptr(), next_value.data_ptr(),
            T, N, gamma, gamma * gae_lambda, advantages.data_ptr(), returns.data_ptr(), L.ptr(value_rms),
            L.ptr(norm_stats), ws_ptr, ws_bytes, L.stream(),
        ),
        "gae",
    )  # fmt: skip
    return advantages, returns


# --------------------------------------------------------------------------------------------------
# actor-critic MLP, PPO minibatch gradient, Adam  (reference cleanrl/ppo.py:71-123,294-354)
# --------------------------------------------------------------------------------------------------
def make_dims(obs_dim: int, act_dim: int, hidden=(512, 256, 128), precision: str | None = None) -> L.MlpDims:
    """`precision`: "tf32" (default; env CATB200_GEMM_PREC) or "bf16" -- operand type of the hidden-layer GEMMs."""
    d = L.MlpDims()
    d.obs_dim, d.act_dim = int(obs_dim), int(act_dim)
    d.h1, d.h2, d.h3 = (int(h) for h in hidden)
    d.obs_pad = (int(obs_dim) + 63) // 64 * 64
    d.prec = L.PREC_NAMES[(precision or L.default_precision()).lower()]
    return d


def weight_copies(dims, layout, device) -> torch.Tensor:
    """Zeroed buffer for the operand-precision compute copies of the hidden-layer weights."""
    return torch.zeros(layout.n_wc, dtype=L.operand_dtype(dims.prec), device=device)


def mlp_layout(dims: L.MlpDims) -> L.MlpLayout:
    layout = L.MlpLayout()
    L.check(L.load().catb200_mlp_layout(dims, layout), "mlp_layout")
    return layout


def _check_operand(dims, t: torch.Tensor, what: str) -> None:
    L.require_cuda(t, what)
    if t.dtype != L.operand_dtype(dims.prec) or not t.is_contiguous():
        raise TypeError(f"{what} must be a contiguous {L.operand_dtype(dims.prec)} tensor for this precision, got {t.dtype}")


def cast_weights(dims, params: torch.Tensor, wc: torch.Tensor) -> None:
    _f32c(params, "params")
    _check_operand(dims, wc, "wc")
    L.check(L.load().catb200_mlp_cast_weights(dims, params.data_ptr(), wc.data_ptr(), L.stream()), "mlp_cast_weights")


def obs_to_operand(dims, obs: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """fp32 [..., obs_dim] -> operand [..., obs_pad] (bf16, or fp32 rounded to tf32), zero padded."""
    _f32c(obs, "obs")
    if obs.shape[-1] != dims.obs_dim:
        raise ValueError(f"obs has {obs.shape[-1]} features, expected {dims.obs_dim}")
    rows = obs.numel() // dims.obs_dim
    if out is None:
        out = torch.empty((*obs.shape[:-1], dims.obs_pad), dtype=L.operand_dtype(dims.prec), device=obs.device)
    _check_operand(dims, out, "obs_op")
    if out.numel() != rows * dims.obs_pad:
        raise ValueError("obs_op output must hold [rows, obs_pad] elements")
    L.check(L.load().catb200_obs_to_operand(dims, obs.data_ptr(), rows, out.data_ptr(), L.stream()), "obs_to_operand")
    return out


def mlp_workspace(dims, rows: int, training: bool, device) -> torch.Tensor:
    need = L.load().catb200_mlp_workspace_bytes(dims, rows, int(training))
    if need == 0:
        raise RuntimeError("unsupported MLP dimensions for the fused kernels (h1, h2 multiples of 128, h3 == 128)")
    return L.zeros_workspace(need, device)


def mlp_act(dims, obs_op, params, wc, ws, noise=None, action_in=None, action=None, logprob=None, value=None, mean_out=None,
            rng_state=None, validate: bool = True):  # fmt: skip
    """`noise` [rows, act_dim] supplies Normal.sample()'s eps; `rng_state` (2 x int64 on the device: seed, offset) lets the
    head kernel draw it from Philox instead (and advances the offset)."""
    rows = obs_op.numel() // dims.obs_pad
    if validate:
        _check_operand(dims, obs_op, "obs_op")
        _check_operand(dims, wc, "wc")
        for t, name in ((noise, "noise"), (action_in, "action_in"), (action, "action"), (logprob, "logprob"), (value, "value"), (mean_out, "mean_out")):  # fmt: skip
            if t is not None:
                _f32c(t, name)
        if rng_state is not None and (rng_state.dtype != torch.int64 or rng_state.numel() != 2 or not rng_state.is_cuda):
            raise TypeError("rng_state must be a CUDA int64 tensor {seed, offset}")
    L.check(
        L.load().catb200_mlp_act(
            dims, obs_op.data_ptr(), rows, params.data_ptr(), wc.data_ptr(), L.ptr(noise), L.ptr(rng_state), L.ptr(action_in),
            L.ptr(action), L.ptr(logprob), L.ptr(value), L.ptr(mean_out), ws.data_ptr(), ws.numel() * 8, L.stream(),
        ),
        "mlp_act",
    )  # fmt: skip


def make_hparams(clip_coef=0.2, ent_coef=0.001, vf_coef=2.0, norm_adv=True, clip_vloss=True) -> L.PpoHparams:
    hp = L.PpoHparams()
    hp.clip_coef, hp.ent_coef, hp.vf_coef = clip_coef, ent_coef, vf_coef
    hp.norm_adv, hp.clip_vloss = int(norm_adv), int(clip_vloss)
    return hp


def ppo_minibatch_grad(dims, hp, mb_inds, obs_op_all, actions_all, logprobs_all, advantages_all, returns_all, values_all,
                       norm_stats, params, wc, grads, loss_acc, ws) -> None:  # fmt: skip
    L.require_cuda(mb_inds, "mb_inds")
    _check_operand(dims, obs_op_all, "obs_op_all")
    _check_operand(dims, wc, "wc")
    if mb_inds.dtype != torch.int64 or not mb_inds.is_contiguous():
        raise TypeError("mb_inds must be a contiguous int64 tensor")
    for t, name in ((actions_all, "actions"), (logprobs_all, "logprobs"), (advantages_all, "advantages"), (returns_all, "returns"),
                    (values_all, "values"), (norm_stats, "norm_stats"), (params, "params"), (grads, "grads"), (loss_acc, "loss_acc")):  # fmt: skip
        _f32c(t, name)
    L.check(
        L.load().catb200_ppo_minibatch_grad(
            dims, hp, mb_inds.numel(), mb_inds.data_ptr(), obs_op_all.data_ptr(), actions_all.data_ptr(),
            logprobs_all.data_ptr(), advantages_all.data_ptr(), returns_all.data_ptr(), values_all.data_ptr(),
            norm_stats.data_ptr(), params.data_ptr(), wc.data_ptr(), grads.data_ptr(), loss_acc.data_ptr(),
            ws.data_ptr(), ws.numel() * 8, L.stream(),
        ),
        "ppo_minibatch_grad",
    )  # fmt: skip


def ppo_minibatch_update(dims, hp, mb_inds, obs_op_all, actions_all, logprobs_all, advantages_all, returns_all, values_all,
                         norm_stats, params, wc, grads, loss_acc, ws, exp_avg, exp_avg_sq, lr_dev, step_dev, opt_ws,
                         max_grad_norm=1.0, betas=(0.9, 0.999), eps=1e-5, grad_scale=1.0, grad_norm_out=None) -> None:  # fmt: skip
    """`ppo_minibatch_grad` + `adam_step` of one minibatch (single GPU) with the fold / norm / clip / Adam tail as one launch."""
    L.require_cuda(mb_inds, "mb_inds")
    _check_operand(dims, obs_op_all, "obs_op_all")
    _check_operand(dims, wc, "wc")
    if mb_inds.dtype != torch.int64 or not mb_inds.is_contiguous():
        raise TypeError("mb_inds must be a contiguous int64 tensor")
    for t, name in ((actions_all, "actions"), (logprobs_all, "logprobs"), (advantages_all, "advantages"), (returns_all, "returns"),
                    (values_all, "values"), (norm_stats, "norm_stats"), (params, "params"), (grads, "grads"), (loss_acc, "loss_acc"),
                    (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq"), (lr_dev, "lr")):  # fmt: skip
        _f32c(t, name)
    L.check(
        L.load().catb200_ppo_minibatch_update(
            dims, hp, mb_inds.numel(), mb_inds.data_ptr(), obs_op_all.data_ptr(), actions_all.data_ptr(),
            logprobs_all.data_ptr(), advantages_all.data_ptr(), returns_all.data_ptr(), values_all.data_ptr(),
            norm_stats.data_ptr(), params.data_ptr(), wc.data_ptr(), grads.data_ptr(), loss_acc.data_ptr(),
            ws.data_ptr(), ws.numel() * 8, exp_avg.data_ptr(), exp_avg_sq.data_ptr(), lr_dev.data_ptr(), step_dev.data_ptr(),
            max_grad_norm, betas[0], betas[1], eps, grad_scale, L.ptr(grad_norm_out), opt_ws.data_ptr(), L.stream(),
        ),
        "ppo_minibatch_update",
    )  # fmt: skip


def adam_step(dims, params, grads, exp_avg, exp_avg_sq, wc, lr_dev, step_dev, opt_ws, max_grad_norm=1.0,
              betas=(0.9, 0.999), eps=1e-5, grad_scale=1.0, grad_norm_out=None) -> None:  # fmt: skip
    for t, name in ((params, "params"), (grads, "grads"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq"), (lr_dev, "lr")):
        _f32c(t, name)
    L.check(
        L.load().catb200_adam_step(
            dims, params.data_ptr(), grads.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), wc.data_ptr(),
            lr_dev.data_ptr(), step_dev.data_ptr(), max_grad_norm, betas[0], betas[1], eps, grad_scale,
            L.ptr(grad_norm_out), opt_ws.data_ptr(), L.stream(),
        ),
        "adam_step",
    )  # fmt: skip


def adam_apply(dims, params, grads, exp_avg, exp_avg_sq, wc, lr_dev, opt_ws, betas=(0.9, 0.999), eps=1e-5, grad_scale=1.0) -> None:
    """The Adam + operand-copy half of `adam_step`; the clip coefficient / bias corrections are already in `opt_ws`
    (written by the peer-memory all-reduce, dist.PeerGradExchange.reduce)."""
    for t, name in ((params, "params"), (grads, "grads"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq"), (lr_dev, "lr")):
        _f32c(t, name)
    L.check(
        L.load().catb200_adam_apply(
            dims, params.data_ptr(), grads.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), wc.data_ptr(),
            lr_dev.data_ptr(), betas[0], betas[1], eps, grad_scale, opt_ws.data_ptr(), L.stream(),
        ),
        "adam_apply",
    )  # fmt: skip


# --------------------------------------------------------------------------------------------------
# single-`dones` GAE of the rl_games / skrl front-ends  (reference rl_games/cat_common.py:96-104,
# skrl/ppo.py:397-442)
# --------------------------------------------------------------------------------------------------
def gae_float_dones(
    variant: int,
    rewards: torch.Tensor,
    values: torch.Tensor,
    dones: torch.Tensor,
    last_values: torch.Tensor,
    gamma: float,
    coef: float,
    last_dones: torch.Tensor | None = None,
    normalize: bool = False,
    advantages: torch.Tensor | None = None,
    returns: torch.Tensor | None = None,
    workspace: Workspace | None = None,
):
    """rewards / values / dones: [T, N] (trailing singleton dims allowed), last_values / last_dones: [N].
    `variant` is L.GAE_RLGAMES (coef = gamma * tau, dones observed before each step, `last_dones` after the
    last one) or L.GAE_SKRL (coef = lambda, dones = `terminated` of each step, optional global advantage
    normalisation).  Returns (advantages, returns) shaped like `rewards`."""
    shape = rewards.shape
    T = shape[0]
    N = rewards.numel() // max(T, 1)
    for t, name in ((rewards, "rewards"), (values, "values"), (dones, "dones"), (last_values, "last_values")):
        _f32c(t, name)
    if values.numel() != T * N or dones.numel() != T * N or last_values.numel() != N:
        raise ValueError("gae_float_dones: rewards / values / dones must hold T*N entries and last_values N")
    if variant == L.GAE_RLGAMES:
        if last_dones is None:
            raise ValueError("gae_float_dones: the rl_games variant needs last_dones")
        _f32c(last_dones, "last_dones")
        if last_dones.numel() != N:
            raise ValueError("gae_float_dones: last_dones must hold N entries")
    advantages = torch.empty_like(rewards) if advantages is None else _f32c(advantages, "advantages")
    returns = torch.empty_like(rewards) if returns is None else _f32c(returns, "returns")
    lib = L.load()
    ws_ptr, ws_bytes = None, 0
    if normalize:
        ws = (workspace or Workspace(rewards.device)).get(lib.catb200_gae_float_dones_workspace_bytes())
        ws_ptr, ws_bytes = ws.data_ptr(), ws.numel() * 8
    L.check(
        lib.catb200_gae_float_dones(
            variant, rewards.data_ptr(), values.data_ptr(), dones.data_ptr(), L.ptr(last_dones), last_values.data_ptr(),
            T, N, gamma, coef, advantages.data_ptr(), returns.data_ptr(), 1 if normalize else 0, ws_ptr, ws_bytes, L.stream(),
        ),
        "gae_float_dones",
    )  # fmt: skip
    return advantages, returns


# --------------------------------------------------------------------------------------------------
# device-side random draws (Philox4x32-10)
# --------------------------------------------------------------------------------------------------
def make_rng_state(seed: int, device, offset: int = 0) -> torch.Tensor:
    """{seed, offset} as two int64 on the device (the kernels read them as uint64 and advance the offset)."""
    return torch.tensor([int(seed) & 0x7FFFFFFFFFFFFFFF, int(offset)], dtype=torch.int64, device=device)


def random_permutation(n: int, rng_state: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Pseudo-random permutation of 0..n-1 (replaces torch.randperm, reference cleanrl/ppo.py:295): keyed Feistel
    bijection + cycle walking, one launch, no sort."""
    L.require_cuda(rng_state, "rng_state")
    if out is None:
        out = torch.empty(n, dtype=torch.int64, device=rng_state.device)
    if out.dtype != torch.int64 or not out.is_contiguous() or out.numel() != n:
        raise TypeError("out must be a contiguous int64 tensor with n entries")
    L.check(L.load().catb200_random_permutation(n, rng_state.data_ptr(), out.data_ptr(), L.stream()), "random_permutation")
    return out


def bernoulli_mask(p: torch.Tensor, rng_state: torch.Tensor, with_ids: bool = False):
    """mask = (u < p) with Philox uniforms; with_ids also returns (ids buffer [n], count [1]) where ids[:count] are the
    ascending positions of the set entries."""
    _f32c(p, "p")
    n = p.numel()
    mask = torch.empty(n, dtype=torch.uint8, device=p.device)
    ids = count = None
    if with_ids:
        ids = torch.empty(n, dtype=torch.int64, device=p.device)
        count = torch.zeros(1, dtype=torch.int32, device=p.device)
    L.check(
        L.load().catb200_bernoulli_mask(p.data_ptr(), n, rng_state.data_ptr(), mask.data_ptr(), L.ptr(ids), L.ptr(count), L.stream()),
        "bernoulli_mask",
    )
    return (mask.view(torch.bool), ids, count) if with_ids else mask.view(torch.bool)
