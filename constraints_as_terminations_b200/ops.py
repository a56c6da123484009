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
) -> torch.Tensor | None:
    """x: [rows, dim] or [rows] (dim = 1).  Updates (mean, var, count) in place when `update`, then
    returns (x - mean) / sqrt(var + eps) with the updated statistics (written to `out` if given)."""
    _f32c(x, "x")
    dim = 1 if x.ndim == 1 else x.shape[-1]
    rows = x.numel() // dim
    for t, name in ((mean, "mean"), (var, "var"), (count, "count")):
        _f32c(t, name)
    if mean.numel() != dim or var.numel() != dim or count.numel() != 1:
        raise ValueError("running statistics do not match the feature dimension of x")
    if normalize:
        if out is None:
            out = torch.empty_like(x)
        _f32c(out, "out")
        if out.numel() != x.numel():
            raise ValueError("out must have as many elements as x")
    lib = L.load()
    ws_ptr, ws_bytes = None, 0
    if update:
        need = lib.catb200_rms_workspace_bytes(dim)
        ws = (workspace or Workspace(x.device)).get(need)
        ws_ptr, ws_bytes = ws.data_ptr(), ws.numel() * 8
    L.check(
        lib.catb200_rms_forward(
            x.data_ptr(), rows, dim, mean.data_ptr(), var.data_ptr(), count.data_ptr(), eps, int(update),
            out.data_ptr() if normalize else None, ws_ptr, ws_bytes, L.stream(),
        ),
        "rms_forward",
    )  # fmt: skip
    return out if normalize else None


# --------------------------------------------------------------------------------------------------
# rollout append  (reference cleanrl/ppo.py:203-205,215-216)
# --------------------------------------------------------------------------------------------------
def rollout_append(reward, done, time_out, rewards_t, dones_t1, true_dones_t1) -> None:
    """rewards[t] = reward; dones[t+1] = done; true_dones[t+1] = float(time_out)."""
    n = reward.numel()
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
