"""Multi-GPU plumbing: one process per GPU, envs sharded by rank, one gradient exchange per optimizer step.

The reference's CleanRL path is single-process; its rl_games / skrl front-ends shard envs
across processes and all-reduce the policy gradient (`U/skrl/ppo.py:126-131,534-537`).  This module is the
equivalent for the B200-native trainer: `torch.distributed` (NCCL over NVLink on GPUs, gloo in the CPU tests)
carries exactly two kinds of traffic -- the initial weight broadcast and a sum-allreduce of the flat fp32
gradient bucket (377 241 floats = 1.509 MB) per optimizer step; the 1/world factor is folded into the
clip + Adam kernels (`catb200_adam_step(grad_scale=...)`), so every rank clips with the same global norm.
"""

from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise the default process group from torchrun's env vars. Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kwargs)
    return rank, world, local_rank


def bind_to_gpu_numa_node(device_index: int) -> list[int]:
    """Pin the calling thread to the CPUs of the NUMA node the GPU hangs off, so that pinned host buffers allocated and
    first touched afterwards are node-local (8 ranks feeding 118 MB per iteration each through the wrong socket's memory
    controller is what limited the end-to-end scaling).  NVML's own affinity helper when available, sysfs otherwise; returns
    the CPU list (empty: nothing was changed -- no NVML, no sysfs, or a container without the right to set affinity)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))
    except Exception:
        pass
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        node = int(open(f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node").read())
        if node < 0:
            return []
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return []


def world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def broadcast_params(flat: torch.Tensor, src: int = 0) -> None:
    """Rank `src` seeds every rank's flat parameter vector (done once, before the first rollout)."""
    if world_size() > 1:
        dist.broadcast(flat, src=src)


def allreduce_grads(flat_grads: torch.Tensor) -> float:
    """Sum-allreduce the flat gradient bucket in place; returns the scale (1/world) the optimizer applies."""
    w = world_size()
    if w > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return 1.0 / w


def pooled_moments(mean: torch.Tensor, var: torch.Tensor, count: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Statistics of the union of every rank's samples from the per-rank running (mean, population variance, count):
    the parallel-moments merge `RunningMeanStd.update_from_moments` itself uses (reference U/cleanrl/ppo.py:41-62), applied
    across ranks in rank order, in float64.  Collective (all ranks must call it); the inputs are not modified.  Used when
    a checkpoint is written: training keeps the normalisers rank-local, as the reference's multi-process front-ends do,
    but the saved policy should carry the statistics of all env shards, not rank 0's."""
    if world_size() == 1:
        return mean.clone(), var.clone(), count.clone()
    packed = torch.cat([mean.reshape(-1), var.reshape(-1), count.reshape(-1)]).to(torch.float64)
    gathered = [torch.empty_like(packed) for _ in range(world_size())]
    dist.all_gather(gathered, packed)
    d = mean.numel()
    m, v, n = gathered[0][:d].clone(), gathered[0][d : 2 * d].clone(), gathered[0][2 * d].clone()
    for g in gathered[1:]:
        mb, vb, nb = g[:d], g[d : 2 * d], g[2 * d]
        delta, tot = mb - m, n + nb
        m2 = v * n + vb * nb + delta * delta * n * nb / tot
        m, v, n = m + delta * nb / tot, m2 / tot, tot
    return m.to(mean.dtype).view_as(mean), v.to(var.dtype).view_as(var), n.to(count.dtype).view_as(count)


def shard_seed(base_seed: int) -> int:
    """Per-rank seed, like the reference's distributed front-ends (`scripts/skrl/train.py:116-117`)."""
    return base_seed + rank()


class _RawCuda:
    """Zero-copy torch view of raw device memory (the __cuda_array_interface__ protocol)."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class PeerGradExchange:
    """One-shot gradient all-reduce over NVLink peer memory, fused with the gradient norm (csrc/peer.cu).

    Every rank cudaMallocs one peer-visible block ([flags][arena 0][arena 1][summed gradient of the reduce-scatter path]); the 64-byte IPC handles travel once
    through `torch.distributed.all_gather_object` (host plumbing) and every rank maps its peers' blocks.  After that the
    exchange is ONE kernel per optimizer step and rank (`reduce`), with no library collective and no host
    synchronisation, so a whole epoch of minibatches replays as a single CUDA graph on every rank."""

    def __init__(self, n_params: int, device: torch.device):
        from . import _lib as L

        self.L, self.lib = L, L.load()
        self.rank, self.world = rank(), world_size()
        if self.world > 8:
            raise RuntimeError("PeerGradExchange supports up to 8 ranks (one NVSwitch domain)")
        self.n = int(n_params)
        self.n_pad = (self.n + 63) // 64 * 64
        self.device = device
        nbytes = self.lib.catb200_peer_arena_bytes(self.n)
        own = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        L.check(self.lib.catb200_peer_alloc(nbytes, C.byref(own), handle), "peer_alloc")
        self._own = own.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle))
        self._bases = (C.c_void_p * self.world)()
        self._opened = []
        for p, h in enumerate(handles):
            if p == self.rank:
                self._bases[p] = self._own
                continue
            ptr = C.c_void_p()
            buf = (C.c_uint8 * 64).from_buffer_copy(h)
            L.check(self.lib.catb200_peer_open(buf, C.byref(ptr)), f"peer_open(rank {p})")
            self._bases[p] = ptr.value
            self._opened.append(ptr.value)
        arenas = self._own + 64 * 4
        self.arena = [torch.as_tensor(_RawCuda(arenas + i * self.n_pad * 4, self.n), device=device) for i in range(2)]
        self.grad_sum = torch.zeros(self.n, dtype=torch.float32, device=device)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        dist.barrier()  # nobody touches a peer block before everybody has mapped everything

    def reduce(self, parity: int, step_dev, opt_ws, max_grad_norm=1.0, betas=(0.9, 0.999), grad_norm_out=None) -> torch.Tensor:
        """Sum of every rank's arena `parity` -> self.grad_sum (+ clip coefficient / bias corrections in opt_ws)."""
        L = self.L
        L.check(
            self.lib.catb200_grad_allreduce_norm(
                self._bases, self.rank, self.world, self.n, int(parity), self.grad_sum.data_ptr(), 1.0 / self.world,
                max_grad_norm, betas[0], betas[1], step_dev.data_ptr(), L.ptr(grad_norm_out), opt_ws.data_ptr(),
                self.epoch.data_ptr(), self.err.data_ptr(), L.stream(),
            ),
            "grad_allreduce_norm",
        )  # fmt: skip
        return self.grad_sum

    def minibatch_update(self, parity: int, dims, hp, mb_inds, obs_op_all, actions_all, logprobs_all, advantages_all, returns_all,
                         values_all, norm_stats, params, wc, loss_acc, ws, exp_avg, exp_avg_sq, lr_dev, step_dev, opt_ws,
                         max_grad_norm=1.0, betas=(0.9, 0.999), eps=1e-5, grad_norm_out=None) -> None:  # fmt: skip
        """Forward + backward of one minibatch into arena `parity`, then fold + exchange + norm + clip + Adam + operand
        refresh as ONE launch (catb200_ppo_minibatch_update_peer): `reduce` and the two launches around it in one."""
        L = self.L
        L.check(
            self.lib.catb200_ppo_minibatch_update_peer(
                dims, hp, mb_inds.numel(), mb_inds.data_ptr(), obs_op_all.data_ptr(), actions_all.data_ptr(),
                logprobs_all.data_ptr(), advantages_all.data_ptr(), returns_all.data_ptr(), values_all.data_ptr(),
                norm_stats.data_ptr(), params.data_ptr(), wc.data_ptr(), loss_acc.data_ptr(), ws.data_ptr(), ws.numel() * 8,
                exp_avg.data_ptr(), exp_avg_sq.data_ptr(), lr_dev.data_ptr(), step_dev.data_ptr(), max_grad_norm, betas[0],
                betas[1], eps, L.ptr(grad_norm_out), opt_ws.data_ptr(), self._bases, self.rank, self.world, int(parity),
                self.grad_sum.data_ptr(), self.epoch.data_ptr(), self.err.data_ptr(), L.stream(),
            ),
            "ppo_minibatch_update_peer",
        )  # fmt: skip

    def check(self) -> None:
        """Raise if a handshake timed out or the arena parity went out of step (one device->host read)."""
        code = int(self.err.item())
        if code:
            raise RuntimeError(f"peer gradient exchange failed on rank {self.rank}: " + {1: "a peer did not arrive within 2 s", 2: "arena parity out of step"}.get(code, str(code)))

    def close(self) -> None:
        if getattr(self, "_own", None):
            torch.cuda.synchronize(self.device)
            if dist.is_initialized():
                dist.barrier()
            for ptr in self._opened:
                self.lib.catb200_peer_close(ptr)
            self.arena = None
            self.lib.catb200_peer_free(self._own)
            self._own = None
