"""CleanRL-style PPO trainer of the reference, re-built on the libcatb200 kernels.

Public names and behaviour follow `exts/cat_envs/cat_envs/tasks/utils/cleanrl/ppo.py`:

    RunningMeanStd(shape, epsilon)            :12-45   buffers running_mean / running_var / count
    Agent(envs)                               :71-123  critic / actor_mean / actor_logstd / obs_rms / value_rms,
                                                       get_value, get_action_and_value, forward; identical
                                                       state_dict keys (checkpoints interchange with play.py)
    PPO(envs, ppo_cfg, run_path)              :126-372 the training loop (float `dones`, time-out aware GAE)

Execution model (one process per GPU):
  rollout step  : pre  = 3 batched tcgen05 GEMM launches -> head kernel (Philox sample, log-prob, value)
                  env.step(action)   (Isaac Lab + the fused ConstraintManager kernels)
                  post = append -> running moments -> normalise (+ operand copy for the first GEMM)
  update        : 1 GAE(+value statistics) launch, then per epoch a Feistel/Philox permutation and per minibatch
                  `catb200_ppo_minibatch_grad` (gather, 3 fwd GEMMs, head/loss, 3 wgrad + 2 dgrad GEMMs, fold)
                  [-> NCCL allreduce of the flat 1.5 MB gradient] -> `catb200_adam_step` (norm, clip, Adam,
                  operand-precision weight refresh)
Hidden-layer GEMM operands are tf32 by default (fp32 storage; the reference's GPU numerics,
scripts/clean_rl/train.py:86-87) or bf16 (`precision="bf16"` / CATB200_GEMM_PREC=bf16).
No tensor math runs in python; there is no eager fallback (CPU tensors raise).
"""

from __future__ import annotations

import os
import time

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import dist as cdist
from . import ops


class RunningMeanStd(nn.Module):
    """Running mean / variance with the reference's semantics (count starts at 1, Chan merge, eps 1e-8)."""

    def __init__(self, shape=(), epsilon=1e-08):
        super().__init__()
        self.register_buffer("running_mean", torch.zeros(shape))
        self.register_buffer("running_var", torch.ones(shape))
        self.register_buffer("count", torch.ones(()))
        self.epsilon = epsilon
        self._ws = None

    def _workspace(self):
        if self._ws is None or self._ws.device != self.running_mean.device:
            self._ws = ops.Workspace(self.running_mean.device)
        return self._ws

    def forward(self, obs, update=True, out=None, out_op=None, validate=True):
        obs = obs if obs.is_contiguous() else obs.contiguous()
        return ops.rms_forward(
            obs, self.running_mean, self.running_var, self.count, self.epsilon, update=update, out=out,
            workspace=self._workspace(), out_op=out_op, validate=validate,
        )  # fmt: skip

    def update(self, x):
        x = x if x.is_contiguous() else x.contiguous()
        ops.rms_forward(
            x, self.running_mean, self.running_var, self.count, self.epsilon, update=True, normalize=False,
            workspace=self._workspace(),
        )  # fmt: skip


def layer_init(layer, std=np.sqrt(2), bias_const=0.0):
    torch.nn.init.orthogonal_(layer.weight, std)
    torch.nn.init.constant_(layer.bias, bias_const)
    return layer


class Agent(nn.Module):
    """Actor-critic with the reference's module tree; parameters are views into one flat fp32 vector."""

    HIDDEN = (512, 256, 128)

    def __init__(self, envs, device=None, precision=None):
        super().__init__()
        obs_shape = envs.unwrapped.single_observation_space["policy"].shape
        act_shape = envs.unwrapped.single_action_space.shape
        self.obs_dim = int(np.array(obs_shape).prod())
        self.act_dim = int(np.prod(act_shape))
        h1, h2, h3 = self.HIDDEN

        def mlp(out_dim, last_std):
            return nn.Sequential(
                layer_init(nn.Linear(self.obs_dim, h1)), nn.ELU(),
                layer_init(nn.Linear(h1, h2)), nn.ELU(),
                layer_init(nn.Linear(h2, h3)), nn.ELU(),
                layer_init(nn.Linear(h3, out_dim), std=last_std),
            )  # fmt: skip

        self.critic = mlp(1, 1.0)
        self.actor_mean = mlp(self.act_dim, 0.01)
        self.actor_logstd = nn.Parameter(torch.zeros(1, self.act_dim))
        self.obs_rms = RunningMeanStd(shape=obs_shape)
        self.value_rms = RunningMeanStd(shape=())

        self.dims = ops.make_dims(self.obs_dim, self.act_dim, self.HIDDEN, precision=precision)
        self.precision = "tf32" if self.dims.prec == L.PREC_TF32 else "bf16"
        self.layout = ops.mlp_layout(self.dims)
        self._flat = None
        self._wc = None  # operand-precision compute copies of the hidden-layer weights (W and W^T)
        self._act_ws = None
        self._ws_need: dict = {}
        self._noise = None
        self._bind_flat(torch.device("cpu"))
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.sync_weights())
        if device is not None:
            self.to(device)

    # -- flat parameter storage ----------------------------------------------------------------------
    def _param_slots(self):
        lay = self.layout
        for z, net in enumerate((self.critic, self.actor_mean)):
            for l, idx in enumerate((0, 2, 4, 6)):
                yield net[idx].weight, lay.w[z][l]
                yield net[idx].bias, lay.b[z][l]
        yield self.actor_logstd, lay.logstd

    def _bind_flat(self, device):
        """(Re)create the flat vector on `device` from the current parameter values and re-point every
        nn.Parameter at its slice, so optimiser kernels and the module tree share storage."""
        flat = torch.zeros(self.layout.n_params, dtype=torch.float32, device=device)
        for p, off in self._param_slots():
            flat[off : off + p.numel()] = p.detach().reshape(-1).to(device)
        for p, off in self._param_slots():
            p.data = flat[off : off + p.numel()].view(p.shape)
        self._flat = flat
        self._wc = ops.weight_copies(self.dims, self.layout, device)
        self._act_ws = None
        if device.type == "cuda":
            self.sync_weights()

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        new_device = self.actor_logstd.device
        if self._flat is None or new_device != self._flat.device or self.actor_logstd.data_ptr() != (
            self._flat.data_ptr() + self.layout.logstd * 4
        ):
            self._bind_flat(new_device)
        return self

    def parameters_flat(self) -> torch.Tensor:
        return self._flat

    def sync_weights(self) -> None:
        """Refresh the operand-precision compute copies after the fp32 parameters were written from outside
        (load_state_dict, manual edits).  The Adam kernel keeps them fresh during training."""
        if self._flat is not None and self._flat.is_cuda:
            ops.cast_weights(self.dims, self._flat, self._wc)

    # -- inference ---------------------------------------------------------------------------------
    def _workspace(self, rows):
        need = self._ws_need.get(rows)
        if need is None:
            need = self._ws_need[rows] = L.load().catb200_mlp_workspace_bytes(self.dims, rows, 0)
        if self._act_ws is None or self._act_ws.numel() * 8 < need:
            self._act_ws = L.zeros_workspace(need, self._flat.device)
        return self._act_ws

    def _as_operand(self, x):
        """Observations as the first GEMM reads them: [rows, obs_pad] in the operand precision (already converted
        rows -- the trainer's padded rollout copy -- pass through)."""
        if x.shape[-1] == self.dims.obs_pad and x.dtype == L.operand_dtype(self.dims.prec):
            return x
        x = x.float() if x.dtype != torch.float32 else x
        return ops.obs_to_operand(self.dims, x if x.is_contiguous() else x.contiguous())

    def get_value(self, x):
        obs_op = self._as_operand(x)
        rows = obs_op.numel() // self.dims.obs_pad
        value = torch.empty((rows, 1), dtype=torch.float32, device=obs_op.device)
        ops.mlp_act(self.dims, obs_op, self._flat, self._wc, self._workspace(rows), value=value)
        return value

    def get_action_and_value(self, x, action=None, deterministic=False, out=None, validate=True, rng_state=None, workspace=None):
        """-> (action, log-prob summed over action dims, entropy summed over action dims, value [N,1]).

        `out=(action, logprob, value)` lets the trainer write straight into its rollout buffers; `rng_state`
        (ops.make_rng_state) draws Normal.sample()'s noise inside the head kernel instead of torch's generator;
        `workspace` is caller-owned activation scratch (the trainer's CUDA graphs must not share the Agent's,
        which is regrown on demand)."""
        obs_op = self._as_operand(x)
        rows = obs_op.numel() // self.dims.obs_pad
        dev = obs_op.device
        if out is None:
            act_out = torch.empty((rows, self.act_dim), dtype=torch.float32, device=dev)
            logprob = torch.empty(rows, dtype=torch.float32, device=dev)
            value = torch.empty((rows, 1), dtype=torch.float32, device=dev)
        else:
            act_out, logprob, value = out
        noise = None
        sample = action is None and not deterministic
        if sample and rng_state is None:  # Normal.sample's eps, from torch's generator
            if self._noise is None or self._noise.shape[0] != rows or self._noise.device != dev:
                self._noise = torch.empty((rows, self.act_dim), dtype=torch.float32, device=dev)
            noise = self._noise.normal_()
        ops.mlp_act(
            self.dims, obs_op, self._flat, self._wc, workspace if workspace is not None else self._workspace(rows),
            noise=noise, rng_state=rng_state if sample else None,
            action_in=None if action is None else action.contiguous(), action=act_out, logprob=logprob, value=value,
            validate=validate,
        )  # fmt: skip
        if out is not None:
            return act_out, logprob, None, value  # trainer fast path: the rollout never uses the entropy
        # entropy of a diagonal Gaussian does not depend on the state: sum_j (0.5 + 0.5 log 2pi + logstd_j)
        entropy = (0.5 + 0.5 * np.log(2 * np.pi) + self.actor_logstd.detach()).sum().expand(rows)
        return act_out, logprob, entropy, value

    def forward(self, x, deterministic=True):
        action, _, _, _ = self.get_action_and_value(self.obs_rms(x, update=False), deterministic=deterministic)
        return action

    def export_module(self) -> nn.Module:
        """Plain torch module computing `forward(x)` (normalise -> actor mean) from the same weights, for the
        ONNX / TorchScript export of the reference's play script (scripts/clean_rl/play.py:123-137)."""

        class _Export(nn.Module):
            def __init__(self, agent):
                super().__init__()
                self.actor_mean = agent.actor_mean
                self.register_buffer("mean", agent.obs_rms.running_mean)
                self.register_buffer("var", agent.obs_rms.running_var)
                self.eps = agent.obs_rms.epsilon

            def forward(self, x):
                return self.actor_mean((x - self.mean) / torch.sqrt(self.var + self.eps))

        return _Export(self)


class PPOTrainer:
    """State and steps of the training loop; `PPO()` below drives it exactly like the reference function."""

    def __init__(self, envs, cfg, device=None, seed=None, use_graphs=True, distributed=True, precision=None):
        self.envs = envs
        self.cfg = cfg
        env = envs.unwrapped
        self.num_envs = int(env.num_envs)
        self.T = int(cfg.num_steps)
        if device is None:
            device = getattr(env, "device", None) or torch.device("cuda")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("PPOTrainer needs a CUDA device: the hot path exists only as sm_100a kernels")
        # distributed=False keeps this trainer rank-local even inside an initialised process group (bench probes)
        self.world, self.rank = (cdist.world_size(), cdist.rank()) if distributed else (1, 0)

        self.agent = Agent(envs, device=self.device, precision=precision)
        if self.world > 1:  # rank 0 seeds everybody (the reference's multi-process front-ends do the same)
            cdist.broadcast_params(self.agent.parameters_flat(), src=0)
            self.agent.sync_weights()
        dims, lay = self.agent.dims, self.agent.layout
        N, T, O, A, dev = self.num_envs, self.T, self.agent.obs_dim, self.agent.act_dim, self.device
        self.batch_size = N * T
        self.minibatch_size = min(int(cfg.minibatch_size), self.batch_size)

        f32 = dict(dtype=torch.float32, device=dev)
        self.obs = torch.zeros((T + 1, N, O), **f32)
        # the normalised observations once more as the first GEMM reads them: zero-padded rows in the operand precision
        self.obs_op = torch.zeros((T + 1, N, dims.obs_pad), dtype=L.operand_dtype(dims.prec), device=dev)
        self.actions = torch.zeros((T, N, A), **f32)
        self.logprobs = torch.zeros((T, N), **f32)
        self.rewards = torch.zeros((T, N), **f32)
        self.dones = torch.zeros((T + 1, N), **f32)
        self.true_dones = torch.zeros((T + 1, N), **f32)
        self.values = torch.zeros((T, N), **f32)
        self.advantages = torch.zeros((T, N), **f32)
        self.returns = torch.zeros((T, N), **f32)
        self.next_value = torch.zeros(N, **f32)
        self.norm_stats = torch.zeros(4, **f32)
        self.value_rms_state = torch.tensor([0.0, 1.0, 1.0], **f32)  # mean, var, count of agent.value_rms

        self.grads = torch.zeros(lay.n_params, **f32)
        # several ranks: the gradient of minibatch k is accumulated straight into arena k & 1 of a peer-visible block and
        # exchanged by ONE kernel per optimizer step over NVLink (dist.PeerGradExchange); CATB200_PEER_ALLREDUCE=0 or an
        # odd number of minibatches per epoch falls back to torch.distributed's NCCL all-reduce between two graphs
        self.peer = None
        n_mb = -(-self.batch_size // self.minibatch_size)
        if self.world > 1 and dev.type == "cuda" and os.environ.get("CATB200_PEER_ALLREDUCE", "1") != "0" and n_mb % 2 == 0:
            self.peer = cdist.PeerGradExchange(lay.n_params, dev)
        self.exp_avg = torch.zeros(lay.n_params, **f32)
        self.exp_avg_sq = torch.zeros(lay.n_params, **f32)
        self.lr_dev = torch.tensor(float(cfg.learning_rate), **f32)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.opt_ws = torch.zeros(8, dtype=torch.int64, device=dev)
        self.loss_acc = torch.zeros(8, **f32)
        self.grad_norm = torch.zeros(1, **f32)
        self.train_ws = ops.mlp_workspace(dims, self.minibatch_size, True, dev)
        self.gae_ws = ops.Workspace(dev)
        self.perm = torch.zeros(self.batch_size, dtype=torch.int64, device=dev)
        # private activation scratch of the rollout policy / bootstrap value: CUDA graphs bake its address in, so it
        # must not be the Agent's own workspace (public Agent calls with other batch sizes regrow that one)
        self.act_ws = ops.mlp_workspace(dims, N, False, dev)
        # device-side Philox state {seed, offset}: action noise and minibatch permutations (graph-capturable)
        base_seed = int(seed if seed is not None else getattr(cfg, "seed", 0))
        self.rng_state = ops.make_rng_state(base_seed * 1000003 + 7919 * self.rank + 1, dev)
        self.device_rng = os.environ.get("CATB200_DEVICE_RNG", "1") != "0"
        self.fused_opt = os.environ.get("CATB200_FUSED_OPT", "1") != "0"
        self.hp = ops.make_hparams(cfg.clip_coef, cfg.ent_coef, cfg.vf_coef, cfg.norm_adv, cfg.clip_vloss)
        self.global_step = 0
        self.iteration = 0
        # CUDA graphs: one whole-epoch graph on a single GPU; with several ranks the gradient and optimizer parts of
        # every minibatch replay as graphs around the eager NCCL all-reduce (CATB200_MULTI_GPU_GRAPHS=0 -> all eager).
        self.use_graphs = bool(use_graphs) and (self.world == 1 or os.environ.get("CATB200_MULTI_GPU_GRAPHS", "1") != "0")
        self._split_graphs = None  # multi-GPU: (per-minibatch gradient graphs, optimizer graph)
        self._epoch_graphs: dict = {}  # in-graph shuffle? -> (graph, kernels recorded)
        self._captured_launches = 0  # kernels recorded (not executed) while capturing epoch graphs
        self._replayed_launches = 0  # kernels executed by epoch-graph replays
        self._eager_epochs = 0
        self._policy_graphs: dict = {}
        self._step_graphs: dict = {}  # (slot, state-pointer key) -> (graph, kernels recorded, reset-statistics buffer)
        self.step_graphs = os.environ.get("CATB200_STEP_GRAPHS", "1") != "0"
        self._policy_graph_launches = 0
        self._policy_captures = 0
        self._policy_replays = 0
        self._split_launches = 0
        self._split_captured = 0
        self._split_replays = 0
        self._validate = True  # argument checks of the ops wrappers; switched off after the first iteration
        # thread_local: with a process group alive, the NCCL watchdog thread may query events while this thread captures
        self._capture_mode = "thread_local" if self.world > 1 else "global"

    # -- rollout -------------------------------------------------------------------------------------
    def start(self):
        """envs.reset() and the first observation normalisation (reference ppo.py:186-189)."""
        first = self.envs.reset()[0]["policy"]
        self._ingest_obs(first, 0)
        self.dones[0].zero_()
        self.true_dones[0].zero_()

    def _ingest_obs(self, raw_obs, slot):
        raw_obs = raw_obs.float() if raw_obs.dtype != torch.float32 else raw_obs
        # update + normalise (ppo.py:187,225); the same kernel also emits the bf16 rows the MLP reads
        self.agent.obs_rms(raw_obs, update=True, out=self.obs[slot], out_op=self.obs_op[slot], validate=self._validate)

    def _policy(self, t):
        """Agent.get_action_and_value on slot t -> actions[t], logprobs[t], values[t] (ppo.py:208-212);
        after the first iteration the launches (3 GEMMs + head with in-kernel Philox noise) replay as one CUDA graph
        per slot."""
        rng = self.rng_state if self.device_rng else None
        if not self.use_graphs or self.iteration < 2:
            self.agent.get_action_and_value(
                self.obs_op[t], out=(self.actions[t], self.logprobs[t], self.values[t]), validate=self._validate,
                rng_state=rng, workspace=self.act_ws,
            )  # fmt: skip
            return
        graph = self._policy_graphs.get(t)
        if graph is None:
            graph = torch.cuda.CUDAGraph()
            before = L.launch_count()
            with torch.cuda.graph(graph, capture_error_mode=self._capture_mode):
                self.agent.get_action_and_value(
                    self.obs_op[t], out=(self.actions[t], self.logprobs[t], self.values[t]), validate=False,
                    rng_state=rng, workspace=self.act_ws,
                )  # fmt: skip
            self._policy_graph_launches = L.launch_count() - before
            self._policy_captures += 1
            self._policy_graphs[t] = graph
        graph.replay()
        self._policy_replays += 1

    def _post_step(self, t, next_obs, reward, next_done, timeouts):
        """rewards[t], dones / true_dones[t + 1] (ppo.py:215-224) and the normalised next observation (ppo.py:225)."""
        if next_done.dtype != torch.float32:
            next_done = next_done.to(torch.float)
        if reward.dtype != torch.float32 or not reward.is_contiguous():
            reward = reward.float().contiguous()
        if timeouts.dtype not in (torch.bool, torch.uint8) or not timeouts.is_contiguous():
            timeouts = (timeouts != 0).contiguous()
        ops.rollout_append(reward, next_done, timeouts, self.rewards[t], self.dones[t + 1], self.true_dones[t + 1], validate=self._validate)
        self._ingest_obs(next_obs["policy"], t + 1)

    def _graphable_env(self):
        env = self.envs.unwrapped
        return all(hasattr(env, m) for m in ("step_host", "step_device", "step_finish_host", "pointer_key")) and env.pointer_key() is not None

    def _rollout_step_graphed(self, t):
        """The whole env step -- policy, the env's device half (fused constraint step), append, observation statistics
        and normalisation -- as ONE CUDA-graph launch.  Needs an env that separates its host bookkeeping from its kernels
        (`step_host` / `step_device` / `step_finish_host`) and names the state tensors the step reads (`pointer_key`):
        the synthetic env does; with Isaac Lab's own `step` in between only the policy is graphed."""
        env = self.envs.unwrapped
        env.step_host()
        key = (t, env.pointer_key())
        entry = self._step_graphs.get(key)
        if entry is None:
            if len(self._step_graphs) >= 512:
                self._step_graphs.clear()
            graph = torch.cuda.CUDAGraph()
            stats = torch.empty(2 * max(1, len(getattr(env.constraint_manager, "active_terms", []))), dtype=torch.float, device=self.device)
            before = L.launch_count()
            with torch.cuda.graph(graph, capture_error_mode=self._capture_mode):
                self.agent.get_action_and_value(
                    self.obs_op[t], out=(self.actions[t], self.logprobs[t], self.values[t]), validate=False,
                    rng_state=self.rng_state if self.device_rng else None, workspace=self.act_ws,
                )  # fmt: skip
                env.step_device(uniform=True, fused_out=stats)
                self._post_step(t, env.obs_buf, env.reward_buf, env._dones, env.reset_time_outs)
            entry = self._step_graphs[key] = (graph, L.launch_count() - before, stats)
            self._captured_launches += entry[1]
        graph, n, stats = entry
        graph.replay()
        self._replayed_launches += n
        if env._resetting and env.constraint_manager is not None:  # the statistics of this step's resets: this graph's own buffer
            env.extras["log"] = env.constraint_manager.fused_reset_stats(stats)
        env.step_finish_host()
        info = env.extras
        info["true_dones"] = env.reset_time_outs
        return info

    def rollout_step(self, t):
        """One env step of the rollout (reference ppo.py:201-230)."""
        self.global_step += self.num_envs * self.world
        # obs[t], dones[t], true_dones[t] are already in place (slot t was filled by the previous post-step)
        if self.use_graphs and self.step_graphs and self.iteration >= 2 and not self._validate and self._graphable_env():
            return self._rollout_step_graphed(t)
        self._policy(t)
        next_obs, reward, next_done, timeouts, info = self.envs.step(self.actions[t])
        self._post_step(t, next_obs, reward, next_done, timeouts)
        info["true_dones"] = timeouts
        return info

    def collect_rollout(self):
        ep_infos = []
        for t in range(self.T):
            info = self.rollout_step(t)
            if "episode" in info:
                ep_infos.append(info["episode"])
            elif "log" in info:
                ep_infos.append(info["log"])
        return ep_infos

    # -- update --------------------------------------------------------------------------------------
    def compute_gae(self, bootstrap=True):
        """Bootstrap value, GAE with float dones, value-normalisation statistics (ppo.py:251-288).
        `bootstrap=False` keeps whatever `self.next_value` holds (replaying a recorded rollout)."""
        a = self.agent
        if bootstrap:
            ops.mlp_act(a.dims, self.obs_op[self.T], a.parameters_flat(), a._wc, self.act_ws, value=self.next_value)
        vr = a.value_rms
        self.value_rms_state[0:1].copy_(vr.running_mean.reshape(1))
        self.value_rms_state[1:2].copy_(vr.running_var.reshape(1))
        self.value_rms_state[2:3].copy_(vr.count.reshape(1))
        ops.gae(
            self.rewards, self.values, self.dones, self.true_dones, self.next_value, self.cfg.gamma, self.cfg.gae_lambda,
            advantages=self.advantages, returns=self.returns, value_rms=self.value_rms_state, norm_stats=self.norm_stats,
            workspace=self.gae_ws,
        )  # fmt: skip
        vr.running_mean.copy_(self.value_rms_state[0])
        vr.running_var.copy_(self.value_rms_state[1])
        vr.count.copy_(self.value_rms_state[2])

    def _minibatch_grad(self, mb_inds, grads=None):
        a = self.agent
        B = self.batch_size
        ops.ppo_minibatch_grad(
            a.dims, self.hp, mb_inds, self.obs_op.view(-1, a.dims.obs_pad), self.actions.view(B, -1),
            self.logprobs.view(B), self.advantages.view(B), self.returns.view(B), self.values.view(B), self.norm_stats,
            a.parameters_flat(), a._wc, self.grads if grads is None else grads, self.loss_acc, self.train_ws,
        )  # fmt: skip

    def _minibatch_opt(self):
        a = self.agent
        ops.adam_step(
            a.dims, a.parameters_flat(), self.grads, self.exp_avg, self.exp_avg_sq, a._wc, self.lr_dev, self.step_dev,
            self.opt_ws, max_grad_norm=self.cfg.max_grad_norm, eps=1e-5, grad_scale=1.0 / self.world,
            grad_norm_out=self.grad_norm,
        )  # fmt: skip

    def _minibatch(self, mb_inds, index=0):
        """One optimizer step on minibatch number `index` of the epoch (ppo.py:298-354)."""
        if self.peer is not None:
            # gradient into the peer-visible arena of this minibatch's parity; ONE kernel exchanges it with every rank over
            # NVLink, sums in rank order and produces the clip coefficient; Adam then consumes the private sum
            a, parity, B = self.agent, index & 1, self.batch_size
            if self.fused_opt:  # ... and the fold / exchange / norm / clip / Adam / operand refresh tail is ONE launch
                self.peer.minibatch_update(
                    parity, a.dims, self.hp, mb_inds, self.obs_op.view(-1, a.dims.obs_pad), self.actions.view(B, -1),
                    self.logprobs.view(B), self.advantages.view(B), self.returns.view(B), self.values.view(B), self.norm_stats,
                    a.parameters_flat(), a._wc, self.loss_acc, self.train_ws, self.exp_avg, self.exp_avg_sq, self.lr_dev,
                    self.step_dev, self.opt_ws, max_grad_norm=self.cfg.max_grad_norm, eps=1e-5, grad_norm_out=self.grad_norm,
                )  # fmt: skip
                return
            self._minibatch_grad(mb_inds, grads=self.peer.arena[parity])
            gsum = self.peer.reduce(parity, self.step_dev, self.opt_ws, max_grad_norm=self.cfg.max_grad_norm, grad_norm_out=self.grad_norm)
            ops.adam_apply(a.dims, a.parameters_flat(), gsum, self.exp_avg, self.exp_avg_sq, a._wc, self.lr_dev, self.opt_ws,
                           eps=1e-5, grad_scale=1.0 / self.world)  # fmt: skip
            return
        if self.world == 1 and self.fused_opt:
            # single GPU: gradient + optimizer as one call whose tail (fold, norm, clip, Adam, operand copies) is one launch
            a, B = self.agent, self.batch_size
            ops.ppo_minibatch_update(
                a.dims, self.hp, mb_inds, self.obs_op.view(-1, a.dims.obs_pad), self.actions.view(B, -1),
                self.logprobs.view(B), self.advantages.view(B), self.returns.view(B), self.values.view(B), self.norm_stats,
                a.parameters_flat(), a._wc, self.grads, self.loss_acc, self.train_ws, self.exp_avg, self.exp_avg_sq,
                self.lr_dev, self.step_dev, self.opt_ws, max_grad_norm=self.cfg.max_grad_norm, eps=1e-5,
                grad_norm_out=self.grad_norm,
            )  # fmt: skip
            return
        self._minibatch_grad(mb_inds)
        if self.world > 1:  # the one exchange step: sum-allreduce of the flat 1.5 MB gradient over NVLink (NCCL)
            cdist.allreduce_grads(self.grads)
        self._minibatch_opt()

    def _capture(self, fn):
        graph = torch.cuda.CUDAGraph()
        before = L.launch_count()
        # thread_local: the NCCL watchdog thread may query events while this thread captures
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            fn()
        return graph, L.launch_count() - before

    def _epoch_split_graphs(self):
        """Multi-GPU epoch: the gradient part of every minibatch and the optimizer part replay as CUDA graphs;
        the NCCL all-reduce between them stays an ordinary stream operation."""
        if self._split_graphs is None:
            grads, n = [], 0
            for start in range(0, self.batch_size, self.minibatch_size):
                g, k = self._capture(lambda s=start: self._minibatch_grad(self.perm[s : s + self.minibatch_size]))
                grads.append(g)
                n += k
            opt, k = self._capture(self._minibatch_opt)
            self._split_graphs = (grads, opt)
            self._split_launches = n + k * len(grads)
            self._split_captured = n + k
        grads, opt = self._split_graphs
        for g in grads:
            g.replay()
            cdist.allreduce_grads(self.grads)
            opt.replay()
        self._split_replays += 1

    def _shuffle(self):
        """b_inds = randperm(batch) of this epoch (ppo.py:295): one Feistel/Philox launch into self.perm."""
        if self.device_rng:
            ops.random_permutation(self.batch_size, self.rng_state, out=self.perm)
        else:
            self.perm.copy_(torch.randperm(self.batch_size, device=self.device))

    def _epoch(self, shuffle=False):
        if shuffle:
            self._shuffle()
        for index, start in enumerate(range(0, self.batch_size, self.minibatch_size)):
            self._minibatch(self.perm[start : start + self.minibatch_size], index)

    def update(self, perms=None):
        """All epochs x minibatches of one iteration (ppo.py:290-354).  `perms` (one index permutation per
        epoch) replaces the `torch.randperm` draws, for replaying a recorded run."""
        self.loss_acc.zero_()
        # the permutation launch is part of the epoch graph when it is drawn on the device (its Philox offset lives
        # in device memory); recorded permutations (`perms`) and torch.randperm are copied in before the replay
        in_graph_shuffle = perms is None and self.device_rng
        for epoch in range(int(self.cfg.updates_epochs)):
            if perms is not None:
                self.perm.copy_(perms[epoch])
            elif not in_graph_shuffle:
                self._shuffle()
            if self.use_graphs and self.world > 1 and self.peer is None and self._eager_epochs >= 1:
                if in_graph_shuffle:
                    self._shuffle()
                self._epoch_split_graphs()
                self._eager_epochs += 1
                continue
            key = bool(in_graph_shuffle)
            if self.use_graphs and key not in self._epoch_graphs and self._eager_epochs >= 1:
                # the first epoch ever ran eagerly (it also set the kernel attributes); record the identical
                # launch sequence once and replay it from now on
                graph = torch.cuda.CUDAGraph()
                before = L.launch_count()
                with torch.cuda.graph(graph, capture_error_mode=self._capture_mode):
                    self._epoch(shuffle=key)
                self._epoch_graphs[key] = (graph, L.launch_count() - before)
                self._captured_launches += L.launch_count() - before
            if key in self._epoch_graphs:
                graph, n = self._epoch_graphs[key]
                graph.replay()
                self._replayed_launches += n
                self._eager_epochs += 1
            else:
                self._epoch(shuffle=key)
                self._eager_epochs += 1

    def kernel_launches(self) -> int:
        """libcatb200 kernels executed so far by this process (graph replays included)."""
        captured = self._captured_launches
        captured += self._policy_graph_launches * self._policy_captures
        captured += self._split_captured
        replayed = self._replayed_launches + self._policy_graph_launches * self._policy_replays
        replayed += self._split_launches * self._split_replays
        return L.launch_count() - captured + replayed

    def finish_iteration(self):
        """Slot T (the bootstrap observation / dones) becomes slot 0 of the next rollout (ppo.py:203-205)."""
        T = self.T
        self.obs[0].copy_(self.obs[T])
        self.obs_op[0].copy_(self.obs_op[T])
        self.dones[0].copy_(self.dones[T])
        self.true_dones[0].copy_(self.true_dones[T])

    def set_lr(self, lr: float):
        self.lr_dev.fill_(lr)

    def train_iteration(self):
        """rollout + GAE + update; returns the episode infos gathered during the rollout."""
        self.iteration += 1
        cfg = self.cfg
        if cfg.anneal_lr:  # ppo.py:196-199
            frac = 1.0 - (self.iteration - 1.0) / cfg.num_iterations
            self.set_lr(frac * cfg.learning_rate)
        ep_infos = self.collect_rollout()
        self.compute_gae()
        self.update()
        self.finish_iteration()
        self._validate = False
        return ep_infos

    def checkpoint_state(self) -> dict:
        """`agent.state_dict()` (the reference's keys, ppo.py:357) for a checkpoint.  With several ranks the observation /
        value normalisers in it are the pooled statistics of all env shards (dist.pooled_moments; every rank must call
        this), while the live ones stay rank-local."""
        state = {k: v.detach().clone() for k, v in self.agent.state_dict().items()}
        if self.world > 1:
            for name in ("obs_rms", "value_rms"):
                rms = getattr(self.agent, name)
                m, v, n = cdist.pooled_moments(rms.running_mean, rms.running_var, rms.count)
                state[f"{name}.running_mean"], state[f"{name}.running_var"], state[f"{name}.count"] = m, v, n
        return state

    def losses_async(self):
        """Enqueue the device->host copy of the last update's loss accumulators (into pinned memory, behind the work
        submitted so far) and return a handle for `read_losses`.  A logger that reads the handle of iteration i while
        iteration i + 1 is being submitted never drains the GPU queue, which a blocking `losses()` per iteration does."""
        ring = getattr(self, "_loss_ring", None)
        if ring is None:
            ring = self._loss_ring = [(torch.empty(8, dtype=torch.float32).pin_memory(), torch.cuda.Event()) for _ in range(4)]
            self._loss_seq = 0
        buf, ev = ring[self._loss_seq % len(ring)]
        self._loss_seq += 1
        buf.copy_(self.loss_acc, non_blocking=True)
        ev.record()
        return buf, ev

    def read_losses(self, handle) -> dict:
        """Wait for the copy `losses_async` enqueued (not for anything submitted after it) and return the mean losses."""
        buf, ev = handle
        ev.synchronize()
        return self._loss_dict(buf.clone())

    def losses(self) -> dict:
        """Mean losses over the minibatches of the last update (one blocking device->host read)."""
        return self._loss_dict(self.loss_acc.cpu())

    def _loss_dict(self, acc) -> dict:
        if self.peer is not None:
            self.peer.check()
        n = max(float(acc[7]), 1.0)
        return {
            "mean_pg_loss": float(acc[0]) / n,
            "mean_v_loss": float(acc[1]) / n,
            "mean_entropy_loss": float(acc[2]) / n,
            "approx_kl": float(acc[3]) / n,
            "clipfrac": float(acc[4]) / n,
            "mean_surrogate_loss": float(acc[6]) / n,
        }


def _make_writer(ppo_cfg, run_path):
    if ppo_cfg.logger == "wandb":
        from rsl_rl.utils.wandb_utils import WandbSummaryWriter

        return WandbSummaryWriter(log_dir=run_path, flush_secs=10, cfg=ppo_cfg.to_dict())
    if ppo_cfg.logger == "tensorboard":
        from torch.utils.tensorboard import SummaryWriter as TensorboardSummaryWriter

        return TensorboardSummaryWriter(log_dir=run_path)
    if ppo_cfg.logger is None:  # extension: no logging (benchmarks)
        return None
    raise AssertionError("logger type not found")


def PPO(envs, ppo_cfg, run_path):
    """Train; same signature, logging keys and checkpoint files as the reference (ppo.py:126-372)."""
    # one writer / log directory per job: rank 0's (every rank of the reference's distributed front-ends logs for itself)
    writer = _make_writer(ppo_cfg, run_path) if cdist.rank() == 0 else None
    if cdist.rank() == 0 and not os.path.exists(run_path):
        os.makedirs(run_path)
    trainer = PPOTrainer(envs, ppo_cfg)
    device = trainer.device
    trainer.start()
    print(f"Starting training for {ppo_cfg.num_iterations} steps")
    start_time = time.time()
    for iteration in range(1, ppo_cfg.num_iterations + 1):
        ep_infos = trainer.train_iteration()
        if writer is not None and ep_infos:  # adapted from rsl_rl like the reference (ppo.py:232-248)
            for key in ep_infos[0]:
                infotensor = torch.tensor([], device=device)
                for ep_info in ep_infos:
                    if key not in ep_info:
                        continue
                    if not isinstance(ep_info[key], torch.Tensor):
                        ep_info[key] = torch.Tensor([ep_info[key]])
                    if len(ep_info[key].shape) == 0:
                        ep_info[key] = ep_info[key].unsqueeze(0)
                    infotensor = torch.cat((infotensor, ep_info[key].to(device)))
                value = torch.mean(infotensor)
                writer.add_scalar(key if "/" in key else "Episode/" + key, value, iteration)
        if writer is not None:
            losses = trainer.losses()
            for name in ("mean_pg_loss", "mean_entropy_loss", "mean_v_loss", "mean_surrogate_loss"):
                writer.add_scalar("Loss/" + name, losses[name], iteration)
            writer.add_scalar("Loss/learning_rate", float(trainer.lr_dev), iteration)
        if (iteration + 1) % ppo_cfg.save_interval == 0:
            state = trainer.checkpoint_state()  # collective with several ranks: pooled normaliser statistics
            if trainer.rank == 0:
                torch.save(state, f"{run_path}/model_{iteration}.pt")
                print("Saved model")
    torch.cuda.synchronize()
    elapsed = time.time() - start_time
    print(f"Trained {trainer.global_step} env steps in {elapsed:.1f} s")
    return trainer
