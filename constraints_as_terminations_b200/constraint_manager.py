"""`ConstraintManager` / `CaT`: the reference's manager API on top of the fused sm_100a CaT kernels.

Mirrors the public surface of `exts/cat_envs/cat_envs/tasks/utils/cat/constraint_manager.py`
(`CaT` :22-116, `ConstraintManager` :119-264): constructor signature, `compute()`, `reset(env_ids)`,
`active_terms`, `get_term_cfg` / `set_term_cfg`, `__str__`, the `cat` attribute with its dict views and
helper getters, and the same `TypeError` / `ValueError` behaviour.  What changed is the execution model:

* the per-term python loop with ~10 eager kernels and one device->host sync per term is replaced by
  one `catb200_cat_step` call (two kernel launches, no sync) that evaluates every term of every env,
  updates the Polyak running maxima, produces `cstr_prob` and accumulates the episode statistics;
* persistent state is a handful of flat device tensors (`running_max[K]`, `[S, N]` statistics) that the
  dict-style attributes of the reference (`cat.running_maxes[name]`, `_episode_sums[name]`, ...) view.

There is no torch/CPU fallback: on a non-CUDA env `compute()` raises.
"""

from __future__ import annotations

import math
from collections.abc import Sequence

import numpy as np
import torch

from . import _lib as L
from ._isaaclab_compat import ManagerBase, ManagerTermBase
from .constraints import SourceRef, TermSpec, _ids_list
from .manager_constraint_cfg import ConstraintTermCfg


# --------------------------------------------------------------------------------------------------
# plan construction (host side of catb200_plan_t)
# --------------------------------------------------------------------------------------------------
def _as_source_tensor(t: torch.Tensor, what: str):
    """Normalise a tensor read by a term into (tensor, row_len, row_stride, dtype, bodies)."""
    L.require_cuda(t, what)
    if t.dtype == torch.bool:
        t = t.view(torch.uint8)
    elif t.dtype not in (torch.float32, torch.uint8):
        t = t.float()  # same cast CaT.add applies to non-float constraints (reference :45-47)
    if t.ndim == 1:
        t = t.unsqueeze(1)
    bodies = t.shape[2] if t.ndim == 4 else 0
    row_len = int(math.prod(t.shape[1:]))
    inner_contig = t[0].is_contiguous() if t.shape[0] > 0 else True
    if not inner_contig or (t.shape[0] > 1 and t.stride(0) < row_len):
        t = t.contiguous()
    row_stride = int(t.stride(0)) if t.shape[0] > 1 else row_len
    dtype = L.F32 if t.dtype == torch.float32 else L.U8
    return t, row_len, row_stride, dtype, bodies


class _BuiltPlan:
    """catb200_plan_t plus the python-side bookkeeping needed to refresh its pointers every step."""

    def __init__(self):
        self.plan = L.Plan()
        self.fetchers: list = []  # per source: callable(env) -> tensor
        self.keys: list[str] = []
        self.live: list = []  # tensors currently referenced by plan.sources[i].ptr
        self.raw: list = []  # the objects the fetchers returned (identity check for the fast path)
        self.cache: list = []  # per source: id(tensor) -> (tensor, data_ptr, row_stride) of tensors already checked
        self.n_cols = 0
        self.term_cols: list[tuple[int, int]] = []  # per slot: (col_begin, col_end)

    def refresh(self, env) -> None:
        """Re-point the plan at the tensors the env exposes this step.  Simulators hand out the same buffer objects
        every step (pointer compare only); envs that rotate between a few state sets (the synthetic one, a host-fed
        staging ring) hit a small per-source cache of already normalised tensors, so the full layout check runs only
        the first time a tensor object is seen."""
        srcs = self.plan.sources
        for i, fetch in enumerate(self.fetchers):
            t = fetch(env)
            ptr = t.data_ptr()
            if t is self.raw[i] and ptr == srcs[i].ptr:
                continue
            hit = self.cache[i].get(id(t))
            if hit is not None and hit[0] is t and hit[1] == ptr:
                row_stride = hit[2]
                nt = t
            else:
                nt, row_len, row_stride, dtype, _ = _as_source_tensor(t, self.keys[i])
                if row_len != srcs[i].row_len or dtype != srcs[i].dtype:
                    raise RuntimeError(
                        f"source '{self.keys[i]}' changed layout (row_len {srcs[i].row_len}->{row_len}, "
                        f"dtype {srcs[i].dtype}->{dtype}); rebuild the manager"
                    )
                # only tensors used in place are remembered: a converted copy would go stale
                if nt.data_ptr() == ptr and len(self.cache[i]) < 64:
                    self.cache[i][id(t)] = (t, ptr, row_stride)
            self.raw[i] = t
            self.live[i] = nt
            srcs[i].ptr = nt.data_ptr()
            srcs[i].row_stride = row_stride


def build_plan(env, term_specs):
    """term_specs: list of (name, TermSpec, stat_slot).  Returns (plan, live tensors, n_cols)."""
    built = _build(env, term_specs)
    return built.plan, built.live, built.n_cols


def _build(env, term_specs) -> _BuiltPlan:
    built = _BuiltPlan()
    plan = built.plan
    index: dict[str, int] = {}

    def source(ref: SourceRef | None) -> int:
        if ref is None:
            return L.NO_SOURCE
        if ref.key in index:
            return index[ref.key]
        i = len(built.fetchers)
        if i >= L.MAX_SOURCES:
            raise RuntimeError(f"more than {L.MAX_SOURCES} distinct source tensors in one constraint plan")
        raw = ref.fetch(env)
        t, row_len, row_stride, dtype, bodies = _as_source_tensor(raw, ref.key)
        s = plan.sources[i]
        s.ptr, s.row_len, s.row_stride, s.dtype = t.data_ptr(), row_len, row_stride, dtype
        s.aux = bodies if ref.bodies else 0
        if ref.bodies and bodies == 0:
            raise RuntimeError(f"source '{ref.key}' must be a [N, H, B, 3] contact-force history")
        index[ref.key] = i
        built.fetchers.append(ref.fetch)
        built.keys.append(ref.key)
        built.live.append(t)
        built.raw.append(raw)
        built.cache.append({})
        return i

    n_terms = 0
    col = 0
    slot_begin: dict[int, int] = {}
    for name, spec, slot in term_specs:
        s0 = source(spec.src0)
        src = plan.sources[s0]
        id_space = src.aux if src.aux else src.row_len
        ids = _ids_list(spec.ids, id_space)
        if any(i < 0 or i >= id_space for i in ids):
            raise ValueError(f"term '{name}': ids {ids} out of range for a source with {id_space} entries")
        if any(i > 255 for i in ids):  # catb200_term_t.ids is uint8: ctypes would truncate silently
            raise ValueError(f"term '{name}': element index {max(ids)} > 255 cannot be selected by a fused term (ids are 8-bit); "
                             "slice the source tensor first or use a python term")
        chunks = [ids] if spec.single_column else [ids[k : k + L.MAX_IDS] for k in range(0, len(ids), L.MAX_IDS)]
        if spec.single_column and len(ids) > L.MAX_IDS:
            raise RuntimeError(f"term '{name}': more than {L.MAX_IDS} bodies in one reducing term")
        s1, s2 = source(spec.src1), source(spec.src2)
        slot_begin.setdefault(slot, col)
        for chunk in chunks:
            if n_terms >= L.MAX_TERMS:
                raise RuntimeError(f"more than {L.MAX_TERMS} fused term blocks in one constraint plan")
            t = plan.terms[n_terms]
            t.op = spec.op
            t.n_cols = 1 if spec.single_column else len(chunk)
            t.n_ids = len(chunk)
            t.src0, t.src1, t.src2 = s0, s1, s2
            t.stat_slot = slot
            t.p0, t.p1, t.p2 = float(spec.p0), float(spec.p1), float(spec.p2)
            for k, v in enumerate(chunk):
                t.ids[k] = v
            col += t.n_cols
            n_terms += 1
    plan.n_sources = len(built.fetchers)
    plan.n_terms = n_terms
    L.check(L.load().catb200_cat_plan_finalize(plan), "cat_plan_finalize")
    built.n_cols = plan.n_cols
    bounds = list(plan.slot_col_begin[: plan.n_slots + 1])
    built.term_cols = [(bounds[s], bounds[s + 1]) for s in range(plan.n_slots)]
    return built


class _LazyStats(dict):
    """dict of 0-d statistics tensors that splits the packed `[2 * n_terms]` device vector only when somebody looks
    (a training loop that does not log this step never pays for 26 tensor views)."""

    def __init__(self, keys, packed):
        super().__init__()
        self._pending = (keys, packed)

    def _fill(self):
        if self._pending is not None:
            keys, packed = self._pending
            self._pending = None
            super().update(zip(keys, packed.unbind(0)))

    def __getitem__(self, k):
        self._fill()
        return super().__getitem__(k)

    def __iter__(self):
        self._fill()
        return super().__iter__()

    def __len__(self):
        self._fill()
        return super().__len__()

    def __contains__(self, k):
        self._fill()
        return super().__contains__(k)

    def keys(self):
        self._fill()
        return super().keys()

    def values(self):
        self._fill()
        return super().values()

    def items(self):
        self._fill()
        return super().items()

    def get(self, k, default=None):
        self._fill()
        return super().get(k, default)

    def __eq__(self, other):
        self._fill()
        return super().__eq__(other)

    def __repr__(self):
        self._fill()
        return super().__repr__()


def _generic_spec(name: str, value: torch.Tensor) -> TermSpec:
    """Spec of a python term that has already been evaluated to a tensor [N] or [N, J]."""
    return TermSpec(L.OP_GENERIC, SourceRef(f"generic:{name}", lambda env, v=value: v), None)


# --------------------------------------------------------------------------------------------------
# CaT: probability engine facade (reference constraint_manager.py:22-116)
# --------------------------------------------------------------------------------------------------
class _TermDict:
    """Read-only dict-like view keyed by term name whose values are materialised on access."""

    def __init__(self, names_fn, get_fn):
        self._names_fn, self._get_fn = names_fn, get_fn

    def __getitem__(self, name):
        if name not in self._names_fn():
            raise KeyError(name)
        return self._get_fn(name)

    def __contains__(self, name):
        return name in self._names_fn()

    def __iter__(self):
        return iter(self._names_fn())

    def __len__(self):
        return len(self._names_fn())

    def keys(self):
        return list(self._names_fn())

    def values(self):
        return [self._get_fn(n) for n in self._names_fn()]

    def items(self):
        return [(n, self._get_fn(n)) for n in self._names_fn()]

    def clear(self):
        pass

    def __bool__(self):
        return len(self) > 0


class CaT:
    """Termination probabilities from constraint violations.

    Two uses, same attributes as the reference class:

    * owned by a `ConstraintManager` (the hot path): nothing is computed here; `probs`,
      `raw_constraints`, `running_maxes`, `max_p` are views / lazily rebuilt matrices over the manager's
      fused buffers, valid for the most recent `compute()`.
    * stand-alone: `add(name, constraint, max_p)` (reference :39-76) runs the same kernels on a
      one-term generic plan with its own per-name running max, for code that drives CaT by hand.
    """

    def __init__(self, tau: float = 0.95, min_p: float = 0.0):
        self.tau = tau
        self.min_p = min_p
        self._device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self._manager = None
        self._own: dict[str, dict] = {}  # stand-alone terms: name -> state
        self._order: list[str] = []
        self.probs = _TermDict(self._names, self._probs_of)
        self.raw_constraints = _TermDict(self._names, self._raw_of)
        self.running_maxes = _TermDict(self._rm_names, self._rm_of)
        self.max_p = _TermDict(self._names, self._max_p_of)

    # -- name sets -------------------------------------------------------------------------------
    def _names(self):
        if self._manager is not None:
            return self._manager._term_names if self._manager._computed else []
        return [n for n in self._order if self._own[n]["fresh"]]

    def _rm_names(self):
        if self._manager is not None:
            return self._manager._term_names if self._manager._computed else []
        return list(self._order)

    # -- per-term getters --------------------------------------------------------------------------
    def _probs_of(self, name):
        if self._manager is not None:
            c0, c1 = self._manager._cols_of(name)
            return self._manager._probs_matrix()[:, c0:c1]
        return self._own[name]["probs"]

    def _raw_of(self, name):
        if self._manager is not None:
            c0, c1 = self._manager._cols_of(name)
            return self._manager._raw_matrix()[:, c0:c1]
        return self._own[name]["raw"]

    def _rm_of(self, name):
        if self._manager is not None:
            c0, c1 = self._manager._cols_of(name)
            return self._manager._running_max[c0:c1].unsqueeze(0)
        return self._own[name]["running_max"].unsqueeze(0)

    def _max_p_of(self, name):
        if self._manager is not None:
            c0, c1 = self._manager._cols_of(name)
            mp = float(self._manager.get_term_cfg(name).max_p)
            return torch.full((c1 - c0,), mp, dtype=torch.float, device=self._manager._device)
        st = self._own[name]
        return torch.full((st["cols"],), st["max_p"], dtype=torch.float, device=st["raw"].device)

    # -- reference API ---------------------------------------------------------------------------
    def reset(self):
        """Forget this step's probabilities / raw constraints; running maxima persist (reference :34-37)."""
        for st in self._own.values():
            st["fresh"] = False

    def add(self, name: str, constraint: torch.Tensor, max_p: float = 0.1):
        """Process one constraint tensor [N] or [N, J] (any dtype) exactly like reference :39-76."""
        if self._manager is not None:
            raise RuntimeError("this CaT is owned by a ConstraintManager; terms are added through its cfg")
        if constraint.device != self._device:
            constraint = constraint.to(self._device)
        L.require_cuda(constraint, f"constraint '{name}'")
        raw = constraint if torch.is_floating_point(constraint) else constraint.float()
        if raw.ndim == 1:
            raw = raw.unsqueeze(1)
        n, cols = raw.shape
        st = self._own.get(name)
        if st is None or st["cols"] != cols or st["n"] != n:
            st = {
                "cols": cols,
                "n": n,
                "running_max": torch.zeros(cols, dtype=torch.float32, device=raw.device),
                "rm_init": torch.zeros(cols, dtype=torch.int32, device=raw.device),
                "stats": torch.zeros((2, 1, n), dtype=torch.float32, device=raw.device),
                "cstr": torch.empty(n, dtype=torch.float32, device=raw.device),
                "ws": None,
            }
            self._own[name] = st
            if name not in self._order:
                self._order.append(name)
        built = _build(None, [(name, _generic_spec(name, raw), 0)])
        params = _cat_params(self.tau, self.min_p, [max_p])
        lib = L.load()
        need = lib.catb200_cat_workspace_bytes(n, cols)
        if st["ws"] is None or st["ws"].numel() * 8 < need:
            st["ws"] = L.zeros_workspace(need, raw.device)
        L.check(
            lib.catb200_cat_step(
                built.plan, params, n, st["running_max"].data_ptr(), st["rm_init"].data_ptr(),
                st["stats"][0].data_ptr(), st["stats"][1].data_ptr(), st["cstr"].data_ptr(),
                None, None, None, None, st["ws"].data_ptr(), st["ws"].numel() * 8, L.stream(),
            ),
            "cat_step",
        )  # fmt: skip
        probs = torch.empty((n, cols), dtype=torch.float32, device=raw.device)
        rawm = torch.empty((n, cols), dtype=torch.float32, device=raw.device)  # the values as CaT.add casts them (:45-47)
        L.check(lib.catb200_cat_eval_terms(built.plan, n, rawm.data_ptr(), L.stream()), "cat_eval_terms")
        L.check(
            lib.catb200_cat_probs(built.plan, params, n, st["running_max"].data_ptr(), probs.data_ptr(), rawm.data_ptr(), L.stream()),
            "cat_probs",
        )
        st.update(raw=raw, probs=probs, max_p=float(max_p), fresh=True)

    def get_probs(self) -> torch.Tensor:
        """Max over every column of every term -> [N] (reference :78-82)."""
        if self._manager is not None:
            if not self._manager._computed:
                return torch.tensor([], device=self._device)
            return self._manager._cstr_prob_buf
        names = self._names()
        if not names:
            return torch.tensor([], device=self._device)
        return torch.cat([self._own[n]["probs"] for n in names], dim=1).max(1).values

    def get_raw_constraints(self) -> torch.Tensor:
        vals = self.raw_constraints.values()
        return torch.cat(vals, dim=1) if vals else torch.tensor([], device=self._device)

    def get_running_maxes(self) -> torch.Tensor:
        vals = self.running_maxes.values()
        return torch.cat(vals, dim=1) if vals else torch.tensor([], device=self._device)

    def get_max_p(self) -> torch.Tensor:
        vals = self.max_p.values()
        return torch.cat(vals) if vals else torch.tensor([], device=self._device)

    def get_names(self) -> list[str]:
        return list(self._names())

    def get_vals(self) -> list[float]:
        """Percentage of envs violating each term this step (reference :114-116)."""
        return [100.0 * p.max(1).values.gt(0.0).float().mean().item() for p in self.probs.values()]

    def get_str(self, names: list[str] | None = None) -> str:
        names = names or self.get_names()
        parts = []
        for name in names:
            pct = 100.0 * self.probs[name].max(1).values.gt(0.0).float().mean().item()
            parts.append(f"{name}: {pct:.1f}")
        return " ".join(parts)

    def log_all(self, episode_sums: dict[str, torch.Tensor]):
        """Accumulate per-term violation indicators under `cstr_<name>` (reference :102-109)."""
        for name, probs in self.probs.items():
            key = f"cstr_{name}"
            values = probs.max(1).values.gt(0.0).float()
            if key not in episode_sums:
                episode_sums[key] = torch.zeros_like(values)
            episode_sums[key].add_(values)


def _cat_params(tau: float, min_p: float, max_ps: Sequence[float]) -> L.CatParams:
    prm = L.CatParams()
    prm.tau = tau  # ctypes rounds the python double to fp32, like torch does for tensor * python_float
    prm.one_minus_tau = 1.0 - tau  # double subtraction first (reference :59)
    prm.min_p = min_p
    prm.floor_max = 1e-6  # reference :55
    for i, mp in enumerate(max_ps):
        prm.span[i] = mp - min_p  # double subtraction first (reference :72)
    return prm


# --------------------------------------------------------------------------------------------------
# ConstraintManager (reference constraint_manager.py:119-264)
# --------------------------------------------------------------------------------------------------
class ConstraintManager(ManagerBase):
    """Manager computing the per-env termination probability from the configured constraint terms."""

    def __init__(self, cfg: object, env, tau: float = 0.95, min_p: float = 0.0):
        self.cat = CaT(tau, min_p)
        self.cat._manager = self
        self._device = env.device
        self._term_names: list[str] = []
        self._term_cfgs: list[ConstraintTermCfg] = []
        self._class_term_cfgs: list[ConstraintTermCfg] = []
        self._computed = False
        self._built: _BuiltPlan | None = None
        self._param_snapshot = None
        self._generic: list = []
        self._param_watch: list = []

        super().__init__(cfg, env)  # parses the terms through _prepare_terms()

        self._term_index = {name: i for i, name in enumerate(self._term_names)}
        self._stat_keys = [
            k for name in self._term_names for k in (f"Episode_Constraint_violation/{name}", f"Episode_Constraint_probability/{name}")
        ]
        n, s = self.num_envs, max(1, len(self._term_names))
        dev = self._device
        self._stats = torch.zeros((2, s, n), dtype=torch.float, device=dev)
        self._episode_sums = {name: self._stats[0, i] for i, name in enumerate(self._term_names)}
        self._cstr_mean_values = {name: self._stats[1, i] for i, name in enumerate(self._term_names)}
        self._cstr_prob_buf = torch.zeros(n, dtype=torch.float, device=dev)
        self._reward_buf = torch.zeros(n, dtype=torch.float, device=dev)
        self._dones_buf = torch.zeros(n, dtype=torch.float, device=dev)
        self._running_max = None
        self._rm_init = None
        self._workspace = None
        self._params = _cat_params(tau, min_p, [])
        self._max_p_cache: list[float | None] = [None] * len(self._term_names)
        self._probs_cache = None
        self._raw_cache = None
        self._reset_ws = None
        self._fused_out = None
        # max_p - min_p per term in device memory (the kernels read it there: a CUDA graph of the step stays current when
        # the curriculum changes max_p), refreshed from a pinned host mirror with an asynchronous copy
        self._span_dev = torch.zeros(L.MAX_TERMS, dtype=torch.float, device=dev)
        self._span_host = torch.zeros(L.MAX_TERMS, dtype=torch.float)
        if torch.device(dev).type == "cuda":
            self._span_host = self._span_host.pin_memory()
        self._params.span_dev = self._span_dev.data_ptr()

    # -- reference API -----------------------------------------------------------------------------
    def __str__(self) -> str:
        msg = f"<ConstraintManager> contains {len(self._term_names)} active terms.\n"
        rows = []
        for index, (name, term_cfg) in enumerate(zip(self._term_names, self._term_cfgs)):
            limit_value = term_cfg.params.get("limit", "-")
            names_value = "-"
            asset_cfg = term_cfg.params.get("asset_cfg")
            if asset_cfg is not None:
                names_value = asset_cfg.body_names or asset_cfg.joint_names or "-"
            elif "names" in term_cfg.params:
                import warnings

                warnings.warn(
                    "Using 'names' parameter is deprecated. Use 'asset_cfg' instead.", DeprecationWarning, stacklevel=2
                )
                names_value = term_cfg.params["names"]
            rows.append([index, name, limit_value, names_value, term_cfg.max_p])
        fields = ["Index", "Name", "Limit", "Names", "Max p"]
        try:
            from prettytable import PrettyTable

            table = PrettyTable()
            table.title = "Active Constraint Terms"
            table.field_names = fields
            table.align["Name"] = "l"
            table.align["Limit"] = "r"
            table.align["Max p"] = "r"
            for r in rows:
                table.add_row(r)
            body = table.get_string()
        except ImportError:
            cells = [fields] + [[str(c) for c in r] for r in rows]
            widths = [max(len(r[i]) for r in cells) for i in range(len(fields))]
            line = "+" + "+".join("-" * (w + 2) for w in widths) + "+"
            fmt = lambda r: "| " + " | ".join(c.ljust(w) for c, w in zip(r, widths)) + " |"  # noqa: E731
            body = "\n".join(
                ["Active Constraint Terms", line, fmt(cells[0]), line] + [fmt(r) for r in cells[1:]] + [line]
            )
        return msg + body + "\n"

    @property
    def active_terms(self) -> list[str]:
        return self._term_names

    def set_term_cfg(self, term_name: str, cfg: ConstraintTermCfg):
        i = self._term_index.get(term_name)
        if i is None:
            raise ValueError(f"Constraint term '{term_name}' not found.")
        if cfg is not self._term_cfgs[i]:
            self._built = None  # a swapped cfg may change func / params: re-plan on the next compute
        self._term_cfgs[i] = cfg

    def get_term_cfg(self, term_name: str) -> ConstraintTermCfg:
        i = self._term_index.get(term_name)
        if i is None:
            raise ValueError(f"Constraint term '{term_name}' not found.")
        return self._term_cfgs[i]

    def compute(self) -> torch.Tensor:
        """Termination probability of every env for this step -> Tensor[N] (reference :213-229)."""
        self._launch(None, None)
        return self._cstr_prob_buf

    def compute_step(self, raw_reward: torch.Tensor, reset_buf: torch.Tensor | None, fuse_reset: bool = False, fused_out=None):
        """`compute()` fused with the reward / dones lines of `CaTEnv.step` (reference cat_env.py:100-107,118-121).

        Returns `(reward_buf, dones)`: reward = clip(raw_reward * (1 - cstr_prob), min=0) and
        dones = cstr_prob with 1.0 where `reset_buf` is set.  Both are manager-owned buffers that the
        next call overwrites.

        `fuse_reset=True` additionally performs `reset(ids of the envs flagged in reset_buf)` inside the same two
        launches (the order of `CaTEnv.step`: compute at cat_env.py:100, `constraint_manager.reset` at :181 via
        `_reset_idx`), reading `env.episode_length_buf` before anybody zeroes it; `fused_reset_stats()` then returns
        what `reset()` would have.  `fused_out` (float32 [2 * n_terms]) receives those statistics instead of a fresh
        tensor (a caller that replays the step from a CUDA graph needs the address to be its own).
        """
        L.require_cuda(raw_reward, "raw_reward")
        if raw_reward.dtype != torch.float32 or not raw_reward.is_contiguous():
            raw_reward = raw_reward.float().contiguous()
        if reset_buf is not None:
            if reset_buf.dtype == torch.bool:
                reset_buf = reset_buf.view(torch.uint8)
            elif reset_buf.dtype != torch.uint8:
                reset_buf = (reset_buf != 0).view(torch.uint8)
            if not reset_buf.is_contiguous():
                reset_buf = reset_buf.contiguous()
        if fuse_reset and reset_buf is None:
            raise ValueError("compute_step(fuse_reset=True) needs reset_buf")
        self._launch(raw_reward, reset_buf, fuse_reset=fuse_reset, fused_out=fused_out)
        return self._reward_buf, self._dones_buf

    def sample_terminations(self, rng_state: torch.Tensor, probs: torch.Tensor | None = None, with_ids: bool = True):
        """OPTIONAL stochastic termination mode (default off; the reference itself never samples: it hands the
        probability to the trainer as a float `dones`, U/cat/cat_env.py:107, and GAE uses it as a soft discount).
        Draws mask[i] ~ Bernoulli(p[i]) for the probabilities of the last `compute()` / `compute_step()` (or `probs`)
        from the device Philox stream `rng_state` (ops.make_rng_state): mask[i] = (u_i < p[i]).  Returns
        (mask bool [N], ids int64 [N] buffer, count int32 [1]) -- ids[:count] are the terminated env indices in ascending
        order, bit-exact against oracle/philox_oracle.py -- or just the mask."""
        from . import ops

        p = self._dones_buf if probs is None and self._computed else (self._cstr_prob_buf if probs is None else probs)
        return ops.bernoulli_mask(p, rng_state, with_ids=with_ids)

    def fused_reset_stats(self, packed=None) -> dict[str, torch.Tensor]:
        """Episode statistics gathered by the last `compute_step(..., fuse_reset=True)`: same keys / values as
        `reset(env_ids)` for the envs that were flagged in its `reset_buf` (NaN if none was).  `packed`: the `fused_out`
        buffer of a graph-replayed step."""
        packed = self._fused_out if packed is None else packed
        if packed is None:
            raise RuntimeError("no compute_step(fuse_reset=True) has run yet")
        return _LazyStats(self._stat_keys, packed)

    def reset(self, env_ids: Sequence[int] | None = None) -> dict[str, torch.Tensor]:
        """Episode statistics of the envs being reset, then clear them (reference :190-211)."""
        extras = {}
        if self._term_names:
            extras = dict(zip(self._stat_keys, self._reset_stats(env_ids=env_ids, mask=None).unbind(0)))
        ids = slice(None) if env_ids is None else env_ids
        for term_cfg in self._class_term_cfgs:
            term_cfg.func.reset(env_ids=ids)
        return extras

    def reset_masked(self, mask: torch.Tensor) -> dict[str, torch.Tensor]:
        """Like `reset` but selects envs with a device-side bool mask: no `nonzero()`, no host sync."""
        if not self._term_names:
            return {}
        return dict(zip(self._stat_keys, self._reset_stats(env_ids=None, mask=mask).unbind(0)))

    # -- internals ---------------------------------------------------------------------------------
    def _prepare_terms(self):
        cfg_items = self.cfg.items() if isinstance(self.cfg, dict) else self.cfg.__dict__.items()
        for term_name, term_cfg in cfg_items:
            if term_cfg is None:
                continue
            if not isinstance(term_cfg, ConstraintTermCfg):
                raise TypeError(
                    f"Configuration for term '{term_name}' is not ConstraintTermCfg. Received: '{type(term_cfg)}'."
                )
            if not isinstance(term_cfg.max_p, (float, int)):
                raise TypeError(
                    f"Limit for term '{term_name}' must be float or int. Received: '{type(term_cfg.max_p)}'."
                )
            self._resolve_common_term_cfg(term_name, term_cfg, min_argc=1)
            self._term_names.append(term_name)
            self._term_cfgs.append(term_cfg)
            if isinstance(term_cfg.func, ManagerTermBase):
                self._class_term_cfgs.append(term_cfg)

    @staticmethod
    def _scalar_keys(term_cfg):
        return [k for k, v in term_cfg.params.items() if isinstance(v, (int, float))]

    def _param_values(self):
        """Current scalar params of every term (cheap per-step check that a cfg was not edited in place)."""
        return [[d[k] for k in keys] for d, keys in self._param_watch]

    def _specs(self):
        specs, generic = [], []
        for slot, (name, term_cfg) in enumerate(zip(self._term_names, self._term_cfgs)):
            spec_fn = getattr(term_cfg.func, "fused_spec", None)
            if spec_fn is not None:
                specs.append((name, spec_fn(self._env, **term_cfg.params), slot))
            else:  # a user's python term: evaluate it now, feed the tensor to the kernel
                holder = {"value": self._call_python_term(term_cfg)}
                generic.append((slot, term_cfg, holder))
                spec = TermSpec(L.OP_GENERIC, SourceRef(f"generic:{name}", lambda env, h=holder: h["value"]), None)
                specs.append((name, spec, slot))
        return specs, generic

    def _call_python_term(self, term_cfg) -> torch.Tensor:
        value = term_cfg.func(self._env, **term_cfg.params)
        if value.device != self._device:  # reference :42-43 moves stray tensors to the CaT device
            value = value.to(self._device)
        return value

    def _ensure_plan(self):
        if self._built is not None and self._param_values() == self._param_snapshot:
            for _, term_cfg, holder in self._generic:
                holder["value"] = self._call_python_term(term_cfg)
            self._built.refresh(self._env)
            return
        specs, self._generic = self._specs()
        built = _build(self._env, specs)
        if self._running_max is not None and built.n_cols != self._running_max.numel():
            raise RuntimeError("the number of constraint columns changed; create a new ConstraintManager")
        self._built = built
        self._param_watch = [(c.params, self._scalar_keys(c)) for c in self._term_cfgs]
        self._param_snapshot = self._param_values()
        if self._running_max is None:
            k = built.n_cols
            self._running_max = torch.zeros(k, dtype=torch.float, device=self._device)
            self._rm_init = torch.zeros(k, dtype=torch.int32, device=self._device)
            need = L.load().catb200_cat_workspace_bytes(self.num_envs, k)
            self._workspace = L.zeros_workspace(need, self._device)

    def _refresh_max_p(self):
        min_p = self.cat.min_p
        changed = False
        for i, term_cfg in enumerate(self._term_cfgs):  # re-read every step like the reference (:217)
            mp = term_cfg.max_p
            if mp != self._max_p_cache[i]:
                self._max_p_cache[i] = mp
                self._params.span[i] = mp - min_p
                self._span_host[i] = self._params.span[i]  # the fp32 value ctypes rounded to
                changed = True
        if changed:  # stream-ordered before the next launch / graph replay; no host synchronisation
            self._span_dev.copy_(self._span_host, non_blocking=True)

    def refresh_params(self) -> None:
        """Host-side part of a step for callers that replay the launches of `compute_step` from a CUDA graph: pick up
        `max_p` changes (curriculum, `set_term_cfg`) into the device-side table the captured kernels read."""
        self._refresh_max_p()

    def _episode_lengths(self) -> torch.Tensor:
        ep_len = self._env.episode_length_buf
        if ep_len.dtype != torch.int64 or not ep_len.is_contiguous():
            ep_len = ep_len.to(torch.int64).contiguous()
        L.require_cuda(ep_len, "episode_length_buf")
        return ep_len

    def _launch(self, raw_reward, reset_buf, fuse_reset=False, fused_out=None):
        if not self._term_names:
            self._cstr_prob_buf = torch.tensor([], device=self._device)
            self._computed = False
            return
        self._ensure_plan()
        self._refresh_max_p()
        ws = self._workspace
        if fuse_reset:
            dev = self._device
            if self._reset_ws is None:
                self._reset_ws = L.zeros_workspace(L.load().catb200_cat_reset_workspace_bytes(), dev)
            # a fresh output per call: the dict handed out by fused_reset_stats() may be kept by the caller (extras["log"])
            self._fused_out = fused_out if fused_out is not None else torch.empty(2 * len(self._term_names), dtype=torch.float, device=dev)
            ep_len = self._episode_lengths()
            L.check(
                L.load().catb200_cat_step_reset(
                    self._built.plan, self._params, self.num_envs,
                    self._running_max.data_ptr(), self._rm_init.data_ptr(),
                    self._stats[0].data_ptr(), self._stats[1].data_ptr(), self._cstr_prob_buf.data_ptr(),
                    L.ptr(raw_reward), L.ptr(reset_buf), self._reward_buf.data_ptr(), self._dones_buf.data_ptr(),
                    ws.data_ptr(), ws.numel() * 8, ep_len.data_ptr(), self._fused_out.data_ptr(),
                    self._reset_ws.data_ptr(), self._reset_ws.numel() * 8, L.stream(),
                ),
                "cat_step_reset",
            )  # fmt: skip
            self._computed = True
            self._probs_cache = None
            self._raw_cache = None
            return
        L.check(
            L.load().catb200_cat_step(
                self._built.plan, self._params, self.num_envs,
                self._running_max.data_ptr(), self._rm_init.data_ptr(),
                self._stats[0].data_ptr(), self._stats[1].data_ptr(), self._cstr_prob_buf.data_ptr(),
                L.ptr(raw_reward), L.ptr(reset_buf),
                self._reward_buf.data_ptr() if raw_reward is not None else None,
                self._dones_buf.data_ptr() if raw_reward is not None else None,
                ws.data_ptr(), ws.numel() * 8, L.stream(),
            ),
            "cat_step",
        )  # fmt: skip
        self._computed = True
        self._probs_cache = None
        self._raw_cache = None

    def _reset_stats(self, env_ids, mask):
        dev = self._device
        out = torch.empty(2 * len(self._term_names), dtype=torch.float, device=dev)
        ids_t, n_ids, mask_t = None, 0, None
        if env_ids is not None and not (isinstance(env_ids, slice) and env_ids == slice(None)):
            if isinstance(env_ids, torch.Tensor):
                ids_t = env_ids.to(device=dev, dtype=torch.int64).contiguous()
            else:
                ids_t = torch.as_tensor(np.asarray(list(env_ids), dtype=np.int64), device=dev)
            n_ids = ids_t.numel()
            if n_ids == 0:  # empty selection (a NULL id pointer would mean "all envs" to the C ABI): all-false mask
                ids_t, mask = None, torch.zeros(self.num_envs, dtype=torch.bool, device=dev)
        if ids_t is None and mask is not None:
            mask_t = mask.view(torch.uint8) if mask.dtype == torch.bool else (mask != 0).view(torch.uint8)
            mask_t = mask_t.contiguous()
        if self._reset_ws is None:
            self._reset_ws = L.zeros_workspace(L.load().catb200_cat_reset_workspace_bytes(), dev)
        ep_len = self._episode_lengths()
        L.check(
            L.load().catb200_cat_reset_stats(
                L.ptr(ids_t), n_ids, L.ptr(mask_t), ep_len.data_ptr(), self.num_envs, len(self._term_names),
                self._stats[0].data_ptr(), self._stats[1].data_ptr(), out.data_ptr(),
                self._reset_ws.data_ptr(), self._reset_ws.numel() * 8, L.stream(),
            ),
            "cat_reset_stats",
        )  # fmt: skip
        return out

    def _cols_of(self, name: str) -> tuple[int, int]:
        return self._built.term_cols[self._term_names.index(name)]

    def _probs_matrix(self) -> torch.Tensor:
        """[N, K] per-column probabilities of the last compute(), rebuilt on demand (debug / logging)."""
        if self._probs_cache is None:
            out = torch.empty((self.num_envs, self._built.n_cols), dtype=torch.float, device=self._device)
            L.check(
                L.load().catb200_cat_probs(
                    self._built.plan, self._params, self.num_envs, self._running_max.data_ptr(), out.data_ptr(),
                    self._raw_matrix().data_ptr(), L.stream(),
                ),
                "cat_probs",
            )  # fmt: skip
            self._probs_cache = out
        return self._probs_cache

    def _raw_matrix(self) -> torch.Tensor:
        """[N, K] raw constraint values, re-evaluated from the current state tensors (debug / logging)."""
        if self._raw_cache is None:
            out = torch.empty((self.num_envs, self._built.n_cols), dtype=torch.float, device=self._device)
            L.check(
                L.load().catb200_cat_eval_terms(self._built.plan, self.num_envs, out.data_ptr(), L.stream()),
                "cat_eval_terms",
            )
            self._raw_cache = out
        return self._raw_cache


ConstraintsManager = ConstraintManager
