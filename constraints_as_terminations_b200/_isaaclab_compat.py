"""Isaac Lab symbols the CaT hot path touches, with local shims when Isaac Lab is absent.

The reference subclasses / imports exactly five Isaac Lab names on this path
(reference `exts/cat_envs/cat_envs/tasks/utils/cat/constraint_manager.py:15`,
`manager_constraint_cfg.py:14-15`, `constraints.py:17`):

    isaaclab.managers.manager_base.ManagerBase / ManagerTermBase
    isaaclab.managers.manager_term_cfg.ManagerTermBaseCfg
    isaaclab.utils.configclass
    isaaclab.managers.SceneEntityCfg

When Isaac Lab is importable the real classes are used, so the Solo12 env cfg
drops in unchanged.  Otherwise the shims below provide the same members the
reference uses (and nothing more), which is what the synthetic Solo12 env and
the tests run against.
"""

from __future__ import annotations

import copy
import dataclasses
import inspect
import re
from collections.abc import Callable, Sequence
from dataclasses import MISSING, field
from typing import Any

try:  # pragma: no cover - Isaac Lab is not installed in the build container
    from isaaclab.managers import SceneEntityCfg  # type: ignore
    from isaaclab.managers.manager_base import ManagerBase, ManagerTermBase  # type: ignore
    from isaaclab.managers.manager_term_cfg import ManagerTermBaseCfg  # type: ignore
    from isaaclab.utils import configclass  # type: ignore

    HAVE_ISAACLAB = True
except Exception:  # noqa: BLE001 - any import failure means "use the shims"
    HAVE_ISAACLAB = False

    def configclass(cls=None, **kwargs):
        """Minimal stand-in for `isaaclab.utils.configclass`.

        Like the original it turns un-annotated class attributes into dataclass
        fields and deep-copies mutable defaults per instance, which is what the
        reference task cfg relies on (`cat_flat_env_cfg.py:259-355` assigns term
        cfg instances as bare class attributes).
        """

        def wrap(c):
            ann = dict(c.__dict__.get("__annotations__", {}))
            for name, value in list(c.__dict__.items()):
                if name.startswith("__") or name in ann:
                    continue
                if isinstance(value, (staticmethod, classmethod, property)) or inspect.isfunction(value):
                    continue
                if inspect.isclass(value):
                    continue
                ann[name] = type(value) if value is not None else Any
            c.__annotations__ = ann
            for name in ann:
                if name not in c.__dict__:
                    continue
                value = c.__dict__[name]
                if isinstance(value, dataclasses.Field):
                    continue
                if not isinstance(value, (int, float, str, bool, type(None), tuple, frozenset)) and value is not MISSING:
                    setattr(c, name, field(default_factory=lambda v=value: copy.deepcopy(v)))
            kwargs.setdefault("kw_only", True)  # fields without defaults (MISSING) may follow defaulted ones
            c = dataclasses.dataclass(c, **kwargs)

            def to_dict(self):
                return {f.name: getattr(self, f.name) for f in dataclasses.fields(self)}

            if not hasattr(c, "to_dict"):
                c.to_dict = to_dict
            return c

        return wrap if cls is None else wrap(cls)

    @configclass
    class ManagerTermBaseCfg:
        func: Callable = MISSING
        params: dict = field(default_factory=dict)

    class SceneEntityCfg:
        """Stand-in for `isaaclab.managers.SceneEntityCfg` (name + joint/body selectors)."""

        def __init__(
            self,
            name: str,
            joint_names: str | Sequence[str] | None = None,
            joint_ids: Sequence[int] | slice = slice(None),
            body_names: str | Sequence[str] | None = None,
            body_ids: Sequence[int] | slice = slice(None),
            preserve_order: bool = False,
        ):
            self.name = name
            self.joint_names = joint_names
            self.joint_ids = joint_ids
            self.body_names = body_names
            self.body_ids = body_ids
            self.preserve_order = preserve_order

        @staticmethod
        def _match(patterns, names, preserve_order):
            if isinstance(patterns, str):
                patterns = [patterns]
            if preserve_order:
                ids = []
                for p in patterns:
                    ids += [i for i, n in enumerate(names) if re.fullmatch(p, n) and i not in ids]
            else:
                ids = [i for i, n in enumerate(names) if any(re.fullmatch(p, n) for p in patterns)]
            if not ids:
                raise ValueError(f"No match for {patterns} in {names}")
            return ids

        def resolve(self, scene):
            entity = scene[self.name]
            if self.joint_names is not None:
                names = list(entity.joint_names)
                ids = self._match(self.joint_names, names, self.preserve_order)
                # Isaac Lab collapses "everything selected" to slice(None)
                self.joint_ids = slice(None) if len(ids) == len(names) and ids == sorted(ids) else ids
            if self.body_names is not None:
                names = list(entity.body_names)
                ids = self._match(self.body_names, names, self.preserve_order)
                self.body_ids = slice(None) if len(ids) == len(names) and ids == sorted(ids) else ids

        def __repr__(self):
            return (
                f"SceneEntityCfg(name={self.name!r}, joint_names={self.joint_names!r}, "
                f"body_names={self.body_names!r})"
            )

    class ManagerTermBase:
        """Stand-in for class-based manager terms (`reset(env_ids)` + `__call__`)."""

        def __init__(self, cfg, env):
            self.cfg = cfg
            self._env = env

        def reset(self, env_ids=None):
            pass

        def __call__(self, *args, **kwargs):
            raise NotImplementedError

    class ManagerBase:
        """Stand-in for `isaaclab.managers.ManagerBase` (the 4 members CaT uses)."""

        def __init__(self, cfg, env):
            self.cfg = copy.deepcopy(cfg)
            self._env = env
            self._prepare_terms()

        @property
        def num_envs(self) -> int:
            return self._env.num_envs

        @property
        def device(self):
            return self._env.device

        def _resolve_common_term_cfg(self, term_name: str, term_cfg, min_argc: int = 1):
            if not isinstance(term_cfg, ManagerTermBaseCfg):
                raise TypeError(f"Configuration for the term '{term_name}' is not of type ManagerTermBaseCfg.")
            for value in term_cfg.params.values():
                if isinstance(value, SceneEntityCfg):
                    value.resolve(self._env.scene)
            if inspect.isclass(term_cfg.func):
                if not issubclass(term_cfg.func, ManagerTermBase):
                    raise TypeError(f"Term '{term_name}' class must inherit ManagerTermBase.")
                term_cfg.func = term_cfg.func(cfg=term_cfg, env=self._env)
            if not callable(term_cfg.func):
                raise AttributeError(f"The term '{term_name}' is not callable. Received: {term_cfg.func}")
            # the static check Isaac Lab performs: the function's parameters beyond the first `min_argc` must be exactly
            # the keys of `params` (plus whatever has a default)
            func_static = term_cfg.func.__call__ if isinstance(term_cfg.func, ManagerTermBase) else term_cfg.func
            sig = inspect.signature(func_static).parameters
            with_defaults = [a for a in sig if sig[a].default is not inspect.Parameter.empty]
            without_defaults = [a for a in sig if sig[a].default is inspect.Parameter.empty]
            args = without_defaults + with_defaults
            term_params = list(term_cfg.params.keys())
            if len(args) > min_argc and set(args[min_argc:]) != set(term_params + with_defaults):
                raise ValueError(
                    f"The term '{term_name}' expects mandatory parameters: {without_defaults[min_argc:]}"
                    f" and optional parameters: {with_defaults}, but received: {term_params}."
                )
