"""ORACLE tooling: import the *real* reference modules.

The reference is pure Python on top of Isaac Lab, which is not installed, so its CaT modules are loaded
from where they lie under a synthetic package with `sys.modules` stubs for the five Isaac Lab names and
`prettytable`.  `cleanrl/ppo.py` needs no stubs.

Where the files come from, in this order: $CAT_REFERENCE_ROOT, /root/reference (the build container),
oracle/_ref/ref_hotpath.tar.gz (a git-ignored archive that `oracle/build_ref.py` -- run by `__graft_entry__.build()`
whenever /root/reference is present -- packs from the handful of reference files on the hot path, so that the unmodified
reference also travels to the GPU box; nothing of it is committed; it is unpacked into a temporary directory here).  Used by
`oracle/make_golden.py` (fixtures), `oracle/ref_runner.py` (the reference arm of bench.py) and by tests that skip
when no reference tree is found.
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types

_REL_UTILS = "exts/cat_envs/cat_envs/tasks/utils"
LOCAL_ARCHIVE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "ref_hotpath.tar.gz")


def _unpack_local() -> str | None:
    if not os.path.isfile(LOCAL_ARCHIVE):
        return None
    import atexit
    import shutil
    import tarfile
    import tempfile

    root = tempfile.mkdtemp(prefix="cat_ref_")
    atexit.register(shutil.rmtree, root, ignore_errors=True)
    with tarfile.open(LOCAL_ARCHIVE, "r:gz") as tar:
        tar.extractall(root)
    return root


def _find_root() -> str:
    for root in (os.environ.get("CAT_REFERENCE_ROOT"), "/root/reference"):
        if root and os.path.isfile(os.path.join(root, _REL_UTILS, "cat/constraint_manager.py")):
            return root
    return _unpack_local() or os.environ.get("CAT_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find_root()
_UTILS = os.path.join(REFERENCE_ROOT, _REL_UTILS)


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_UTILS, "cat/constraint_manager.py"))


def _install_stubs():
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if repo not in sys.path:
        sys.path.insert(0, repo)
    from constraints_as_terminations_b200 import _isaaclab_compat as shim

    if shim.HAVE_ISAACLAB:
        return

    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("isaaclab")
    mod("isaaclab.utils", configclass=shim.configclass)
    mod("isaaclab.managers", SceneEntityCfg=shim.SceneEntityCfg)
    mod("isaaclab.managers.manager_base", ManagerBase=shim.ManagerBase, ManagerTermBase=shim.ManagerTermBase)
    mod("isaaclab.managers.manager_term_cfg", ManagerTermBaseCfg=shim.ManagerTermBaseCfg)
    if "prettytable" not in sys.modules:
        try:
            import prettytable  # noqa: F401
        except ImportError:

            class PrettyTable:
                def __init__(self):
                    self.title = ""
                    self.field_names = []
                    self.align = {}
                    self._rows = []

                def add_row(self, row):
                    self._rows.append(row)

                def get_string(self):
                    lines = [str(self.title), " | ".join(map(str, self.field_names))]
                    lines += [" | ".join(map(str, r)) for r in self._rows]
                    return "\n".join(lines)

            mod("prettytable", PrettyTable=PrettyTable)


def _load(pkg_name: str, sub: str, files: list[str]):
    pkg = types.ModuleType(pkg_name)
    pkg.__path__ = [os.path.join(_UTILS, sub)]
    sys.modules[pkg_name] = pkg
    out = {}
    for f in files:
        full = f"{pkg_name}.{f}"
        spec = importlib.util.spec_from_file_location(full, os.path.join(_UTILS, sub, f + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[full] = m
        spec.loader.exec_module(m)
        setattr(pkg, f, m)
        out[f] = m
    return out


_CACHE: dict = {}


def load_reference_cat():
    """Returns dict with the reference modules: manager_constraint_cfg, constraint_manager, constraints, curriculums."""
    if "cat" not in _CACHE:
        if not reference_available():
            raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
        _install_stubs()
        _CACHE["cat"] = _load(
            "_ref_cat", "cat", ["manager_constraint_cfg", "constraint_manager", "constraints", "curriculums"]
        )
    return _CACHE["cat"]


def load_reference_ppo():
    """Returns the reference `cleanrl/ppo.py` module (imports only os, time, numpy, torch)."""
    if "ppo" not in _CACHE:
        if not reference_available():
            raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
        _CACHE["ppo"] = _load("_ref_cleanrl", "cleanrl", ["ppo"])["ppo"]
    return _CACHE["ppo"]
