"""CPU-side checks of the C-ABI boundary: the library builds, loads, and exports every symbol that
include/catb200.h declares; plan validation works without a GPU (host-only function)."""

import ctypes
import os
import re

import pytest

from constraints_as_terminations_b200 import _lib as L


def _declared_symbols():
    text = open(os.path.join(L.INCLUDE_DIR, "catb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(catb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_typed(lib):
    declared = _declared_symbols()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in catb200.h but not exported"
        assert name in L.SIGNATURES, f"{name} has no ctypes signature in _lib.SIGNATURES"
    for name in L.SIGNATURES:
        assert name in declared, f"{name} bound in python but not declared in catb200.h"


def test_version_and_error_strings(lib):
    assert lib.catb200_version() == 100
    assert lib.catb200_error_string(0) == b"ok"
    assert b"workspace" in lib.catb200_error_string(-3)


def test_struct_sizes_match_header():
    # the kernels receive the plan by value as a launch parameter: must stay well below 4 KiB
    assert ctypes.sizeof(L.Source) == 32
    assert ctypes.sizeof(L.Term) == 56
    assert ctypes.sizeof(L.Plan) < 3072 + 64
    assert ctypes.sizeof(L.CatParams) == 16 + 4 * L.MAX_TERMS + 8


def test_ctypes_layout_matches_header_compiled_as_c(tmp_path):
    """include/catb200.h is plain C (gcc -std=c99 -pedantic) and every ctypes mirror has the header's size and field offsets."""
    import subprocess

    mirrors = {
        "catb200_source_t": L.Source,
        "catb200_term_t": L.Term,
        "catb200_plan_t": L.Plan,
        "catb200_cat_params_t": L.CatParams,
        "catb200_mlp_dims_t": L.MlpDims,
        "catb200_mlp_layout_t": L.MlpLayout,
        "catb200_ppo_hparams_t": L.PpoHparams,
        "catb200_command_cfg_t": L.CommandCfg,
        "catb200_obs_term_t": L.ObsTerm,
        "catb200_obs_plan_t": L.ObsPlan,
    }
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "catb200.h"', "int main(void) {"]
    for cname, cls in mirrors.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for field, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{field} %zu\\n", offsetof({cname}, {field}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", L.INCLUDE_DIR, str(src), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in mirrors.items():
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for field, _ in cls._fields_:
            assert int(out[f"{cname}.{field}"]) == getattr(cls, field).offset, f"{cname}.{field}"


def test_plan_finalize_validates(lib):
    plan = L.Plan()
    plan.n_sources, plan.n_terms = 1, 1
    plan.sources[0].row_len, plan.sources[0].row_stride, plan.sources[0].dtype = 12, 12, L.F32
    t = plan.terms[0]
    t.op, t.n_cols, t.n_ids, t.src0, t.src1, t.src2, t.stat_slot = L.OP_ABS_MINUS, 3, 3, 0, 0xFF, 0xFF, 0
    for k, v in enumerate([0, 5, 11]):
        t.ids[k] = v
    assert lib.catb200_cat_plan_finalize(plan) == 0
    assert plan.n_cols == 3 and plan.n_slots == 1 and plan.smem_bytes == 32 * 12 * 4 + 3 * 32 * 4 and plan.n_peaks == 0
    assert list(plan.slot_col_begin[:2]) == [0, 3]
    t.ids[2] = 12  # out of the 12-wide row
    assert lib.catb200_cat_plan_finalize(plan) == -1
    t.ids[2] = 11
    t.op = L.OP_ACTION_RATE  # needs a second source
    assert lib.catb200_cat_plan_finalize(plan) == -1
    t.op = 99
    assert lib.catb200_cat_plan_finalize(plan) == -2
    assert lib.catb200_cat_workspace_bytes(4096, 78) >= 4096 * 78 * 4


def test_ops_fail_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from constraints_as_terminations_b200 import ConstraintManager
    from constraints_as_terminations_b200 import synthetic_env as se

    env = se.SyntheticSolo12Env(8, device="cpu", pool=1)
    mgr = ConstraintManager(se.solo12_constraints_cfg(), env)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mgr.compute()
