"""ctypes binding of libcatb200.so (the C ABI declared in include/catb200.h).

This is the stub a reference maintainer would add: plain pointers and sizes,
no torch types in any signature.  torch is used only to own device memory and to name the current
CUDA stream.  There is deliberately no CPU or eager fallback: if the shared library is missing or a
kernel is asked to run without a CUDA device, the call raises.
"""

from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libcatb200.so")
INCLUDE_DIR = os.path.join(REPO_ROOT, "include")

MAX_SOURCES, MAX_TERMS, MAX_IDS, MAX_COLS = 16, 32, 32, 256
F32, U8 = 0, 1

# catb200_op
OP_GENERIC = 0
OP_ABS_MINUS = 1
OP_ABSDIFF_MINUS = 2
OP_ABSDIFF_MINUS_GATE_Y = 3
OP_ACTION_RATE = 4
OP_COMPONENT_GT = 5
OP_CONTACT_ANY = 6
OP_NORM2_MINUS = 7
OP_AIR_TIME = 8
OP_N_CONTACT = 9
OP_FORCE_PEAK_MINUS = 10
OP_LIMIT_MINUS = 11
OP_ABS_MINUS_GATE_STILL = 12

NO_SOURCE = 0xFF

# catb200_gae_variant
GAE_RLGAMES, GAE_SKRL = 0, 1

# catb200_prec: operand precision of the hidden-layer GEMMs
PREC_BF16, PREC_TF32 = 0, 1
PREC_NAMES = {"bf16": PREC_BF16, "tf32": PREC_TF32}


def default_precision() -> str:
    """`CATB200_GEMM_PREC=tf32|bf16` (default tf32: the reference's GPU numerics, scripts/clean_rl/train.py:86-87)."""
    name = os.environ.get("CATB200_GEMM_PREC", "tf32").lower()
    if name not in PREC_NAMES:
        raise ValueError(f"CATB200_GEMM_PREC must be tf32 or bf16, got {name!r}")
    return name


def operand_dtype(prec: int) -> torch.dtype:
    return torch.float32 if prec == PREC_TF32 else torch.bfloat16


class Source(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p),
        ("row_len", C.c_int32),
        ("row_stride", C.c_int32),
        ("dtype", C.c_int32),
        ("aux", C.c_int32),
        ("smem_off", C.c_int32),
        ("magic", C.c_uint32),
    ]


class Term(C.Structure):
    _fields_ = [
        ("op", C.c_uint8),
        ("n_cols", C.c_uint8),
        ("n_ids", C.c_uint8),
        ("src0", C.c_uint8),
        ("src1", C.c_uint8),
        ("src2", C.c_uint8),
        ("stat_slot", C.c_uint8),
        ("reserved", C.c_uint8),
        ("col_offset", C.c_uint16),
        ("reserved2", C.c_uint16),
        ("p0", C.c_float),
        ("p1", C.c_float),
        ("p2", C.c_float),
        ("ids", C.c_uint8 * MAX_IDS),
    ]


class Plan(C.Structure):
    _fields_ = [
        ("n_sources", C.c_int32),
        ("n_terms", C.c_int32),
        ("n_cols", C.c_int32),
        ("n_slots", C.c_int32),
        ("smem_bytes", C.c_int32),
        ("n_peaks", C.c_int32),
        ("smem_peak_off", C.c_int32),
        ("smem_ctile_off", C.c_int32),
        ("sources", Source * MAX_SOURCES),
        ("terms", Term * MAX_TERMS),
        ("col_term", C.c_uint8 * MAX_COLS),
        ("slot_col_begin", C.c_uint16 * (MAX_TERMS + 2)),
        ("peak_src", C.c_uint8 * 32),
        ("peak_body", C.c_uint8 * 32),
    ]


class CatParams(C.Structure):
    _fields_ = [
        ("tau", C.c_float),
        ("one_minus_tau", C.c_float),
        ("min_p", C.c_float),
        ("floor_max", C.c_float),
        ("span", C.c_float * MAX_TERMS),
        ("span_dev", C.c_void_p),
    ]


class MlpDims(C.Structure):
    _fields_ = [
        ("obs_dim", C.c_int32),
        ("act_dim", C.c_int32),
        ("h1", C.c_int32),
        ("h2", C.c_int32),
        ("h3", C.c_int32),
        ("obs_pad", C.c_int32),
        ("prec", C.c_int32),
    ]


class MlpLayout(C.Structure):
    _fields_ = [
        ("n_params", C.c_int64),
        ("w", (C.c_int64 * 4) * 2),
        ("b", (C.c_int64 * 4) * 2),
        ("logstd", C.c_int64),
        ("n_wc", C.c_int64),
        ("wc", (C.c_int64 * 3) * 2),
        ("wtc", (C.c_int64 * 3) * 2),
    ]


class CommandCfg(C.Structure):
    _fields_ = [
        ("lin_vel_x", C.c_float * 2),
        ("lin_vel_y", C.c_float * 2),
        ("ang_vel_z", C.c_float * 2),
        ("heading", C.c_float * 2),
        ("velocity_deadzone", C.c_float),
        ("heading_control_stiffness", C.c_float),
        ("rel_heading_envs", C.c_float),
        ("rel_standing_envs", C.c_float),
        ("p_step", C.c_float),
        ("heading_command", C.c_int32),
    ]


class ObsTerm(C.Structure):
    _fields_ = [
        ("src", C.c_void_p),
        ("row_stride", C.c_int32),
        ("n_cols", C.c_int32),
        ("n_min", C.c_float),
        ("n_max", C.c_float),
        ("noise_span", C.c_float),
        ("ids", C.c_uint8 * 32),
        ("scale", C.c_float * 32),
    ]


class ObsPlan(C.Structure):
    _fields_ = [
        ("n_terms", C.c_int32),
        ("n_cols", C.c_int32),
        ("terms", ObsTerm * 8),
        ("col_term", C.c_uint8 * 64),
        ("col_idx", C.c_uint8 * 64),
    ]


class PpoHparams(C.Structure):
    _fields_ = [
        ("clip_coef", C.c_float),
        ("ent_coef", C.c_float),
        ("vf_coef", C.c_float),
        ("norm_adv", C.c_int32),
        ("clip_vloss", C.c_int32),
    ]


NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-lineinfo",
    "-std=c++17",
    "-Xcompiler",
    "-fPIC",
    "-shared",
]


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(INCLUDE_DIR, "catb200.h"))
    return any(os.path.getmtime(d) > built for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into lib/libcatb200.so (nvcc cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        raise RuntimeError("nvcc not found: cannot build libcatb200.so")
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc, *NVCC_FLAGS, f"-I{INCLUDE_DIR}", f"-I{CSRC}", *sources(), "-o", LIB_PATH + ".tmp"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{proc.stdout}\n{proc.stderr}")
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    if verbose:
        print(proc.stderr)
    return LIB_PATH


_P = C.c_void_p
_I32 = C.c_int32
_I64 = C.c_int64
_F = C.c_float
_SZ = C.c_size_t

# name -> (restype, argtypes); one entry per symbol declared in include/catb200.h
SIGNATURES = {
    "catb200_version": (C.c_int, []),
    "catb200_launch_count": (C.c_uint64, []),
    "catb200_error_string": (C.c_char_p, [C.c_int]),
    "catb200_cat_plan_finalize": (C.c_int, [C.POINTER(Plan)]),
    "catb200_cat_workspace_bytes": (_SZ, [_I32, _I32]),
    "catb200_cat_step": (
        C.c_int,
        [C.POINTER(Plan), C.POINTER(CatParams), _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P],
    ),
    "catb200_cat_step_reset": (
        C.c_int,
        [C.POINTER(Plan), C.POINTER(CatParams), _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P, _P, _P, _SZ, _P],
    ),
    "catb200_cat_eval_terms": (C.c_int, [C.POINTER(Plan), _I32, _P, _P]),
    "catb200_cat_probs": (C.c_int, [C.POINTER(Plan), C.POINTER(CatParams), _I32, _P, _P, _P, _P]),
    "catb200_cat_reset_workspace_bytes": (_SZ, []),
    "catb200_cat_reset_stats": (C.c_int, [_P, _I32, _P, _P, _I32, _I32, _P, _P, _P, _P, _SZ, _P]),
    "catb200_rms_workspace_bytes": (_SZ, [_I32]),
    "catb200_rms_forward": (C.c_int, [_P, _I64, _I32, _P, _P, _P, _F, _I32, _P, _P, _I32, _I32, _P, _SZ, _P]),
    "catb200_rollout_append": (C.c_int, [_P, _P, _P, _I32, _P, _P, _P, _P]),
    "catb200_gae_workspace_bytes": (_SZ, []),
    "catb200_gae": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _F, _F, _P, _P, _P, _P, _P, _SZ, _P]),
    "catb200_gae_float_dones_workspace_bytes": (_SZ, []),
    "catb200_gae_float_dones": (C.c_int, [_I32, _P, _P, _P, _P, _P, _I32, _I32, _F, _F, _P, _P, _I32, _P, _SZ, _P]),
    "catb200_mlp_layout": (C.c_int, [C.POINTER(MlpDims), C.POINTER(MlpLayout)]),
    "catb200_mlp_cast_weights": (C.c_int, [C.POINTER(MlpDims), _P, _P, _P]),
    "catb200_obs_to_operand": (C.c_int, [C.POINTER(MlpDims), _P, _I64, _P, _P]),
    "catb200_mlp_workspace_bytes": (_SZ, [C.POINTER(MlpDims), _I32, _I32]),
    "catb200_mlp_act": (C.c_int, [C.POINTER(MlpDims), _P, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "catb200_ppo_minibatch_grad": (
        C.c_int,
        [C.POINTER(MlpDims), C.POINTER(PpoHparams), _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P],
    ),
    "catb200_adam_step": (
        C.c_int,
        [C.POINTER(MlpDims), _P, _P, _P, _P, _P, _P, _P, _F, _F, _F, _F, _F, _P, _P, _P],
    ),
    "catb200_ppo_minibatch_update": (
        C.c_int,
        [C.POINTER(MlpDims), C.POINTER(PpoHparams), _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ,
         _P, _P, _P, _P, _F, _F, _F, _F, _F, _P, _P, _P],
    ),
    "catb200_adam_apply": (C.c_int, [C.POINTER(MlpDims), _P, _P, _P, _P, _P, _P, _F, _F, _F, _F, _P, _P]),
    "catb200_peer_arena_bytes": (_SZ, [_I64]),
    "catb200_peer_alloc": (C.c_int, [_SZ, C.POINTER(_P), _P]),
    "catb200_peer_open": (C.c_int, [_P, C.POINTER(_P)]),
    "catb200_peer_close": (C.c_int, [_P]),
    "catb200_peer_free": (C.c_int, [_P]),
    "catb200_grad_allreduce_norm": (C.c_int, [_P, _I32, _I32, _I64, _I32, _P, _F, _F, _F, _F, _P, _P, _P, _P, _P, _P]),
    "catb200_ppo_minibatch_update_peer": (
        C.c_int,
        [C.POINTER(MlpDims), C.POINTER(PpoHparams), _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ,
         _P, _P, _P, _P, _F, _F, _F, _F, _P, _P, _P, _I32, _I32, _I32, _P, _P, _P, _P],
    ),
    "catb200_philox4x32_10": (C.c_int, [_P, _P, _P]),
    "catb200_random_permutation_host": (C.c_int, [_I64, C.c_uint64, C.c_uint64, _P]),
    "catb200_random_permutation": (C.c_int, [_I64, _P, _P, _P]),
    "catb200_bernoulli_mask": (C.c_int, [_P, _I32, _P, _P, _P, _P, _P]),
    "catb200_command_update": (C.c_int, [C.POINTER(CommandCfg), _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "catb200_push_select": (C.c_int, [_I32, _F, _P, _P, _P, _P, _P, _P, _P]),
    "catb200_obs_assemble": (C.c_int, [C.POINTER(ObsPlan), _I32, _P, _P, _P, _P]),
}

_lib = None


def load():
    """dlopen libcatb200.so and type every entry point.  Raises if the library is absent."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback for the CaT hot path."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def launch_count() -> int:
    return int(load().catb200_launch_count())


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().catb200_error_string(status).decode()
        raise RuntimeError(f"libcatb200 {what} failed: {msg} (status {status})")


def require_cuda(t: torch.Tensor, what: str = "tensor") -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{what} lives on {t.device}: the CaT hot path runs only as CUDA kernels on a B200 "
            "(no CPU fallback is provided)."
        )


def ptr(t: torch.Tensor | None):
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream() -> int:
    """cudaStream_t of torch's current stream on the current device (the fast C accessor when available)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def f32(x: float) -> float:
    """Round a python double to fp32 like torch does when a python scalar meets an fp32 tensor."""
    return float(torch.tensor(x, dtype=torch.float64).to(torch.float32).item())


def zeros_workspace(nbytes: int, device) -> torch.Tensor:
    """Zero-initialised scratch the kernels keep clean between calls."""
    return torch.zeros((int(nbytes) + 7) // 8, dtype=torch.int64, device=device)
