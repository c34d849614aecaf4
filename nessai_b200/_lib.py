"""ctypes binding of ``libnessai_b200.so`` (the C ABI in include/nessai_b200.h).

There is NO fallback: if the library is missing or a call fails this raises.
Loading the library does not create a CUDA context.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.environ.get("NB200_LIB", os.path.join(LIB_DIR, "libnessai_b200.so"))  # NB200_LIB: experiment builds
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]

_lib = None


class B200LibraryError(RuntimeError):
    pass


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree with nvcc for sm_100a."""
    srcs = [os.path.join(CSRC, "nessai_b200.cu")]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(INCLUDE, "nessai_b200.h")
    ]
    if (
        not force
        and os.path.exists(LIB_PATH)
        and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps)
    ):
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    # NB200_NVCC_FLAGS: extra -D switches of an experiment build (together with NB200_LIB)
    extra = os.environ.get("NB200_NVCC_FLAGS", "").split()
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-lcuda", "-o", LIB_PATH, *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise B200LibraryError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr, file=sys.stderr)
    return LIB_PATH


_SIGS = {
    "nb200_version": (C.c_int, []),
    "nb200_last_error": (C.c_char_p, []),
    "nb200_launch_count": (C.c_int64, []),
    "nb200_reset_launch_count": (None, []),
    "nb200_set_tensor_core_path": (C.c_int, [C.c_int]),
    "nb200_flow_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]),
    "nb200_flow_destroy": (C.c_int, [C.c_void_p]),
    "nb200_flow_set_program": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_double],
    ),
    "nb200_flow_inverse": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p],
    ),
    "nb200_flow_forward": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p],
    ),
    "nb200_flow_set_base_variance": (C.c_int, [C.c_void_p, C.c_double]),
    "nb200_sample_latent": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_uint64, C.c_double, C.c_void_p],
    ),
    "nb200_populate_draw": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_float, C.c_float,
         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "nb200_populate_accept": (
        C.c_int,
        [C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_uint64, C.c_uint64, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
         C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "nb200_populate_accept_x64": (
        C.c_int,
        [C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_uint64, C.c_uint64, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
         C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "nb200_reparam_tail": (
        C.c_int,
        [C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p],
    ),
    "nb200_sum_exp": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "nb200_coupling_transform": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int,
         C.c_int, C.c_int, C.c_void_p],
    ),
    "nb200_xchg_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p]),
    "nb200_xchg_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "nb200_xchg_close": (C.c_int, [C.c_void_p]),
    "nb200_xchg_destroy": (C.c_int, [C.c_void_p]),
    "nb200_xchg_allgather": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p],
    ),
    "nb200_trainer_create": (
        C.c_int,
        [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int],
    ),
    "nb200_trainer_destroy": (C.c_int, [C.c_void_p]),
    "nb200_trainer_set_param_mask": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nb200_trainer_set_itab": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "nb200_trainer_copy_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "nb200_train_epoch": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
         C.c_double, C.c_double, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "nb200_train_run": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
         C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
         C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "nb200_eval_loss": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
         C.c_void_p, C.c_void_p],
    ),
}

EXPORTED_SYMBOLS = tuple(_SIGS)


def load():
    """Load the library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200LibraryError(
            f"{LIB_PATH} not found: the nessai_b200 CUDA library has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
            "There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().nb200_last_error().decode(errors="replace")
        raise B200LibraryError(f"{what or 'nessai_b200 call'} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().nb200_launch_count())


def reset_launch_count() -> None:
    load().nb200_reset_launch_count()
