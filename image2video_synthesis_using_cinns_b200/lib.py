"""Build + ctypes binding of the C-ABI library (include/i2v_b200.h).

The shared object is built IN-TREE (``libi2v_b200.so`` next to this file) with plain ``nvcc`` for
sm_100a only.  There is no CPU fallback and no alternative backend: if the library cannot be loaded
every entry point of the product path raises.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libi2v_b200.so")
BUILD_DIR = os.path.join(HERE, "csrc", "build")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]

_lib = None


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return exe


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    paths.append(os.path.join(os.path.dirname(HERE), "include", "i2v_b200.h"))
    return max(os.path.getmtime(p) for p in paths if os.path.exists(p))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu -> libi2v_b200.so (sm_100a).  Cross-compiles without a GPU."""
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _deps_mtime():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(BUILD_DIR, exist_ok=True)
    srcs = sources()
    hdr_mtime = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)
                    if f.endswith((".cuh", ".h")))
    hdr_mtime = max(hdr_mtime, os.path.getmtime(os.path.join(os.path.dirname(HERE), "include", "i2v_b200.h")))

    def compile_one(src):
        obj = os.path.join(BUILD_DIR, os.path.basename(src)[:-3] + ".o")
        if (not force and os.path.exists(obj)
                and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_mtime)):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(" ".join(cmd))
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


_c = ctypes
_P, _I, _F, _SZ, _I64 = _c.c_void_p, _c.c_int, _c.c_float, _c.c_size_t, _c.c_int64

# name -> (restype, argtypes): every symbol include/i2v_b200.h declares
SIGNATURES = {
    "i2v_abi_version": (_I, []),
    "i2v_last_error": (_c.c_char_p, []),
    "i2v_launch_count": (_c.c_longlong, []),
    "i2v_prof_enable": (None, [_I]),
    "i2v_prof_is_enabled": (_I, []),
    "i2v_prof_collect": (_I, [_P, _P, _P, _P]),
    "i2v_prof_dump_path": (None, [_c.c_char_p]),
    "i2v_set_option": (_I, [_c.c_char_p, _c.c_double]),
    "i2v_flow_create": (_P, [_I, _I, _I, _I, _I, _P]),
    "i2v_flow_set_tensor": (_I, [_P, _c.c_char_p, _P, _SZ]),
    "i2v_flow_workspace_bytes": (_SZ, [_P, _I]),
    "i2v_flow_reverse": (_I, [_P, _P, _P, _P, _I, _P, _SZ, _P]),
    "i2v_flow_forward": (_I, [_P, _P, _P, _P, _P, _I, _P, _SZ, _P]),
    "i2v_flow_destroy": (None, [_P]),
    "i2v_embedder_create": (_P, [_I, _I]),
    "i2v_embedder_set_tensor": (_I, [_P, _c.c_char_p, _P, _SZ]),
    "i2v_embedder_set_scalar": (_I, [_P, _c.c_char_p, _c.c_double]),
    "i2v_embedder_workspace_bytes": (_SZ, [_P, _I, _I, _I]),
    "i2v_embedder_forward": (_I, [_P, _P, _P, _I, _I, _I, _P, _SZ, _P]),
    "i2v_embedder_destroy": (None, [_P]),
    "i2v_decoder_create": (_P, [_I, _I, _P, _P, _I]),
    "i2v_decoder_set_tensor": (_I, [_P, _c.c_char_p, _P, _SZ]),
    "i2v_decoder_set_scalar": (_I, [_P, _c.c_char_p, _c.c_double]),
    "i2v_decoder_workspace_bytes": (_SZ, [_P, _I, _I, _I]),
    "i2v_decoder_forward": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _SZ, _P]),
    "i2v_decoder_destroy": (None, [_P]),
    "i2v_encoder3d_create": (_P, [_P, _P, _P, _I]),
    "i2v_encoder3d_set_tensor": (_I, [_P, _c.c_char_p, _P, _SZ]),
    "i2v_encoder3d_set_scalar": (_I, [_P, _c.c_char_p, _c.c_double]),
    "i2v_encoder3d_workspace_bytes": (_SZ, [_P, _I, _I, _I, _I]),
    "i2v_encoder3d_forward": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _SZ, _P]),
    "i2v_encoder3d_destroy": (None, [_P]),
    "i2v_op_conv": (_I, [_P] * 5 + [_I] * 21 + [_P]),
    "i2v_op_spade_conv3": (_I, [_P] * 5 + [_F, _I, _I, _I, _I, _P]),
    "i2v_op_conv_tc": (_I, [_P] * 5 + [_I] * 17 + [_F, _F, _P, _SZ, _P]),
    "i2v_op_conv_tc_phase": (_I, [_P] * 4 + [_I] * 9 + [_F, _F, _P, _SZ, _P]),
    "i2v_op_conv_tc_side": (_I, [_P] * 6 + [_I] * 12 + [_F, _F, _P, _SZ, _P]),
    "i2v_debug_conv_tc_timestamps": (_I, [_P, _I]),
    "i2v_debug_flow_timestamps": (_I, [_P]),
    "i2v_op_channel_stats": (_I, [_P, _P, _I, _I64, _I, _P]),
    "i2v_op_norm_coeffs": (_I, [_P, _P, _I, _I, _I64, _I, _F, _P, _P, _P, _P]),
    "i2v_op_modulate": (_I, [_P] * 6 + [_I] * 9 + [_P]),
    "i2v_op_modulate_split": (_I, [_P] * 5 + [_I] * 9 + [_F, _P, _P, _P, _P]),
    "i2v_op_linear": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "i2v_op_resize_bilinear": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "i2v_op_maxpool3x3s2": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "i2v_op_preprocess_u8": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "i2v_op_frames_max": (_I, [_P, _P, _I64, _P]),
    "i2v_op_frames_to_u8": (_I, [_P, _P, _P, _I, _I, _I, _I, _I64, _I64, _I64, _P]),
}


def load():
    """dlopen the in-tree library and type every exported symbol.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU / PyTorch fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError here = header and library disagree
        fn.restype, fn.argtypes = res, args
    if lib.i2v_abi_version() != 1:
        raise RuntimeError("libi2v_b200.so ABI version mismatch")
    _lib = lib
    return lib


def set_option(name: str, value: float):
    """Process-wide tuning switch (i2v_set_option): A/B measurements only."""
    check(load().i2v_set_option(name.encode(), float(value)), f"set_option({name})")


def check(rc: int, what: str = "i2v"):
    if rc != 0:
        msg = load().i2v_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")
