"""ctypes binding of libtbi_sm100.so (the C ABI in include/tbi_sm100.h) + the in-tree nvcc build.

There is no CPU fallback and no alternative backend: if the library is missing or the device is
not sm_100, every op raises.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_CSRC = _PKG / "csrc"
LIB_PATH = _PKG / "libtbi_sm100.so"
HASH_PATH = _PKG / "libtbi_sm100.so.hash"
SOURCES = ["c_api.cu", "tapgemm_simt.cu", "tapgemm_tc.cu", "tapgemm_halo.cu", "tapgemm_xpack.cu", "tapwgrad_tc.cu", "tapwgrad_tc2.cu", "tapwgrad_small.cu", "direct_small.cu", "splitatt_fused.cu", "bandwidth.cu", "variant_b.cu", "vit.cu", "data.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

F32, BF16 = 0, 1
ACT_NONE, ACT_ELU, ACT_LRELU, ACT_RELU = 0, 1, 2, 3
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05 = 0, 1, 2
MAX_TAPS = 16


def _source_hash() -> str:
    """content hash of everything the library is built from (mtimes do not survive a snapshot copy)"""
    h = hashlib.sha256()
    deps = sorted(_CSRC.glob("*.cu")) + sorted(_CSRC.glob("*.cuh")) + [_PKG.parent / "include" / "tbi_sm100.h"]
    for d in deps:
        h.update(d.name.encode()); h.update(d.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_stale() -> bool:
    return not (LIB_PATH.exists() and HASH_PATH.exists() and HASH_PATH.read_text().strip() == _source_hash())


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu into libtbi_sm100.so for sm_100a (nvcc cross-compiles without a GPU).
    Rebuilds only when the sources' content hash differs from the one recorded beside the .so.
    Safe under torchrun: one process builds (exclusive flock), into a temporary file that is renamed over the library, so no
    rank can dlopen a half-written .so; the others wait on the lock and then find the library fresh."""
    import fcntl
    if not force and not is_stale():
        return LIB_PATH
    with open(str(LIB_PATH) + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():         # another process built it while we waited
                return LIB_PATH
            nvcc = os.environ.get("NVCC", "nvcc")
            objdir = _PKG / "build"
            objdir.mkdir(exist_ok=True)
            hdr = hashlib.sha256()
            for d in sorted(_CSRC.glob("*.cuh")) + [_PKG.parent / "include" / "tbi_sm100.h"]:
                hdr.update(d.name.encode()); hdr.update(d.read_bytes())
            hdr.update(" ".join(NVCC_FLAGS).encode())
            cflags = [f for f in NVCC_FLAGS if f != "-shared"]

            def compile_one(name):
                """one translation unit -> object, reused while the file and every header are unchanged"""
                src, obj = _CSRC / name, objdir / (name + ".o")
                tag = hashlib.sha256(hdr.digest() + src.read_bytes()).hexdigest()
                stamp = objdir / (name + ".hash")
                if obj.exists() and stamp.exists() and stamp.read_text().strip() == tag:
                    return
                cmd = [nvcc, *cflags, "-c", "-o", str(obj), str(src)]
                if verbose:
                    print(" ".join(cmd), flush=True)
                subprocess.run(cmd, check=True, cwd=str(_CSRC))
                stamp.write_text(tag + "\n")

            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
                list(pool.map(compile_one, SOURCES))
            tmp = LIB_PATH.with_name(f".{LIB_PATH.name}.{os.getpid()}.tmp")
            cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(tmp), *[str(objdir / (n + ".o")) for n in SOURCES]]
            if verbose:
                print(" ".join(cmd), flush=True)
            try:
                subprocess.run(cmd, check=True, cwd=str(_CSRC))
                os.replace(tmp, LIB_PATH)
            finally:
                if tmp.exists():
                    tmp.unlink()
            HASH_PATH.write_text(_source_hash() + "\n")
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


class View(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("cstride", C.c_int32), ("coff", C.c_int32)]


class Epilogue(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("drop_keep", C.c_void_p), ("act", C.c_int32),
                ("residual", View), ("dact", C.c_int32), ("dact_ref", View), ("dact_keep", C.c_void_p),
                ("out", View), ("out_f32", C.c_int32), ("out_stride", C.c_int32), ("out_off_y", C.c_int32),
                ("out_off_x", C.c_int32), ("split_c", C.c_int32), ("out2", View), ("residual2", View)]


class TapGemm(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("impl", C.c_int32), ("n", C.c_int32), ("gh", C.c_int32), ("gw", C.c_int32),
                ("groups", C.c_int32), ("cin_g", C.c_int32), ("cout_g", C.c_int32), ("src", View * 2),
                ("in_stride", C.c_int32), ("ntaps", C.c_int32), ("dy", C.c_int32 * MAX_TAPS),
                ("dx", C.c_int32 * MAX_TAPS), ("w", C.c_void_p), ("nphase", C.c_int32),
                ("ph_dy", (C.c_int32 * 4) * 4), ("ph_dx", (C.c_int32 * 4) * 4), ("ph_off_y", C.c_int32 * 4),
                ("ph_off_x", C.c_int32 * 4), ("epi", Epilogue)]


class TapWgrad(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("impl", C.c_int32), ("n", C.c_int32), ("gh", C.c_int32), ("gw", C.c_int32),
                ("groups", C.c_int32), ("cin_g", C.c_int32), ("cout_g", C.c_int32), ("a_src", View * 2),
                ("b_src", View), ("a_stride", C.c_int32), ("b_stride", C.c_int32), ("ntaps", C.c_int32),
                ("a_dy", C.c_int32 * MAX_TAPS), ("a_dx", C.c_int32 * MAX_TAPS), ("b_dy", C.c_int32 * MAX_TAPS),
                ("b_dx", C.c_int32 * MAX_TAPS), ("dw", C.c_void_p), ("tap_stride", C.c_int64),
                ("ci_stride", C.c_int64), ("co_stride", C.c_int64), ("dbias", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64)]


class PrepItem(C.Structure):
    _fields_ = [("src", C.c_void_p), ("out", C.c_void_p), ("gamma", C.c_void_p), ("var", C.c_void_p),
                ("ntaps", C.c_int32), ("A", C.c_int32), ("B", C.c_int32), ("co_is_a", C.c_int32), ("co_base", C.c_int32),
                ("tile_begin", C.c_int32), ("tiles_a", C.c_int32), ("tiles_b", C.c_int32),
                ("src_tap", C.c_int64), ("src_a", C.c_int64), ("out_a", C.c_int64), ("out_b", C.c_int64),
                ("src_tap_index", C.c_int32 * MAX_TAPS), ("out_tap", C.c_int64 * MAX_TAPS)]


class FoldItem(C.Structure):
    _fields_ = [("c", C.c_int32), ("pad_", C.c_int32), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("mean", C.c_void_p),
                ("var", C.c_void_p), ("bias", C.c_void_p), ("scale", C.c_void_p), ("fbias", C.c_void_p)]


class SplitAtt(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("kpaths", C.c_int32),
                ("radix", C.c_int32), ("c", C.c_int32), ("act", C.c_int32), ("bn_eps", C.c_float),
                ("w1", C.c_void_p), ("b1", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("mean", C.c_void_p), ("var", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
                ("gap", C.c_void_p), ("h1", C.c_void_p), ("att", C.c_void_p)]


_VP, _I, _I64, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float
_PV, _PE = C.POINTER(View), C.POINTER(Epilogue)

# name -> (restype, argtypes); every symbol include/tbi_sm100.h declares
SIGNATURES = {
    "tbi_version": (_I, []),
    "tbi_last_error": (C.c_char_p, []),
    "tbi_device_ok": (_I, []),
    "tbi_fallback_stats": (_I, [C.POINTER(C.c_int64), C.POINTER(C.c_int64), _I]),
    "tbi_last_fallback": (C.c_char_p, []),
    "tbi_tapgemm_run": (_I, [C.POINTER(TapGemm), _VP]),
    "tbi_tapwgrad_run": (_I, [C.POINTER(TapWgrad), _VP]),
    "tbi_workspace_bytes": (_I64, [C.POINTER(TapWgrad)]),
    "tbi_conv2d_fwd": (_I, [_I, _I, _I, _I, _I, _I, _I, _I, _PV, _PV, _I, _VP, _PE, _VP]),
    "tbi_conv2d_dgrad": (_I, [_I, _I, _I, _I, _I, _I, _I, _I, _PV, _I, _VP, _PE, _VP]),
    "tbi_conv2d_wgrad": (_I, [_I, _I, _I, _I, _I, _I, _I, _I, _PV, _PV, _PV, _VP, _VP, _VP, _I64, _VP]),
    "tbi_conv2d_transpose_s2_fwd": (_I, [_I, _I, _I, _I, _I, _I, _PV, _PV, _I, _VP, _PE, _VP]),
    "tbi_conv2d_transpose_s2_dgrad": (_I, [_I, _I, _I, _I, _I, _I, _PV, _I, _VP, _PE, _VP]),
    "tbi_conv2d_transpose_s2_wgrad": (_I, [_I, _I, _I, _I, _I, _I, _PV, _PV, _PV, _I, _VP, _VP, _VP, _I64, _VP]),
    "tbi_pack_conv_weights": (_I, [_I, _I, _I, _I, _I, _I, _VP, _VP, _VP, _VP]),
    "tbi_pack_convt_weights": (_I, [_I, _I, _I, _I, _I, _I, _VP, _VP, _VP, _VP]),
    "tbi_convt_gather_dz": (_I, [_I, _I, _I, _I, _I, _I, _PV, _PV, _VP]),
    "tbi_convt_scatter_y": (_I, [_I, _I, _I, _I, _I, _I, _PV, _VP, _PV, _VP]),
    "tbi_convt_phase_taps": (_I, [_I, _I, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "tbi_prepare_run": (_I, [_I, _VP, _I, _I, _F, _VP]),
    "tbi_bn_fold_multi": (_I, [_VP, _I, _F, _VP]),
    "tbi_bn_fold": (_I, [_I, _VP, _VP, _VP, _VP, _VP, _F, _VP, _VP, _VP]),
    "tbi_bn_param_grad": (_I, [_I, _I64, _I64, _I64, _I64, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _F, _VP, _VP, _VP]),
    "tbi_avgpool2x2_fwd": (_I, [_I, _I, _I, _I, _PV, _PV, _VP]),
    "tbi_avgpool2x2_bwd": (_I, [_I, _I, _I, _I, _PV, _PV, _I, _I, _PV, _VP]),
    "tbi_split_attention_fwd": (_I, [C.POINTER(SplitAtt), _PV, _PV, _VP]),
    "tbi_split_attention_bwd": (_I, [C.POINTER(SplitAtt), _PV, _PV, _PV, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "tbi_splitatt_gap": (_I, [C.POINTER(SplitAtt), _PV, _VP]),
    "tbi_splitatt_combine": (_I, [C.POINTER(SplitAtt), _PV, _PV, _VP]),
    "tbi_softmax_loss_fwd_bwd": (_I, [_I, _I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _I, _VP]),
    "tbi_softmax_loss_fwd_bwd_taps": (_I, [_I, _I, _I, _I, _I, _PV, _VP, _VP, _VP, _VP, _VP, _VP, _I, _VP]),
    "tbi_act_bwd": (_I, [_I, _I64, _I, _PV, _PV, _VP, _PV, _VP]),
    "tbi_accumulate": (_I, [_I, _I64, _PV, _PV, _VP]),
    "tbi_colsum": (_I, [_I, _I64, _PV, _VP, _VP]),
    "tbi_conv_dense_expand": (_I, [_I, _I, _I, _I]),
    "tbi_conv_packed_elems": (_I64, [_I, _I, _I, _I, _I]),
    "tbi_conv2d_wgrad_workspace": (_I64, [_I, _I, _I, _I, _I]),
    "tbi_layernorm_c_fwd": (_I, [_I, _I64, _I, _PV, _VP, _VP, _F, _I, _PV, _VP]),
    "tbi_layernorm_c_bwd": (_I, [_I, _I64, _I, _PV, _PV, _PV, _VP, _F, _I, _PV, _VP, _VP, _VP]),
    "tbi_splitatt_shared_fwd": (_I, [_I, _I, _I, _I, _I, _I, _I, _PV, _PV, _VP, _VP, _VP, _VP, _F, _I, _VP, _VP, _VP, _I, _VP]),
    "tbi_splitatt_shared_bwd": (_I, [_I, _I, _I, _I, _I, _I, _I, _PV, _PV, _PV, _VP, _VP, _VP, _VP, _F, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _VP]),
    "tbi_attention_fwd": (_I, [_I, _I, _I, _I, _I, _F, _VP, _VP, _VP, _VP, _VP, _VP]),
    "tbi_attention_bwd": (_I, [_I, _I, _I, _I, _I, _F, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "tbi_gelu_fwd": (_I, [_I, _I64, _VP, _VP, _VP]),
    "tbi_gelu_bwd": (_I, [_I, _I64, _VP, _VP, _VP, _VP]),
    "tbi_softmax_cce_fwd_bwd": (_I, [_I64, _I, _F, _F, _VP, _VP, _VP, _VP, _VP, _VP]),
    "tbi_label2vec": (_I, [_I64, _I, _VP, _VP, _VP]),
    "tbi_data_aug": (_I, [_I, _I, _I, _I, _VP, _VP, _VP, C.c_uint64, _VP, _VP, _VP]),
    "tbi_softmax_prob_maps": (_I, [_I64, _I, _VP, _VP, _VP, _VP, _VP]),
    "tbi_apply_brain_mask": (_I, [_I64, _I, _I, _VP, _VP, _VP]),
    "tbi_dropout_mask": (_I, [_VP, _I64, C.c_uint64, _VP, _VP]),
    "tbi_cast": (_I, [_I, _I, _I64, _VP, _VP, _VP]),
    "tbi_adam_multi": (_I, [_I64, _VP, _VP, _VP, _VP, _VP, _F, _F, _F, _F, _F, _VP]),
    "tbi_adam_advance": (_I, [_VP, _VP]),
    "tbi_adam_multi_dev": (_I, [_I64, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _F, _F, _F, _VP]),
    "tbi_sumsq": (_I, [_I64, _VP, _VP, _VP]),
}

_lib = None


class TbiError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (building first if the .so is absent and nvcc is present).  Raises if it cannot."""
    global _lib
    if _lib is None:
        if is_stale():                      # sources changed (or never built): rebuild in-tree; raises if nvcc fails
            build()
        L = C.CDLL(os.environ.get("TBI_LIB", str(LIB_PATH)))      # TBI_LIB: load an experimental build (scratch/ only)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if L.tbi_version() != 100:
            raise TbiError("libtbi_sm100.so version mismatch")
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().tbi_last_error().decode(errors="replace")
        raise TbiError(f"{what}: tbi error {rc}: {msg}")
