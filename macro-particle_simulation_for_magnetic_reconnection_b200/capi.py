"""ctypes binding of the C ABI declared in include/mrg_fulmov.h.

This is the product path: if the CUDA library is missing or no GPU is present
the calls raise -- there is no CPU fallback (the CPU oracle under oracle/ is
test infrastructure and is never imported from here).
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

dp = C.POINTER(C.c_double)
UNIQUE_ID_BYTES = 128

# every symbol include/mrg_fulmov.h declares
SYMBOLS = [
    "mrg_create", "mrg_destroy", "mrg_last_error", "mrg_build_info", "mrg_comm_unique_id",
    "mrg_comm_init", "mrg_upload_particles", "mrg_download_particles", "mrg_num_local",
    "mrg_loadpt", "mrg_set_fields", "mrg_set_fields_device", "mrg_fulmov", "mrg_get_moments",
    "mrg_get_moments_device", "mrg_get_prepared_fields", "mrg_prefld", "mrg_update_b", "mrg_get_fields", "mrg_sort", "mrg_set_option",
    "mrg_get_counters", "mrg_last_kernel_ms", "mrg_event_record", "mrg_event_elapsed_ms",
    "mrg_synchronize", "mrg_bind_fields_device", "mrg_renew_fields", "mrg_pass_ms", "mrg_get_prep_stats",
    "mrg_set_moment_sink", "mrg_plane_sets", "mrg_compact_layout", "mrg_self_check", "mrg_dfma_peak", "mrg_set_fields_lazy", "mrg_renew_fields_host", "mrg_phase_ms", "mrg_phase_detail", "mrg_peer_export", "mrg_peer_import", "mrg_peer_pushes", "mrg_split_pushes",
]


class StepParams(C.Structure):
    """mrg_step_params: the scalars fulmov reads from COMMON (F:1085-1110)."""

    _fields_ = [
        ("dt", C.c_double), ("adt", C.c_double), ("hdt", C.c_double), ("aimpl", C.c_double),
        ("bxc", C.c_double), ("byc", C.c_double), ("bzc", C.c_double),
        ("ifilx", C.c_int32), ("ifily", C.c_int32), ("ifilz", C.c_int32), ("drive_on", C.c_int32),
        ("Ez00", C.c_double), ("zcent", C.c_double), ("ycent1", C.c_double), ("ycent2", C.c_double),
    ]


class MrgError(RuntimeError):
    pass


_lib = None


def library_path():
    return _build.LIB


def load(build_if_missing=True):
    """Load libmrg_fulmov.so (building it with nvcc if the sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("MRG_LIB") or _build.LIB      # MRG_LIB: an experimental build of the same sources (tools/)
    if build_if_missing and path == _build.LIB:
        try:
            _build.build_cuda()
        except Exception:
            if not os.path.exists(_build.LIB):
                raise
    if not os.path.exists(path):
        raise MrgError("%s is missing: run __graft_entry__.build() (no CPU fallback exists)" % path)
    L = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.mrg_last_error.restype = C.c_char_p
    L.mrg_build_info.restype = C.c_char_p
    L.mrg_create.argtypes = [C.POINTER(vp), i32, i32, i32, C.c_double, C.c_double, C.c_double, i32, i32, i32, i32]
    L.mrg_destroy.argtypes = [vp]
    L.mrg_comm_unique_id.argtypes = [C.c_char_p]
    L.mrg_comm_init.argtypes = [vp, C.c_char_p]
    L.mrg_upload_particles.argtypes = [vp, i32] + [dp] * 6 + [i64, i64, i64]
    L.mrg_download_particles.argtypes = [vp, i32] + [dp] * 6 + [i64, i64, i64]
    L.mrg_num_local.restype = i64
    L.mrg_num_local.argtypes = [vp, i32]
    L.mrg_loadpt.argtypes = [vp, i32, i32, C.c_double, C.c_double, C.c_double, C.POINTER(i32), C.POINTER(i32)]
    L.mrg_set_fields.argtypes = [vp, C.c_uint32, C.POINTER(dp)]
    L.mrg_set_fields_device.argtypes = [vp, C.c_uint32, C.POINTER(vp)]
    L.mrg_bind_fields_device.argtypes = [vp, C.c_uint32, C.POINTER(vp)]
    L.mrg_renew_fields.argtypes = [vp]
    L.mrg_set_fields_lazy.argtypes = [vp, C.c_uint32, C.POINTER(dp)]
    L.mrg_renew_fields_host.argtypes = [vp, C.POINTER(dp)]
    L.mrg_compact_layout.argtypes = [i32, i32, i32, C.POINTER(C.c_uint8), C.POINTER(i32)]
    L.mrg_plane_sets.argtypes = [i32, C.POINTER(C.c_uint8), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.mrg_set_moment_sink.argtypes = [vp, i32, dp, dp, dp, dp]
    L.mrg_pass_ms.argtypes = [vp, i32, i32, dp]
    L.mrg_get_prep_stats.argtypes = [vp, C.POINTER(i64), i32]
    L.mrg_fulmov.argtypes = [vp, i32, C.c_double, C.c_double, i32, C.POINTER(StepParams), C.POINTER(i32), dp, dp]
    L.mrg_get_moments.argtypes = [vp, i32, dp, dp, dp, dp, i32]
    L.mrg_get_moments_device.argtypes = [vp, i32, C.POINTER(vp)]
    L.mrg_get_prepared_fields.argtypes = [vp, C.POINTER(StepParams), C.POINTER(dp)]
    L.mrg_prefld.argtypes = [vp, C.c_double, C.c_double]
    L.mrg_update_b.argtypes = [vp, C.c_double, C.c_double, i32]
    L.mrg_get_fields.argtypes = [vp, C.c_uint32, C.POINTER(dp)]
    L.mrg_sort.argtypes = [vp, i32, C.c_double]
    L.mrg_set_option.argtypes = [vp, C.c_char_p, i64]
    L.mrg_get_counters.argtypes = [vp, C.POINTER(i64), i32]
    L.mrg_last_kernel_ms.argtypes = [vp, dp]
    L.mrg_synchronize.argtypes = [vp]
    L.mrg_event_record.argtypes = [vp, i32]
    L.mrg_event_elapsed_ms.argtypes = [vp, i32, i32, dp]
    L.mrg_self_check.argtypes = [vp, i32, dp, C.POINTER(i64)]
    L.mrg_dfma_peak.argtypes = [vp, dp]
    L.mrg_phase_ms.argtypes = [vp, dp, C.POINTER(i64), i32]
    L.mrg_phase_detail.argtypes = [vp, i32, i32, dp]
    L.mrg_peer_export.argtypes = [vp, i32, C.c_char_p]
    L.mrg_peer_import.argtypes = [vp, i32, i32, C.c_char_p]
    L.mrg_peer_pushes.argtypes = [vp, i32]
    L.mrg_peer_pushes.restype = i64
    L.mrg_split_pushes.argtypes = [vp, i32]
    L.mrg_split_pushes.restype = i64
    for name in SYMBOLS:
        fn = getattr(L, name)
        if name not in ("mrg_last_error", "mrg_build_info", "mrg_num_local", "mrg_peer_pushes", "mrg_split_pushes"):
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise MrgError("mrg error %d: %s" % (rc, load().mrg_last_error().decode()))


def as_dp(a):
    if a is None:
        return dp()
    if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
        raise TypeError("need a contiguous float64 array")
    return a.ctypes.data_as(dp)
