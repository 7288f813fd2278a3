"""B200-native /fulmov/ particle hot path of Tanaka's implicit macro-particle
PIC code (@mrg37-080A.f03): CUDA kernels for sm_100a behind a C ABI
(include/mrg_fulmov.h), plus host-side mirrors of the reference's subroutine
interface (csrc/mrg_host.cpp in C++, host.py in Python, fortran/mrg_gpu.f03 as
the ISO_C_BINDING shim)."""
from . import build, capi
from .capi import MrgError, StepParams
from .host import Common, Fulmov, MrgContext, broadcast_unique_id, mxyzA, owned_count, owned_slice

__all__ = ["build", "capi", "MrgError", "StepParams", "Common", "Fulmov", "MrgContext", "mxyzA",
           "owned_count", "owned_slice", "broadcast_unique_id"]
