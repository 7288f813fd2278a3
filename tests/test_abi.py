"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/mrg_fulmov.h declares; compute calls fail loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

from tests.conftest import HAS_GPU, ROOT


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "mrg_fulmov.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mrg_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_exported():
    import mrg_b200
    lib = mrg_b200.capi.load()
    names = header_symbols()
    assert len(names) >= 20
    assert sorted(mrg_b200.capi.SYMBOLS) == names          # the binding covers the whole header
    for n in names:
        assert hasattr(lib, n), n
    assert b"sm_100a" in lib.mrg_build_info()


def test_library_is_sm100a_only():
    import subprocess
    import mrg_b200
    out = subprocess.run(["cuobjdump", "-lelf", mrg_b200.capi.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_host_mirror_exports_reference_signature():
    import mrg_b200
    mrg_b200.build.build_host()
    lib = C.CDLL(mrg_b200.build.HOSTLIB)
    for n in ("mrg_host_fulmov", "mrg_host_bind", "mrg_host_pull_particles", "mrg_host_particles_changed",
              "mrg_host_fields_changed", "mrg_host_set_unique_id"):
        assert hasattr(lib, n), n
    # ... and everything else the two host headers declare (the marks of the hint protocol, the abort hook, the restart records)
    csrc = os.path.join(ROOT, "macro-particle_simulation_for_magnetic_reconnection_b200", "csrc")
    for hdr in ("mrg_host.h", "mrg_restart.h"):
        txt = re.sub(r"/\*.*?\*/", "", open(os.path.join(csrc, hdr)).read(), flags=re.S)
        names = sorted(set(re.findall(r"\b(mrg_(?:host|f77|restart)_[a-z_0-9]+)\s*\(", txt)))
        assert len(names) >= 4, hdr
        for n in names:
            assert hasattr(lib, n), (hdr, n)


@pytest.mark.skipif(HAS_GPU, reason="checks the CPU-only failure mode")
def test_no_cpu_fallback():
    import mrg_b200
    with pytest.raises(mrg_b200.MrgError, match="no CUDA device"):
        mrg_b200.MrgContext(8, 6, 8, 60.0, 50.0, 60.0)


def test_bad_arguments_are_rejected_before_touching_the_gpu():
    import mrg_b200
    lib = mrg_b200.capi.load()
    h = C.c_void_p()
    assert lib.mrg_create(C.byref(h), 2, 6, 8, 60.0, 50.0, 60.0, 2, 0, 1, 0) == 1      # mx < 4
    assert b"grid too small" in lib.mrg_last_error()
    assert lib.mrg_create(C.byref(h), 8, 6, 8, 60.0, 50.0, 60.0, 9, 0, 1, 0) == 1      # nspecies
    assert lib.mrg_create(C.byref(h), 8, 6, 8, 60.0, 50.0, 60.0, 2, 3, 2, 0) == 1      # rank >= nranks
    assert lib.mrg_synchronize(None) == 1
