"""The particle records of the reference's restart file (F:9722-9725), written from HBM by libmrg_host (csrc/mrg_restart.*).
CPU part: the Fortran unformatted framing (4-byte markers, subrecord split with signed markers).  GPU part: the four records
round-trip through a second context and carry exactly what mrg_download_particles returns.  The reference cannot be run to
produce a file here (its restrt I/O needs a Fortran run time), so this format is restated, not pinned: see the header."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from tests import util as U


def host_lib():
    import mrg_b200 as mrg
    mrg.build.build_host()
    L = C.CDLL(mrg.build.HOSTLIB)
    L.mrg_f77_set_max_subrecord.argtypes = [C.c_uint64]
    L.mrg_restart_append_particles.argtypes = [C.c_void_p, C.c_char_p] + [C.c_double] * 4 + [C.c_int64] * 4
    L.mrg_restart_read_particles.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int64)] + [C.c_int64] * 3
    return L


def write_record(L, path, arrays, mode="ab"):
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    L.mrg_f77_write_record.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_int32]
    f = libc.fopen(path.encode(), mode.encode())
    parts = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
    sizes = (C.c_uint64 * len(arrays))(*[a.nbytes for a in arrays])
    rc = L.mrg_f77_write_record(f, parts, sizes, len(arrays))
    libc.fclose(f)
    return rc


def test_unformatted_record_framing(tmp_path):
    L = host_lib()
    path = str(tmp_path / "rec.bin")
    a = np.arange(5, dtype=np.float64)
    b = np.arange(3, dtype=np.int32)
    assert write_record(L, path, [a, b], "wb") == 0
    raw = open(path, "rb").read()
    n = a.nbytes + b.nbytes
    assert len(raw) == n + 8
    assert struct.unpack("<i", raw[:4])[0] == n and struct.unpack("<i", raw[-4:])[0] == n      # [len][payload][len]
    assert raw[4:4 + a.nbytes] == a.tobytes() and raw[4 + a.nbytes:-4] == b.tobytes()
    # a record longer than the maximum subrecord length: leading marker negative while more follows, trailing marker
    # negative when a subrecord precedes (gfortran's convention)
    L.mrg_f77_set_max_subrecord(16)
    assert write_record(L, path, [a], "wb") == 0                   # 40 bytes -> 16 + 16 + 8
    L.mrg_f77_set_max_subrecord(0)
    raw = open(path, "rb").read()
    m = [struct.unpack("<i", raw[o:o + 4])[0] for o in (0, 20, 24, 44, 48, 60)]
    assert m == [-16, 16, -16, -16, 8, -8], m
    body = raw[4:20] + raw[28:44] + raw[52:60]
    assert body == a.tobytes()


@pytest.mark.gpu
def test_restart_particle_records_round_trip(tmp_path):
    import mrg_b200 as mrg
    L = host_lib()
    p = U.make_parm(8, 6, 8)
    sp, _ = U.load_species(p, 9)
    npr = len(sp[1][0])
    np0 = npr + 37                                               # declared array length of param_080A.h
    ctx = mrg.MrgContext(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax)
    for k in (1, 2):
        ctx.upload(k, *sp[k])
        ctx.sort(k, p.hdt)                                       # device order differs from l order
    path = str(tmp_path / "forta.12")
    head = np.arange(7, dtype=np.float64)
    assert write_record(L, path, [head], "wb") == 0              # stands for the host's records 1-8
    L.mrg_f77_set_max_subrecord(100000)                          # force subrecords inside the particle records
    assert L.mrg_restart_append_particles(ctx.h, path.encode(), 1.0, 100.0, -1.0, 1.0, npr, np0, 1, 1) == 0
    ctx2 = mrg.MrgContext(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax)
    qw = (C.c_double * 4)()
    n_out = C.c_int64()
    assert L.mrg_restart_read_particles(ctx2.h, path.encode(), 1, qw, C.byref(n_out), np0, 1, 1) == 0
    L.mrg_f77_set_max_subrecord(0)
    assert list(qw) == [1.0, 100.0, -1.0, 1.0] and n_out.value == npr
    for k in (1, 2):
        got = ctx2.download(k, npr)
        for c in range(6):
            np.testing.assert_array_equal(got[c], sp[k][c])     # bit-exact, original l order
    ctx.close()
    ctx2.close()
