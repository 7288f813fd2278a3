"""The particle records of the reference's restart file (F:9722-9725), written from HBM by libmrg_host (csrc/mrg_restart.*).
CPU part: the Fortran unformatted framing (4-byte markers, subrecord split with signed markers), and the pin: the
reference's own restrt (translated, oracle/_ref) writes its unit 12 after a step of its time cycle -- the last four records
of that file are byte-identical to what the product-side writer produces, and the reference's restrt(iresrt=1) reads a file
whose particle records came from that writer.  GPU part: the four records round-trip through a second context, carry exactly
what mrg_download_particles returns, and the device-written file equals the reference-written file byte for byte."""
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


# ---- the reference's own restrt (F:9522-9856), through the translated code ------------------------------------------------
def read_records(path):
    """[(payload bytes)] of a Fortran unformatted sequential file (sub-records joined)"""
    raw, out, o, cur = open(path, "rb").read(), [], 0, b""
    while o < len(raw):
        m0 = struct.unpack("<i", raw[o:o + 4])[0]
        n = abs(m0)
        cur += raw[o + 4:o + 4 + n]
        m1 = struct.unpack("<i", raw[o + 4 + n:o + 8 + n])[0]
        assert abs(m1) == n
        o += 8 + n
        if m0 >= 0:
            out.append(cur)
            cur = b""
    return out


def reference_restart_file(tmp_path, nranks=2, steps=1):
    """two ranks run the reference's time cycle for a step, then its own restrt(iresrt=2) writes unit 12 (rank 0, io_pe = 1)"""
    from oracle import pyref as PR
    grid = (8, 6, 8)
    p = U.make_parm(*grid)
    path = str(tmp_path / "ref.12")
    A = PR.ReferenceLoop(grid, (p.xmax, p.ymax, p.zmax), nranks)
    A.startup()
    for _ in range(steps):
        A.begin_step(); A.fulmov(1); A.emfild(); A.fulmov(0); A.renew()
    R = A.R
    R.L.ref_set_unit_path(12, path.encode())
    R.arr("iope66", "io_pe", 0, "restrt")[0] = 1                      # rank 0 writes
    want = A.particles()
    np0 = R.np0
    xs = [[A.parts[r][k][c] for r in range(nranks)] for k in (1, 2) for c in range(6)]
    args = xs[:6] + [1.0, 100.0] + xs[6:] + [-1.0, 1.0] + [A.npr, PR.IPAR, PR.SIZE, 2]
    R.call("restrt", *args)
    return A, path, want, np0


needs_ref = pytest.mark.skipif(not __import__("oracle.pyref", fromlist=["x"]).available(), reason="oracle/_ref is not built and /root/reference is not here")


@needs_ref
def test_particle_records_are_what_the_reference_restrt_writes(tmp_path):
    """The four records csrc/mrg_restart.cpp produces (F:9722-9725) against the file the reference's own restrt writes:
    12 records, the last four are (qmulti,wmulti,qmulte,wmulte), npr, six ion arrays of np0 doubles, six electron arrays --
    the strided ownership of the ranks merged by the twelve mpi_allreduce (F:9644-9668).  The same payloads written with
    mrg_f77_write_record give the same bytes, and the reference's restrt(iresrt=1) reads a file whose particle records
    came from that writer."""
    from oracle import pyref as PR
    A, path, want, np0 = reference_restart_file(tmp_path)
    try:
        recs = read_records(path)
        assert len(recs) == 12                                            # F:9700-9725
        assert len(recs[0]) == 4 * 19 and struct.unpack("<i", recs[0][:4])[0] == A.it          # it, ldec, ... nhist (F:9700-9702)
        n_grid = (8 + 4) * (6 + 3) * (8 + 4)
        assert len(recs[3]) == 12 * 8 * n_grid and len(recs[4]) == 11 * 8 * n_grid            # /fields/, /srimp7/
        assert np.frombuffer(recs[8], dtype=np.float64).tolist() == [1.0, 100.0, -1.0, 1.0]    # F:9722
        assert struct.unpack("<i", recs[9])[0] == A.npr                                        # F:9723
        for k, rec in ((1, recs[10]), (2, recs[11])):                                          # F:9724-9725
            a = np.frombuffer(rec, dtype=np.float64).reshape(6, np0)
            for c in range(6):
                np.testing.assert_array_equal(a[c][:A.npr], want[k][c])
        # the product-side record writer on the same payloads: byte-identical tail of the file
        L = host_lib()
        mine = str(tmp_path / "mine.bin")
        pay = [np.array([1.0, 100.0, -1.0, 1.0]), np.array([A.npr], dtype=np.int32)]
        assert write_record(L, mine, [pay[0]], "wb") == 0 and write_record(L, mine, [pay[1]]) == 0
        for k in (1, 2):
            arrs = [np.zeros(np0) for _ in range(6)]
            for c in range(6):
                arrs[c][:A.npr] = want[k][c]
            assert write_record(L, mine, arrs) == 0
        raw_ref, raw_mine = open(path, "rb").read(), open(mine, "rb").read()
        assert raw_ref[-len(raw_mine):] == raw_mine
        # and back: records 1-8 of the reference + OUR four records, read by the reference's restrt(iresrt = 1)
        mixed = str(tmp_path / "mixed.12")
        with open(mixed, "wb") as f:
            f.write(raw_ref[:len(raw_ref) - len(raw_mine)] + raw_mine)
        R = A.R
        R.L.ref_set_unit_path(12, mixed.encode())
        nr = A.nranks
        back = [{k: [np.zeros(np0) for _ in range(6)] for k in (1, 2)} for _ in range(nr)]
        xs = [[back[r][k][c] for r in range(nr)] for k in (1, 2) for c in range(6)]
        q = [[C.c_double(0.0) for _ in range(nr)] for _ in range(4)]
        npr = [C.c_int32(0) for _ in range(nr)]
        R.set("parm1", "it", -1, unit="restrt")
        R.call("restrt", *(xs[:6] + [q[0], q[1]] + xs[6:] + [q[2], q[3]] + [npr, PR.IPAR, PR.SIZE, 1]))
        assert int(R.get("parm1", "it", unit="restrt")) == A.it and npr[0].value == A.npr
        assert [q[i][0].value for i in range(4)] == [1.0, 100.0, -1.0, 1.0]
        for r in range(nr):                                                # "complete set" on every rank (F:9749-9750)
            for k in (1, 2):
                for c in range(6):
                    np.testing.assert_array_equal(back[r][k][c][:A.npr], want[k][c])
    finally:
        A.close()


@needs_ref
def test_reference_runtime_splits_long_records_like_the_product_writer(tmp_path):
    """two independent implementations of the gfortran sub-record convention (oracle/ref_runtime.c for the translated
    reference, csrc/mrg_restart.cpp for the product) write the same bytes when a record exceeds the sub-record limit"""
    A, path, want, np0 = reference_restart_file(tmp_path)
    try:
        whole = read_records(path)
        A.R.L.ref_set_max_subrecord(4096)
        _, path2, _, _ = None, str(tmp_path / "split.12"), None, None
        A.R.L.ref_set_unit_path(12, path2.encode())
        nr = A.nranks
        from oracle import pyref as PR
        xs = [[A.parts[r][k][c] for r in range(nr)] for k in (1, 2) for c in range(6)]
        A.R.call("restrt", *(xs[:6] + [1.0, 100.0] + xs[6:] + [-1.0, 1.0] + [A.npr, PR.IPAR, PR.SIZE, 2]))
        A.R.L.ref_set_max_subrecord(0)
        assert os.path.getsize(path2) > os.path.getsize(path)              # more markers
        assert read_records(path2) == whole                                # same payloads
        L = host_lib()
        mine = str(tmp_path / "mine_split.bin")
        L.mrg_f77_set_max_subrecord(4096)
        arrs = [np.zeros(np0) for _ in range(6)]
        for c in range(6):
            arrs[c][:A.npr] = want[2][c]
        assert write_record(L, mine, arrs, "wb") == 0
        L.mrg_f77_set_max_subrecord(0)
        raw2, rawm = open(path2, "rb").read(), open(mine, "rb").read()
        assert raw2[-len(rawm):] == rawm
    finally:
        A.close()


@pytest.mark.gpu
@needs_ref
def test_device_restart_records_equal_the_reference_file(tmp_path):
    """mrg_restart_append_particles, from particles resident (and sorted) in HBM, against the file the reference's own restrt
    wrote for the same particles: the last four records are byte-identical; and the reference's restrt(iresrt=1) reads them."""
    import mrg_b200 as mrg
    from oracle import pyref as PR
    A, path, want, np0 = reference_restart_file(tmp_path)
    try:
        L = host_lib()
        p = U.make_parm(8, 6, 8)
        ctx = mrg.MrgContext(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax)
        for k in (1, 2):
            ctx.upload(k, *want[k])
            ctx.sort(k, p.hdt)                                   # device order differs from l order
        raw_ref = open(path, "rb").read()
        tail = sum(len(r) + 8 for r in read_records(path)[8:])
        mine = str(tmp_path / "device.12")
        with open(mine, "wb") as f:
            f.write(raw_ref[:len(raw_ref) - tail])               # the host's records 1-8
        assert L.mrg_restart_append_particles(ctx.h, mine.encode(), 1.0, 100.0, -1.0, 1.0, A.npr, np0, 1, 1) == 0
        ctx.close()
        assert open(mine, "rb").read() == raw_ref
    finally:
        A.close()
