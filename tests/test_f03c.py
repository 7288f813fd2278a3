"""oracle/f03c.py on its own: the translator that turns the reference's Fortran into the C behind oracle/_ref is held to
Fortran semantics on small sources written for this purpose (tests/f03c_cases/semantics.f03) -- integer division and mod,
default-real literals, x**n, parentheses, DO trip counts and the value of the DO variable afterwards, shared terminal labels,
GO TO, COMMON blocks viewed through different member lists, by-reference arguments (array elements, expression
temporaries), functions, SAVE/DATA, EQUIVALENCE, ENTRY, and the simulated MPI ranks (isend/irecv/wait ring, rank-ordered
allreduce), and unformatted sequential records.  Expected values are worked out by hand from the Fortran standard's
rules.  Needs gcc only."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from oracle import pyref as PR          # noqa: E402

CASES = os.path.join(ROOT, "tests", "f03c_cases")
UNITS = ["arith", "loops", "blocks_a", "blocks_b", "caller", "overlay", "twodoors", "sidedoor", "ring", "dump"]


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    import f03c
    out = tmp_path_factory.mktemp("f03c")
    src, _ = f03c.translate(os.path.join(CASES, "semantics.f03"), [CASES], UNITS, param_include="case_sizes.h")
    with open(os.path.join(out, "mrgref_gen.c"), "w") as f:
        f.write(src)
    so = os.path.join(out, "libcases.so")
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fwrapv", "-fno-strict-aliasing", "-w", "-std=gnu99", "-fPIC", "-shared", "-pthread",
           "-I" + str(out), "-o", so, os.path.join(ROOT, "oracle", "ref_runtime.c"), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[:3000]
    return PR.load_library(so)


def run1(lib, unit, *args, nranks=1):
    with PR.RefRun(4, 3, 4, 10, nranks=nranks, npc=2, lib=lib) as R:
        R.call(unit, *args)


def test_arithmetic_follows_fortran_rules(lib):
    iout, rout = np.zeros(10, dtype=np.int32), np.zeros(10)
    run1(lib, "arith", iout, rout)
    assert list(iout[:8]) == [3, -3, -1, 1024, -2, 0, 6, 12]           # nint(2.5) = 3, nint(-2.5) = -3; nn = mx*my
    f32 = np.float32
    y = 1.1
    want = [1.0 / 3, float(f32(0.1)), float(f32(0.1) * f32(3)), 0.1 * 3, y * y * y, 1.0 / (y * y), 2.0 ** 0.5,
            -3.0 + 2.5 + 2.0 + 2, float(f32(7) / f32(2)), 0.0]
    np.testing.assert_array_equal(rout, np.array(want))


def test_loops_and_branches(lib):
    iout = np.zeros(10, dtype=np.int32)
    run1(lib, "loops", iout)
    # zero-trip loop leaves i = 5; 10,7,4,1 then i = -2; sum i*k = 18; while: 1 + 3 = 4; three passes of the GO TO loop
    assert list(iout[:8]) == [0, 5, 22, -2, 18, 4, 30, 1]


def test_common_is_a_storage_sequence(lib):
    sout = np.zeros(4)
    with PR.RefRun(4, 3, 4, 10, nranks=1, npc=2, lib=lib) as R:
        R.call("blocks_a")
        R.call("blocks_b", sout)
        a = R.arr("blk", "a", unit="blocks_a")
        assert a.size == 8 * 4 and R.arr("blk", "v", unit="blocks_b").size == 32
    # a(-2,0) = -2; a(0,my=3) = 300 is element (mx+4)*my + 3 of the vector view; the tail b(1:3), kk
    np.testing.assert_array_equal(sout, [-2.0, 300.0, 321.0, 42.0])


def test_arguments_functions_save(lib):
    rout = np.zeros(8)
    with PR.RefRun(4, 3, 4, 10, nranks=1, npc=2, lib=lib) as R:
        R.call("caller", rout)
        first = rout.copy()
        R.call("caller", rout)
    # x+1 = 2, n*2 = 6, arr(2) of the element actual is w(3) -> -3; the expression actual leaves x alone, n = 12, w(2) = -2
    np.testing.assert_array_equal(first, [2.0, 6.0, -3.0, 2.0, 12.0, -2.0, 7.0, 1.0])
    assert rout[7] == 2.0                                                  # SAVE + DATA: counts the calls


def test_equivalence_and_entry(lib):
    rout = np.zeros(3)
    run1(lib, "overlay", rout)
    np.testing.assert_array_equal(rout, [2.0, 5.0, -4.0])                   # w1(2,2) is w0(4)
    r2 = np.zeros(2)
    with PR.RefRun(4, 3, 4, 10, nranks=1, npc=2, lib=lib) as R:
        R.call("twodoors", 5, r2)
        R.call("sidedoor")
        R.call("twodoors", 1, r2)
    assert r2[0] == 1006.0


def test_simulated_ranks_exchange_and_reduce_in_rank_order(lib):
    outs = [np.zeros(3) for _ in range(2)]
    with PR.RefRun(4, 3, 4, 10, nranks=2, npc=2, lib=lib) as R:
        R.call("ring", [0, 1], outs)
    # rank 0 receives rank 1's (20, -1), rank 1 receives rank 0's (10, 0); the sum is formed in rank order on every rank
    np.testing.assert_array_equal(outs[0][:2], [20.0, -1.0])
    np.testing.assert_array_equal(outs[1][:2], [10.0, 0.0])
    tot = 0.0 + 0.1 * 1
    tot = tot + 0.1 * 2
    assert outs[0][2] == outs[1][2] == tot


def test_unformatted_records(lib, tmp_path):
    """write(u) list / read(u) list: one record per statement, implied DO lists in storage order, mixed kinds; the file is
    what a Fortran run time writes (4-byte markers)"""
    import struct
    path = str(tmp_path / "unit31.bin")
    lib.ref_set_unit_path(31, path.encode())
    with PR.RefRun(4, 3, 4, 10, nranks=1, npc=2, lib=lib) as R:
        R.call("dump", 2)
        raw = open(path, "rb").read()
        n1 = 4 + 16 + 8                       # n, kk(1:4), s
        n2 = 6 * 8 + 3 * 4                    # a(2,3) in column-major order, r4(3)
        assert len(raw) == n1 + n2 + 16
        assert struct.unpack("<i", raw[:4])[0] == n1 == struct.unpack("<i", raw[4 + n1:8 + n1])[0]
        assert struct.unpack("<i5id", raw[:8 + n1 - 4])[1:] == (7, 10, 20, 30, 40, 2.5)
        rec2 = raw[8 + n1 + 4:8 + n1 + 4 + n2]
        assert struct.unpack("<6d3f", rec2) == (11.0, 12.0, 21.0, 22.0, 31.0, 32.0, 0.5, 1.0, 1.5)
        R.call("dump", 1)
        assert R.arr("dumpc", "b", unit="dump").tolist() == [11.0, 12.0, 21.0, 22.0, 31.0, 32.0]
        assert float(R.get("dumpc", "sback", unit="dump")) == 2.5 and int(R.get("dumpc", "nback", unit="dump")) == 7
        assert R.arr("dumpc", "q4", unit="dump").tolist() == [0.5, 1.0, 1.5] and R.arr("dumpc", "ll", unit="dump").tolist() == [10, 20, 30, 40]
