"""The C++ host mirror of fulmov (csrc/mrg_host.cpp, the Fortran by-reference ABI) on the CPU: linked against a recording
stand-in of the C ABI (tests/host_emul/stub_abi.cpp), so that what it asks of the device library -- and when -- can be
checked without a GPU: the upload masks of the field-hint protocol, the renewal on the device, the it = 0 sequence
(ADVICE r1), the prefld / emfild marks that keep bx,by,bz off PCIe, COMMON /wkinel/ and edec, the sort cadence."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "macro-particle_simulation_for_magnetic_reconnection_b200", "csrc")
_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)


class View(C.Structure):
    """mirror of mrg_common_view (csrc/mrg_host.h)"""
    _fields_ = ([("mx", C.c_int32), ("my", C.c_int32), ("mz", C.c_int32)]
                + [(n, _dp) for n in ("ex", "ey", "ez", "bx", "by", "bz", "ex0", "ey0", "ez0", "bx0", "by0", "bz0")]
                + [(n, _dp) for n in ("qix", "qiy", "qiz", "qex", "qey", "qez", "qi", "qe")]
                + [(n, _ip) for n in ("it", "ldec", "ifilx", "ifily", "ifilz", "nha")]
                + [(n, _dp) for n in ("xmax", "ymax", "zmax", "dt", "aimpl", "adt", "hdt", "bxc", "byc", "bzc", "edec")]
                + [(n, _dp) for n in ("wkix", "wkih", "zcent", "ycent1", "ycent2", "Ez00")]
                + [("ranfb", _ip), ("io_pe", _ip)])


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostlogic")
    so = os.path.join(out, "libhostlogic.so")
    srcs = [os.path.join(CSRC, "mrg_host.cpp"), os.path.join(ROOT, "tests", "host_emul", "stub_abi.cpp")]
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so] + srcs, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[:3000]
    L = C.CDLL(so)
    L.stub_trace.restype = C.c_char_p
    L.mrg_host_fulmov.argtypes = [_dp] * 8 + [_ip] * 5
    L.mrg_host_fulmov.restype = None
    L.mrg_host_bind.argtypes = [C.POINTER(View), C.c_int32]
    return L


class Host:
    """a miniature of the reference's COMMON blocks + the call sequence of trans"""

    def __init__(self, L, auto_fields):
        self.L = L
        n = 6 * 5 * 6
        self.keep = {}
        v = View()
        v.mx, v.my, v.mz = 2, 2, 2
        for k, name in enumerate(("ex", "ey", "ez", "bx", "by", "bz", "ex0", "ey0", "ez0", "bx0", "by0", "bz0")):
            self.keep[name] = np.full(n, float(k))
            setattr(v, name, self.keep[name].ctypes.data_as(_dp))
        for name in ("qix", "qiy", "qiz", "qex", "qey", "qez", "qi", "qe"):
            self.keep[name] = np.zeros(n)
            setattr(v, name, self.keep[name].ctypes.data_as(_dp))
        ints = {"it": 0, "ldec": 1, "ifilx": 1, "ifily": 1, "ifilz": 1, "nha": 5, "ranfb": 100, "io_pe": 1}
        for name, val in ints.items():
            self.keep[name] = np.array([val], dtype=np.int32)
            setattr(v, name, self.keep[name].ctypes.data_as(_ip))
        dbl = {"xmax": 1.0, "ymax": 1.0, "zmax": 1.0, "dt": 1.2, "aimpl": 0.6, "adt": 0.72, "hdt": 0.6, "bxc": 0.2, "byc": 0.0,
               "bzc": 0.0, "wkix": 0.0, "wkih": 0.0, "zcent": 0.5, "ycent1": 0.3, "ycent2": 0.7, "Ez00": 0.0025}
        for name, val in dbl.items():
            self.keep[name] = np.array([val])
            setattr(v, name, self.keep[name].ctypes.data_as(_dp))
        self.keep["edec"] = np.zeros(3000 * 12)
        v.edec = self.keep["edec"].ctypes.data_as(_dp)
        self.v = v
        L.mrg_host_unbind()
        L.stub_reset()
        assert L.mrg_host_bind(C.byref(v), 0) == 0
        L.mrg_host_set_auto_fields(1 if auto_fields else 0)
        self.x = np.zeros(8)

    def fulmov(self, ipc, ksp):
        a = self.x.ctypes.data_as(_dp)
        q, w = np.array([1.0 if ksp == 1 else -1.0]), np.array([100.0 if ksp == 1 else 1.0])
        i = [np.array([val], dtype=np.int32) for val in (8, ipc, ksp, 1, 1)]
        self.L.mrg_host_fulmov(a, a, a, a, a, a, q.ctypes.data_as(_dp), w.ctypes.data_as(_dp), *[j.ctypes.data_as(_ip) for j in i])

    def trace(self):
        t = self.L.stub_trace().decode().splitlines()
        self.L.stub_reset()
        return t


def test_auto_mode_uploads_everything_on_every_ion_call(lib):
    h = Host(lib, auto_fields=True)
    h.keep["it"][0] = 1
    h.fulmov(1, 1); h.fulmov(1, 2); h.fulmov(0, 1); h.fulmov(0, 2)
    t = h.trace()
    assert t[0].startswith("create 2 2 2") and t[1].startswith("upload ksp=1 npr=8 first=1 stride=1")
    assert [x for x in t if x.startswith("set_fields")] == ["set_fields mask=0xfff ex=0 bx=3"] * 2       # once per ion call
    assert [x.split()[0] + x.split()[1] + x.split()[2] for x in t if x.startswith("fulmov")] == \
        ["fulmovksp=1ipc=1", "fulmovksp=2ipc=1", "fulmovksp=1ipc=0", "fulmovksp=2ipc=0"]
    assert [x for x in t if x.startswith("get_moments")] == ["get_moments ksp=1 folded=1", "get_moments ksp=2 folded=1"]
    assert [x for x in t if x.startswith("sort")] == ["sort ksp=1 lookahead=0.6", "sort ksp=2 lookahead=0.6"]
    assert h.keep["qix"][0] == 101.0 and h.keep["qex"][0] == 102.0          # COMMON /srimp7/ filled per species
    assert h.keep["wkix"][0] == 20.0 and h.keep["ranfb"][0] == 102            # COMMON /wkinel/ of the last call; two ipc = 0 calls drew


def test_marks_keep_b_off_pcie_and_survive_it0(lib):
    """hints mode through a start-up and two steps: the it = 0 pair, emfld0 + renewal (no mark exists for emfld0), then
    prefld_done / emfild_done / fields_renewed per step -- the second step is a smoothing step of emfild (it = 6)"""
    h = Host(lib, auto_fields=False)
    L = lib
    h.keep["it"][0] = 0
    for nm in ("dt", "adt", "hdt"):
        h.keep[nm][0] = 0.0
    h.fulmov(1, 1); h.fulmov(1, 2)                         # it = 0 pair (F:684-689)
    h.keep["dt"][0], h.keep["adt"][0], h.keep["hdt"][0] = 1.2, 0.72, 0.6
    L.mrg_host_fields_renewed()                            # F:796-807 after emfld0
    t0 = h.trace()
    assert [x for x in t0 if x.startswith("set_fields")] == ["set_fields mask=0xfff ex=0 bx=3"]
    assert any(x.startswith("fulmov ksp=1 ipc=1 dt=0 hdt=0") for x in t0)
    steps = []
    for it in (5, 6):
        h.keep["it"][0] = it
        L.mrg_host_prefld_done()
        h.fulmov(1, 1); h.fulmov(1, 2)
        L.mrg_host_emfild_done()
        h.fulmov(0, 1); h.fulmov(0, 2)
        L.mrg_host_fields_renewed()
        steps.append(h.trace())
    a, b = steps
    # first step after it = 0: everything but bx,by,bz is uploaded (emfld0 rewrote it all), the stale device renewal is dropped,
    # B comes from the device's prefld; after emfild only ex,ey,ez cross, B is recomputed without smoothing (it = 5)
    assert [x for x in a if x.startswith(("set_fields", "renew", "update_b"))] == \
        ["set_fields mask=0xfc7 ex=0 bx=-1", "update_b dt=1.2 aimpl=0.6 smooth=0", "set_fields mask=0x007 ex=0 bx=-1",
         "update_b dt=1.2 aimpl=0.6 smooth=0"]
    # second step: the renewal runs on the device, nothing is uploaded for the predictor, and emfild's smoothing step (mod(it,5) = 1)
    assert [x for x in b if x.startswith(("set_fields", "renew", "update_b"))] == \
        ["renew", "update_b dt=1.2 aimpl=0.6 smooth=0", "set_fields mask=0x007 ex=0 bx=-1", "update_b dt=1.2 aimpl=0.6 smooth=1"]
    # edec rows (F:1320-1328): it = 5 is a history step (nha = 5, io_pe = 1): columns 5,6 (ions) and 7,8 (electrons) of row ldec
    e = h.keep["edec"].reshape(12, 3000)
    assert e[4, 0] == 10.0 and e[5, 0] == 0.5 and e[6, 0] == 20.0 and e[7, 0] == 0.5


def test_auto_mode_turns_the_marks_into_uploads(lib):
    h = Host(lib, auto_fields=True)
    h.keep["it"][0] = 3
    lib.mrg_host_prefld_done()
    h.fulmov(1, 1)
    lib.mrg_host_emfild_done()
    h.fulmov(0, 1)
    t = h.trace()
    assert not any(x.startswith("update_b") for x in t)
    assert [x for x in t if x.startswith("set_fields")] == ["set_fields mask=0xfff ex=0 bx=3"] * 2
