"""ctypes binding of oracle/_ref/libmrg_ref.so: the reference's own routines (F = @mrg37-080A.f03), translated to C
by oracle/f03c.py and compiled by oracle/build_ref.py.  TEST INFRASTRUCTURE: only tests/, bench.py's CPU legs and the
golden-vector generators use it.

    with RefRun(mx, my, mz, np0, nranks=4) as R:      # param_080A.h sizes; one thread per simulated MPI rank
        R.set("parm2", "dt", 1.2)                     # COMMON members by the name the reference gives them
        ex = R.arr("fields", "ex", rank=0)            # numpy view of rank 0's COMMON /fields/ ex(-2:mx+1,-1:my+1,-2:mz+1)
        R.call("fulmov", x, y, z, vx, vy, vz, qmult, wmult, npr, ipc, ksp, IPAR, SIZE)

`call` runs the unit on every rank at once (the ranks meet in the simulated mpi_allreduce).  Arguments: numpy arrays are
passed as they are (shared by the ranks unless a list of per-rank arrays is given), Python floats / ints become
by-reference temporaries, IPAR stands for rank+1 and SIZE for the number of ranks (F:219).
"""
import ctypes as C
import os

import numpy as np

from . import build_ref

_lib = None
IPAR, SIZE = object(), object()


def available():
    return build_ref.available() or build_ref.can_build()


def load_library(path):
    """bind one translated library (oracle/_ref's, or a test's own translation of other Fortran sources)"""
    L = C.CDLL(path)
    _bind(L)
    return L


def load():
    global _lib
    if _lib is None:
        path = build_ref.build()
        if not path or not os.path.exists(path):
            raise RuntimeError("oracle/_ref/libmrg_ref.so is missing and /root/reference is not here to build it from")
        _lib = load_library(path)
    return _lib


def _bind(L):
    L.ref_set_params.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_long]
    L.ref_param.argtypes = [C.c_char_p]
    L.ref_param.restype = C.c_long
    L.ref_pool_start.argtypes = [C.c_int]
    L.ref_common.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_long), C.POINTER(C.c_int)]
    L.ref_common.restype = C.c_void_p
    L.ref_call.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.ref_has_unit.argtypes = [C.c_char_p]
    L.ref_set_unit_path.argtypes = [C.c_int, C.c_char_p]
    L.ref_set_max_subrecord.argtypes = [C.c_ulonglong]
    L.ref_collective_seconds.argtypes = [C.c_int, C.c_int]
    L.ref_collective_seconds.restype = C.c_double


_NP = {1: np.int32, 2: np.float32, 3: np.float64}


class RefRun:
    """One run of the translated reference: sizes of param_080A.h + a pool of simulated MPI ranks.  Only one can be
    alive at a time (the reference's sizes are process-wide, as its PARAMETERs are)."""

    def __init__(self, mx, my, mz, np0, nranks=1, npc=None, lib=None):
        self.L = lib or load()
        self.L.ref_pool_stop()
        npc = npc or nranks
        if self.L.ref_set_params(npc, mx, my, mz, np0):
            raise RuntimeError("ref_set_params failed")
        if self.L.ref_pool_start(nranks):
            raise RuntimeError("ref_pool_start failed")
        self.mx, self.my, self.mz, self.np0, self.nranks = mx, my, mz, np0, nranks
        self.last_secs = [0.0] * nranks

    def close(self):
        self.L.ref_pool_stop()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def param(self, name):
        return self.L.ref_param(name.encode())

    def has(self, unit):
        return self.L.ref_has_unit(unit.encode()) >= 0

    # -- COMMON ----------------------------------------------------------------------------------
    def arr(self, block, name, rank=0, unit=None):
        """numpy view (flat, Fortran storage order) of a COMMON member of one rank"""
        cnt, typ = C.c_long(), C.c_int()
        p = self.L.ref_common(rank, block.encode(), unit.encode() if unit else None, name.encode(), C.byref(cnt), C.byref(typ))
        if not p:
            raise KeyError("COMMON /%s/ %s (unit %s)" % (block, name, unit))
        ct = {1: C.c_int32, 2: C.c_float, 3: C.c_double}[typ.value]
        return np.ctypeslib.as_array((ct * cnt.value).from_address(p))

    def set(self, block, name, value, unit=None):
        for r in range(self.nranks):
            self.arr(block, name, r, unit)[...] = value

    def get(self, block, name, rank=0, unit=None):
        a = self.arr(block, name, rank, unit)
        return a[0] if a.size == 1 else a.copy()

    # -- calls -----------------------------------------------------------------------------------
    def call(self, unit, *args):
        n = self.L.ref_has_unit(unit.encode())
        if n < 0:
            raise KeyError("unit %r is not in the translated reference" % unit)
        if n != len(args):
            raise TypeError("%s takes %d arguments, %d given" % (unit, n, len(args)))
        keep, table = [], (C.c_void_p * (self.nranks * max(n, 1)))()
        for r in range(self.nranks):
            for i, a in enumerate(args):
                if a is IPAR:
                    a = r + 1
                elif a is SIZE:
                    a = self.nranks
                elif isinstance(a, (list, tuple)):
                    a = a[r]
                if isinstance(a, np.ndarray):
                    assert a.flags["C_CONTIGUOUS"] or a.flags["F_CONTIGUOUS"]
                    table[r * n + i] = a.ctypes.data
                    keep.append(a)
                elif isinstance(a, (int, np.integer)):
                    b = C.c_int32(int(a))
                    keep.append(b)
                    table[r * n + i] = C.addressof(b)
                elif isinstance(a, (float, np.floating)):
                    b = C.c_double(float(a))
                    keep.append(b)
                    table[r * n + i] = C.addressof(b)
                elif isinstance(a, (C.c_int32, C.c_double)):
                    keep.append(a)
                    table[r * n + i] = C.addressof(a)
                else:
                    raise TypeError("argument %d of %s: %r" % (i + 1, unit, type(a)))
        ret = (C.c_double * self.nranks)()
        secs = (C.c_double * self.nranks)()
        rc = self.L.ref_call(unit.encode(), table, ret, secs)
        if rc:
            raise RuntimeError("ref_call(%s) -> %d" % (unit, rc))
        self.last_secs = list(secs)
        return list(ret)

    def collective_seconds(self, reset=True):
        return [self.L.ref_collective_seconds(r, 1 if reset else 0) for r in range(self.nranks)]


# ------------------------------------------------------------------------------------------------
# Conveniences shared by the pin tests, the golden-vector generator and bench.py's reference leg
# ------------------------------------------------------------------------------------------------
def setup_run(R, xmax, ymax, zmax, dt=1.2, aimpl=0.6, wce_by_wpe=0.2, Ez00=0.25e-2, veth=0.2, te_by_ti=1.0,
              qspec=(1.0, -1.0), wspec=(100.0, 1.0), vbeam=(0.35e-2, -0.35e-2), vdr=(0.0, 0.0), nha=5, io_pe=0):
    """What `program` does before `init` (F:184-374): the namelist values of rec_3d80A into COMMON, filters forced to 1."""
    for name, v in (("xmax", xmax), ("ymax", ymax), ("zmax", zmax), ("dt", dt), ("aimpl", aimpl), ("wce_by_wpe", wce_by_wpe),
                    ("veth", veth), ("te_by_ti", te_by_ti), ("pi", 3.141592653589), ("thb", 90.0), ("rwd", 50.0),
                    ("epsln1", 1e-5)):
        R.set("parm2", name, v, unit="fulmov")
    for r in range(R.nranks):
        R.arr("parm2", "qspec", r, "fulmov")[:2] = qspec
        R.arr("parm2", "wspec", r, "fulmov")[:2] = wspec
        R.arr("parm2", "vbeam", r, "fulmov")[:2] = vbeam
        R.arr("parm2", "vdr", r, "fulmov")[:2] = vdr
    R.set("profl", "ez00", Ez00, unit="fulmov")
    for name, v in (("ifilx", 1), ("ifily", 1), ("ifilz", 1), ("nha", nha), ("it", 0), ("ldec", 1), ("iloadp", 0)):   # F:368-370
        R.set("parm1", name, v, unit="fulmov")
    R.set("iope66", "io_pe", io_pe, unit="fulmov")


def ref_init(R):
    """call the reference's own init (F:8244-8731) on every rank: tables, constants, loadpt of both species, xe = xi.
    Returns per-rank particle arrays {ksp: [x,y,z,vx,vy,vz]} (every rank loads ALL particles, F:121-122)."""
    n = R.np0
    parts = []
    for r in range(R.nranks):
        parts.append({k: [np.zeros(n) for _ in range(6)] for k in (1, 2)})
    args = []
    for k in (1, 2):
        args += [[parts[r][k][c] for r in range(R.nranks)] for c in range(6)]
        args += [0.0, 0.0]
    qm = [[C.c_double(), C.c_double(), C.c_double(), C.c_double()] for _ in range(R.nranks)]
    npr = [C.c_int32(0) for _ in range(R.nranks)]
    a = ([[parts[r][1][c] for r in range(R.nranks)] for c in range(6)] + [[q[0] for q in qm], [q[1] for q in qm]]
         + [[parts[r][2][c] for r in range(R.nranks)] for c in range(6)] + [[q[2] for q in qm], [q[3] for q in qm]]
         + [npr, 0])
    R.call("init", *a)
    return parts, npr[0].value, [(q[0].value, q[1].value, q[2].value, q[3].value) for q in qm][0]


FIELD_NAMES = ("ex", "ey", "ez", "bx", "by", "bz", "ex0", "ey0", "ez0", "bx0", "by0", "bz0")
MOMENT_NAMES = {1: ("qix", "qiy", "qiz", "qi"), 2: ("qex", "qey", "qez", "qe")}


def reference_startup(grid, box, nranks=1, qspec=(1.0, -1.0), wspec=(100.0, 1.0), dt=1.2, aimpl=0.6, wce_by_wpe=0.2, Ez00=0.25e-2):
    """The reference's own start-up (F:664-706): init (tables, constants, loadpt of both species at 32 per cell), then the
    it = 0 block of trans -- dt = adt = hdt = 0, fulmov(ipc=1) for ions and electrons, emfld0 (static B by Ampere's law from
    the accumulated moments + Poisson solve, F:3384-3703, 6596-7304) -- executed by `nranks` simulated ranks.  Returns the
    particles as init left them, the it = 0 moments and wkix/wkih, and the twelve COMMON /fields/ arrays emfld0 defined: the
    reference's own initial condition, for parity cases on real fields instead of synthetic ones."""
    mx, my, mz = grid
    np0 = 32 * mx * my * mz
    with RefRun(mx, my, mz, np0, nranks=nranks) as R:
        setup_run(R, box[0], box[1], box[2], dt=dt, aimpl=aimpl, wce_by_wpe=wce_by_wpe, Ez00=Ez00, qspec=qspec, wspec=wspec)
        parts, npr, _ = ref_init(R)
        ranfb = int(R.get("ranfb", "ir", unit="ranfp"))
        particles = {k: [parts[0][k][c][:npr].copy() for c in range(6)] for k in (1, 2)}
        sav = {nm: float(R.get("parm2", nm, unit="fulmov")) for nm in ("dt", "adt", "hdt")}
        for nm in sav:                                                             # F:673-679
            R.set("parm2", nm, 0.0, unit="fulmov")
        R.set("parm1", "it", 0, unit="fulmov")
        mom, wk = {}, {}
        for k in (1, 2):                                                           # F:684-689
            xs = [[parts[r][k][c] for r in range(nranks)] for c in range(6)]
            R.call("fulmov", *xs, float(qspec[k - 1]), float(wspec[k - 1]), npr, 1, k, IPAR, SIZE)
            mom[k] = [R.get("srimp7", nm, unit="fulmov") for nm in MOMENT_NAMES[k]]
            wk[k] = (float(R.get("wkinel", "wkix", unit="fulmov")), float(R.get("wkinel", "wkih", unit="fulmov")))
        R.call("emfld0")                                                           # F:694
        for nm, v in sav.items():                                                  # F:700-702
            R.set("parm2", nm, v, unit="fulmov")
        fields = [R.get("fields", nm, unit="fulmov") for nm in FIELD_NAMES]
        same = all(np.array_equal(R.get("fields", nm, rank=r, unit="fulmov"), f) for r in range(1, nranks) for nm, f in zip(FIELD_NAMES, fields))
        unmoved = all(np.array_equal(parts[0][k][c][:npr], particles[k][c]) for k in (1, 2) for c in range(6))
    return {"particles": particles, "npr": npr, "ranfb": ranfb, "mom0": mom, "wk0": wk, "fields": fields,
            "ranks_agree": bool(same), "particles_unmoved": bool(unmoved)}


def reference_steps(grid, box, particles, field_sets, nranks=1, ranfb_in=None, qspec=(1.0, -1.0), wspec=(100.0, 1.0),
                    dt=1.2, aimpl=0.6, wce_by_wpe=0.2, Ez00=0.25e-2, nha=5, it0=1):
    """The reference's own call sequence of trans (F:749-807) around the particle path, with the field solve replaced by
    given fields: for every (f_pred, f_corr) in field_sets
        COMMON /fields/ <- f_pred;  fulmov(ions, ipc=1); fulmov(electrons, ipc=1)      -> folded moments, wkix/wkih
        COMMON /fields/ <- f_corr;  fulmov(ions, ipc=0); fulmov(electrons, ipc=0)      -> particles, ranfb
    executed by `nranks` simulated MPI ranks (every rank holds all particles, touches l = rank+1, rank+1+nranks, ...).
    particles = {ksp: [x,y,z,vx,vy,vz]} (not modified).  Returns a dict of per-step results and the final state
    (owned subsets merged back into one set of arrays).  `init` runs first, so every COMMON constant is the reference's."""
    mx, my, mz = grid
    n = len(particles[1][0])
    np0 = max(n, 32 * mx * my * mz)           # init's loadpt loads 32 per cell into arrays of np0 (F:8941, Q6)
    out = {"mom": [], "wk_pred": [], "wk_corr": []}
    with RefRun(mx, my, mz, np0, nranks=nranks) as R:
        setup_run(R, box[0], box[1], box[2], dt=dt, aimpl=aimpl, wce_by_wpe=wce_by_wpe, Ez00=Ez00, qspec=qspec, wspec=wspec, nha=nha)
        parts, _, _ = ref_init(R)
        ranfb_after_init = int(R.get("ranfb", "ir", unit="ranfp"))
        if ranfb_in is None:
            ranfb_in = ranfb_after_init
        for r in range(nranks):
            R.arr("ranfb", "ir", r, "ranfp")[0] = ranfb_in
            for k in (1, 2):
                for c in range(6):
                    parts[r][k][c][:n] = particles[k][c]
        it = it0
        for f_pred, f_corr in field_sets:
            R.set("parm1", "it", it, unit="fulmov")
            for name, a in zip(FIELD_NAMES, f_pred):
                R.set("fields", name, a, unit="fulmov")
            mom, wk = {}, {}
            for k in (1, 2):
                xs = [[parts[r][k][c] for r in range(nranks)] for c in range(6)]
                R.call("fulmov", *xs, float(qspec[k - 1]), float(wspec[k - 1]), n, 1, k, IPAR, SIZE)
                mom[k] = [R.get("srimp7", nm, unit="fulmov") for nm in MOMENT_NAMES[k]]
                wk[k] = (float(R.get("wkinel", "wkix", unit="fulmov")), float(R.get("wkinel", "wkih", unit="fulmov")))
            out["mom"].append(mom)
            out["wk_pred"].append(wk)
            for name, a in zip(FIELD_NAMES, f_corr):
                R.set("fields", name, a, unit="fulmov")
            wk = {}
            for k in (1, 2):
                xs = [[parts[r][k][c] for r in range(nranks)] for c in range(6)]
                R.call("fulmov", *xs, float(qspec[k - 1]), float(wspec[k - 1]), n, 0, k, IPAR, SIZE)
                wk[k] = (float(R.get("wkinel", "wkix", unit="fulmov")), float(R.get("wkinel", "wkih", unit="fulmov")))
            out["wk_corr"].append(wk)
            it += 1
        final = {}
        for k in (1, 2):
            final[k] = [np.empty(n) for _ in range(6)]
            for r in range(nranks):
                for c in range(6):
                    final[k][c][r::nranks] = parts[r][k][c][:n][r::nranks]
        out["final"] = final
        out["ranfb"] = [int(R.get("ranfb", "ir", rank=r, unit="ranfp")) for r in range(nranks)]
        out["ranfb_in"] = ranfb_in
        out["edec"] = R.get("parm2", "edec", unit="fulmov")
        out["consts"] = {nm: float(R.get("parm2", nm, unit="fulmov")) for nm in ("hxi", "hyi", "hzi", "xmaxe", "zmaxe", "adt", "hdt", "bxc")}
    return out


class ReferenceLoop:
    """The reference's time cycle (trans, F:664-807) driven from Python around the translated units, by `nranks` >= 2
    simulated ranks (the field solver's halo exchange needs a neighbour, F:6411-6501):

        startup()                        init, the it = 0 moment pass with dt = 0, emfld0, renewal          F:664-706, 796-807
        per step:  begin_step()          it = it + 1; prefld (B predicted from the last E)                 F:749-759
                   fulmov(1)             the reference's own particle path, ions then electrons ...        F:761-766
                   emfild()              implicit field solve: emcoef, cfpsol, bcgstb, wwstb*, sendrev*     F:771
                   fulmov(0)             ... or any other particle path in their place (set_moments)       F:781-786
                   renew()               ex0 <- ex ...                                                     F:796-807
    What `program` sets up before trans is restated here as data: the namelist values (setup_run), the index tables of
    COMMON /array1d/ (F:318-331) and the solver's block ranges np1, np2, nz1, nz2 (F:294-300).  The particle path is
    pluggable: fulmov() runs the reference's; a test may instead compute the moments elsewhere (C oracle, CUDA) from
    fields() and hand them to set_moments(), which is how a drop-in replacement of fulmov is exercised against the
    reference's own field solver."""

    def __init__(self, grid, box, nranks=2, qspec=(1.0, -1.0), wspec=(100.0, 1.0), dt=1.2, aimpl=0.6, wce_by_wpe=0.2,
                 Ez00=0.25e-2, itermx=1, iterfx=150, itersx=150, **setup):
        if nranks < 2:
            raise ValueError("the reference's field solver exchanges halos with a neighbour rank: nranks >= 2")
        mx, my, mz = grid
        if mz % nranks:
            raise ValueError("mz must be a multiple of the number of ranks (kd = mz/npc, param_080A.h)")
        self.grid, self.nranks, self.qspec, self.wspec = grid, nranks, qspec, wspec
        self.R = R = RefRun(mx, my, mz, 32 * mx * my * mz, nranks=nranks)
        setup_run(R, box[0], box[1], box[2], dt=dt, aimpl=aimpl, wce_by_wpe=wce_by_wpe, Ez00=Ez00, qspec=qspec, wspec=wspec, **setup)
        for nm, v in (("itermx", itermx), ("iterfx", iterfx), ("itersx", itersx)):      # rec_3d80A
            R.set("parm1", nm, v, unit="fulmov")
        self.parts, self.npr, _ = ref_init(R)
        kk, jj, ii = np.meshgrid(np.arange(mz), np.arange(my + 1), np.arange(mx), indexing="ij")      # F:318-331
        for r in range(nranks):
            R.arr("array1d", "arrayx", r, "emfild")[:] = ii.ravel()
            R.arr("array1d", "arrayy", r, "emfild")[:] = jj.ravel()
            R.arr("array1d", "arrayz", r, "emfild")[:] = kk.ravel()
        kd = mz // nranks                                                                             # F:294-300
        self.np1 = np.array([k * 3 * mx * (my + 1) * kd + 1 for k in range(nranks)], dtype=np.int32)
        self.np2 = np.array([(k + 1) * 3 * mx * (my + 1) * kd for k in range(nranks)], dtype=np.int32)
        self.nz1 = np.array([k * kd + 1 for k in range(nranks)], dtype=np.int32)
        self.nz2 = np.array([(k + 1) * kd for k in range(nranks)], dtype=np.int32)
        self.it = 0

    def close(self):
        self.R.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- state ---------------------------------------------------------------------------------------
    def scalar(self, name):
        return float(self.R.get("parm2", name, unit="fulmov"))

    def ranfb(self):
        return [int(self.R.get("ranfb", "ir", rank=r, unit="ranfp")) for r in range(self.nranks)]

    def fields(self, rank=0):
        return [self.R.get("fields", nm, rank=rank, unit="fulmov") for nm in FIELD_NAMES]

    def ranks_agree(self):
        f0 = self.fields(0)
        return all(np.array_equal(a, b) for r in range(1, self.nranks) for a, b in zip(f0, self.fields(r)))

    def moments(self, ksp):
        return [self.R.get("srimp7", nm, unit="fulmov") for nm in MOMENT_NAMES[ksp]]

    def set_moments(self, ksp, mom4):
        """the summed, folded moments of one species into every rank's COMMON /srimp7/ (what fulmov leaves, F:1377-1386)"""
        for nm, a in zip(MOMENT_NAMES[ksp], mom4):
            self.R.set("srimp7", nm, a, unit="fulmov")

    def particles(self):
        """{ksp: [x,y,z,vx,vy,vz]} with every rank's owned subset l = rank+1, rank+1+N, ... merged (F:1162)"""
        n, N = self.npr, self.nranks
        out = {}
        for k in (1, 2):
            out[k] = [np.empty(n) for _ in range(6)]
            for r in range(N):
                for c in range(6):
                    out[k][c][r::N] = self.parts[r][k][c][:n][r::N]
        return out

    # -- the cycle -----------------------------------------------------------------------------------
    def fulmov(self, ipc):
        R, N = self.R, self.nranks
        wk = {}
        for k in (1, 2):
            xs = [[self.parts[r][k][c] for r in range(N)] for c in range(6)]
            R.call("fulmov", *xs, float(self.qspec[k - 1]), float(self.wspec[k - 1]), self.npr, ipc, k, IPAR, SIZE)
            wk[k] = (float(R.get("wkinel", "wkix", unit="fulmov")), float(R.get("wkinel", "wkih", unit="fulmov")))
        return wk

    def renew(self):
        mx, my, mz = self.grid
        for r in range(self.nranks):
            for a, b in zip(FIELD_NAMES[:6], FIELD_NAMES[6:]):
                src = self.R.arr("fields", a, r, "fulmov").reshape(mz + 4, my + 3, mx + 4)
                dst = self.R.arr("fields", b, r, "fulmov").reshape(mz + 4, my + 3, mx + 4)
                dst[2:mz + 2, 1:my + 2, 2:mx + 2] = src[2:mz + 2, 1:my + 2, 2:mx + 2]

    def startup(self, particle_pass=None):
        """F:664-706; particle_pass(loop) may replace the it = 0 pair of fulmov calls (it must call set_moments)"""
        R = self.R
        sav = {nm: self.scalar(nm) for nm in ("dt", "adt", "hdt")}
        for nm in sav:
            R.set("parm2", nm, 0.0, unit="fulmov")
        R.set("parm1", "it", 0, unit="fulmov")
        if particle_pass is None:
            self.fulmov(1)
        else:
            particle_pass(self)
        R.call("emfld0")
        for nm, v in sav.items():
            R.set("parm2", nm, v, unit="fulmov")
        self.renew()

    def begin_step(self):
        self.it += 1
        self.R.set("parm1", "it", self.it, unit="fulmov")
        self.R.call("prefld")

    def emfild(self):
        self.R.call("emfild", self.np1, self.np2, self.nz1, self.nz2, IPAR)
