"""Host-side Python mirror of the reference's fulmov interface.

`Fulmov` keeps the call shape of `subroutine fulmov(x,y,z,vx,vy,vz,qmult,
wmult,npr,ipc,ksp,ipar,size)` (F:1044 of @mrg37-080A.f03) and of the COMMON
values it reads (a `Common` object stands for /fields/, /srimp7/, /parm1/,
/parm2/, /wkinel/, /profl/, /ranfb/), and drives the CUDA library through the
C ABI.  Particles stay resident in HBM between calls.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import StepParams, as_dp, check

FIELD_NAMES = ("ex", "ey", "ez", "bx", "by", "bz", "ex0", "ey0", "ez0", "bx0", "by0", "bz0")


def mxyzA(mx, my, mz):
    """param_080A.h:33"""
    return (mx + 4) * (my + 3) * (mz + 4)


def owned_count(npr, ipar, size):
    """Number of l in `do l= ipar,npr,size` (F:1162); ipar = rank+1 (F:219)."""
    return 0 if npr < ipar else (npr - ipar) // size + 1


def owned_slice(ipar, size):
    """0-based numpy slice of the particles rank ipar-1 owns."""
    return slice(ipar - 1, None, size)


def broadcast_unique_id(rank, make_id=None, device=None):
    """Rank 0 creates the NCCL unique id, torch.distributed carries the 128
    bytes to every rank (the Fortran host would MPI_Bcast them).  Works on any
    initialised backend (gloo: CPU tensor, nccl: pass the cuda device)."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(capi.UNIQUE_ID_BYTES, dtype=torch.uint8, device=device if device is not None else "cpu")
    if rank == 0:
        raw = (make_id or MrgContext.unique_id)()
        buf.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


class Common:
    """The COMMON-block values fulmov reads/writes (F:1066-1110), as numpy."""

    def __init__(self, mx, my, mz, xmax, ymax, zmax, dt=1.2, aimpl=0.6, wce_by_wpe=0.2,
                 Ez00=0.25e-2, nha=5):
        self.mx, self.my, self.mz = mx, my, mz
        self.xmax, self.ymax, self.zmax = float(xmax), float(ymax), float(zmax)
        n = mxyzA(mx, my, mz)
        for name in FIELD_NAMES:                                   # common/fields/
            setattr(self, name, np.zeros(n))
        for name in ("qix", "qiy", "qiz", "qex", "qey", "qez", "qi", "qe"):   # common/srimp7/
            setattr(self, name, np.zeros(n))
        self.it, self.ldec, self.nha = 0, 1, nha                   # common/parm1/
        self.ifilx = self.ifily = self.ifilz = 1                   # F:368-370
        self.dt, self.aimpl = float(dt), float(aimpl)              # common/parm2/
        self.adt, self.hdt = aimpl * dt, 0.5 * dt                  # F:8579-8580
        self.bxc, self.byc, self.bzc = float(wce_by_wpe), 0.0, 0.0  # F:8601-8603
        self.edec = np.zeros((12, 3000))                           # edec(3000,12) column-major
        self.wkix = self.wkih = 0.0                                # common/wkinel/
        self.Ez00 = float(Ez00)                                    # common/profl/, F:9001-9006
        self.zcent, self.ycent1, self.ycent2 = 0.5 * zmax, 0.30 * ymax, 0.70 * ymax
        self.ranfb = 7331                                          # common/ranfb/, F:553
        self.io_pe = 1

    def step_params(self, drive_on=True):
        return StepParams(self.dt, self.adt, self.hdt, self.aimpl, self.bxc, self.byc, self.bzc,
                          self.ifilx, self.ifily, self.ifilz, 1 if drive_on else 0,
                          self.Ez00, self.zcent, self.ycent1, self.ycent2)

    def fields(self):
        return [getattr(self, n) for n in FIELD_NAMES]


class MrgContext:
    """Thin object wrapper over mrg_ctx (one per rank / GPU)."""

    def __init__(self, mx, my, mz, xmax, ymax, zmax, nspecies=2, rank=0, nranks=1, device=0):
        self.lib = capi.load()
        self.h = C.c_void_p()
        check(self.lib.mrg_create(C.byref(self.h), mx, my, mz, xmax, ymax, zmax, nspecies, rank, nranks, device))
        self.mx, self.my, self.mz = mx, my, mz
        self.n_grid = mxyzA(mx, my, mz)
        self.rank, self.nranks = rank, nranks

    def close(self):
        if self.h:
            self.lib.mrg_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- communicator ------------------------------------------------------
    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(capi.UNIQUE_ID_BYTES)
        check(capi.load().mrg_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, uid):
        check(self.lib.mrg_comm_init(self.h, uid))

    def map_peers(self, nspecies=2, device=None):
        """NVLink peer memory for the slab-wise moment exchange: all-gather the cudaIpc handles of the raw-moment arrays
        over torch.distributed (a Fortran host would MPI_Allgather the 64 bytes) and map every peer's array.  Collective."""
        import torch
        import torch.distributed as dist
        if self.nranks == 1:
            return
        for ksp in range(1, nspecies + 1):
            buf = C.create_string_buffer(64)
            check(self.lib.mrg_peer_export(self.h, ksp, buf))
            mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(device if device is not None else "cpu")
            allh = [torch.zeros_like(mine) for _ in range(self.nranks)]
            dist.all_gather(allh, mine)
            for r in range(self.nranks):
                if r != self.rank:
                    check(self.lib.mrg_peer_import(self.h, ksp, r, bytes(allh[r].cpu().numpy().tobytes())))

    def peer_pushes(self, reset=False):
        return self.lib.mrg_peer_pushes(self.h, 1 if reset else 0)

    def split_pushes(self, reset=False):
        """predictor calls that ran as a split launch with an early partial push (option split_push)"""
        return self.lib.mrg_split_pushes(self.h, 1 if reset else 0)

    # -- particles ---------------------------------------------------------
    def upload(self, ksp, x, y, z, vx, vy, vz, first=1, stride=1):
        check(self.lib.mrg_upload_particles(self.h, ksp, *[as_dp(a) for a in (x, y, z, vx, vy, vz)],
                                            len(x), first, stride))

    def download(self, ksp, npr, first=1, stride=1, out=None):
        arrs = out if out is not None else [np.zeros(npr) for _ in range(6)]
        check(self.lib.mrg_download_particles(self.h, ksp, *[as_dp(a) for a in arrs], npr, first, stride))
        return arrs

    def num_local(self, ksp):
        return self.lib.mrg_num_local(self.h, ksp)

    def loadpt(self, ksp, ppc, vth, vdr, vbeam, ranfa=3021, ranfb=7331):
        a, b = C.c_int32(ranfa), C.c_int32(ranfb)
        check(self.lib.mrg_loadpt(self.h, ksp, ppc, vth, vdr, vbeam, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- fields ------------------------------------------------------------
    def set_fields(self, f12, mask=0xFFF):
        arr = (capi.dp * 12)(*[as_dp(a) for a in f12])
        check(self.lib.mrg_set_fields(self.h, mask, arr))

    def set_fields_device(self, ptrs, mask=0xFFF):
        arr = (C.c_void_p * 12)(*ptrs)
        check(self.lib.mrg_set_fields_device(self.h, mask, arr))

    def bind_fields_device(self, ptrs, mask=0xFFF):
        arr = (C.c_void_p * 12)(*ptrs)
        check(self.lib.mrg_bind_fields_device(self.h, mask, arr))

    def set_fields_lazy(self, f12, mask=0xFFF):
        arr = (capi.dp * 12)(*[as_dp(a) for a in f12])
        check(self.lib.mrg_set_fields_lazy(self.h, mask, arr))

    def renew_fields(self, old6=None):
        if old6 is None:
            check(self.lib.mrg_renew_fields(self.h))
        else:
            arr = (capi.dp * 6)(*[as_dp(a) for a in old6])
            check(self.lib.mrg_renew_fields_host(self.h, arr))

    def prefld(self, dt, aimpl):
        """entry prefld of emfild (F:3820-3873) on the device copies: bx,by,bz <- b0 - dt curl(ea)"""
        check(self.lib.mrg_prefld(self.h, dt, aimpl))

    def update_b(self, dt, aimpl, smooth):
        """bx,by,bz as emfild leaves them after its solve (F:4238-4302), from the device's new ex,ey,ez"""
        check(self.lib.mrg_update_b(self.h, dt, aimpl, 1 if smooth else 0))

    def get_fields(self, mask=0xFFF):
        """the device copies of COMMON /fields/ ({index: array} for the members in mask)"""
        out = {k: np.zeros(self.n_grid) for k in range(12) if (mask >> k) & 1}
        arr = (capi.dp * 12)(*[as_dp(out[k]) if k in out else None for k in range(12)])
        check(self.lib.mrg_get_fields(self.h, mask, arr))
        return out

    def prepared_fields(self, params):
        out = [np.zeros(self.n_grid) for _ in range(6)]
        arr = (capi.dp * 6)(*[as_dp(a) for a in out])
        check(self.lib.mrg_get_prepared_fields(self.h, C.byref(params), arr))
        return out

    # -- the hot path --------------------------------------------------------
    def fulmov(self, ksp, qmult, wmult, ipc, params, ranfb=7331):
        st = C.c_int32(ranfb)
        wkix, wkih = C.c_double(), C.c_double()
        check(self.lib.mrg_fulmov(self.h, ksp, qmult, wmult, ipc, C.byref(params), C.byref(st),
                                  C.byref(wkix), C.byref(wkih)))
        return wkix.value, wkih.value, st.value

    def fulmov_deferred(self, ksp, qmult, wmult, params):
        """ipc=1 call in deferred mode (option "defer"): returns a pair of c_double that the library fills
        with wkix/wkih when the host next waits for the species (moments(), synchronize())."""
        st = C.c_int32(0)
        wk = (C.c_double(), C.c_double())
        check(self.lib.mrg_fulmov(self.h, ksp, qmult, wmult, 1, C.byref(params), C.byref(st),
                                  C.byref(wk[0]), C.byref(wk[1])))
        return wk

    def pass_ms(self, ksp, ipc):
        ms = C.c_double()
        check(self.lib.mrg_pass_ms(self.h, ksp, ipc, C.byref(ms)))
        return ms.value

    def prep_stats(self, reset=False):
        out = (C.c_int64 * 4)()
        check(self.lib.mrg_get_prep_stats(self.h, out, 1 if reset else 0))
        return {"preps": out[0], "restricted": out[1], "planes": out[2], "compact_sums": out[3]}

    def moments(self, ksp, folded=True, out=None):
        arrs = out if out is not None else [np.zeros(self.n_grid) for _ in range(4)]
        check(self.lib.mrg_get_moments(self.h, ksp, *[as_dp(a) for a in arrs], 1 if folded else 0))
        return arrs

    def moments_device(self, ksp):
        arr = (C.c_void_p * 4)()
        check(self.lib.mrg_get_moments_device(self.h, ksp, arr))
        return list(arr)

    def sort(self, ksp, lookahead=0.0):
        check(self.lib.mrg_sort(self.h, ksp, lookahead))

    def set_option(self, name, value):
        check(self.lib.mrg_set_option(self.h, name.encode(), int(value)))

    def counters(self, reset=False):
        out = (C.c_int64 * 3)()
        check(self.lib.mrg_get_counters(self.h, out, 1 if reset else 0))
        return {"launches": out[0], "h2d_bytes": out[1], "d2h_bytes": out[2]}

    def last_kernel_ms(self):
        ms = C.c_double()
        check(self.lib.mrg_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def synchronize(self):
        check(self.lib.mrg_synchronize(self.h))

    def self_check(self, ksp):
        """invariants of the resident state (mrg_self_check): raw moment sums, particle count, cell-index end,
        and whether the slots still hold a permutation of the original indices"""
        sums = (C.c_double * 4)()
        cnt = (C.c_int64 * 4)()
        check(self.lib.mrg_self_check(self.h, ksp, sums, cnt))
        n = cnt[0]
        m64 = (1 << 64) - 1
        want1 = (n * (n - 1) // 2) & m64
        want2 = ((n - 1) * n * (2 * n - 1) // 6) & m64 if n > 0 else 0
        return {"sums": list(sums), "n": n, "cell_end_last": cnt[1],
                "permutation_ok": (cnt[2] & m64) == want1 and ((cnt[3] & m64) == want2 or n <= 0)}

    def phase_ms(self, reset=False):
        """device milliseconds per phase of the mrg_fulmov calls since the last reset (option "phases" must be 1)"""
        out = (C.c_double * 6)()
        calls = C.c_int64()
        check(self.lib.mrg_phase_ms(self.h, out, C.byref(calls), 1 if reset else 0))
        return dict(zip(("prep", "setup", "kernel", "rank_sum", "fold", "kick"), out)), calls.value

    def phase_detail(self, ksp, ipc):
        """the same timers for one species and kind of call; the rank sum of ipc >= 1 calls split into strips / push / barrier"""
        out = (C.c_double * 9)()
        check(self.lib.mrg_phase_detail(self.h, ksp, ipc, out))
        return dict(zip(("prep", "setup", "kernel", "rank_sum", "fold", "kick", "strips", "push", "barrier"), out))

    def dfma_peak(self):
        v = C.c_double()
        check(self.lib.mrg_dfma_peak(self.h, C.byref(v)))
        return v.value

    def event_record(self, slot):
        check(self.lib.mrg_event_record(self.h, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_double()
        check(self.lib.mrg_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value


class Fulmov:
    """Drop-in for the reference's `fulmov` on one rank.

    fm = Fulmov(common, ipar, size); then, as in trans (F:761-787):
        fm(xi,yi,zi,vxi,vyi,vzi, qspec1, wspec1, npr, ipc, 1)
        fm(xe,ye,ze,vxe,vye,vze, qspec2, wspec2, npr, ipc, 2)
    Side effects mirror the subroutine: qix..qi / qex..qe in `common` after
    ipc>=1 (F:1377-1386), wkix/wkih (F:1316-1317), edec(ldec,5..8) when
    mod(it,nha)==0 on io_pe==1 (F:1320-1328), ranfb advanced by the kick.
    The host particle arrays are only read on the first call per species (or
    after particles_changed); use pull() before host code looks at them.

    Field traffic.  Without hints every ksp==1 call uploads all of COMMON /fields/ (the mirror cannot see
    what the host changed).  A host that marks its three field updates avoids most of it:
        fields_changed(MASK_B)      after prefld (F:759, writes bx,by,bz)
        fields_changed(MASK_NEW)    after emfild (F:771, writes ex..bz)
        fields_renewed()            after the renewal loop ex0 <- ex (F:796-807): done on the device
    With `defer=True` the two ipc=1 calls only queue their work: the NCCL sum, the fold and the copy of
    the moments into COMMON /srimp7/ of one species overlap the particle kernel of the next;
    `finish_moments()` (call it before emfild reads /srimp7/) waits and fills wkix/wkih/edec.
    """

    MASK_NEW, MASK_B, MASK_OLD, MASK_ALL = 0x03F, 0x038, 0xFC0, 0xFFF

    def __init__(self, common, ipar=1, size=1, device=0, uid=None, sort_interval=1, ctx=None, hints=False,
                 defer=False, lazy=False, share_moments=False, nspecies=2):
        self.c = common
        self.ipar, self.size = ipar, size
        self.nspecies = nspecies      # the reference handles ksp = 1|2 (F:1321-1327); more only on request (qspec(4), F:1100)
        self.extra = {}               # moment arrays of species 3, 4: the reference has no COMMON member for them
        self.resident = {}
        if ctx is not None:      # adopt a context whose particles are already resident (device loader)
            self.ctx = ctx
            self.resident = {k: ctx.num_local(k) for k in range(1, nspecies + 1) if ctx.num_local(k) > 0}
        else:
            self.ctx = MrgContext(common.mx, common.my, common.mz, common.xmax, common.ymax, common.zmax,
                                  nspecies=nspecies, rank=ipar - 1, nranks=size, device=device)
            if size > 1:
                if uid is None:
                    raise ValueError("size > 1 needs the NCCL unique id broadcast from rank 0")
                self.ctx.comm_init(uid)
        self.hints = hints
        # lazy: COMMON /fields/ is not uploaded when it changes; each field preparation fetches the planes it reads
        # (a rank that owns a z slab then uploads its slab instead of the replicated arrays).  share_moments: the
        # COMMON /srimp7/ arrays are shared by the ranks of the node and every rank delivers its own z block.
        self.lazy = lazy
        self.ctx.set_option("sink_share", 1 if share_moments else 0)
        self.dirty = self.MASK_ALL
        self.renew = False
        self.b_pending = None
        self.it0 = False
        self.sort_interval = sort_interval
        self.ncorr = {k: 0 for k in range(1, 5)}
        self.defer = defer
        self.inflight = {}
        self.ctx.set_option("defer", 1 if defer else 0)
        self.sinks = {}

    def fields_changed(self, mask=0xFFF):
        self.dirty |= mask

    def prefld_done(self):
        """The host has called prefld (F:759, rewrites bx,by,bz).  With hints on and whole device arrays the same entry is
        repeated on the device (mrg_prefld, bit-identical) instead of uploading the three arrays -- with lazily held fields on
        exactly the planes each preparation reads; without hints this is fields_changed(MASK_B)."""
        if self.hints:
            self.b_pending = False
        else:
            self.dirty |= self.MASK_B

    def emfild_done(self):
        """The host has called emfild (F:771, rewrites ex..bz).  With hints on and whole device arrays only ex,ey,ez are
        uploaded; bx,by,bz are recomputed on the device exactly as emfild does behind its solve (F:4238-4302, smoothed on
        the steps with mod(it,5) = 1); otherwise this is fields_changed(MASK_NEW)."""
        smooth = self.c.it % 5 == 1
        if self.hints and not (self.lazy and smooth):      # the smoothing needs whole arrays: lazily held fields upload B then
            self.dirty |= 0x007
            self.b_pending = smooth
        else:
            self.dirty |= self.MASK_NEW

    def fields_renewed(self):
        """The host has copied ex..bz into ex0..bz0 (F:796-807)."""
        if self.hints:
            self.renew = True
        else:
            self.dirty |= self.MASK_OLD

    def particles_changed(self, ksp):
        self.resident.pop(ksp, None)

    def _moment_arrays(self, ksp):
        c = self.c
        if ksp > 2:
            if ksp not in self.extra:
                self.extra[ksp] = [np.zeros(self.ctx.n_grid) for _ in range(4)]
            return self.extra[ksp]
        return [c.qix, c.qiy, c.qiz, c.qi] if ksp == 1 else [c.qex, c.qey, c.qez, c.qe]

    def _push_fields(self, ksp):
        if not self.hints and ksp == 1:
            self.dirty = self.MASK_ALL
        # it = 0 (F:664-706): after the dt = 0 pair, emfld0 rewrites all of COMMON /fields/ (F:691) and the renewal loop
        # runs; the three optional marks do not describe that, so the first call afterwards uploads everything and drops
        # the pending device renewal
        if self.c.it == 0:
            self.it0 = True
        elif self.it0:
            self.it0, self.dirty, self.renew = False, self.MASK_ALL, False
        if self.renew:
            self.ctx.renew_fields(self.c.fields()[6:] if self.lazy else None)
            self.renew = False
            self.dirty &= ~self.MASK_OLD
        if self.b_pending is not None:
            self.dirty &= ~self.MASK_B               # computed below from what the device holds
        if self.dirty:
            (self.ctx.set_fields_lazy if self.lazy else self.ctx.set_fields)(self.c.fields(), mask=self.dirty)
            self.dirty = 0
        if self.b_pending is not None:               # None = nothing to do, False = prefld, True = + the smoothing of emfild
            self.ctx.update_b(self.c.dt, self.c.aimpl, self.b_pending)
            self.b_pending = None

    def _record_wk(self, ksp, wkix, wkih):
        c = self.c
        c.wkix, c.wkih = wkix, wkih
        if c.it % c.nha == 0 and c.io_pe == 1 and ksp <= 2:
            col = 5 if ksp == 1 else 7
            c.edec[col - 1, c.ldec - 1] = wkix
            c.edec[col, c.ldec - 1] = wkih

    def finish_moments(self):
        """Deferred mode: wait for the queued ipc=1 calls; COMMON /srimp7/, wkix/wkih, edec are then set."""
        for ksp in sorted(self.inflight):
            wk = self.inflight[ksp]
            self.ctx.moments(ksp, folded=True, out=self._moment_arrays(ksp))
            self._record_wk(ksp, wk[0].value, wk[1].value)
        self.inflight = {}

    def __call__(self, x, y, z, vx, vy, vz, qmult, wmult, npr, ipc, ksp):
        c = self.c
        if not 1 <= ksp <= self.nspecies:
            raise ValueError("ksp must be 1 or 2 (F:1321-1327) unless the mirror was created with nspecies > 2")
        if not self.resident.get(ksp):
            self.ctx.upload(ksp, x[:npr], y[:npr], z[:npr], vx[:npr], vy[:npr], vz[:npr], self.ipar, self.size)
            self.resident[ksp] = npr
        if ipc == 0 and self.inflight:
            self.finish_moments()
        self._push_fields(ksp)
        if ipc >= 1 and self.defer:
            out = self._moment_arrays(ksp)
            key = tuple(a.ctypes.data for a in out)
            if self.sinks.get(ksp) != key:
                check(self.ctx.lib.mrg_set_moment_sink(self.ctx.h, ksp, *[as_dp(a) for a in out]))
                self.sinks[ksp] = key
            self.inflight[ksp] = self.ctx.fulmov_deferred(ksp, qmult, wmult, c.step_params())
            return
        wkix, wkih, c.ranfb = self.ctx.fulmov(ksp, qmult, wmult, ipc, c.step_params(), c.ranfb)
        self._record_wk(ksp, wkix, wkih)
        if ipc >= 1:
            self.ctx.moments(ksp, folded=True, out=self._moment_arrays(ksp))
        else:
            self.ncorr[ksp] += 1
            if self.sort_interval and self.ncorr[ksp] % self.sort_interval == 0:
                self.ctx.sort(ksp, c.hdt)     # key = cell of the next gather position x + hdt*v

    def pull(self, ksp, x, y, z, vx, vy, vz, npr):
        self.ctx.download(ksp, npr, self.ipar, self.size, out=[x, y, z, vx, vy, vz])
