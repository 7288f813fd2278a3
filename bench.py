#!/usr/bin/env python
"""Benchmark of the /fulmov/ particle hot path (BASELINE.json metric:
particle-pushes/sec incl. J/chi deposition, % of the HBM roofline).

One "step" = one time step of the particle path for BOTH species of the
synthetic two-flux-bundle load, in the call order of trans (F:761-787):
  predictor  fulmov(ions, ipc=1), fulmov(electrons, ipc=1)   gather+push+deposit, NCCL sum, fold
  [fields change: emfild]                                     field prep runs again
  corrector  fulmov(ions, ipc=0), fulmov(electrons, ipc=0)   gather+push+partbc+drive kick
  cell sort of both species every --sort-every steps (maintenance, inside the timed region)
particle-pushes/s = particles of all species x steps / time  (one push = predictor + corrector pass).

  python bench.py [--gpus N --steps K --warmup W]            CUDA path (this repository)
  python bench.py --impl reference ...                        the reference's CPU path (C restatement,
                                                              all host threads, bounded sample)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HX, HY, HZ = 300.0 / 40, 600.0 / 72, 300.0 / 40        # rec_3d80A:4 with param_080A.h:14
QSPEC, WSPEC = {1: 1.0, 2: -1.0}, {1: 100.0, 2: 1.0}   # rec_3d80A:5-6
VBEAM = {1: 0.35e-2, 2: -0.35e-2}
VETH, DT, AIMPL, WCE, EZ00 = 0.2, 1.2, 0.6, 0.2, 0.25e-2
# weak scaling: 268 M particles (both species) per GPU, 64 ppc; N=1 is BASELINE configs[1], N=8 is configs[2]
GRIDS = {1: (128, 128, 128), 2: (256, 128, 128), 4: (256, 256, 128), 8: (256, 256, 256)}
METRIC = "particle-pushes/sec incl. J/chi deposition"
UNIT = "particle-pushes/s"


SPECIES = (1, 2)      # ksp values of a step; --config 5 adds a heavy positive third species (qspec(4), wspec(4): F:1100)


def vth(ksp):
    return VETH / np.sqrt(1.0 * WSPEC[ksp]) if QSPEC[ksp] > 0 else VETH       # F:8589-8594 (te_by_ti = 1)


def apply_config(args):
    """BASELINE.json configs by number (1-based as in BASELINE.md): 2 = default (128^3 x 64 ppc per GPU, weak: 256^3 at 8);
    4 = weak sweep 256x128x128 per GPU, 100 ppc; 5 = large-Dt run 512x256x256 (whole job), dt*wce > 10, heavy third species."""
    global WCE, SPECIES
    if args.config == 4:
        args.grid = args.grid or [256, 128, 128 * max(args.gpus, 1)]
        args.ppc = args.ppc or 100
    elif args.config == 5:
        args.grid = args.grid or [512, 256, 256]
        args.ppc = args.ppc or 16
        WCE = 9.0                                   # dt * wce = 10.8
        QSPEC[3], WSPEC[3], VBEAM[3] = 1.0, 1600.0, 0.0
        SPECIES = (1, 2, 3)
    args.ppc = args.ppc or 64


def synth_fields(xp, mx, my, mz, seed, **kw):
    """12 smooth flux-bundle-like field arrays (ex..bz, ex0..bz0) in the reference layout
    (E ~ 1e-2, B ~ 3e-2 on top of bxc = 0.2, the magnitudes of EMfields.pdf).  xp = numpy or torch."""
    rs = np.random.RandomState(seed)
    i = xp.arange(-2, mx + 2, **kw).reshape(1, 1, -1)
    j = xp.arange(-1, my + 2, **kw).reshape(1, -1, 1)
    k = xp.arange(-2, mz + 2, **kw).reshape(-1, 1, 1)
    X, Y, Z = 2 * np.pi * i / mx, np.pi * j / my, 2 * np.pi * k / mz
    out = []
    for c in range(12):
        amp = 1e-2 if (c % 6) < 3 else 3e-2
        a = rs.normal(size=4)
        ph = rs.uniform(0, 2 * np.pi, size=3)
        f = amp * (a[0] * xp.sin(X + ph[0]) * xp.cos(Y) + a[1] * xp.cos(2 * Z + ph[1]) * xp.sin(Y)
                   + a[2] * xp.sin(X + Z + ph[2]) * xp.cos(2 * Y) + 0.3 * a[3] * xp.cos(3 * X) * xp.sin(Z) * xp.cos(Y))
        out.append(f.reshape(-1))
    return out


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc = device, None
        self.path = tempfile.mktemp(prefix="mrg_clocks_", suffix=".csv")

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            parts = [t.strip() for t in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def _mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return float(line.split()[1]) / 1e6
    except OSError:
        pass
    return 16.0


CPU_SAMPLE = {"grid": (64, 64, 64), "ppc": 32}      # bounded sample of the workload for the CPU legs (see cpu_reference_rate)


def cpu_sample_text():
    g, ppc = CPU_SAMPLE["grid"], CPU_SAMPLE["ppc"]
    return ("CPU legs run a bounded sample of this workload: %dx%dx%d grid, %d ppc/species (the per-cell count loadpt hard-codes, "
            "F:8941), same box proportions, dt, aimpl, species; every simulated MPI rank holds all particles as the reference "
            "does (F:121-122), so the full 128^3 x 64 ppc load does not fit a host" % (g[0], g[1], g[2], ppc))


def cpu_reference_rate(steps, warmup, nthreads=None):
    """The reference's CPU path on a bounded sample of the workload.

    kind "reference": the reference's own init/loadpt/fulmov/srimp1/srimp2/... (oracle/_ref: @mrg37-080A.f03 translated to C by
    oracle/f03c.py because no Fortran compiler exists, built by oracle/build_ref.py), one thread per simulated MPI rank with
    private COMMON storage and private copies of all particles, field preparation inside every fulmov call and the
    mpi_allreduce of the moments -- exactly what `mpiexec -n N` of the reference does per step, minus the field solve.
    kind "port" (only when oracle/_ref is not there): the C restatement oracle/fulmov_oracle.c.
    Returns a dict: value (particle-pushes/s), per-step seconds, ful(1)/ful(0) split (the reference's own timer split,
    F:813-822), threads, kind, description."""
    if not nthreads:      # all the host cores this process may use, whatever OMP_NUM_THREADS says (torchrun sets it to 1)
        nthreads = min(len(os.sched_getaffinity(0)), 64)
    mx, my, mz = CPU_SAMPLE["grid"]
    ppc = CPU_SAMPLE["ppc"]
    try:
        from oracle import pyref as PR
        have_ref = PR.available()
    except Exception:
        have_ref = False
    t_pred, t_corr = [], []
    t_em = []
    if have_ref:
        # the reference's own time cycle (oracle/pyref.ReferenceLoop): start-up solve, then per step prefld -> fulmov x2 ->
        # emfild (implicit field solve) -> fulmov x2 -> renewal, so every fulmov call sees the reference's own fields and the
        # step splits into the reference's three timers ful(1), em, ful(0) (F:813-822).  Only the two ful() parts are the path.
        np0 = 32 * mx * my * mz                          # init loads 32 per cell (F:8941)
        per_rank_gb = 18 * np0 * 8 / 1e9 + 60 * 3 * mx * (my + 1) * mz * 8 / 1e9 + 0.2      # particles + the solver's band arrays
        limit = int(max(2, min(nthreads, 0.6 * _mem_available_gb() / per_rank_gb)))
        nranks = max(d for d in range(2, limit + 1) if mz % d == 0) if any(mz % d == 0 for d in range(2, limit + 1)) else 2
        with PR.ReferenceLoop((mx, my, mz), (HX * mx, HY * my, HZ * mz), nranks, qspec=(QSPEC[1], QSPEC[2]), wspec=(WSPEC[1], WSPEC[2]),
                              dt=DT, aimpl=AIMPL, wce_by_wpe=WCE, Ez00=EZ00, veth=VETH, vbeam=(VBEAM[1], VBEAM[2])) as A:
            npr = A.npr
            A.startup()
            for s in range(warmup + steps):
                A.begin_step()
                t0 = time.perf_counter()
                A.fulmov(1)
                t1 = time.perf_counter()
                A.emfild()
                t2 = time.perf_counter()
                A.fulmov(0)
                t3 = time.perf_counter()
                A.renew()
                if s >= warmup:
                    t_pred.append(t1 - t0)
                    t_em.append(t2 - t1)
                    t_corr.append(t3 - t2)
        npart, kind, nthreads = 2 * npr, "reference", nranks
        what = ("the reference's own time cycle from @mrg37-080A.f03 (init/loadpt, emfld0, prefld, fulmov with partbc, srimp1/2, outmesh3, "
                "filt3e, vmesh, and emfild/cfpsol/bcgstb between the fulmov pairs), translated to C (oracle/f03c.py: no Fortran compiler "
                "in the image), gcc -O2, %d simulated MPI ranks = %d threads; timed: the four fulmov calls of a step" % (nranks, nranks))
    else:
        from oracle import pyoracle as O
        fa = [np.ascontiguousarray(a, dtype=np.float64) for a in synth_fields(np, mx, my, mz, 1, dtype=np.float64)]
        fb = [np.ascontiguousarray(a, dtype=np.float64) for a in synth_fields(np, mx, my, mz, 2, dtype=np.float64)]
        O.set_num_threads(nthreads)
        nthreads = min(O.num_threads(), nthreads)
        p = O.make_parm(mx, my, mz, HX * mx, HY * my, HZ * mz, DT, AIMPL, WCE, EZ00)
        sp = {}
        for ksp in (1, 2):
            sp[ksp], _, _ = O.loadpt(p, ppc, vth(ksp), 0.0, VBEAM[ksp])
        for c in range(3):
            sp[2][c][:] = sp[1][c]
        npart = 2 * len(sp[1][0])
        for s in range(warmup + steps):
            tp = tc = 0.0
            for ksp in (1, 2):       # the reference prepares the fields inside every call (F:1127-1148)
                t0 = time.perf_counter()
                a6p = O.field_prep(p, fa)
                a6c = O.field_prep(p, fb)
                half = 0.5 * (time.perf_counter() - t0)
                _, t1, t0c = O.time_step(p, a6p, a6c, sp[ksp], QSPEC[ksp], WSPEC[ksp], nthreads)
                tp += t1 + half
                tc += t0c + half
            if s >= warmup:
                t_pred.append(tp)
                t_corr.append(tc)
        kind = "port"
        what = "C restatement of the reference CPU path (oracle/fulmov_oracle.c), %d OpenMP threads as simulated ranks" % nthreads
    times = [a + b for a, b in zip(t_pred, t_corr)]
    rate = npart * len(times) / sum(times)
    desc = "%dx%dx%d grid, %d ppc, 2 species = %d particles, %d steps; %s" % (mx, my, mz, ppc, npart, len(times), what)
    return {"value": rate, "times": times, "cores": nthreads, "kind": kind, "sample": desc,
            "ful1_s_per_step": float(np.mean(t_pred)), "ful0_s_per_step": float(np.mean(t_corr)),
            "em_s_per_step": float(np.mean(t_em)) if t_em else None}


def cpu_baseline_record(r):
    return {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
            "ful(1)_s_per_step": r["ful1_s_per_step"], "em_s_per_step": r.get("em_s_per_step"), "ful(0)_s_per_step": r["ful0_s_per_step"],
            "note": "ful(1), em, ful(0) = the reference's own timer split of a step (F:813-822); value counts ful(1) + ful(0) only: "
                    "'em' (the implicit field solve) is not on this path"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_rate(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(r["times"])), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, args),
        "cpu_baseline": cpu_baseline_record(r),
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n, args):
    if getattr(args, "slab_of", None):
        n = args.slab_of[0]
    mx, my, mz = GRIDS[n] if not args.grid else tuple(args.grid)
    slab = args.shard == "slab" and n > 1
    if n == 1:
        sharding = "one GPU: no rank sum"
    elif slab:
        sharding = ("z-slab ownership by initial position (NOT the reference's partition): replicated grids; J/chi rank sum = slab-wise "
                    "exchange (ncclSend/Recv of boundary strips + in-place ncclAllGather + ghost-plane ncclBroadcast) whenever every rank "
                    "agrees its deposits stay near its slab, whole-grid ncclAllReduce(fp64) otherwise -- see config.rank_sum for what ran; "
                    "drive-kick draws are taken by particle index (no reference stream exists for this ownership), so the kicked set "
                    "is statistically, not bitwise, the reference's")
    else:
        sharding = ("the reference's round-robin particle ownership l = rank+1 (mod N) (F:1162): replicated grids, one in-place "
                    "ncclAllReduce(fp64) of [qjx|qjy|qjz|q|wkix,wkih] per species per step (F:2379-2384, 2533, 1312-1315), "
                    "drive kick in the reference's serial per-rank ranfp order")
    return {"workload": "two-flux-bundle equilibrium (rec_3d80A / param_080A.h), %dx%dx%d grid, %d ppc/species, "
                        "ions+electrons mi/me=100%s, dt=1.2, aimpl=0.6, wce/wpe=%g (BASELINE configs[%s])"
                        % (mx, my, mz, args.ppc, " + heavy ions q=+1 m=1600" if len(SPECIES) == 3 else "", WCE,
                           {4: "3", 5: "4"}.get(getattr(args, "config", 2), "1" if n == 1 else ("2" if n == 8 else "1 scaled"))),
            "grid": [mx, my, mz], "ppc": args.ppc, "species": len(SPECIES), "particles_per_gpu": len(SPECIES) * mx * my * mz * args.ppc // n,
            "sharding": sharding, "shard": args.shard if n > 1 else "none",
            "sort_every": args.sort_every, "sort_every_ions": args.sort_every_ions or args.sort_every, "deposit": args.deposit, "iters": args.iters, "tile": args.tile,
            "fused_keys": args.fused_keys, "fused_sort": args.fused_sort, "defer": args.defer, "planes": args.planes,
            "l2": "inputs larger than L2 (12.9 GB of particle arrays per GPU vs 126 MB)",
            "cpu_sample": cpu_sample_text()}


def multi_rank_parity(mrg, dist, torch, rank, world, local, dev):
    """Small parity pass on ALL ranks of this job before the timed loop (so that a scaling record carries 2/4/8-rank
    evidence): 16 x 12 x 32N grid, 8 ppc, two steps, both particle partitions.  Checker = the CPU oracle on rank 0
    (pinned bit for bit to the reference's own fulmov, tests/test_ref_pin.py); it is not on any measured path.
      roundrobin  the reference's partition and its serial per-rank kick streams: summed folded moments of step 1 and
                  step 2 (step 2 sees the corrector and the kick of step 1) against the oracle run as N ranks;
      slab        z-slab ownership, slab-wise exchange when the ranks agree: the same with the drive off (that
                  ownership has no reference kick stream)."""
    from oracle import pyoracle as O
    mx, my, mz, ppc = 16, 12, 32 * world, 8
    out = {}
    for part in ("roundrobin", "slab"):
        p = O.make_parm(mx, my, mz, HX * mx, HY * my, HZ * mz, DT, AIMPL, WCE, EZ00 if part == "roundrobin" else 0.0)
        ctx = mrg.MrgContext(mx, my, mz, p.xmax, p.ymax, p.zmax, nspecies=2, rank=rank, nranks=world, device=local)
        ctx.comm_init(mrg.broadcast_unique_id(rank, device=dev))
        ctx.set_option("shard", 1 if part == "slab" else 0)
        if part == "slab":
            ctx.set_option("defer", 1)
            ctx.map_peers(2, device=dev)
        par = mrg.StepParams(p.dt, p.adt, p.hdt, p.aimpl, p.bxc, p.byc, p.bzc, 1, 1, 1, 1, p.Ez00, p.zcent, p.ycent1, p.ycent2)
        st = 7331
        for ksp in (1, 2):
            _, st = ctx.loadpt(ksp, ppc, vth(ksp), 0.0, VBEAM[ksp])
            ctx.sort(ksp, p.hdt)
        if rank == 0:
            sp = {}
            for ksp in (1, 2):
                sp[ksp], _, st_o = O.loadpt(p, ppc, vth(ksp), 0.0, VBEAM[ksp])
            for cc in range(3):
                sp[2][cc][:] = sp[1][cc]
            nr = world if part == "roundrobin" else 1
            sto = np.full(nr, st_o, dtype=np.int32)
        # device loader: electrons share the ions' positions (ipleql, F:8657) by construction of the seeds
        errs = []
        for step in range(2):
            fa = [np.ascontiguousarray(a, dtype=np.float64) for a in synth_fields(np, mx, my, mz, 11 + 2 * step, dtype=np.float64)]
            fb = [np.ascontiguousarray(a, dtype=np.float64) for a in synth_fields(np, mx, my, mz, 12 + 2 * step, dtype=np.float64)]
            ctx.set_fields(fa)
            if rank == 0:
                a6 = O.field_prep(p, fa)
            if part == "slab":       # both predictor calls queued before any moment is read, as in the timed loop
                wkd = [ctx.fulmov_deferred(ksp, QSPEC[ksp], WSPEC[ksp], par) for ksp in (1, 2)]   # filled when the moments are read
            for ksp in (1, 2):
                if part != "slab":
                    ctx.fulmov(ksp, QSPEC[ksp], WSPEC[ksp], 1, par, st)
                mom = ctx.moments(ksp)
                if rank == 0:
                    r = O.fulmov(p, a6, *sp[ksp], QSPEC[ksp], WSPEC[ksp], 1, nranks=nr)
                    errs.append(max(float(np.linalg.norm(mom[cc] - r["mom"][cc]) / np.linalg.norm(r["mom"][cc])) for cc in range(4)))
            ctx.set_fields(fb)
            if rank == 0:
                a6 = O.field_prep(p, fb)
            for ksp in (1, 2):
                _, _, st = ctx.fulmov(ksp, QSPEC[ksp], WSPEC[ksp], 0, par, st)
                ctx.sort(ksp, p.hdt)
                if rank == 0:
                    O.fulmov(p, a6, *sp[ksp], QSPEC[ksp], WSPEC[ksp], 0, nranks=nr, ranfb=sto)
        stats = ctx.prep_stats()
        pushes = ctx.peer_pushes()
        splits = ctx.split_pushes()
        rng_ok = True
        if part == "roundrobin":          # every rank's ranfp state after the reference-order kicks
            t = torch.tensor([float(st)], dtype=torch.float64, device=dev)
            allst = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allst, t)
            if rank == 0:
                rng_ok = [int(v.item()) for v in allst] == [int(v) for v in sto]
        ctx.close()
        if rank == 0:
            out[part] = {"moments_rel_l2_step1": max(errs[:2]), "moments_rel_l2_step2": max(errs[2:]),
                         "ok": bool(max(errs) < 1e-10 and rng_ok), "ranks": world, "grid": [mx, my, mz], "ppc": ppc,
                         "slabwise_sums": int(stats["compact_sums"]), "peer_pushes": int(pushes), "split_launches": int(splits), "ranfp_states_equal_reference": bool(rng_ok) if part == "roundrobin" else None}
        dist.barrier()
    return out


def measure_reference_partition(mrg, dist, torch, args, rank, world, local, dev, grid, params, steps=3, warmup=2):
    """The same resident step on the REFERENCE's partition (round-robin ownership l = rank+1 mod N, F:1162; one whole-grid
    ncclAllReduce per species; the reference's serial kick streams), timed like `value`: a short secondary leg so that every
    N > 1 record carries the number for the reference decomposition next to the z-slab headline."""
    mx, my, mz = grid
    ctx = mrg.MrgContext(mx, my, mz, HX * mx, HY * my, HZ * mz, nspecies=len(SPECIES), rank=rank, nranks=world, device=local)
    ctx.comm_init(mrg.broadcast_unique_id(rank, device=dev))
    for name in ("deposit", "iters", "group_min", "tile", "fused_keys", "fused_sort", "defer"):
        ctx.set_option(name, getattr(args, name))
    ctx.set_option("shard", 0)
    ctx.set_option("compact", 0)
    st = 7331
    for ksp in SPECIES:
        _, st = ctx.loadpt(ksp, args.ppc, vth(ksp), 0.0, VBEAM[ksp])
    fsets = [[t.contiguous() for t in synth_fields(torch, mx, my, mz, seed, dtype=torch.float64, device=dev)] for seed in (1, 2)]
    fptr = [[t.data_ptr() for t in fs] for fs in fsets]
    for ksp in SPECIES:
        ctx.sort(ksp, params.hdt)

    def step():
        nonlocal st
        ctx.bind_fields_device(fptr[0])
        for ksp in SPECIES:
            if args.defer:
                ctx.fulmov_deferred(ksp, QSPEC[ksp], WSPEC[ksp], params)
            else:
                ctx.fulmov(ksp, QSPEC[ksp], WSPEC[ksp], 1, params, st)
        ctx.bind_fields_device(fptr[1])
        for ksp in SPECIES:
            _, _, st = ctx.fulmov(ksp, QSPEC[ksp], WSPEC[ksp], 0, params, st)
        for ksp in SPECIES:
            ctx.sort(ksp, params.hdt)

    for _ in range(warmup):
        step()
    ctx.synchronize(); torch.cuda.synchronize(); dist.barrier()
    ctx.event_record(0)
    for _ in range(steps):
        step()
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1)
    ctx.synchronize(); torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    stats = ctx.prep_stats()
    ctx.close()
    del fsets
    torch.cuda.empty_cache()
    ms = float(t[0])
    return {"value": float(len(SPECIES)) * mx * my * mz * args.ppc * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
            "warmup": warmup, "sharding": "round-robin l = rank+1 (mod N), whole-grid ncclAllReduce(fp64) per species, reference kick order",
            "slabwise_sums": int(stats["compact_sums"])}


def bind_to_gpu_numa(local):
    """Pin this rank's host threads (and hence the first-touch placement of the pinned field / moment arrays it allocates)
    to the CPUs next to its GPU, what `mpiexec --bind-to` / numactl does for an MPI host.  Without it eight ranks' PCIe
    traffic crosses the socket interconnect at random.  Returns the number of CPUs bound to, or 0 when NVML is not usable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    import mrg_b200 as mrg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback exists); use --impl reference for the CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = bind_to_gpu_numa(local) if (world > 1 and args.numa_bind) else 0
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mx, my, mz = GRIDS[args.gpus] if not args.grid else tuple(args.grid)
    if args.slab_of:      # one GPU holds what one rank of an N-GPU job holds (no NCCL): development aid, not a bench line
        mx, my, mz = GRIDS[args.slab_of[0]]
    ppc = args.ppc
    mrp = None
    if world > 1 and not args.no_rank_parity:
        try:
            mrp = multi_rank_parity(mrg, dist, torch, rank, world, local, dev)
        except Exception as ex:
            mrp = {"error": repr(ex)}
    ctx = mrg.MrgContext(mx, my, mz, HX * mx, HY * my, HZ * mz, nspecies=len(SPECIES), rank=rank, nranks=world, device=local)
    if world > 1:
        uid = mrg.broadcast_unique_id(rank, device=dev)
        ctx.comm_init(uid)
    ctx.set_option("deposit", args.deposit)
    ctx.set_option("iters", args.iters)
    ctx.set_option("group_min", args.group_min)
    ctx.set_option("tile", args.tile)
    ctx.set_option("fused_keys", args.fused_keys)
    ctx.set_option("fused_sort", args.fused_sort)
    ctx.set_option("shard", 1 if args.shard == "slab" else 0)
    if args.slab_of:
        ctx.set_option("shard", 1)
        ctx.set_option("slab_of", args.slab_of[0])
        ctx.set_option("slab_index", args.slab_of[1])
    ctx.set_option("planes", args.planes)
    ctx.set_option("defer", args.defer)
    ctx.set_option("phases", 1)
    # synthetic two-flux-bundle load generated on the device (same values as loadpt, F:8937-9040)
    ranfb = 7331
    for ksp in SPECIES:
        _, ranfb = ctx.loadpt(ksp, ppc, vth(ksp), 0.0, VBEAM[ksp])
    if world > 1 and args.peer_push:
        ctx.set_option("peer_push", args.peer_push)
        ctx.set_option("peer_push_last", args.peer_push_last)
        ctx.set_option("split_push", args.split_push)
        ctx.map_peers(len(SPECIES), device=dev)          # NVLink peer memory for the slab-wise exchange (cudaIpc handles over torch.distributed)
    nloc = sum(ctx.num_local(k) for k in SPECIES)
    ntot_particles = len(SPECIES) * mx * my * mz * ppc if not args.slab_of else nloc
    n_grid = ctx.n_grid
    c = mrg.Common(4, 4, 4, 1.0, 1.0, 1.0, dt=DT, aimpl=AIMPL, wce_by_wpe=WCE, Ez00=EZ00)   # scalars only
    c.mx, c.my, c.mz, c.xmax, c.ymax, c.zmax = mx, my, mz, HX * mx, HY * my, HZ * mz
    c.zcent, c.ycent1, c.ycent2 = 0.5 * c.zmax, 0.30 * c.ymax, 0.70 * c.ymax
    params = c.step_params()
    # two field sets resident in HBM: "before emfild" and "after emfild"
    fsets = []
    for seed in (1, 2):
        f = synth_fields(torch, mx, my, mz, seed, dtype=torch.float64, device=dev)
        fsets.append([t.contiguous() for t in f])
    torch.cuda.synchronize()
    for ksp in SPECIES:
        ctx.sort(ksp, c.hdt)          # sort key = cell of the gather position x + hdt*v

    state = {"ranfb": ranfb, "step": 0, "tp": [], "tc": []}

    fptr = [[t.data_ptr() for t in fs] for fs in fsets]

    def step_resident():
        # the field arrays are resident in HBM (as a device-side emfild would leave them) and read in place
        ctx.bind_fields_device(fptr[0])
        wk = [ctx.fulmov_deferred(ksp, QSPEC[ksp], WSPEC[ksp], params) if args.defer
              else ctx.fulmov(ksp, QSPEC[ksp], WSPEC[ksp], 1, params, state["ranfb"]) for ksp in SPECIES]
        ctx.bind_fields_device(fptr[1])        # queued behind the moment sums: new fields need the moments
        for ksp in SPECIES:
            _, _, state["ranfb"] = ctx.fulmov(ksp, QSPEC[ksp], WSPEC[ksp], 0, params, state["ranfb"])
        for ksp in SPECIES:
            state["tp"].append(ctx.pass_ms(ksp, 1))
            state["tc"].append(ctx.pass_ms(ksp, 0))
        state["wk"] = wk
        state["step"] += 1
        for ksp in SPECIES:
            every = args.sort_every_ions if (ksp == 1 and args.sort_every_ions) else args.sort_every
            if every and state["step"] % every == 0:
                ctx.sort(ksp, c.hdt)

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local)
    sampler.start()      # nvidia-smi needs ~0.1 s to produce its first line: started under the warm-up steps (same load)
    for _ in range(args.warmup):
        step_resident()
    state["tp"], state["tc"] = [], []
    ctx.counters(reset=True)
    ctx.phase_ms(reset=True)
    barrier()
    ctx.event_record(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_resident()
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1)
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    barrier()
    detail = {"%d/%s" % (ksp, "pred" if ipc else "corr"): {k: round(v / args.steps, 4) for k, v in ctx.phase_detail(ksp, ipc).items() if v > 0.0}
              for ksp in SPECIES for ipc in (1, 0)}
    phases, _ = ctx.phase_ms(reset=True)
    pt = torch.tensor([phases[k] / args.steps for k in ("prep", "setup", "kernel", "rank_sum", "fold", "kick")], dtype=torch.float64, device=dev)
    if world > 1:
        pmax = pt.clone(); dist.all_reduce(pmax, op=dist.ReduceOp.MAX)
        pmin = pt.clone(); dist.all_reduce(pmin, op=dist.ReduceOp.MIN)
    else:
        pmax = pmin = pt
    phases_rec = {"unit": "device ms per step (CUDA events on the stream each phase runs on; rank_sum and fold run on the second stream "
                          "in deferred mode and overlap the next kernel, so the phases need not add up to ms_per_step)",
                  "max_over_ranks": dict(zip(("prep", "setup", "kernel", "rank_sum", "fold", "kick"), [float(v) for v in pmax])),
                  "min_over_ranks": dict(zip(("prep", "setup", "kernel", "rank_sum", "fold", "kick"), [float(v) for v in pmin]))}
    if world > 1:      # every rank's own breakdown (species/pass; the rank sum split into strips, push, barrier): who waits for whom
        per_rank = [None] * world
        dist.all_gather_object(per_rank, {"ms": ms, "detail": detail})
        phases_rec["per_rank"] = per_rank
    else:
        phases_rec["detail"] = detail
    cnt = ctx.counters(reset=True)
    tms = torch.tensor([ms, float(cnt["launches"])], dtype=torch.float64, device=dev)
    if world > 1:
        mx_t = tms.clone(); dist.all_reduce(mx_t, op=dist.ReduceOp.MAX)
        sm_t = tms.clone(); dist.all_reduce(sm_t, op=dist.ReduceOp.SUM)
        ms, launches = float(mx_t[0]), int(sm_t[1])
    else:
        launches = int(cnt["launches"])
    value = ntot_particles * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (per launch, algorithmic bytes) --------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 (B200_PROFILING.md)"
    n_sp = nloc / float(len(SPECIES))                     # particles per launch (one species on this GPU)
    bytes_pred = 48.0 * n_sp + (6 + 2 * 4) * 8.0 * n_grid
    bytes_corr = 96.0 * n_sp + 6 * 8.0 * n_grid
    tp, tc = float(np.mean(state["tp"])), float(np.mean(state["tc"]))
    gb_pred, gb_corr = bytes_pred / (tp * 1e-3) / 1e9, bytes_corr / (tc * 1e-3) / 1e9
    kn = {0: ("k_predict_run", "k_correct"), 1: ("k_predict_tile", "k_correct_tile")}[args.tile]
    dominant = kn[0] if tp >= tc else kn[1]
    ach = gb_pred if tp >= tc else gb_corr
    step_bytes = len(SPECIES) * (bytes_pred + bytes_corr)
    traffic = None                                       # measured DRAM bytes per launch of the dominant kernel (one ncu capture)
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if tj.get("grid") == [mx, my, mz] and tj.get("ppc") == ppc and tj.get("n_gpus") == world and args.fused_sort:
            traffic = tj.get(dominant)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "peak_source": peak_src,
                "predictor": {"ms_per_launch": tp, "algorithmic_bytes": bytes_pred, "gbs": gb_pred, "frac": gb_pred / peak,
                              "particle_passes_per_s": n_sp / (tp * 1e-3)},
                "corrector": {"ms_per_launch": tc, "algorithmic_bytes": bytes_corr, "gbs": gb_corr, "frac": gb_corr / peak,
                              "particle_passes_per_s": n_sp / (tc * 1e-3)},
                "whole_step": {"algorithmic_bytes": step_bytes, "gbs": step_bytes * args.steps / (ms * 1e-3) / 1e9 if world == 1 else None,
                               "frac": step_bytes * args.steps / (ms * 1e-3) / 1e9 / peak if world == 1 else None}}

    # ---- second roofline: fp64 issue.  Instruction counts per particle-pass are static properties of
    # the kernels (ncu source counters of the committed profile); the peak is measured live with a dependent-free DFMA stream.
    fp64 = None
    try:
        ic = json.load(open(os.path.join(ROOT, "profiles", "fp64_inst_per_particle.json")))
        dpeak = ctx.dfma_peak()
        ip, icr = float(ic[kn[0]]["fp64_inst_per_particle"]), float(ic[kn[1]]["fp64_inst_per_particle"])
        fp64 = {"peak_dfma_per_s": dpeak, "peak_source": "measured live (mrg_dfma_peak: 8 independent DFMA chains per thread, all SMs)",
                "unit": "thread-level fp64 instructions/s", "inst_source": ic.get("source"),
                "predictor": {"inst_per_particle": ip, "achieved": ip * n_sp / (tp * 1e-3), "frac": ip * n_sp / (tp * 1e-3) / dpeak},
                "corrector": {"inst_per_particle": icr, "achieved": icr * n_sp / (tc * 1e-3), "frac": icr * n_sp / (tc * 1e-3) / dpeak},
                "whole_step": {"inst_per_particle_step": ip + icr,
                               "frac": (ip + icr) * nloc * args.steps / (ms * 1e-3) / dpeak}}
    except Exception as ex:
        fp64 = {"error": repr(ex)}
    roofline["fp64"] = fp64

    # ---- self-checks of the timed configuration (no CPU reference needed) ----------------------------
    parity = {"checked": True}
    try:
        worst_q, perm_ok, cells_ok = 0.0, True, True
        for ksp in SPECIES:
            sc = ctx.self_check(ksp)
            ntot_sp = ntot_particles // len(SPECIES)
            parity["sum_q_raw_%d" % ksp] = sc["sums"][3]
            worst_q = max(worst_q, abs(sc["sums"][3] - QSPEC[ksp] * ntot_sp) / ntot_sp)
            perm_ok = perm_ok and sc["permutation_ok"]
            cells_ok = cells_ok and sc["cell_end_last"] == sc["n"] == ctx.num_local(ksp)
        parity.update({"sum_q_raw_rel_err": worst_q, "sum_q_ok": worst_q < 1e-9, "particles_conserved": perm_ok and cells_ok,
                       "wk_finite": all(np.isfinite(float(w[i].value if hasattr(w[i], "value") else w[i])) for w in state["wk"] for i in (0, 1))})
        if world == 1 and not args.no_check_direct:
            # the same predictor call through the simplest kernel of the library (one thread per particle, 72 red.global per
            # particle, no tiles, no pre-reduction, no sort dependence) must give the same moments as the tiled kernel: both run
            # here, back to back, on the state the timed loop left behind
            ctx.set_option("defer", 0)
            ctx.bind_fields_device(fptr[0])
            tiled = {}
            for ksp in SPECIES:
                ctx.fulmov(ksp, QSPEC[ksp], WSPEC[ksp], 1, params, state["ranfb"])
                tiled[ksp] = ctx.moments(ksp, folded=False)
            ctx.set_option("tile", 0); ctx.set_option("deposit", 0)
            worst = 0.0
            for ksp in SPECIES:
                ctx.fulmov(ksp, QSPEC[ksp], WSPEC[ksp], 1, params, state["ranfb"])
                direct = ctx.moments(ksp, folded=False)
                for cidx in range(4):
                    den = float(np.linalg.norm(direct[cidx]))
                    worst = max(worst, float(np.linalg.norm(tiled[ksp][cidx] - direct[cidx])) / (den if den > 0 else 1.0))
            ctx.set_option("tile", args.tile); ctx.set_option("deposit", args.deposit); ctx.set_option("defer", args.defer)
            del tiled, direct
            parity["moments_tiled_vs_direct_rel_l2"] = worst
            parity["moments_ok"] = worst < 1e-10
        parity["ok"] = bool(parity["sum_q_ok"] and parity["particles_conserved"] and parity["wk_finite"] and parity.get("moments_ok", True))
    except Exception as ex:
        parity = {"checked": False, "error": repr(ex)}

    # ---- end to end through the reference-facing call with HOST buffers ------------------------------
    e2e = None
    host_ok = 0.0
    hsets = []
    share = world > 1 and bool(args.share_moments)
    lazy = world > 1 and bool(args.lazy_fields)
    shm_keep = []
    if not args.no_e2e:
        def pinned(n):
            return torch.empty(n, dtype=torch.float64).pin_memory().numpy()

        def shared_pinned(name, n):
            # one POSIX shared-memory array for all ranks of the node, page-locked in every process: COMMON /srimp7/ as the
            # ranks of one node would share it (each rank delivers its own z block, option "sink_share")
            path = "/dev/shm/mrg_bench_%s_%s" % (os.environ.get("MASTER_PORT", "0"), name)
            if rank == 0:
                with open(path, "wb") as f:
                    f.truncate(n * 8)
            dist.barrier()
            t = torch.from_file(path, shared=True, size=n, dtype=torch.float64)
            # first touch: every rank faults in the z block it will deliver, so those pages sit on the NUMA node next to its GPU
            nz = mz + 4
            nxy = n // nz
            e0 = 0 if rank == 0 else mz * rank // world + 2
            e1 = nz if rank == world - 1 else mz * (rank + 1) // world + 2
            t[e0 * nxy:e1 * nxy].zero_()
            dist.barrier()
            rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), n * 8, 0)
            if int(rc) != 0:
                raise RuntimeError("cudaHostRegister failed: %r" % (rc,))
            shm_keep.append((t, path))
            return t.numpy()
        try:      # 32 page-locked grid arrays per rank (4.5 GB at 256^3): a host that cannot pin them loses this leg, not the line
            for fs in fsets:
                hs = []
                for t in fs:
                    a = pinned(n_grid)
                    a[:] = t.cpu().numpy()
                    hs.append(a)
                hsets.append(hs)
            for name in ("qix", "qiy", "qiz", "qex", "qey", "qez", "qi", "qe"):
                setattr(c, name, shared_pinned(name, n_grid) if share else pinned(n_grid))
        except (RuntimeError, MemoryError) as ex:
            host_ok = 1.0
            sys.stderr.write("bench: e2e leg skipped on rank %d: %r\n" % (rank, ex))
        flag = torch.tensor([host_ok], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)     # every rank takes the same branch
        host_ok = float(flag[0])
    if not args.no_e2e and host_ok == 0.0:
        c.ranfb = state["ranfb"]
        fm = mrg.Fulmov(c, ipar=rank + 1, size=world, device=local, sort_interval=args.sort_every, ctx=ctx,
                        hints=bool(args.hints), defer=bool(args.defer), lazy=lazy, share_moments=share, nspecies=len(SPECIES))
        dummy = [np.zeros(1)] * 6
        npr = ntot_particles // len(SPECIES)
        FN = mrg.host.FIELD_NAMES
        for name, arr in zip(FN, hsets[0]):
            setattr(c, name, arr)
        estate = {"n": 0}

        def step_e2e():
            # the trans protocol (F:749-807) with host arrays: prefld rewrites bx,by,bz; the two ipc=1 calls fill
            # COMMON /srimp7/; emfild rewrites ex..bz; the two ipc=0 calls; renewal ex0 <- ex.  "Rewrites" = the
            # COMMON member now is another pinned array (the host solver is not part of the timed path).
            new = hsets[(estate["n"] + 1) % 2]
            for i in (3, 4, 5):
                setattr(c, FN[i], new[i])
            if args.device_prefld:
                fm.prefld_done()        # whole device arrays (N = 1): prefld is repeated on the device, no upload; else = MASK_B
            else:
                fm.fields_changed(fm.MASK_B)
            for ksp in SPECIES:
                fm(*dummy, QSPEC[ksp], WSPEC[ksp], npr, 1, ksp)
            fm.finish_moments()
            if share:
                dist.barrier()          # every rank's block of the shared COMMON /srimp7/ arrays has landed (MPI_Barrier in a Fortran host)
            for i in (0, 1, 2):
                setattr(c, FN[i], new[i])
            if args.device_prefld:
                c.it = estate["n"] + 2     # emfild_done looks at mod(it,5): every fifth step smooths B on the device too (F:4298)
                fm.emfild_done()           # whole device arrays (N = 1): only ex,ey,ez are uploaded; else = MASK_NEW
            else:
                fm.fields_changed(fm.MASK_NEW)
            for ksp in SPECIES:
                fm(*dummy, QSPEC[ksp], WSPEC[ksp], npr, 0, ksp)
            for i in range(6):
                setattr(c, FN[i + 6], getattr(c, FN[i]))
            fm.fields_renewed()
            estate["n"] += 1

        ne = max(2, min(args.steps, 5))
        step_e2e()
        ctx.counters(reset=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(ne):
            step_e2e()
        ctx.synchronize()
        te = time.perf_counter() - t0
        cnt_e = ctx.counters(reset=True)
        tt = torch.tensor([te], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        te = float(tt[0])
        # what landed in the host's COMMON /srimp7/ arrays during the timed steps must be what the device holds: with shared
        # arrays every rank delivered one z block, and rank 0 checks all of them against its own complete folded arrays
        shared_ok = None
        if share:
            dist.barrier()
            ctx.set_option("sink_share", 0)
            ok = 1.0
            if rank == 0:
                for ksp in (1, 2):
                    full = ctx.moments(ksp, folded=True)
                    for a, b in zip(full, fm._moment_arrays(ksp)):
                        ok = ok if np.array_equal(a, b) else 0.0
            ctx.set_option("sink_share", 1)
            flag = torch.tensor([ok], dtype=torch.float64, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            shared_ok = bool(float(flag[0]) == 1.0)
        e2e = {"value": ntot_particles * ne / te, "unit": UNIT, "h2d_bytes_per_step": cnt_e["h2d_bytes"] // ne,
               "d2h_bytes_per_step": cnt_e["d2h_bytes"] // ne, "steps": ne, "ms_per_step": 1e3 * te / ne,
               "pcie_gb_per_s_per_rank": (cnt_e["h2d_bytes"] + cnt_e["d2h_bytes"]) / te / 1e9,
               "lazy_fields": lazy, "shared_moment_arrays": share, "shared_arrays_equal_device_moments": shared_ok,
               "note": "host fields in pinned memory -> mrg_set_fields (H2D), moments -> COMMON /srimp7/ arrays (D2H) every "
                       "step through the Fulmov mirror of the reference call; particles stay resident in HBM by design; "
                       + ("each rank fetches only the z planes its field preparation reads (mrg_set_fields_lazy) and delivers only its "
                          "own z block of the moments into host arrays shared by the ranks of the node (POSIX shm, option sink_share); "
                          if (lazy or share) else "")
                       + ("the host marks its field updates (prefld: bx..bz, emfild: ex..bz, renewal on the device)"
                          if args.hints else "no field hints: all of COMMON /fields/ is uploaded in both phases")
                       + ("; prefld and emfild's B update are repeated on the device (mrg_update_b, bit-identical to the host's), so bx,by,bz are never uploaded"
                          if (args.hints and args.device_prefld and not lazy) else "")}
        barrier()
        for t, path in shm_keep:
            torch.cuda.cudart().cudaHostUnregister(t.data_ptr())
            if rank == 0:
                try:
                    os.unlink(path)
                except OSError:
                    pass

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu = cpu_baseline_record(cpu_reference_rate(args.cpu_steps, 1))
        except Exception as ex:                             # the oracle is a reported baseline only
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (ex,)}

    prep_stats = ctx.prep_stats()
    prep_stats["peer_pushes"] = int(ctx.peer_pushes())
    prep_stats["split_launches"] = int(ctx.split_pushes())
    refpart = None
    if world > 1 and args.shard == "slab" and not args.slab_of and not args.no_reference_partition:
        ctx.close()
        del fsets
        torch.cuda.empty_cache()
        try:
            refpart = measure_reference_partition(mrg, dist, torch, args, rank, world, local, dev, (mx, my, mz), params)
        except Exception as ex:
            refpart = {"error": repr(ex)}
    if rank == 0:
        cfg = workload_config(args.gpus, args)
        cfg["parallelism"] = "particle-sharded x%d" % world
        cfg["host_binding"] = ("each rank bound to the %d CPUs next to its GPU (NVML affinity)" % numa_cpus) if numa_cpus else "none"
        if args.slab_of:
            cfg["emulated_rank"] = "slab %d of %d on one GPU, no NCCL sum (development aid, not a bench line)" % (args.slab_of[1], args.slab_of[0])
        cfg["prep"] = prep_stats
        cfg["rank_sum"] = ("none (1 GPU)" if world == 1 else
                           "%d of %d moment sums went through the slab-wise exchange (%d of them finished by the fused add+push kernel over "
                           "NVLink peer memory -- %d of those after a split launch that pushed part of the block under the rest of the "
                           "particle kernel -- the others by ncclAllGather), the rest through ncclAllReduce"
                           % (cfg["prep"]["compact_sums"], len(SPECIES) * (args.steps + args.warmup + (e2e["steps"] + 1 if e2e else 0)),
                              cfg["prep"]["peer_pushes"], cfg["prep"]["split_launches"]))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.config == 5 else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg, "roofline": roofline, "cpu_baseline": cpu,
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "parity": parity, "multi_rank_parity": mrp, "reference_partition": refpart, "phases": phases_rec,
                "wall_ms_per_step": 1e3 * wall / args.steps,
                "pct_hbm_roofline_whole_step": None if world > 1 else 100.0 * roofline["whole_step"]["frac"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


def run_verify(args):
    """GPU vs oracle at a BASELINE size (default configs[1]: 128^3 x 64 ppc): the device-loaded particles are downloaded, the
    C oracle (bit-identical to the reference's own fulmov, tests/test_ref_pin.py) runs the same predictor and corrector pass on
    all host cores, and the results are compared: raw + folded moments (rel-L2), every particle (max rel error), ranfp state."""
    import torch
    import mrg_b200 as mrg
    from oracle import pyoracle as O
    mx, my, mz = tuple(args.grid) if args.grid else GRIDS[1]
    ppc = args.ppc
    nthreads = min(len(os.sched_getaffinity(0)), 64)
    O.set_num_threads(nthreads)
    ctx = mrg.MrgContext(mx, my, mz, HX * mx, HY * my, HZ * mz, nspecies=len(SPECIES), device=0)
    for name in ("deposit", "iters", "group_min", "tile", "fused_keys", "fused_sort"):
        ctx.set_option(name, getattr(args, name))
    p = O.make_parm(mx, my, mz, HX * mx, HY * my, HZ * mz, DT, AIMPL, WCE, EZ00)
    par = mrg.StepParams(p.dt, p.adt, p.hdt, p.aimpl, p.bxc, p.byc, p.bzc, 1, 1, 1, 1, p.Ez00, p.zcent, p.ycent1, p.ycent2)
    ranfb = 7331
    for ksp in SPECIES:
        _, ranfb = ctx.loadpt(ksp, ppc, vth(ksp), 0.0, VBEAM[ksp])
    fa = [np.ascontiguousarray(a, dtype=np.float64) for a in synth_fields(np, mx, my, mz, 1, dtype=np.float64)]
    fb = [np.ascontiguousarray(a, dtype=np.float64) for a in synth_fields(np, mx, my, mz, 2, dtype=np.float64)]
    rep = {"verify": True, "grid": [mx, my, mz], "ppc": ppc, "oracle_threads": nthreads, "species": {}}
    n = ctx.num_local(1)
    t00 = time.perf_counter()
    for ksp in SPECIES:
        ctx.sort(ksp, p.hdt)
    st_o = np.array([ranfb], dtype=np.int32)
    st_g = ranfb
    worst_m = worst_p = 0.0
    for ksp in SPECIES:           # one species at a time: 6 x n doubles on the host (6.4 GB at configs[1])
        host = ctx.download(ksp, n)
        ctx.set_fields(fa)
        a6 = O.field_prep(p, fa)
        wx, wh, _ = ctx.fulmov(ksp, QSPEC[ksp], WSPEC[ksp], 1, par, st_g)
        r = O.fulmov(p, a6, *host, QSPEC[ksp], WSPEC[ksp], 1, nranks=nthreads, want_raw=True)
        raw, mom = ctx.moments(ksp, folded=False), ctx.moments(ksp, folded=True)
        e_raw = max(float(np.linalg.norm(raw[c] - r["raw"][c]) / np.linalg.norm(r["raw"][c])) for c in range(4))
        e_mom = max(float(np.linalg.norm(mom[c] - r["mom"][c]) / np.linalg.norm(r["mom"][c])) for c in range(4))
        e_wk = max(abs(wx - r["wkix"]) / abs(r["wkix"]), abs(wh - r["wkih"]) / abs(r["wkih"]))
        del raw, mom, r
        ctx.set_fields(fb)
        a6 = O.field_prep(p, fb)
        _, _, st_g = ctx.fulmov(ksp, QSPEC[ksp], WSPEC[ksp], 0, par, st_g)
        O.fulmov(p, a6, *host, QSPEC[ksp], WSPEC[ksp], 0, nranks=1, ranfb=st_o)      # one rank: the serial ranfp order of F:1342-1364
        got = ctx.download(ksp, n)
        e_p = 0.0
        for c in range(6):
            fl = HX if c < 3 else vth(ksp)
            e_p = max(e_p, float(np.max(np.abs(got[c] - host[c]) / np.maximum(np.abs(host[c]), fl))))
        del got, host
        rep["species"][str(ksp)] = {"particles": int(n), "raw_moments_rel_l2": e_raw, "folded_moments_rel_l2": e_mom,
                                    "wkix_wkih_rel": e_wk, "particles_max_rel_err": e_p}
        worst_m, worst_p = max(worst_m, e_raw, e_mom), max(worst_p, e_p)
    rep["ranfb_equal"] = bool(int(st_o[0]) == int(st_g))
    rep["moments_rel_l2_max"], rep["particles_rel_err_max"] = worst_m, worst_p
    rep["ok"] = bool(worst_m < 1e-10 and worst_p < 1e-12 and rep["ranfb_equal"])
    rep["seconds"] = time.perf_counter() - t00
    rep["tolerances"] = {"moments_rel_l2": 1e-10, "particles_rel": 1e-12, "ranfb": "equal"}
    print(json.dumps(rep))
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, nargs=3, default=None, help="override the grid (mx my mz)")
    ap.add_argument("--ppc", type=int, default=None, help="particles per cell per species (default 64; 100 / 16 for --config 4 / 5)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 4, 5],
                    help="BASELINE.md config number: 2 = 128^3 x 64 ppc per GPU (default), 4 = 256x128x128 per GPU x 100 ppc, "
                         "5 = 512x256x256, three species, dt*wce > 10")
    ap.add_argument("--sort-every", type=int, default=1)
    ap.add_argument("--sort-every-ions", type=int, default=0, help="ion sort cadence (0 = same as --sort-every)")
    ap.add_argument("--deposit", type=int, default=2)
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--group-min", type=int, default=2)
    ap.add_argument("--tile", type=int, default=1)
    ap.add_argument("--fused-keys", type=int, default=1)
    ap.add_argument("--fused-sort", type=int, default=1)
    ap.add_argument("--shard", default="slab", choices=["slab", "roundrobin"],
                    help="particle ownership for --gpus > 1: z slabs of the initial positions, or the reference's round-robin")
    ap.add_argument("--planes", type=int, default=-1, help="restricted field preparation: -1 = when N > 1, 0 = off, 1 = on")
    ap.add_argument("--defer", type=int, default=1, help="1 = moment sum + fold on the second stream (overlaps the next kernel)")
    ap.add_argument("--slab-of", type=int, nargs=2, default=None, metavar=("N", "I"),
                    help="development aid: hold z slab I of the N-GPU job's load and grid on ONE GPU (no NCCL); implies --no-e2e")
    ap.add_argument("--hints", type=int, default=1, help="e2e leg: 1 = the host marks which members of COMMON /fields/ it changed")
    ap.add_argument("--peer-push", type=int, default=64, help="N > 1: CTAs of the fused add+push kernel that finishes the slab-wise exchange over NVLink peer memory (0 = ncclAllGather)")
    ap.add_argument("--peer-push-last", type=int, default=296, help="N > 1: CTAs of that kernel for the last species of the step, whose exchange nothing overlaps (0 = --peer-push)")
    ap.add_argument("--split-push", type=int, default=1, help="N > 1: 1 = the last species' predictor runs as two launches and the planes final after the first are pushed to the peers under the second (0 = off, 2 = every species)")
    ap.add_argument("--numa-bind", type=int, default=1, help="N > 1: bind each rank to the CPUs next to its GPU before it allocates pinned host arrays")
    ap.add_argument("--device-prefld", type=int, default=1, help="e2e leg: 1 = prefld and emfild's B update are repeated on the device instead of uploading bx,by,bz (whole device arrays only: N = 1)")
    ap.add_argument("--lazy-fields", type=int, default=1, help="e2e leg at N > 1: upload only the z planes each rank's preparation reads")
    ap.add_argument("--share-moments", type=int, default=1, help="e2e leg at N > 1: ranks share the host moment arrays, each delivers its z block")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-reference-partition", action="store_true", help="skip the secondary round-robin leg of --gpus N > 1")
    ap.add_argument("--no-rank-parity", action="store_true", help="skip the small all-rank parity pass of --gpus N > 1")
    ap.add_argument("--no-check-direct", action="store_true", help="skip the tiled-vs-direct moment cross-check of the parity object")
    ap.add_argument("--verify", action="store_true",
                    help="full-size parity: one predictor + one corrector pass per species at --grid/--ppc against the CPU oracle "
                         "(pinned bit for bit to the reference); prints a JSON report instead of a bench line")
    args = ap.parse_args()
    apply_config(args)
    if args.gpus not in GRIDS and not args.grid:
        raise SystemExit("--gpus must be 1, 2, 4 or 8 (or give --grid)")
    if args.slab_of:
        args.no_e2e = True
    if args.verify:
        run_verify(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
