#!/usr/bin/env python
"""Recipe for oracle/_ref/: the reference's own routines of the /fulmov/ path -- and, around it, the rest of its time
cycle: init/loadpt, the t = 0 solve emfld0/poissn, prefld and the implicit field solve emfild/emcoef/cfpsol/bcgstb --
compiled from the source where it lies (/root/reference/@mrg37-080A.f03 + param_080A.h) -- TEST INFRASTRUCTURE, see
oracle/f03c.py.

The image has no Fortran compiler, so the compile step is  Fortran --(oracle/f03c.py)--> C --(gcc)--> .so :
    oracle/_ref/mrgref_gen.c     generated, derived from the GPL-3.0 reference: git-ignored, never committed
    oracle/_ref/libmrg_ref.so    gcc -O2 -ffp-contract=off -fwrapv (no -march: baseline x86-64 has no FMA, like the
                                 documented `mpif90 -mcmodel=medium -O2`, F:100)
Nothing is copied from the reference into the repository; on a box without /root/reference (the GPU box) the
prebuilt oracle/_ref/ that travelled with the snapshot is used as it is.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.environ.get("MRG_REFERENCE_DIR", "/root/reference")
SRC = os.path.join(REF_DIR, "@mrg37-080A.f03")
OUT = os.path.join(HERE, "_ref")
GEN = os.path.join(OUT, "mrgref_gen.c")
LIB = os.path.join(OUT, "libmrg_ref.so")

# entry units; everything they call is pulled in by the translator
UNITS = ["fulmov", "rantbl", "ranf", "ranfp", "loadpt", "init", "emfld0", "emfild", "prefld", "restrt"]
# out-of-scope callees whose calls are dropped: PostScript plots, labels, wall clocks
STUBS = ["fplot3", "cplot3", "lblbot", "lbltop", "clocks", "clocki", "lplots", "lplot1", "hplot1", "lplmax", "lplmax1"]


def available():
    return os.path.exists(LIB)


def can_build():
    return os.path.exists(SRC)


def build(force=False, units=None, verbose=False):
    """returns the library path, or None when neither the reference source nor a prebuilt library exists"""
    if not can_build():
        return LIB if available() else None
    sys.path.insert(0, HERE)
    import f03c
    deps = [SRC, os.path.join(REF_DIR, "param_080A.h"), os.path.join(HERE, "f03c.py"), os.path.join(HERE, "ref_runtime.c"),
            os.path.abspath(__file__)]
    if not force and available() and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    want = list(units or UNITS)
    ok = []
    for name in want:       # a unit outside the translator's subset is left out, not fatal: say which
        try:
            f03c.translate(SRC, [REF_DIR], ok + [name], stubs=STUBS)
            ok.append(name)
        except Exception as ex:
            sys.stderr.write("build_ref: unit %s left out: %s\n" % (name, ex))
    src, tr = f03c.translate(SRC, [REF_DIR], ok, stubs=STUBS)
    with open(GEN, "w") as f:
        f.write(src)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fwrapv", "-fno-strict-aliasing", "-w", "-std=gnu99", "-fPIC", "-shared", "-pthread",
           "-I" + OUT, "-o", LIB, os.path.join(HERE, "ref_runtime.c"), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed on the translated reference:\n" + r.stderr[:4000])
    if verbose:
        print("translated units:", ", ".join(sorted(tr.externals | set(ok))))
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
