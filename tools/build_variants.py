#!/usr/bin/env python
"""Build experimental variants of libmrg_fulmov.so (compile-time knobs) into gpurun-visible
build/ files: usage build_variants.py name:DEF1=V,DEF2=V ...  -> variants/libmrg_<name>.so"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mrg_b200 as mrg  # noqa: E402

os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    out = os.path.join(ROOT, "variants", "libmrg_%s.so" % name)
    mrg.build.build_cuda(force=True, verbose=True, out=out, defines=[d for d in defs.split(",") if d])
    print(out)
