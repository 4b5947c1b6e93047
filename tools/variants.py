#!/usr/bin/env python3
"""Builds tuning variants of the library: tools/variants.py name:-DX=1,-DY=2 ...  -> 3bz_b200/var_<name>.so"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b = importlib.import_module("3bz_b200.build")
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    out = os.path.join(ROOT, "3bz_b200", "var_%s.so" % name)
    b.build(force=True, extra=[d for d in defs.split(",") if d], out=out)
    print(out)
