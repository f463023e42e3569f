#!/usr/bin/env python3
"""Build A/B variants of the CUDA library next to the product build: zkvm_b200/libzkmsm_<tag>.so with extra -D flags.
usage: tools/ab_build.py tag1:-DFOO=1,-DBAR=2 tag2:-DBAZ=3 ...   (run a tool against one with ZKMSM_DEV=1 ZKMSM_LIB=<path>)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
for spec in sys.argv[1:]:
    tag, _, flags = spec.partition(":")
    so = os.path.join(g.ROOT, "zkvm_b200", f"libzkmsm_{tag}.so")
    g.build_lib(so, tuple(f for f in flags.split(",") if f))
    print(so)
