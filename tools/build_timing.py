"""Per-phase timing of svb_operator_create_counts at a bench configuration, steady state (second build: allocator warm).
usage: SVB_FACT_TIMING=1 python tools/build_timing.py [C3]"""
import os, sys, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import severo_jl_b200 as sv
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
sv.init(0)
B, mu, info = bench.build_workload(sv, cfg, 0, 1)
B.free()
chv, libsize = info["counts_hvg"], info["libsize"]
for i in range(3):
    print(f"---- build {i}", file=sys.stderr, flush=True)
    h, _ = bench.make_counts_operator(sv, chv, libsize, exact=False)
    sv.lib().svb_operator_free(h)
