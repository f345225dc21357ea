"""One rank's share of a sharded C3 solve on ONE GPU (no exchange): the kernels a rank of an N-GPU run launches, at its
shard's size, for an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv ...`).
usage: python tools/shard_profile.py [shards=8] [config=C3]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import severo_jl_b200 as sv

shards = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = dict(bench.CONFIGS[sys.argv[2] if len(sys.argv) > 2 else "C3"])
cfg["m"] = cfg["m"] // shards
sv.init(0)
B, mu, info = bench.build_workload(sv, cfg, 0, 1)
B.free()
op, _ = bench.make_counts_operator(sv, info["counts_hvg"], info["libsize"], exact=False)
print(bench.counts_info(sv, op), file=sys.stderr)
init = np.random.default_rng(bench.SEED).standard_normal(cfg["n"])
import time
for i in range(3):
    sv.lib().svb_synchronize()
    t0 = time.perf_counter()
    r, it, mp, inf = bench.solve_device(sv, op, cfg["nu"], init)
    sv.lib().svb_synchronize()
    print(f"solve {i}: {1e3 * (time.perf_counter() - t0):.3f} ms, restarts {it}, products {mp}, info {inf}", file=sys.stderr)
    sv.lib().svb_result_free(r)
