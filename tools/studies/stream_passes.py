"""GPU study: download the adjoint stream of a count-level operator built WITHOUT replicas (canonical codes) and evaluate on
the host (a) the passes per set of the round-robin order, (b) what the replica matching should reach — to be compared with
the figure the device-side assignment reports (svb_operator_counts_layout) when replicas are on.
usage: SVB_FACT_REPLICAS=0 SVB_FACT_LOG2R=10 python tools/studies/stream_passes.py"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import severo_jl_b200 as sv
from conftest import planted_counts
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

sv.init()
X = planted_counts(9000, 900, 8, seed=21, mean_nnz=170)
hvf = sv.find_variable_features(X, 400)
C = sv.scale_features_counts(X, scale_factor=1e4, scale_max=10.0, features=hvf, levels=16)
info = C.info()
print(info)
nch = info["adj_chunks"]
code = np.zeros(nch * 16, dtype=np.uint16)
meta = np.zeros(nch, dtype=np.uint8)
sv._lib.check(sv.lib().svb_operator_counts_stream(C._op, 1, sv._lib.ptr(code), sv._lib.ptr(meta)))
code = code.reshape(nch, 16).astype(np.int64)
R, L = info["tile_cells"], info["levels"]
RL = R * L
exc = (meta & 2) != 0
print("chunks", nch, "exception chunks", int(exc.sum()), "pad slots", int((code[~exc] == RL).sum()))
if info["adj_replicas"] > 1:
    print("replicated stream: run with SVB_FACT_REPLICAS=0 for the canonical codes")
    sys.exit(0)
import bank_sim as bs
ch = code.copy()
ch[ch == RL] = bs.PAD
ch[exc] = -2          # absent
nblocks = (nch + 15) // 16
tot0 = tot1 = sets = 0
for b in range(nblocks):
    blk = ch[b * 16:(b + 1) * 16]
    for e in range(8):
        addrs = [int(x) for x in blk[:, e] if x != -2]
        if not addrs:
            continue
        sets += 1
        tot0 += bs.passes_of_set(addrs, lambda a: (a & 15) if a != bs.PAD else 0)
        tot1 += bs.matching_partial(addrs, lambda a: (a & 15) if a != bs.PAD else 0, (0, 5, 10, 15), lambda a: a != bs.PAD and (a // R) < 2)
print("host evaluation of the device stream: round-robin order %.4f passes/set, with replica matching %.4f" % (tot0 / sets, tot1 / sets))
