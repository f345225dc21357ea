"""Host-side sample of the C3 HVG count matrix (first `cells` cells, the generator's twin) for layout studies: saves
colptr / rowval / counts of the 2,000 selected genes to /tmp/hvg_sample.npz."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from oracle import severo_oracle as orc

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cfg = bench.CONFIGS[sys.argv[2] if len(sys.argv) > 2 else "C3"]
t0 = time.time()
tab = orc.synth_tables(cfg["m"], cfg["g"], cfg["nnz"], programs=cfg["programs"], fold=bench.FOLD, seed=bench.SEED, rows=(0, cells))
libsize, gene_nnz, mean, var, hist = orc.synth_stats(tab)
sd = np.sqrt(var); ex = sd.copy(); nc = sd > 0; ex[nc] = bench.hvg_trend(mean[nc], sd[nc])
metric = orc.stdvar_clipped_hist(cells, hist, gene_nnz, mean, ex)
hvf = np.argsort(-metric, kind="stable")[:cfg["n"]]
colptr, rowval, counts = orc.synth_columns(tab, hvf)
print("cells", cells, "z", colptr[-1], "nnz/cell", colptr[-1] / cells, "time", time.time() - t0)
h = np.bincount(np.minimum(counts, 40))
tot = h.sum()
for L in (4, 8, 16, 32):
    print("count >", L, ":", h[L + 1:].sum() / tot)
print("level shares:", np.round(h[1:9] / tot, 4))
np.savez("/tmp/hvg_sample.npz", colptr=colptr, rowval=rowval, counts=counts, cells=cells, libsize=libsize)
