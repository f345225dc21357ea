"""Study (CPU, host twin of the generator): how many slots of the count-level FORWARD stream are pads, for the chunk layouts
considered in DESIGN.md section 8.1(a). The forward stream groups the HVG nonzeros of a cell by count level 1..L and pads every
(cell, level) group to whole chunks; with 16 codes per chunk ncu measured DRAM traffic at 1.21x the algorithmic bytes.
Input: the first CELLS cells of the C3 configuration (the real matrix: same generator, same seed), HVGs selected from the
sample's own moments with the bench's trend stand-in (the full-size selection needs all cells; the level mix per cell is what
matters here). Output: slots per coded nonzero for
  (a) 8 codes per chunk, one level per chunk            (round 1)
  (b) 16 codes per chunk, one level per chunk           (round 2, the committed layout)
  (c) 16 codes per chunk, one level per 8-code HALF     (proposed: two level fields in the meta word, cells end on a chunk)
usage: python tests/studies/forward_padding_study.py [cells=40000]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from oracle import severo_oracle as orc

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
cfg = bench.CONFIGS["C3"]
L = 16
tab = orc.synth_tables(cfg["m"], cfg["g"], cfg["nnz"], programs=cfg["programs"], fold=bench.FOLD, seed=bench.SEED, rows=(0, cells))
libsize, gene_nnz, mean, var, hist = orc.synth_stats(tab)
sd = np.sqrt(var)
expected = sd.copy()
nc = sd > 0
expected[nc] = bench.hvg_trend(mean[nc], sd[nc])
metric = orc.stdvar_clipped_hist(cells, hist, gene_nnz, mean, expected)
hvf = np.argsort(-metric, kind="stable")[:cfg["n"]]
colptr, rowval, counts = orc.synth_columns(tab, hvf)
nnz = int(colptr[-1])
rows = rowval[:nnz].astype(np.int64)
cnt = counts[:nnz].astype(np.int64)
coded = (cnt >= 1) & (cnt <= L)
# group sizes per (cell, level)
key = rows[coded] * L + (cnt[coded] - 1)
sizes = np.bincount(key, minlength=cells * L).reshape(cells, L)
n_coded = int(coded.sum())


def slots(unit, cell_unit):
    per_group = -(-sizes // unit) * unit           # every group padded to `unit`
    per_cell = per_group.sum(axis=1)
    per_cell = -(-per_cell // cell_unit) * cell_unit
    return int(per_cell.sum())


a, b, c = slots(8, 8), slots(16, 16), slots(8, 16)
print(f"{cells} cells, {cfg['n']} HVGs: {nnz} nonzeros ({nnz / cells:.0f} per cell), {n_coded} coded at L = {L} "
      f"({100 * (1 - n_coded / nnz):.2f} % exceptions), {np.count_nonzero(sizes)} non-empty (cell, level) groups "
      f"({np.count_nonzero(sizes) / cells:.1f} per cell)")
for name, s_, bytes_per_chunk, per in (("(a) 8 per chunk", a, 17, 8), ("(b) 16 per chunk", b, 33, 16), ("(c) 16 per chunk, level per half", c, 34, 16)):
    print(f"{name:36s} slots / coded nonzero = {s_ / n_coded:.4f}   pads = {100 * (s_ - n_coded) / s_:.1f} % of the slots   "
          f"stream bytes / coded nonzero = {s_ / per * bytes_per_chunk / n_coded:.3f}")
