// Host check of the build-time replica assignment (severo.jl_b200/csrc/assign.cuh — the same source the device compiles):
// random sets of 16 gathers in the forward and adjoint geometries of the count-level operator. For every set and every variant
// (greedy pass at efforts 0..2, greedy + full matching):
//   * every output code is the input entry at one of ITS replicas (replica index < nrep; single-copy entries untouched);
//   * the pads of a set share one address; absent slots are untouched;
//   * the returned number of passes equals an independent recount of the largest bank load;
//   * the full matching is OPTIMAL: it returns 1 pass whenever a conflict-free assignment exists (independent augmenting-path
//     matching on plain arrays). (When no conflict-free assignment exists the matching places what it cannot match on the least
//     loaded candidate bank; that is not always the minimum — 0.06 % of random infeasible sets end one pass above the greedy
//     result — and is reported, not asserted.)
// usage: assign_check <sets per geometry>   (exit code 0 = ok)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../severo.jl_b200/csrc/assign.cuh"

static int cand_bank(int base, int r, int step) { return (base + r * step) & 15; }

// independent maximum bipartite matching (entries -> banks), plain recursive augmenting paths
static int owner_[16];
static bool try_entry(int e, const int *base, const int *nc, int step, const bool *fixedbank, bool *seen) {
    for (int r = 0; r < nc[e]; ++r) {
        const int b = cand_bank(base[e], r, step);
        if (seen[b] || fixedbank[b]) continue;
        seen[b] = true;
        if (owner_[b] < 0 || try_entry(owner_[b], base, nc, step, fixedbank, seen)) {
            owner_[b] = e;
            return true;
        }
    }
    return false;
}

int main(int argc, char **argv) {
    const long nsets = argc > 1 ? atol(argv[1]) : 100000;
    srand(12345);
    long bad = 0, checked = 0, worse = 0;
    for (int adj = 0; adj < 2; ++adj)
        for (int nrep = 1; nrep <= 4; ++nrep)
            for (long it = 0; it < nsets; ++it) {
                const int log2R = 10, L = 16, R = 1 << log2R, nlr = 2, ngenes = 2000;
                const int step = adj ? 1029 : 2005, levstride = 4128, baseB = 8256, pad = 22592;
                const svb::AssignGeom G{adj, nrep, step, adj ? R * L : ngenes, log2R, nlr, baseB, pad, levstride};
                int v[16];
                unsigned present = 0;
                const int mode = rand() % 4;
                for (int c = 0; c < 16; ++c) {
                    if (rand() % 16 != 0) present |= 1u << c;
                    if (adj) {
                        const int l = mode == 0 ? rand() % (L - 1) : (rand() % 4 == 0 ? 2 + rand() % (L - 3) : rand() % 2);
                        const int i = mode == 2 ? (rand() % 64) * 16 + rand() % 3 : (mode == 3 ? (rand() % 8 ? c : rand() % 16) + 16 * (rand() % 64) : rand() % R);
                        v[c] = l * R + i;
                        if (rand() % 10 == 0) v[c] = R * L;
                    } else {
                        int g = mode == 2 ? (rand() % 120) * 16 + rand() % 2 : (mode == 3 ? (rand() % 8 ? c : rand() % 16) + 16 * (rand() % 120) : rand() % ngenes);
                        if (rand() % 10 == 0) g = ngenes;
                        v[c] = g << 3;
                    }
                }
                // canonical description of the set
                int base[16], nc[16];
                bool ispad[16];
                for (int c = 0; c < 16; ++c) {
                    ispad[c] = adj ? v[c] == R * L : (v[c] >> 3) == ngenes;
                    if (!adj) { base[c] = v[c] >> 3; nc[c] = nrep; }
                    else if (ispad[c]) { base[c] = pad; nc[c] = 1; }
                    else {
                        const int l = v[c] >> log2R, il = v[c] & (R - 1);
                        if (l < nlr) { base[c] = l * levstride + il; nc[c] = nrep; } else { base[c] = baseB + ((l - nlr) << log2R) + il; nc[c] = 1; }
                    }
                }
                // does a conflict-free assignment exist? (first pad stands for all pads)
                bool fixedbank[16] = {false};
                bool feasible = true;
                int live[16], nlive = 0, firstpad = -1;
                for (int c = 0; c < 16; ++c) {
                    if (!((present >> c) & 1u)) continue;
                    if (ispad[c]) { if (firstpad >= 0) continue; firstpad = c; }
                    live[nlive++] = c;
                }
                for (int k = 0; k < nlive; ++k)
                    if (nc[live[k]] == 1) {
                        const int b = base[live[k]] & 15;
                        if (fixedbank[b]) feasible = false;
                        fixedbank[b] = true;
                    }
                for (int b = 0; b < 16; ++b) owner_[b] = -1;
                for (int k = 0; k < nlive && feasible; ++k)
                    if (nc[live[k]] > 1) {
                        bool seen[16] = {false};
                        if (!try_entry(live[k], base, nc, step, fixedbank, seen)) feasible = false;
                    }
                int greedy_passes = 0;
                for (int variant = 0; variant < 4; ++variant) {
                    int g[16];
                    memcpy(g, v, sizeof v);
                    const int p = variant < 3 ? svb::assign_set_fast(G, g, present, true, variant) : svb::assign_set(G, g, present);
                    int load[16] = {0}, mx = 0, padphys = -1;
                    bool ok = true;
                    for (int c = 0; c < 16; ++c) {
                        if (!((present >> c) & 1u)) { if (g[c] != v[c]) ok = false; continue; }
                        const int ph = adj ? g[c] : g[c] >> 3, d = ph - base[c];
                        if (d % step || d / step < 0 || d / step >= nc[c]) ok = false;
                        if (ispad[c]) { if (padphys >= 0) { if (ph != padphys) ok = false; continue; } padphys = ph; }
                        if (++load[ph & 15] > mx) mx = load[ph & 15];
                    }
                    if (variant == 2) greedy_passes = p;
                    if (!ok || mx != p) { ++bad; if (getenv("ASSIGN_VERBOSE")) printf("invalid: adj %d nrep %d variant %d p %d recount %d ok %d\n", adj, nrep, variant, p, mx, (int)ok); }
                    if (variant == 3 && feasible && present && p != 1) { ++bad; if (getenv("ASSIGN_VERBOSE")) printf("not optimal: adj %d nrep %d p %d\n", adj, nrep, p); }
                    if (variant == 3 && p > greedy_passes) ++worse;
                    ++checked;
                }
            }
    printf("assign_check: %ld variant runs, %ld failures (%ld infeasible sets one pass above the greedy result)\n", checked, bad, worse);
    return bad != 0;
}
