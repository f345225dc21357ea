// assign.cuh — replica assignment of one SET of gathers (see factored.cu "replica assignment"): host + device, so that the
// matching itself can be checked on the CPU against the layout simulator (tools/studies/).
#pragma once
#include <cstdint>

#ifndef __CUDACC__
#define SVB_HD
#else
#define SVB_HD __host__ __device__
#endif

namespace svb {

struct AssignGeom {
    int adjoint;   // 0: forward stream (codes are byte offsets of xs entries); 1: adjoint stream (canonical (level-1)*R + i)
    int nrep;      // replicas of the replicated region
    int step;      // entries between two replicas (= 5 mod 16)
    int padcanon;  // canonical pad: gene n (forward) / R*L (adjoint)
    int log2R, nlr, baseB, pad;  // adjoint: cells per tile, replicated levels, first single-copy entry, physical pad entry
    int levstride;               // adjoint: entries between two replicated levels (a multiple of 16: replica 0 keeps bank = cell mod 16)
};

// in: v[c] = the 16 codes of the set (canonical), present = which of them exist (not beyond the stream, not an exception chunk).
// out: v[c] = the physical code with the chosen replica. Returns the number of shared-memory passes of the set (largest bank load).
SVB_HD inline int assign_set(const AssignGeom &G, int *v, unsigned present) {
    int base[16], nc[16], choice[16], owner[16], load[16];
    unsigned dup = 0;
    int padrep = -1;
    for (int c = 0; c < 16; ++c) {
        owner[c] = -1;
        load[c] = 0;
        choice[c] = 0;
        base[c] = 0;
        nc[c] = 1;
        if (!((present >> c) & 1u)) continue;
        const int x = v[c];
        if (!G.adjoint) {
            const int idx = x >> 3;
            base[c] = idx;
            nc[c] = G.nrep;
            if (idx == G.padcanon) {
                if (padrep < 0) padrep = c; else dup |= 1u << c;
            }
        } else if (x == G.padcanon) {
            base[c] = G.pad;
            if (padrep < 0) padrep = c; else dup |= 1u << c;
        } else {
            const int l = x >> G.log2R, il = x & ((1 << G.log2R) - 1);
            if (l < G.nlr) {
                base[c] = l * G.levstride + il;
                nc[c] = G.nrep;
            } else {
                base[c] = G.baseB + ((l - G.nlr) << G.log2R) + il;
            }
        }
    }
    const unsigned live = present & ~dup;
    // single-copy entries first
    for (int c = 0; c < 16; ++c)
        if (((live >> c) & 1u) && nc[c] == 1) {
            const int b = base[c] & 15;
            load[b] += 1;
            owner[b] = -2;
        }
    // matching of the replicated entries (Kuhn's augmenting paths, explicit stack)
    unsigned unmatched = 0;
    for (int c = 0; c < 16; ++c) {
        if (!((live >> c) & 1u) || nc[c] == 1) continue;
        int se[17], sr[17], pb[17];
        unsigned seen = 0;
        int sp = 0;
        se[0] = c;
        sr[0] = 0;
        bool found = false;
        while (sp >= 0) {
            const int en = se[sp];
            if (sr[sp] >= nc[en]) {
                --sp;
                continue;
            }
            const int r = sr[sp]++;
            const int b = (base[en] + r * G.step) & 15;
            if ((seen >> b) & 1u) continue;
            seen |= 1u << b;
            if (owner[b] == -2) continue;
            pb[sp] = b;
            if (owner[b] == -1) {
                for (int k = 0; k <= sp; ++k) {
                    owner[pb[k]] = se[k];
                    choice[se[k]] = sr[k] - 1;
                }
                found = true;
                break;
            }
            se[sp + 1] = owner[b];
            sr[sp + 1] = 0;
            ++sp;
        }
        if (!found) unmatched |= 1u << c;
    }
    for (int b = 0; b < 16; ++b)
        if (owner[b] >= 0) load[b] += 1;
    for (int c = 0; c < 16; ++c)
        if ((unmatched >> c) & 1u) {
            int best = 0, bl = 1 << 30;
            for (int r = 0; r < nc[c]; ++r) {
                const int b = (base[c] + r * G.step) & 15;
                if (load[b] < bl) {
                    bl = load[b];
                    best = r;
                }
            }
            choice[c] = best;
            load[(base[c] + best * G.step) & 15] += 1;
        }
    int passes = 0;
    for (int b = 0; b < 16; ++b) passes = load[b] > passes ? load[b] : passes;
    for (int c = 0; c < 16; ++c)
        if ((present >> c) & 1u) {
            const int src = ((dup >> c) & 1u) ? padrep : c;
            const int phys = base[src] + choice[src] * G.step;
            v[c] = G.adjoint ? phys : (phys << 3);
        }
    return passes;
}

}  // namespace svb
