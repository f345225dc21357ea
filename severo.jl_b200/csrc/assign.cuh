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
SVB_HD inline int assign_set_full(const AssignGeom &G, int *v, unsigned present);

SVB_HD inline unsigned svb_get4(unsigned long long p, int i) { return (unsigned)(p >> (4 * i)) & 15u; }
SVB_HD inline unsigned long long svb_set4(unsigned long long p, int i, unsigned val) {
    return (p & ~(15ull << (4 * i))) | ((unsigned long long)val << (4 * i));
}
SVB_HD inline int svb_ctz(unsigned x) {
#ifdef __CUDA_ARCH__
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

// Greedy pass (every set). Bit masks and selects only in the unrolled parts (no data-dependent branch: the lanes of a warp
// work on 32 different sets, and a first version with branches ran with 7.6 of 32 lanes active — ncu, r04); everything in
// registers.
//   1. decode; the single-copy entries take their bank;
//   2. (effort >= 1) every replicated entry whose replica-0 bank is still free KEEPS it — the stream order already spreads
//      the replica-0 banks, so only the colliding entries have to move;
//   3. the others take their first free candidate bank;
//   4. (effort >= 2) an entry with no free candidate tries ONE exchange: move the owner of one of its candidate banks to a
//      free bank of its own (an augmenting path of length 2; a short data-dependent loop, few entries get here);
//   5. what is left goes to the candidate bank holding the fewest entries (1, else 2, else replica 0).
// Returns the passes of the set after rewriting v; with commit_all = false a set that needs more than one pass is left
// untouched and -1 is returned (it then goes through the full matching below).
SVB_HD inline int assign_set_fast(const AssignGeom &G, int *v, unsigned present, bool commit_all = false, int effort = 2) {
    int base[16];
    unsigned flex = 0, dup = 0;
    unsigned m1 = 0, m2 = 0, m3 = 0;  // banks holding >= 1, 2, 3 entries
    unsigned long long b0pack = 0;    // replica-0 bank per entry
    int padrep = -1;
    const unsigned rmask = (1u << G.log2R) - 1u;
    const bool multi = G.nrep > 1;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int c = 0; c < 16; ++c) {
        const unsigned pres = (present >> c) & 1u;
        const int x = v[c];
        const int l = x >> G.log2R, il = (int)((unsigned)x & rmask);
        const bool ispad = G.adjoint ? (x == G.padcanon) : ((x >> 3) == G.padcanon);
        const bool repl = G.adjoint ? (!ispad && l < G.nlr) : true;
        const int badj = ispad ? G.pad : (l < G.nlr ? l * G.levstride + il : G.baseB + ((l - G.nlr) << G.log2R) + il);
        base[c] = pres ? (G.adjoint ? badj : (x >> 3)) : 0;
        b0pack |= (unsigned long long)(base[c] & 15) << (4 * c);
        const unsigned isdup = pres & (unsigned)(ispad && padrep >= 0);
        padrep = (pres && ispad && padrep < 0) ? c : padrep;
        dup |= isdup << c;
        const unsigned active = pres & ~isdup;
        const unsigned isflex = active & (unsigned)(repl && multi);
        const unsigned bit = (active & ~isflex) ? (1u << (base[c] & 15)) : 0u;
        m3 |= m2 & bit;
        m2 |= m1 & bit;
        m1 |= bit;
        flex |= isflex << c;
    }
    const unsigned step16 = (unsigned)G.step & 15u;
    unsigned choice2 = 0;  // two bits per entry
    unsigned todo = flex;
    unsigned long long ownerp = 0;  // entry holding a bank (valid where flexowned is set)
    unsigned flexowned = 0;
    if (effort >= 1) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int c = 0; c < 16; ++c) {
            const unsigned b0 = (unsigned)base[c] & 15u;
            const unsigned take = ((todo >> c) & 1u) & (unsigned)!((m1 >> b0) & 1u);
            m1 |= take << b0;
            flexowned |= take << b0;
            ownerp = take ? svb_set4(ownerp, (int)b0, (unsigned)c) : ownerp;
            todo &= ~(take << c);
        }
    }
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int c = 0; c < 16; ++c) {
        const unsigned f = (todo >> c) & 1u;
        const unsigned b0 = (unsigned)base[c] & 15u;
        const unsigned k1 = (b0 + step16) & 15u, k2 = (b0 + 2u * step16) & 15u, k3 = (b0 + 3u * step16) & 15u;
        const bool f0 = !((m1 >> b0) & 1u), f1 = !((m1 >> k1) & 1u), f2 = !((m1 >> k2) & 1u) && G.nrep > 2,
                   f3 = !((m1 >> k3) & 1u) && G.nrep > 3;
        const unsigned take = f & (unsigned)(f0 || f1 || f2 || f3);
        const unsigned r = f0 ? 0u : f1 ? 1u : f2 ? 2u : 3u;
        const unsigned bk = f0 ? b0 : f1 ? k1 : f2 ? k2 : k3;
        m1 |= take << bk;
        flexowned |= take << bk;
        ownerp = take ? svb_set4(ownerp, (int)bk, (unsigned)c) : ownerp;
        choice2 |= (take ? r : 0u) << (2 * c);
        todo &= ~(take << c);
    }
    if (effort >= 2) {
        unsigned left = todo;
        while (left) {
            const int c = svb_ctz(left);
            left &= left - 1u;
            const unsigned b0c = svb_get4(b0pack, c);
            bool done = false;
            for (int r = 0; r < G.nrep && !done; ++r) {
                const unsigned b = (b0c + (unsigned)r * step16) & 15u;
                if (!((flexowned >> b) & 1u)) continue;
                const unsigned o = svb_get4(ownerp, (int)b), b0o = svb_get4(b0pack, (int)o);
                for (int r2 = 0; r2 < G.nrep; ++r2) {
                    const unsigned b2 = (b0o + (unsigned)r2 * step16) & 15u;
                    if ((m1 >> b2) & 1u) continue;
                    m1 |= 1u << b2;
                    flexowned |= 1u << b2;
                    ownerp = svb_set4(svb_set4(ownerp, (int)b2, o), (int)b, (unsigned)c);
                    choice2 = (choice2 & ~(3u << (2 * o)) & ~(3u << (2 * c))) | ((unsigned)r2 << (2 * o)) | ((unsigned)r << (2 * c));
                    todo &= ~(1u << c);
                    done = true;
                    break;
                }
            }
        }
    }
    if (todo && !commit_all) return -1;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int c = 0; c < 16; ++c) {
        const unsigned f = (todo >> c) & 1u;
        const unsigned b0 = (unsigned)base[c] & 15u;
        const unsigned bit0 = 1u << b0, bit1 = 1u << ((b0 + step16) & 15u), bit2 = 1u << ((b0 + 2u * step16) & 15u),
                       bit3 = 1u << ((b0 + 3u * step16) & 15u);
        const unsigned cand = bit0 | bit1 | (G.nrep > 2 ? bit2 : 0u) | (G.nrep > 3 ? bit3 : 0u);
        // the lowest occupied level that still leaves a candidate
        const unsigned lvl = (cand & ~m2) ? m2 : (cand & ~m3) ? m3 : 0u;
        const unsigned ok = cand & ~lvl;
        const unsigned r = (ok & bit0) ? 0u : (ok & bit1) ? 1u : (ok & bit2) ? 2u : 3u;
        const unsigned bit = f ? ((ok & bit0) ? bit0 : (ok & bit1) ? bit1 : (ok & bit2) ? bit2 : bit3) : 0u;
        m3 |= m2 & bit;
        m2 |= m1 & bit;
        m1 |= bit;
        choice2 |= (f ? r : 0u) << (2 * c);
    }
    if (m2 && !commit_all) return -1;
    // (the pads of the set share the address chosen for the first of them; no dynamically indexed array: registers only)
    const int padphys = (G.adjoint ? G.pad : G.padcanon) + (padrep >= 0 ? (int)((choice2 >> (2 * padrep)) & 3u) * G.step : 0);
    unsigned long long load_lo = 0, load_hi = 0;  // exact bank loads for the statistic, 8 bits per bank
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int c = 0; c < 16; ++c) {
        const int phys = ((dup >> c) & 1u) ? padphys : base[c] + (int)((choice2 >> (2 * c)) & 3u) * G.step;
        v[c] = ((present >> c) & 1u) ? (G.adjoint ? phys : (phys << 3)) : v[c];
        if (m3) {  // (rare: only then can a bank hold more than 2)
            const unsigned live = (present & ~dup) >> c & 1u;
            const unsigned bk = (unsigned)phys & 15u;
            load_lo += (unsigned long long)(live & (unsigned)(bk < 8)) << (8 * (bk & 7u));
            load_hi += (unsigned long long)(live & (unsigned)(bk >= 8)) << (8 * (bk & 7u));
        }
    }
    if (!present) return 0;
    if (!m2) return 1;
    if (!m3) return 2;
    int passes = 0;
    for (int b = 0; b < 8; ++b) {
        const int lo = (int)((load_lo >> (8 * b)) & 255ull), hi = (int)((load_hi >> (8 * b)) & 255ull);
        passes = lo > passes ? lo : passes;
        passes = hi > passes ? hi : passes;
    }
    return passes;
}

SVB_HD inline int assign_set(const AssignGeom &G, int *v, unsigned present) {
    const int p = assign_set_fast(G, v, present);
    return p >= 0 ? p : assign_set_full(G, v, present);
}

// Full matching. All state is PACKED into registers (4 bits per bank / entry in 64-bit words, 2 bits per choice): the only
// array is base[], indexed statically by the unrolled decode and output loops. (With int arrays for owner / choice / stack the
// kernel kept 560 B of local memory per thread and ran at ~1.3 ns per set; half the sets of a C3 stream come here.)

SVB_HD inline int assign_set_full(const AssignGeom &G, int *v, unsigned present) {
    int base[16];
    unsigned long long b0pack = 0;  // bank of replica 0, per entry
    unsigned flex = 0, dup = 0;
    int padrep = -1;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int c = 0; c < 16; ++c) {  // (selects only, as in the greedy pass)
        const unsigned pres = (present >> c) & 1u;
        const int x = v[c];
        const int l = x >> G.log2R, il = (int)((unsigned)x & ((1u << G.log2R) - 1u));
        const bool ispad = G.adjoint ? (x == G.padcanon) : ((x >> 3) == G.padcanon);
        const bool repl = G.adjoint ? (!ispad && l < G.nlr) : true;
        const int badj = ispad ? G.pad : (l < G.nlr ? l * G.levstride + il : G.baseB + ((l - G.nlr) << G.log2R) + il);
        base[c] = pres ? (G.adjoint ? badj : (x >> 3)) : 0;
        b0pack |= (unsigned long long)(base[c] & 15) << (4 * c);
        const unsigned isdup = pres & (unsigned)(ispad && padrep >= 0);
        padrep = (pres && ispad && padrep < 0) ? c : padrep;
        dup |= isdup << c;
        flex |= (pres & ~isdup & (unsigned)repl) << c;
    }
    if (G.nrep <= 1) flex = 0;
    const unsigned live = present & ~dup;
    const unsigned step16 = (unsigned)G.step & 15u;
    // bank loads: 8 bits per bank in two words (a bank can be hit by all 16 entries)
    unsigned long long load_lo = 0, load_hi = 0;
    auto load_inc = [&](unsigned b) {
        if (b < 8) load_lo += 1ull << (8 * b); else load_hi += 1ull << (8 * (b - 8));
    };
    auto load_get = [&](unsigned b) -> unsigned {
        return (unsigned)((b < 8 ? load_lo >> (8 * b) : load_hi >> (8 * (b - 8))) & 255ull);
    };
    // single-copy entries first: their banks are fixed
    unsigned fixedm = 0, ownedm = 0;
    for (int c = 0; c < 16; ++c)
        if (((live & ~flex) >> c) & 1u) {
            const unsigned b = svb_get4(b0pack, c);
            load_inc(b);
            fixedm |= 1u << b;
        }
    // matching of the replicated entries (Kuhn's augmenting paths; the stack is three packed words, depth <= 16), SEEDED by a
    // greedy pass in straight-line code: first free candidate bank per entry. Only the entries the seed could not place (one to
    // three per set, typically) start an augmenting search. (Searching from every entry in turn: 29 + 37 ms at C3.)
    unsigned long long ownerp = 0;
    unsigned choice2 = 0, unmatched = 0;
    unsigned seed_todo = live & flex;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int c = 0; c < 16; ++c) {
        const unsigned f = (seed_todo >> c) & 1u;
        const unsigned occ = fixedm | ownedm;
        const unsigned b0 = (unsigned)base[c] & 15u;
        const unsigned k1 = (b0 + step16) & 15u, k2 = (b0 + 2u * step16) & 15u, k3 = (b0 + 3u * step16) & 15u;
        const bool f0 = !((occ >> b0) & 1u), f1 = !((occ >> k1) & 1u), f2 = !((occ >> k2) & 1u) && G.nrep > 2,
                   f3 = !((occ >> k3) & 1u) && G.nrep > 3;
        const unsigned take = f & (unsigned)(f0 || f1 || f2 || f3);
        const unsigned r = f0 ? 0u : f1 ? 1u : f2 ? 2u : 3u;
        const unsigned bk = f0 ? b0 : f1 ? k1 : f2 ? k2 : k3;
        ownedm |= take << bk;
        ownerp = take ? svb_set4(ownerp, (int)bk, (unsigned)c) : ownerp;
        choice2 |= (take ? r : 0u) << (2 * c);
        seed_todo &= ~(take << c);
    }
    // Augmenting searches of the entries the seed left, breadth-first over BANK SETS: 16-bit masks, one rotation per replica.
    // All entries with the same replica-0 bank x have the same candidates {x, x+s, x+2s, x+3s} (s = step mod 16), so
    // reachability is a property of banks: from a set S of occupied banks, R[r] & S are those whose owner sits on its replica
    // r, rotating them back by r*s gives the owners' replica-0 banks, rotating forward by every r' their candidates.
    // A layer costs ~30 instructions whatever its size; a depth-first search with an explicit stack cost ~100 per step
    // and up to 80 steps for an entry that cannot be matched (18 + 21 ms at C3 even after the seed).
    {
        // R[r]: banks whose (replicated) owner sits on its replica r
        unsigned R0 = 0, R1 = 0, R2 = 0, R3 = 0;
        for (unsigned bk = 0; bk < 16; ++bk) {
            if (!((ownedm >> bk) & 1u)) continue;
            const unsigned r = (choice2 >> (2 * svb_get4(ownerp, (int)bk))) & 3u;
            R0 |= (unsigned)(r == 0u) << bk;
            R1 |= (unsigned)(r == 1u) << bk;
            R2 |= (unsigned)(r == 2u) << bk;
            R3 |= (unsigned)(r == 3u) << bk;
        }
        const unsigned s1 = step16, s2 = (2u * step16) & 15u, s3 = (3u * step16) & 15u;
        auto rotl = [](unsigned x, unsigned k) { return ((x << k) | (x >> (16u - k))) & 0xffffu; };  // (k = 0: x | x)
        auto rotr = [](unsigned x, unsigned k) { return ((x >> k) | (x << (16u - k))) & 0xffffu; };
        const unsigned n2 = G.nrep > 2 ? 0xffffu : 0u, n3 = G.nrep > 3 ? 0xffffu : 0u;
        auto spread = [&](unsigned e0) { return e0 | rotl(e0, s1) | (rotl(e0, s2) & n2) | (rotl(e0, s3) & n3); };
        unsigned todo = seed_todo, dead = 0;
        while (todo) {
            const int c = svb_ctz(todo);
            todo &= todo - 1u;
            const unsigned b0c = svb_get4(b0pack, c);
            // layers of the search: 16 masks of 16 bits in four words
            unsigned long long lay0 = 0, lay1 = 0, lay2 = 0, lay3 = 0;
            auto lay_get = [&](int k) -> unsigned {
                const unsigned long long w = (k >> 2) == 0 ? lay0 : (k >> 2) == 1 ? lay1 : (k >> 2) == 2 ? lay2 : lay3;
                return (unsigned)(w >> (16 * (k & 3))) & 0xffffu;
            };
            auto lay_set = [&](int k, unsigned m) {
                const unsigned long long w = (unsigned long long)m << (16 * (k & 3));
                lay0 |= (k >> 2) == 0 ? w : 0ull;
                lay1 |= (k >> 2) == 1 ? w : 0ull;
                lay2 |= (k >> 2) == 2 ? w : 0ull;
                lay3 |= (k >> 2) == 3 ? w : 0ull;
            };
            const unsigned blocked = fixedm | dead;
            unsigned cur = spread(1u << b0c) & ~blocked, visited = cur;
            int k = 0;
            bool found = false;
            while (cur) {
                lay_set(k, cur);
                if (cur & ~ownedm) {
                    found = true;
                    break;
                }
                const unsigned e0 = rotr(cur & R0, 0u) | rotr(cur & R1, s1) | rotr(cur & R2, s2) | rotr(cur & R3, s3);
                cur = spread(e0) & ~blocked & ~visited;
                visited |= cur;
                if (k >= 15) break;  // (cannot happen: 16 banks)
                ++k;
            }
            if (!found) {
                unmatched |= 1u << c;
                dead |= visited;
                continue;
            }
            // walk back: the free bank goes to an owner of the previous layer, whose bank goes to one of the layer before, ...
            unsigned tgt = (unsigned)svb_ctz(lay_get(k) & ~ownedm);
            for (int j = k; j >= 1; --j) {
                const unsigned prev = lay_get(j - 1);
                unsigned bsel = 0, rsel = 0;
                bool got = false;
                for (unsigned rp = 0; rp < (unsigned)G.nrep && !got; ++rp) {
                    const unsigned x = (tgt - rp * step16) & 15u;  // replica-0 bank of an entry that has tgt as candidate rp
                    const unsigned m = ((R0 & (1u << x)) | (R1 & (1u << ((x + s1) & 15u))) | (R2 & (1u << ((x + s2) & 15u))) |
                                        (R3 & (1u << ((x + s3) & 15u)))) & prev;
                    if (m) {
                        bsel = (unsigned)svb_ctz(m);
                        rsel = rp;
                        got = true;
                    }
                }
                const unsigned o = svb_get4(ownerp, (int)bsel);
                const unsigned tb = ~(1u << tgt);
                R0 = (R0 & tb) | ((unsigned)(rsel == 0u) << tgt);
                R1 = (R1 & tb) | ((unsigned)(rsel == 1u) << tgt);
                R2 = (R2 & tb) | ((unsigned)(rsel == 2u) << tgt);
                R3 = (R3 & tb) | ((unsigned)(rsel == 3u) << tgt);
                ownerp = svb_set4(ownerp, (int)tgt, o);
                ownedm |= 1u << tgt;
                choice2 = (choice2 & ~(3u << (2 * o))) | (rsel << (2 * o));
                tgt = bsel;
            }
            {
                const unsigned rsel = tgt == b0c ? 0u : tgt == ((b0c + s1) & 15u) ? 1u : tgt == ((b0c + s2) & 15u) ? 2u : 3u;
                const unsigned tb = ~(1u << tgt);
                R0 = (R0 & tb) | ((unsigned)(rsel == 0u) << tgt);
                R1 = (R1 & tb) | ((unsigned)(rsel == 1u) << tgt);
                R2 = (R2 & tb) | ((unsigned)(rsel == 2u) << tgt);
                R3 = (R3 & tb) | ((unsigned)(rsel == 3u) << tgt);
                ownerp = svb_set4(ownerp, (int)tgt, (unsigned)c);
                ownedm |= 1u << tgt;
                choice2 = (choice2 & ~(3u << (2 * c))) | (rsel << (2 * c));
            }
        }
    }
    for (unsigned b = 0; b < 16; ++b)
        if ((ownedm >> b) & 1u) load_inc(b);
    for (int c = 0; c < 16; ++c)
        if ((unmatched >> c) & 1u) {
            unsigned best = 0, bl = 1u << 30;
            for (int r = 0; r < G.nrep; ++r) {
                const unsigned b = (svb_get4(b0pack, c) + (unsigned)r * step16) & 15u;
                if (load_get(b) < bl) {
                    bl = load_get(b);
                    best = (unsigned)r;
                }
            }
            choice2 = (choice2 & ~(3u << (2 * c))) | (best << (2 * c));
            load_inc((svb_get4(b0pack, c) + best * step16) & 15u);
        }
    int passes = 0;
    for (unsigned b = 0; b < 16; ++b) passes = (int)load_get(b) > passes ? (int)load_get(b) : passes;
    const int padphys = (G.adjoint ? G.pad : G.padcanon) + (padrep >= 0 ? (int)((choice2 >> (2 * padrep)) & 3u) * G.step : 0);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int c = 0; c < 16; ++c)
        if ((present >> c) & 1u) {
            const int phys = ((dup >> c) & 1u) ? padphys : base[c] + (int)((choice2 >> (2 * c)) & 3u) * G.step;
            v[c] = G.adjoint ? phys : (phys << 3);
        }
    return passes;
}

}  // namespace svb
