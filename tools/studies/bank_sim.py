"""Shared-memory bank-conflict simulator for the count-level stream layouts (csrc/factored.cu), on the host sample written by
make_hvg_sample.py. A half-warp gathers element e of 16 consecutive chunks (an aligned block of 16 chunks): the number of
passes is the largest number of DISTINCT addresses that fall in one 8-byte bank (16 banks). Reports passes per set for
  fwd: the forward stream (cell-major, groups = (cell, level)), bank = gene mod 16, optional bank-shifted replicas of xs
  adj: the adjoint stream (tile, gene) segments, bank = cell mod 16, optional replicas
under the current round-robin placement and under per-set greedy replica assignment."""
import sys
import numpy as np

PAD = -1


def rr_sequence(res):
    """indices of the entries in round-robin order over the residue classes (class after class, j-th entries)"""
    order = np.lexsort((res, np.zeros_like(res)))  # stable by residue
    res_sorted = res[order]
    # rank of each entry inside its class
    starts = np.searchsorted(res_sorted, np.arange(16))
    rank = np.arange(len(res)) - starts[res_sorted]
    key = rank * 16 + res_sorted
    return order[np.argsort(key, kind="stable")]


def place_group(block_codes, g0, g1, seq_codes):
    """write the sequence into the slots of chunks [g0, g1) set by set (rr_slot of factored.cu)"""
    p = 0
    n = len(seq_codes)
    c = g0
    while c < g1 and p < n:
        blk_end = min(g1, ((c >> 4) + 1) << 4)
        wb = blk_end - c
        for e in range(8):
            for k in range(wb):
                if p < n:
                    block_codes[c + k, e] = seq_codes[p]
                    p += 1
        c = blk_end


def passes_of_set(addrs, banks_of):
    """addrs: the 16 addresses of a set (PAD = the shared pad address). banks_of(addr) -> bank."""
    uniq = set(a for a in addrs)
    cnt = [0] * 16
    for a in uniq:
        cnt[banks_of(a)] += 1
    return max(cnt)


def greedy_replicas(addrs, bank_of, shifts):
    """choose a replica per distinct address to minimise the largest bank load (greedy, most constrained first)"""
    uniq = list(set(addrs))
    load = [0] * 16
    # process addresses in order; pick the replica whose bank is least loaded
    for a in uniq:
        b0 = bank_of(a)
        best = min(shifts, key=lambda s: load[(b0 + s) & 15])
        load[(b0 + best) & 15] += 1
    return max(load)


def matching_replicas(addrs, bank_of, shifts):
    """best assignment by augmenting paths: is a perfect (1 pass) assignment possible? else fall back to greedy"""
    uniq = list(set(addrs))
    owner = [-1] * 16

    def try_assign(i, seen):
        b0 = bank_of(uniq[i])
        for s in shifts:
            b = (b0 + s) & 15
            if b in seen:
                continue
            seen.add(b)
            if owner[b] < 0 or try_assign(owner[b], seen):
                owner[b] = i
                return True
        return False

    ok = 0
    for i in range(len(uniq)):
        if try_assign(i, set()):
            ok += 1
    return 1 if ok == len(uniq) else 2 if len(uniq) - ok <= 16 else 3  # unmatched ones go to a second pass


def simulate(chunks, bank_of, label, shift_sets):
    nblocks = len(chunks) // 16
    tot_sets = 0
    base = 0
    res = {k: 0 for k in shift_sets}
    resm = {k: 0 for k in shift_sets}
    for b in range(nblocks):
        blk = chunks[b * 16:(b + 1) * 16]
        for e in range(8):
            addrs = [int(x) for x in blk[:, e]]
            tot_sets += 1
            base += passes_of_set(addrs, bank_of)
            for k, sh in shift_sets.items():
                res[k] += greedy_replicas(addrs, bank_of, sh)
                resm[k] += matching_replicas(addrs, bank_of, sh)
    print(f"{label}: sets {tot_sets}, current placement {base / tot_sets:.3f} passes/set")
    for k in shift_sets:
        print(f"   replicas {k}: greedy {res[k] / tot_sets:.3f}, matching {resm[k] / tot_sets:.3f}")


def forward_chunks(colptr, rowval, counts, cells, ncells, L):
    n = len(colptr) - 1
    gene = np.repeat(np.arange(n), np.diff(colptr))
    sel = rowval < ncells
    order = np.lexsort((gene[sel], rowval[sel]))
    r, g, c = rowval[sel][order], gene[sel][order], counts[sel][order]
    rowptr = np.searchsorted(r, np.arange(ncells + 1))
    out = []
    nex = 0
    for i in range(ncells):
        gi, ci = g[rowptr[i]:rowptr[i + 1]], c[rowptr[i]:rowptr[i + 1]]
        for l in range(1, L + 1):
            gl = gi[ci == l]
            if len(gl) == 0:
                continue
            nch = (len(gl) + 7) // 8
            base = len(out)
            out.extend([[PAD] * 8 for _ in range(nch)])
            arr = np.array(out[base:base + nch])
            seq = gl[rr_sequence(gl & 15)]
            tmp = np.full((base + nch, 8), PAD)
            place_group(tmp, base, base + nch, seq)
            for k in range(nch):
                out[base + k] = list(tmp[base + k])
        nex += int((ci > L).sum())
        out.extend([[PAD] * 8 for _ in range(int((ci > L).sum()))])  # exception chunks: one gather of xs each (ignored)
    ch = np.array(out)
    return ch, nex


def adjoint_chunks(colptr, rowval, counts, R, tiles, L):
    n = len(colptr) - 1
    out = []
    for t in range(tiles):
        lo, hi = t * R, (t + 1) * R
        for j in range(n):
            rv = rowval[colptr[j]:colptr[j + 1]]
            a, b = np.searchsorted(rv, lo), np.searchsorted(rv, hi)
            il = rv[a:b] - lo
            cv = counts[colptr[j] + a:colptr[j] + b]
            ok = cv <= L
            il, cv = il[ok], cv[ok]
            nch = max(1, (len(il) + 7) // 8)
            base = len(out)
            tmp = np.full((base + nch, 8), PAD)
            if len(il):
                codes = (cv - 1) * R + il
                seq = codes[rr_sequence(il & 15)]
                place_group(tmp, base, base + nch, seq)
            out.extend([list(x) for x in tmp[base:base + nch]])
            out.extend([[PAD] * 8 for _ in range(int((~ok).sum()))])
    return np.array(out)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1].isdigit():
    d = np.load("/tmp/hvg_sample.npz")
    colptr, rowval, counts = d["colptr"], d["rowval"], d["counts"]
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    shift_sets = {"x2 {0,8}": (0, 8), "x2 {0,5}": (0, 5), "x4 {0,4,8,12}": (0, 4, 8, 12), "x4 {0,5,10,15}": (0, 5, 10, 15),
                  "x4 {0,1,2,3}": (0, 1, 2, 3)}
    ch, nex = forward_chunks(colptr, rowval, counts, int(d["cells"]), 600, L)
    print("forward: chunks", len(ch), "pads", int((ch == PAD).sum()), "of", ch.size)
    simulate(ch, lambda a: (a & 15) if a != PAD else 0, "fwd", shift_sets)
    R = 16384 // L
    ch = adjoint_chunks(colptr, rowval, counts, R, 1, L)
    print("adjoint: R", R, "chunks", len(ch), "pads", int((ch == PAD).sum()), "of", ch.size)
    simulate(ch, lambda a: (a & 15) if a != PAD else 0, "adj", shift_sets)


def matching_partial(addrs, bank_of, shifts, is_flex):
    """fixed addresses take their bank first (their own collisions cost passes); flexible ones are matched into the rest"""
    uniq = list(set(addrs))
    fixed = [a for a in uniq if not is_flex(a)]
    flex = [a for a in uniq if is_flex(a)]
    load = [0] * 16
    for a in fixed:
        load[bank_of(a)] += 1
    # capacity model: try to fit everything in P passes, P = 1, 2, 3: bank capacity P
    for P in (1, 2, 3, 4):
        if max(load) > P:
            continue
        cap = [P - x for x in load]
        owner = [[] for _ in range(16)]

        def try_assign(i, seen):
            b0 = bank_of(flex[i])
            for s in shifts:
                b = (b0 + s) & 15
                if b in seen:
                    continue
                seen.add(b)
                if len(owner[b]) < cap[b]:
                    owner[b].append(i)
                    return True
                for k, o in enumerate(owner[b]):
                    if try_assign(o, seen):
                        owner[b][k] = i
                        return True
            return False

        if all(try_assign(i, set()) for i in range(len(flex))):
            return P
    return 5


def simulate_partial(chunks, R, label, configs):
    nblocks = len(chunks) // 16
    for name, (shifts, nlev) in configs.items():
        tot = 0
        sets = 0
        for b in range(nblocks):
            blk = chunks[b * 16:(b + 1) * 16]
            for e in range(8):
                addrs = [int(x) for x in blk[:, e]]
                sets += 1
                tot += matching_partial(addrs, lambda a: (a & 15) if a != PAD else 0, shifts,
                                        lambda a: a == PAD or (a // R) < nlev)
        kb = (R * nlev * len(shifts) + R * (16 - nlev)) * 8 / 1024
        print(f"{label} {name}: {tot / sets:.3f} passes/set, table {kb:.0f} KB")


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[2] == "partial":
    for R in (1024, 512):
        ch = adjoint_chunks(colptr, rowval, counts, R, 1 if R == 1024 else 2, L)
        print("adjoint R", R, "chunks", len(ch))
        cfgs = {"lev1-2 x4 {0,5,10,15}": ((0, 5, 10, 15), 2), "lev1-3 x3 {0,5,10}": ((0, 5, 10), 3), "lev1-4 x2 {0,5}": ((0, 5), 4),
                "lev1-2 x3 {0,5,10}": ((0, 5, 10), 2), "lev1 x4": ((0, 5, 10, 15), 1), "all x2 {0,5}": ((0, 5), 16),
                "lev1-4 x3": ((0, 5, 10), 4), "lev1-3 x4": ((0, 5, 10, 15), 3)}
        simulate_partial(ch, R, f"adj R={R}", cfgs)


def device_like(chunks, R, nlev, shifts, pad_flex, dedupe):
    nblocks = len(chunks) // 16
    tot = sets = 0
    for b in range(nblocks):
        blk = chunks[b * 16:(b + 1) * 16]
        for e in range(8):
            addrs = [int(x) for x in blk[:, e]]
            pads = [a for a in addrs if a == PAD]
            real = [a for a in addrs if a != PAD]
            if dedupe:
                real = list(set(real))
            ent = real + ([PAD] if pads else [])
            fixed = [a for a in ent if not ((a == PAD and pad_flex) or (a != PAD and (a // R) < nlev))]
            flex = [a for a in ent if a not in fixed or (a == PAD and pad_flex)]
            flex = [a for a in ent if ((a == PAD and pad_flex) or (a != PAD and (a // R) < nlev))]
            load = [0] * 16
            owner = [-1] * 16
            for a in fixed:
                bb = 0 if a == PAD else (a & 15)
                load[bb] += 1
                owner[bb] = -2
            choice = {}
            unmatched = []
            fl = list(range(len(flex)))
            own = {}

            def bank(i, r):
                a = flex[i]
                return ((0 if a == PAD else a) + shifts[r]) & 15

            def aug(i, seen):
                for r in range(len(shifts)):
                    bb = bank(i, r)
                    if bb in seen:
                        continue
                    seen.add(bb)
                    if owner[bb] == -2:
                        continue
                    if owner[bb] == -1 or aug(owner[bb], seen):
                        owner[bb] = i
                        return True
                return False

            for i in fl:
                if not aug(i, set()):
                    unmatched.append(i)
            for bb in range(16):
                if owner[bb] >= 0:
                    load[bb] += 1
            for i in unmatched:
                r = min(range(len(shifts)), key=lambda r: load[bank(i, r)])
                load[bank(i, r)] += 1
            sets += 1
            tot += max(load) if ent else 0
    return tot / sets


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[2] == "device":
    R = 1024
    ch = adjoint_chunks(colptr, rowval, counts, R, 1, L)
    for pf in (False, True):
        for dd in (False, True):
            print("adj device-like: pad_flex", pf, "dedupe", dd, device_like(ch, R, 2, (0, 5, 10, 15), pf, dd))
