"""CPU study (numpy): sizes of the groups the doubling rounds of the BWT sort work on, per round, weighted by
records, for one level-9 block of each corpus kind — with the packed round-0 key of csrc/bwt_common.cuh."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import corpus

def study(kind, n=899999):
    S = np.frombuffer(corpus.by_name(kind, n).tobytes(), dtype=np.uint8).astype(np.int64)
    syms = np.unique(S); sigma = len(syms)
    code = np.zeros(256, dtype=np.int64); code[syms] = np.arange(sigma)
    k = 5 if sigma > 101 else 6 if sigma > 52 else 7 if sigma > 32 else 8
    L = min(sigma, (1 << 40) // sigma ** k)
    c = code[S]
    idx = np.arange(n)
    key = np.zeros(n, dtype=np.int64)
    for j in range(k):
        key = key * sigma + c[(idx + j) % n]
    key = key * L + (c[(idx + k) % n] * L) // sigma
    order = np.argsort(key, kind="stable")
    ks = key[order]
    head = np.r_[True, ks[1:] != ks[:-1]]
    gid = np.cumsum(head) - 1
    start = np.flatnonzero(head)
    rank = np.empty(n, dtype=np.int64); rank[order] = start[gid]
    h = k
    print(f"{kind}: sigma {sigma}, k {k}, L {L}")
    for rnd in range(1, 8):
        sizes = np.bincount(gid)
        act = sizes[gid] > 1                      # per sorted position
        if not act.any(): break
        sz = sizes[gid][act]
        tot = act.sum()
        bins = [2, 3, 4, 8, 16, 32, 64, 256, 4096, 1 << 30]
        frac = []
        lo = 2
        for b in bins:
            frac.append(((sz >= lo) & (sz <= b)).sum() / tot); lo = b + 1
        print(f"  round {rnd} (h={h}): active {tot / n:.3f} n; records by group size <=2,3,4,8,16,32,64,256,4096,more: " + " ".join(f"{f:.2f}" for f in frac))
        # refine
        r2 = rank[(order + h) % n]
        o2 = np.lexsort((r2, rank[order]))
        order = order[o2]
        k1 = rank[order]; k2 = rank[(order + h) % n]
        head = np.r_[True, (k1[1:] != k1[:-1]) | (k2[1:] != k2[:-1])]
        gid = np.cumsum(head) - 1
        start = np.flatnonzero(head)
        rank = np.empty(n, dtype=np.int64); rank[order] = start[gid]
        h *= 2

for kind in (sys.argv[1:] or ["text", "source", "binary"]):
    study(kind)
