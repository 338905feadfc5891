"""BASELINE.json configs C1..C5: parity / property checks and throughput, one line each.
usage: python tools/run_configs.py [quick]"""
import bz2, hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import corpus, banzai_b200
from oracle import pyoracle as O

quick = len(sys.argv) > 1 and sys.argv[1] == 'quick'
only_c5 = len(sys.argv) > 1 and sys.argv[1] == 'c5'
only_big = len(sys.argv) > 1 and sys.argv[1] == 'big'     # C2, C4, C5 only (the cases too large for the test suite)
ctx = banzai_b200.Context(n_gpus=1)

def run(name, data, level, oracle_check, decode_check=True):
    t0 = time.perf_counter(); out = ctx.encode_bytes(data, level); dt = time.perf_counter() - t0
    t0 = time.perf_counter(); out2 = ctx.encode_bytes(data, level); dt = min(dt, time.perf_counter() - t0)
    st = ctx.stats()
    res = []
    if oracle_check:
        res.append("oracle-identical" if out == O.encode(data, level) else "MISMATCH vs oracle")
    if decode_check:
        ok = bz2.decompress(out) == (data.tobytes() if isinstance(data, np.ndarray) else bytes(data))
        res.append("libbz2 round trip ok" if ok else "ROUND TRIP FAILED")
    res.append("deterministic" if out == out2 else "NONDETERMINISTIC")
    n = len(data)
    print(f"{name}: {n} B L{level} -> {len(out)} B ({len(out) / max(1, n):.4f}), blocks {st['n_blocks']}, "
          f"e2e {n / dt / 1e6:.0f} MB/s (host buffers, unpinned), device stages ms rle {st['rle_ms']:.1f} bwt {st['bwt_ms']:.1f} "
          f"mtf {st['mtf_ms']:.1f} huff {st['huff_ms']:.1f} pack {st['pack_ms']:.1f}; bwt rounds max {st['bwt_max_rounds']}, "
          f"tied blocks {st['bwt_tied_blocks']}; " + ", ".join(res), flush=True)

if only_c5:
    run("C5 random-4GiB-L9", corpus.random_bytes(4 << 30, corpus.SEED_C5), 9, False)
    sys.exit(0)
# C1: 10 MB English-like text, level 9 (the CPU-runnable case; must be byte-identical)
if only_big:
    run("C2 mixed-1GiB", corpus.mixed(1 << 30, corpus.SEED_C2), 9, False)
    run("C4 text-4GiB-L1", corpus.text(4 << 30, corpus.SEED_C4), 1, False)
    run("C5 random-4GiB-L9", corpus.random_bytes(4 << 30, corpus.SEED_C5), 9, False)
    sys.exit(0)
run("C1 text-10MB", corpus.text(10 * 1000 * 1000, corpus.SEED_C1), 9, True)
# C3: degenerate / periodic
unit = corpus.random_bytes(1000, seed=corpus.SEED_C3).tobytes()
big = (8 if quick else 64) << 20
for nm, d in (("zeros", bytes(big)), ("ab", b"ab" * (big // 2)), ("abcdefg", (b"abcdefg" * (big // 7 + 1))[:big]),
              ("period1000", (unit * (big // 1000 + 1))[:big])):
    for lvl in (9, 1):
        run(f"C3 {nm}-{big >> 20}MiB", d, lvl, not (nm != "zeros" and big > (16 << 20) and False))
for nm, d in (("abcdefg x1000 (period | n)", b"abcdefg" * 1000), ("period1000 x1000 (period | n)", unit * 1000), ("aa", b"aa")):
    run(f"C3 {nm}", d, 9, True)
if not quick:
    # C2 (per GPU share is what bench.py measures); here: properties on the full 1 GiB
    run("C2 mixed-1GiB", corpus.mixed(1 << 30, corpus.SEED_C2), 9, False)
    # C4: 4 GiB text at level 1 (many small blocks), C5: 4 GiB random at level 9
    run("C4 text-4GiB-L1", corpus.text(4 << 30, corpus.SEED_C4), 1, False)
    run("C5 random-4GiB-L9", corpus.random_bytes(4 << 30, corpus.SEED_C5), 9, False)
