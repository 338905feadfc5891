"""BASELINE.json configs C1..C5: byte parity against the oracle on EVERY config at its
full size (block-parallel oracle driver), libbz2 round trip where affordable, and throughput from
pinned host buffers (the e2e path of bench.py), one line each.
usage: python tools/run_configs.py [quick|big|small]"""
import bz2, ctypes as C, hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import corpus, banzai_b200
from banzai_b200 import _ffi
from oracle import pyoracle as O

mode = sys.argv[1] if len(sys.argv) > 1 else "all"
lib = _ffi.lib
ctx = banzai_b200.Context(n_gpus=1)


def run(name, data, level, decode_check=True):
    data = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    n = data.size
    h_in = lib.bnz_host_alloc(max(n, 1))
    h_arr = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_uint8)), shape=(max(n, 1),))
    h_arr[:n] = data
    best, shas, out_n = 1e9, set(), 0
    for _ in range(3):
        t0 = time.perf_counter()
        o, out_n = ctx.encode_ptr(h_in, n, level)
        best = min(best, time.perf_counter() - t0)
        shas.add(hashlib.sha256((C.c_uint8 * out_n).from_address(o.value)).hexdigest())
        if decode_check and len(shas) == 1 and _ == 0:
            ok = bz2.decompress(bytes((C.c_uint8 * out_n).from_address(o.value))) == data.tobytes()
        ctx.free_out(o)
    st = ctx.stats()
    sha_o, len_o, nb_o = O.encode_mt(data, level, digest=True)
    res = ["oracle-identical" if (len(shas) == 1 and sha_o in shas and len_o == out_n) else "MISMATCH vs oracle"]
    if decode_check:
        res.append("libbz2 round trip ok" if ok else "ROUND TRIP FAILED")
    res.append("deterministic" if len(shas) == 1 else "NONDETERMINISTIC")
    alg = st["bwt_algorithmic_bytes"]
    print(f"{name}: {n} B L{level} -> {out_n} B ({out_n / max(1, n):.4f}), blocks {st['n_blocks']}, "
          f"e2e {n / best / 1e6:.0f} MB/s (pinned host buffer -> .bz2 in host memory, best of 3), device stages ms "
          f"rle {st['rle_ms']:.1f} bwt {st['bwt_ms']:.1f} mtf {st['mtf_ms']:.1f} huff {st['huff_ms']:.1f} pack {st['pack_ms']:.1f}; "
          f"bwt rounds max {st['bwt_max_rounds']}, sum_active/n {st['bwt_sum_active'] / max(1, st['bwt_n']):.2f}, "
          f"tied blocks {st['bwt_tied_blocks']}, sort {alg / max(st['bwt_ms'], 1e-3) / 1e6:.0f} GB/s algorithmic "
          f"({alg / max(st['bwt_ms'], 1e-3) / 1e6 / 6550.1 * 100:.1f}% of 6550); " + ", ".join(res), flush=True)
    lib.bnz_host_free(h_in)


def small():
    run("C1 text-10MB", corpus.text(10 * 1000 * 1000, corpus.SEED_C1), 9)
    unit = corpus.random_bytes(1000, seed=corpus.SEED_C3).tobytes()
    big = (8 if mode == "quick" else 64) << 20
    for nm, d in (("zeros", bytes(big)), ("ab", b"ab" * (big // 2)), ("abcdefg", (b"abcdefg" * (big // 7 + 1))[:big]),
                  ("period1000", (unit * (big // 1000 + 1))[:big])):
        for lvl in (9, 1):
            run(f"C3 {nm}-{big >> 20}MiB", d, lvl)
    for nm, d in (("abcdefg x1000 (period | n)", b"abcdefg" * 1000), ("period1000 x1000 (period | n)", unit * 1000), ("aa", b"aa")):
        run(f"C3 {nm}", d, 9)


def big():
    run("C2 mixed-1GiB", corpus.mixed(1 << 30, corpus.SEED_C2), 9, decode_check=False)
    run("C4 text-4GiB-L1", corpus.text(4 << 30, corpus.SEED_C4), 1, decode_check=False)
    run("C5 random-4GiB-L9", corpus.random_bytes(4 << 30, corpus.SEED_C5), 9, decode_check=False)


if mode in ("all", "quick", "small"):
    small()
if mode in ("all", "big"):
    big()
