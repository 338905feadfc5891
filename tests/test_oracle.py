"""CPU tests that pin the oracle (oracle/banzai_oracle.c) before anything trusts it.

Every known-answer vector the reference's own tests hold for the hot path is replayed
(tests/golden/reference_kats.json cites file:line), plus the executable semantics of the
reference's debug oracles (debug/bwt.py, debug/rle1.py), libbz2 round trips (the reference's
fuzz oracle, fuzz/fuzz_targets/round_trip.rs) and the SURVEY §8c whole-stream vectors.
"""
import bz2
import hashlib
import itertools
import json
import os

import numpy as np
import pytest

import corpus
from oracle import pyoracle as O
from tests.golden.make_vectors import vector_input

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KATS = json.load(open(os.path.join(GOLD, "reference_kats.json")))
VECS = json.load(open(os.path.join(GOLD, "survey_vectors.json")))


# ---------------------------------------------------------------- reference KATs

def test_crc32_bzip2_check_value():
    k = KATS["crc32_bzip2_check"]
    assert O.crc32(k["input"].encode()) == int(k["crc"], 16)
    assert O.crc32(b"") == 0


def test_out_bitstring_kat():
    """reference lib/out.rs:113-131"""
    k = KATS["out_bitstring"]
    w = O.BitWriter()
    for op in k["ops"]:
        if op[0] == "bits":
            w.write_bits(op[1], op[2])
        elif op[0] == "byte":
            w.write_byte(op[1])
        else:
            w.write_bytes(bytes(op[1]))
    assert w.close().hex() == k["expect_hex"]


def test_bwt_smoke_kat():
    """reference lib/bwt.rs:762-771"""
    k = KATS["bwt_smoke"]
    out, ptr, has = O.bwt(k["input"].encode())
    assert bytes(out).decode() == k["bwt"]
    assert ptr == k["ptr"]
    assert set(np.nonzero(has)[0]) == set(k["input"].encode())


def test_mtf_smoke_kat():
    """reference lib/mtf.rs:138-158 (dead test, vector still valid)"""
    k = KATS["mtf_smoke"]
    buf = np.array(k["input"], dtype=np.uint8)
    has = np.zeros(256, dtype=np.uint8)
    has[buf] = 1
    syms, num_syms, freqs = O.mtf_and_rle(buf, has)
    assert syms.tolist() == k["expected"]
    assert num_syms == 42
    assert int(freqs.sum()) == len(k["expected"])
    assert freqs[41] == 1


# ---------------------------------------------------------------- debug/*.py semantics

def _debug_bwt(line):
    """reference debug/bwt.py:5-27, on bytes"""
    n = len(line)
    l2 = line + line
    sa = sorted(range(2 * n), key=lambda i: l2[i:])
    outs, ptr = [], -1
    for i in sa:
        if i < n:
            if i == 0:
                ptr = len(outs)
                outs.append(line[-1])
            else:
                outs.append(line[i - 1])
    return bytes(outs), ptr


def _debug_rle1(data):
    """reference debug/rle1.py:11-36 (unbounded RLE1)"""
    outbuf = []
    run_count = 0
    cur = -1
    for b in data:
        if b != cur:
            if run_count >= 4:
                outbuf.append(run_count - 4)
            run_count = 1
            cur = b
            outbuf.append(b)
        else:
            run_count += 1
            if run_count <= 4:
                outbuf.append(b)
            if run_count == 256:
                outbuf.append(run_count - 5)
                run_count = 1
                outbuf.append(b)
    if run_count >= 4:
        outbuf.append(run_count - 4)
    return bytes(outbuf)


def test_bwt_matches_debug_oracle_exhaustive_small():
    for n in range(2, 8):
        for tup in itertools.product(b"abc", repeat=n):
            s = bytes(tup)
            exp, ptr = _debug_bwt(s)
            out, p, _ = O.bwt(s)
            assert (bytes(out), p) == (exp, ptr), s


def test_bwt_tie_rule_examples():
    """SURVEY A-Q5: equal rotations in descending index order"""
    assert O.bwt(bytes(10))[1] == 9
    assert O.bwt(b"ab" * 7)[1] == 6
    assert O.bwt(b"ba" * 7)[1] == 13
    assert O.bwt(b"aa")[1] == 1
    out, p, _ = O.bwt(b"x")
    assert bytes(out) == b"x" and p == 0


def test_bwt_sais_vs_naive_random_and_periodic():
    rng = np.random.default_rng(7)
    cases = []
    for _ in range(150):
        n = int(rng.integers(2, 400))
        sigma = int(rng.integers(1, 6))
        cases.append(rng.integers(0, sigma, n).astype(np.uint8).tobytes())
    for _ in range(60):
        unit = rng.integers(0, 4, int(rng.integers(1, 12))).astype(np.uint8).tobytes()
        reps = int(rng.integers(1, 40))
        extra = int(rng.integers(0, len(unit)))
        cases.append(unit * reps + unit[:extra])
    cases.append(corpus.text(50000).tobytes())
    cases.append(corpus.random_bytes(50000).tobytes())
    cases.append(corpus.source(50000).tobytes())
    for s in cases:
        a = O.bwt(s)
        b = O.bwt_naive(s)
        assert bytes(a[0]) == bytes(b[0]) and a[1] == b[1]
        assert (a[2] == b[2]).all()


def test_rle_matches_debug_oracle_when_unbounded():
    rng = np.random.default_rng(3)
    for _ in range(200):
        n = int(rng.integers(1, 3000))
        # run-heavy data
        vals = rng.integers(0, 3, n)
        lens = rng.integers(1, 300, n)
        data = np.repeat(vals, lens)[:n].astype(np.uint8).tobytes()
        out, consumed, crc = O.rle_one(data, 1)
        if consumed == len(data):
            assert bytes(out) == _debug_rle1(data)
            assert crc == O.crc32(data)


def _boundary_case(rng, level):
    """input whose RLE1 image lands near the block capacity with runs around the cut"""
    cap = 100000 * level - 1
    head = rng.integers(0, 256, cap - int(rng.integers(0, 40))).astype(np.uint8)
    # break accidental runs in the random head so its image length is predictable-ish
    tail_vals = rng.integers(0, 2, 64)
    tail_lens = rng.integers(1, 9, 64) * rng.choice([1, 1, 1, 40, 130], 64)
    tail = np.repeat(tail_vals, tail_lens).astype(np.uint8)
    return np.concatenate([head, tail]).tobytes()


def test_rle_literal_loop_equals_canonical_model():
    """SURVEY A-Q1: the 2-byte-hop loop (lib/rle.rs:133-240) == greedy tokens + capacity rule"""
    rng = np.random.default_rng(11)
    for it in range(120):
        data = _boundary_case(rng, 1)
        a_out, a_cons, _ = O.rle_one(data, 1)
        b_out, b_cons = O.rle_canonical(data, 1)
        assert a_cons == b_cons, it
        assert bytes(a_out) == bytes(b_out), it
    # run-dominated streams, multiple consecutive blocks
    for it in range(20):
        vals = rng.integers(0, 4, 40000)
        lens = rng.integers(1, 12, 40000) * rng.choice([1, 1, 30, 100], 40000)
        data = np.repeat(vals, lens).astype(np.uint8)
        off = 0
        while off < data.size:
            a_out, a_cons, _ = O.rle_one(data[off:], 1)
            b_out, b_cons = O.rle_canonical(data[off:], 1)
            assert a_cons == b_cons and bytes(a_out) == bytes(b_out)
            assert a_cons > 0
            off += a_cons


def test_rle_capacity_rule_examples():
    """SURVEY A-Q1 cases B = 5,4,3,2,1 at a run start"""
    cap = 99999
    for B, (emit, cons) in {5: (5, 10), 4: (3, 3), 3: (3, 3), 2: (2, 2), 1: (1, 1)}.items():
        head = (np.arange(cap - B) % 251 + 1).astype(np.uint8)   # no runs, never 0
        data = np.concatenate([head, np.zeros(10, np.uint8), np.array([7], np.uint8)])
        out, consumed, _ = O.rle_one(data, 1)
        assert out.size == cap - B + emit
        assert consumed == cap - B + cons


# ---------------------------------------------------------------- huffman details

def test_build_table_examples():
    """SURVEY A-Q11"""
    assert O.build_table([0, 0, 0]).tolist() == [2, 2, 1]
    assert O.build_table([1, 1, 2, 3, 5]).tolist() == [3, 3, 2, 2, 2]
    # length limit 17 (huffman.rs:13): fibonacci-like weights force rescaling
    fib = [1, 1]
    while len(fib) < 40:
        fib.append(fib[-1] + fib[-2])
    lens = O.build_table(fib)
    assert lens.max() <= 17 and lens.min() >= 1
    assert sum(2.0 ** -int(l) for l in lens) <= 1.0 + 1e-12


def test_huffman_refinement_quirk_closed_form():
    """SURVEY A-Q10: all selectors 0; table0 = build(A0 + 3G), table t = build(A_t)"""
    for data in (corpus.text(120000), corpus.binary(120000), corpus.random_bytes(60000)):
        rle, _, _ = O.rle_one(data, 9)
        bw, _, has = O.bwt(rle)
        syms, num_syms, freqs = O.mtf_and_rle(bw, has)
        nt, tables, sel = O.huffman_model(syms, num_syms, freqs)
        assert nt == (2 if num_syms <= 199 else 3)
        assert sel.size == -(-syms.size // 50)
        assert not sel.any()
        # iteration-0 assignment under the initial range tables
        m = syms.size
        rem, left, ranges = m, 0, []
        for t in range(nt):
            target = rem // (nt - t)
            acc, right = 0, left
            while True:
                acc += int(freqs[right])
                if acc >= target or right + 1 == num_syms:
                    break
                right += 1
            if right > left and t != 0 and t != nt - 1 and t % 2 == 1:
                acc -= int(freqs[right])
                right -= 1
            ranges.append((left, right))
            left = right + 1
            rem -= acc
        A = np.zeros((nt, num_syms), dtype=np.uint64)
        for g in range(0, m, 50):
            grp = syms[g:g + 50]
            costs = [15 * int(((grp >= lo) & (grp <= hi)).sum()) for lo, hi in ranges]
            best = int(np.argmin(costs))          # first minimum == strict '<' scan
            A[best] += np.bincount(grp, minlength=num_syms).astype(np.uint64)
        G = freqs[:num_syms]
        assert (O.build_table(A[0] + np.uint64(3) * G) == tables[0]).all()
        for t in range(1, nt):
            assert (O.build_table(A[t]) == tables[t]).all()


# ---------------------------------------------------------------- whole stream

@pytest.mark.parametrize("name", sorted(VECS.keys() - {"_comment"}, key=lambda s: int(s[1:])))
def test_survey_vectors(name):
    v = VECS[name]
    data = vector_input(name)
    out, infos = O.encode(data, v["level"], with_info=True)
    if "hex" in v:
        assert out.hex() == v["hex"]
    else:
        assert len(out) == v["len"]
        assert hashlib.sha256(out).hexdigest() == v["sha256"]
    assert [[int(i.consumed), int(i.ptr)] for i in infos] == v["blocks"]
    assert bz2.decompress(out) == bytes(data)


@pytest.mark.parametrize("level", [1, 5, 9])
@pytest.mark.parametrize("kind", ["text", "source", "binary", "mixed", "random"])
def test_round_trip_libbz2(kind, level):
    """reference fuzz/fuzz_targets/round_trip.rs:8-22"""
    data = corpus.by_name(kind, 350000)
    out = O.encode(data, level)
    assert bz2.decompress(out) == data.tobytes()
    assert out[:4] == b"BZh" + str(level).encode()


def test_round_trip_degenerate():
    for data in (bytes(300000), b"ab" * 150000, b"abcdefg" * 40000, b"\xff" * 255 + b"\x00" * 256,
                 corpus.periodic(250000, corpus.random_bytes(1000, seed=corpus.SEED_C3)).tobytes(),
                 b"aaaab" * 50000, bytes(range(256)) * 500):
        for level in (1, 9):
            out = O.encode(data, level)
            assert bz2.decompress(out) == data


def test_level_out_of_range_rejected():
    with pytest.raises(ValueError):
        O.encode(b"abc", 0)
    with pytest.raises(ValueError):
        O.encode(b"abc", 10)
