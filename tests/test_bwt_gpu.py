"""GPU parity: K3/K4 BWT rotation sort vs the oracle's `bwt::bwt` (reference lib/bwt.rs:526).
Bit-exact: BWT bytes, origPtr and has_byte must all match."""
import json
import os

import numpy as np
import pytest

import corpus
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))


@pytest.fixture(scope="module")
def ctx():
    import banzai_b200
    c = banzai_b200.Context(n_gpus=1)
    yield c
    c.close()


MODES = [("cluster", 8), ("cluster", 16), ("cluster", 2), ("single", 0)]


def _check(ctx, blocks, level=9):
    for kind, val in MODES:
        ctx.set("bwt_cluster", val)
        got = ctx.stage_bwt(blocks, level, with_stats=True)
        for blk, (bw, ptr, has, st) in zip(blocks, got):
            ebw, eptr, ehas = O.bwt(blk)
            assert ptr == eptr, (len(blk), st)
            assert bytes(bw) == bytes(ebw), (len(blk), st)
            assert (has == ehas).all()
    ctx.set("bwt_cluster", -1)
    return got


def test_reference_kat_sentence(ctx):
    k = KATS["bwt_smoke"]
    got = _check(ctx, [k["input"].encode()])
    assert bytes(got[0][0]).decode() == k["bwt"] and got[0][1] == k["ptr"]


def test_tiny_and_edge_lengths(ctx):
    blocks = [b"x", b"aa", b"ab", b"ba", b"aaa", b"abc", b"abab", b"aaaa", b"abcab", b"abcabc",
              bytes(10), b"ab" * 7, b"ba" * 7, b"abcdefg" * 1000, bytes(range(256))]
    _check(ctx, blocks, level=1)


def test_random_small_alphabets(ctx):
    rng = np.random.default_rng(5)
    blocks = []
    for _ in range(200):
        n = int(rng.integers(1, 3000))
        sigma = int(rng.integers(1, 5))
        blocks.append(rng.integers(0, sigma, n).astype(np.uint8).tobytes())
    _check(ctx, blocks, level=1)


def test_periodic_equal_rotations_and_deep_doubling(ctx):
    unit = corpus.random_bytes(1000, seed=corpus.SEED_C3).tobytes()
    blocks = [
        b"abcdefg" * 1000,                       # period | n  -> equal rotations (V7: ptr 999)
        unit * 100,                              # period 1000 | 100000
        (b"ab" * 50000)[:99999],                 # period 2 does not divide n -> deep doubling
        (b"abcdefg" * 15000)[:99999],
        (unit * 100)[:99999],
        bytes(99999),                            # one symbol
        b"\x00\x00\x00\x00\xfb" * 19999 + b"\x00\x00\x00",   # what RLE1 makes of zeros
    ]
    got = _check(ctx, blocks, level=1)
    assert got[0][1] == 999
    assert got[0][3]["tied"] == 1 and got[1][3]["tied"] == 1
    assert got[2][3]["tied"] == 0


@pytest.mark.parametrize("kind", ["text", "source", "binary", "random"])
def test_full_blocks_level9(ctx, kind):
    data = corpus.by_name(kind, 2 * 899999 + 12345)
    blocks = [data[:899999].tobytes(), data[899999:2 * 899999].tobytes(), data[2 * 899999:].tobytes()]
    _check(ctx, blocks, level=9)


def test_many_blocks_more_than_ctas(ctx):
    data = corpus.mixed(700 * 20000)
    blocks = [data[i * 20000:(i + 1) * 20000].tobytes() for i in range(700)]
    _check(ctx, blocks, level=1)


def test_periodic_runs_closed_form(ctx):
    """Blocks with a long periodic run are ordered in closed form (csrc/bwt_common.cuh: Period) —
    the reference's SA-IS has no bad case for them (README.md:7).  Runs that do not reach the block's
    ends (what RLE1 makes of zero pages: partial first and last run), two runs around an odd byte,
    damaged runs, both directions; the cluster kernel leaves such blocks to the one-CTA kernel."""
    unit = corpus.random_bytes(1000, seed=corpus.SEED_C3).tobytes()
    rng = np.random.default_rng(11)

    def noisy(b, k):
        a = bytearray(b)
        for pos in rng.integers(0, len(a), k):
            a[int(pos)] ^= 0x55
        return bytes(a)

    blocks = [
        (b"ab" * 450000)[:899999],                                   # a level-9 block, period 2, n odd
        (b"ba" * 450000)[:899999],                                   # the other direction
        (b"abcdefg" * 130000)[:899999],
        (unit * 900)[:899999],
        b"\x00\x00\x00\x00\x60" + b"\x00\x00\x00\x00\xfb" * 150000 + b"\x00\x00\x00\x00\x17",
        b"xyz" + b"ab" * 300000 + b"the end",                        # head and tail outside the run
        b"ab" * 100000 + b"c" + b"ab" * 100000,                      # two runs of one period around an odd byte
        noisy((unit * 300)[:299999], 3),
        bytes([1, 0]) * 200000 + bytes([0]),
        (b"aab" * 300000)[:899998],
        b"ab" * 20000,                                               # period | n: identical rotations (tie rule)
        corpus.by_name("text", 60000).tobytes() + b"0123456789" * 30000,   # the run covers the second half only
    ]
    want = [O.bwt(b) for b in blocks]
    for val in (0, 8):
        # (the cluster kernel hands such blocks over only when its clusters need more than one wave)
        rep = 1 if val == 0 else 4
        ctx.set("bwt_cluster", val)
        got = ctx.stage_bwt(blocks * rep, 9, with_stats=True)
        for blk, (bw, ptr, has, st), (ebw, eptr, ehas) in zip(blocks * rep, got, want * rep):
            assert ptr == eptr, (val, len(blk), st)
            assert bytes(bw) == bytes(ebw), (val, len(blk), st)
            assert (has == ehas).all()
        assert got[0][3]["period"] == 2 and got[0][3]["rounds"] <= 4, got[0][3]
        assert got[3][3]["period"] == 1000 and got[3][3]["rounds"] <= 4, got[3][3]
        assert got[4][3]["period"] == 5 and got[4][3]["rounds"] <= 5, got[4][3]
    # the same bytes from plain doubling
    ctx.set("bwt_periodic", 0)
    for val in (0, 8):
        ctx.set("bwt_cluster", val)
        got = ctx.stage_bwt(blocks[:5], 9, with_stats=True)
        for (bw, ptr, has, st), (ebw, eptr, ehas) in zip(got, want):
            assert ptr == eptr and bytes(bw) == bytes(ebw)
        assert got[0][3]["period"] == 0 and got[0][3]["rounds"] > 10
    ctx.set("bwt_periodic", 1)
    ctx.set("bwt_cluster", -1)
