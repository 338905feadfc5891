"""GPU parity: K1/K2 — block cuts, RLE1 bytes and block CRCs vs the oracle's literal
`rle::rle_one` loop driven as `encode` drives it (reference lib/rle.rs:102, lib/lib.rs:101-126)."""
import numpy as np
import pytest

import corpus
from oracle import pyoracle as O
from tests.golden.make_vectors import vector_input

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import banzai_b200
    c = banzai_b200.Context(n_gpus=1)
    yield c
    c.close()


def _oracle_blocks(data, level):
    a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    off, res = 0, []
    while off < a.size:
        out, cons, crc = O.rle_one(a[off:], level)
        assert cons > 0
        res.append({"in_off": off, "consumed": cons, "rle": out, "crc": crc})
        off += cons
    return res


def _check(ctx, data, level):
    got = ctx.stage_rle1(data, level)
    want = _oracle_blocks(data, level)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g["in_off"] == w["in_off"] and g["consumed"] == w["consumed"]
        assert g["crc"] == w["crc"]
        assert bytes(g["rle"]) == bytes(w["rle"])


@pytest.mark.parametrize("name", ["V2", "V3", "V4", "V5", "V6", "V7", "V8", "V9", "V10", "V11"])
def test_survey_vector_inputs(ctx, name):
    level = {"V2": 9, "V6": 9, "V9": 2}.get(name, 1)
    _check(ctx, vector_input(name), level)


def test_empty_input(ctx):
    assert ctx.stage_rle1(b"", 9) == []


@pytest.mark.parametrize("kind", ["text", "source", "binary", "mixed", "random"])
@pytest.mark.parametrize("level", [1, 9])
def test_corpora(ctx, kind, level):
    _check(ctx, corpus.by_name(kind, 2500000), level)


def test_capacity_rule_at_run_starts(ctx):
    """SURVEY A-Q1: B = 0..6 bytes of capacity left when a long run starts"""
    cap = 99999
    for B in range(0, 8):
        head = (np.arange(cap - B) % 251 + 1).astype(np.uint8)
        for runlen in (4, 5, 10, 254, 255, 256, 600):
            data = np.concatenate([head, np.zeros(runlen, np.uint8), (np.arange(5000) % 200 + 3).astype(np.uint8)])
            _check(ctx, data, 1)


def test_run_dominated_streams(ctx):
    rng = np.random.default_rng(21)
    for it in range(6):
        vals = rng.integers(0, 4, 60000)
        lens = rng.integers(1, 12, 60000) * rng.choice([1, 1, 30, 100, 700], 60000)
        _check(ctx, np.repeat(vals, lens).astype(np.uint8), 1)


def test_degenerate(ctx):
    _check(ctx, bytes(70 * 1000 * 1000), 9)         # zeros: one run across many chunks and blocks
    _check(ctx, b"ab" * 1500000, 9)
    _check(ctx, b"aaaab" * 300000, 1)
    _check(ctx, bytes([7]) * 254 + bytes([8]) * 255 + bytes([9]) * 256 + bytes([1]) * 1023 + bytes([2]) * 1025, 1)
