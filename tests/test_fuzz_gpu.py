"""GPU fuzzing in the spirit of the reference's fuzz targets (fuzz/fuzz_targets/encode.rs: no crash
at level 1; round_trip.rs: libbz2 decodes to the input) — plus the stronger check that the stream
is byte-identical to the oracle.  Inputs are biased towards what breaks bzip2 encoders: runs around
the 4/255/256 thresholds, tiny alphabets, periodic data, block-capacity boundaries."""
import bz2

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import pyoracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import banzai_b200
    c = banzai_b200.Context(n_gpus=1)
    yield c
    c.close()


run_piece = st.tuples(st.integers(0, 255), st.sampled_from([1, 2, 3, 4, 5, 6, 254, 255, 256, 257, 509, 510, 511, 1000]))
pieces = st.lists(st.one_of(
    st.binary(min_size=0, max_size=300),
    run_piece.map(lambda t: bytes([t[0]]) * t[1]),
    st.tuples(st.binary(min_size=1, max_size=9), st.integers(1, 400)).map(lambda t: t[0] * t[1]),
), min_size=0, max_size=12).map(b"".join)


@settings(max_examples=150, deadline=None, suppress_health_check=list(HealthCheck))
@given(data=pieces, level=st.sampled_from([1, 9]))
def test_small_inputs_match_oracle_and_round_trip(ctx, data, level):
    got = ctx.encode_bytes(data, level)
    assert got == O.encode(data, level)
    assert bz2.decompress(got) == data


@settings(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2 ** 32 - 1), tail=pieces)
def test_block_boundary_inputs(ctx, seed, tail):
    """level 1: ~one block of noise followed by a fuzzed tail, so the cut lands inside the tail"""
    rng = np.random.default_rng(seed)
    head = rng.integers(0, 256, 99999 - int(rng.integers(0, 60))).astype(np.uint8).tobytes()
    data = head + tail + bytes([7]) * int(rng.integers(0, 600)) + tail
    got = ctx.encode_bytes(data, 1)
    assert got == O.encode(data, 1)
