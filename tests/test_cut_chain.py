"""CPU: the host half of the RLE1 stage — the sequential block-cut chain (`encode` calling
`rle_one` block after block, reference lib/lib.rs:101-126; capacity rule lib/rle.rs:121-240,
SURVEY A-Q1) — against the oracle's block trace.  The chunk tables the GPU kernels normally
produce are computed here with numpy, so no device is involved: bnz_host_cut_chain is a plain
host function of the C-ABI library."""
import ctypes as C

import numpy as np
import pytest

import corpus
from oracle import pyoracle as O

CH = 1024


def chunk_tables(data):
    """o_in[c], P[c] as include/banzai_b200.h defines them"""
    a = np.frombuffer(data, dtype=np.uint8)
    n = a.size
    head = np.ones(n, dtype=bool)
    head[1:] = a[1:] != a[:-1]
    idx = np.arange(n, dtype=np.int64)
    last_head = np.maximum.accumulate(np.where(head, idx, 0))
    o = idx - last_head                                   # run offset of every byte
    r = o % 255
    need = np.where(r < 3, 1, np.where(r == 3, 2, 0)).astype(np.uint64)
    cum = np.concatenate([[0], np.cumsum(need, dtype=np.uint64)]).astype(np.uint64)
    n_chunks = (n + CH - 1) // CH
    starts = np.arange(n_chunks, dtype=np.int64) * CH
    P = np.concatenate([cum[starts], cum[-1:]]).astype(np.uint64)
    o_in = o[starts].astype(np.uint64)
    return np.ascontiguousarray(P), np.ascontiguousarray(o_in), n_chunks


def cut_chain(data, level, final=True):
    from banzai_b200 import _ffi
    P, o_in, n_chunks = chunk_tables(data)
    cap = len(data) // (100000 * level // 300 + 1) + 16
    off = np.zeros(cap, dtype=np.uint64)
    ln = np.zeros(cap, dtype=np.uint64)
    rl = np.zeros(cap, dtype=np.uint32)
    nb, used = C.c_size_t(), C.c_size_t()
    buf = np.frombuffer(data, dtype=np.uint8)
    rc = _ffi.lib.bnz_host_cut_chain(buf.ctypes.data, len(data), level, P.ctypes.data, o_in.ctypes.data, n_chunks,
                                     1 if final else 0, off.ctypes.data, ln.ctypes.data, rl.ctypes.data, cap,
                                     C.byref(nb), C.byref(used))
    assert rc == _ffi.OK
    k = nb.value
    return [(int(off[i]), int(ln[i]), int(rl[i])) for i in range(k)], used.value


def oracle_cuts(data, level):
    _, infos = O.encode(data, level, with_info=True)
    return [(int(b.in_off), int(b.consumed), int(b.rle_len)) for b in infos]


def _cases():
    rng = np.random.default_rng(7)
    yield "mixed", corpus.mixed(1_500_000).tobytes(), 1
    yield "text-L3", corpus.text(1_200_000).tobytes(), 3
    yield "zeros", bytes(3_000_000), 1                      # one block swallows 51x its capacity
    yield "aaaab", b"aaaab" * 60_000, 1                     # RLE1 expands 5 -> 6, cuts inside runs
    yield "runs-255", (b"x" * 255 + b"y" * 256 + b"z" * 4 + b"w" * 3) * 900, 1
    noise = rng.integers(0, 256, 99_990, dtype=np.uint8).tobytes()
    for tail in (b"a" * 3, b"a" * 4, b"a" * 5, b"a" * 9, b"ab" * 8, b"a" * 300):
        # the capacity runs out inside / right before a run: B = 1..5 bytes left (rle.rs:177-208)
        for pad in range(0, 8):
            yield f"edge-{len(tail)}-{pad}", noise + b"q" * pad + tail + noise[:5000], 1


@pytest.mark.parametrize("name,data,level", list(_cases()), ids=[c[0] for c in _cases()])
def test_cut_chain_matches_oracle_blocks(name, data, level):
    got, used = cut_chain(data, level)
    assert got == oracle_cuts(data, level)
    assert used == len(data)


def test_non_final_window_leaves_the_open_block():
    """streaming windows: with more input to come, the block that only ends with the data is not
    cut; the reported blocks are a prefix of the whole input's blocks"""
    data = corpus.mixed(1_000_000).tobytes()
    whole = oracle_cuts(data, 1)
    for cut_at in (250_000, 333_333, 999_999):
        got, used = cut_chain(data[:cut_at], 1, final=False)
        assert got == whole[:len(got)]
        assert used == sum(b[1] for b in got)
        assert len(got) in (len([b for b in whole if b[0] + b[1] < cut_at]),
                            len([b for b in whole if b[0] + b[1] <= cut_at]))


def test_argument_checks():
    from banzai_b200 import _ffi
    nb = C.c_size_t()
    assert _ffi.lib.bnz_host_cut_chain(None, 0, 9, None, None, 0, 1, None, None, None, 0, C.byref(nb), None) == _ffi.OK
    assert nb.value == 0
    assert _ffi.lib.bnz_host_cut_chain(None, 0, 10, None, None, 0, 1, None, None, None, 0, C.byref(nb), None) == _ffi.EINVAL
