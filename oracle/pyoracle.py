"""ctypes binding of the CPU oracle (oracle/banzai_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(banzai_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("banzai_oracle.c", "sais_generic.inc")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class BlockInfo(C.Structure):
    _fields_ = [("in_off", C.c_uint64), ("consumed", C.c_uint64), ("rle_len", C.c_uint64),
                ("mtf_len", C.c_uint64), ("bit_off", C.c_uint64), ("bit_len", C.c_uint64),
                ("crc", C.c_uint32), ("ptr", C.c_uint32), ("num_syms", C.c_uint32),
                ("num_tables", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, u16p, u32p, u64p, szp = (C.POINTER(C.c_uint8), C.POINTER(C.c_uint16),
                                      C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_size_t))
        L.orc_crc32.restype = C.c_uint32
        L.orc_crc32.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_rle_one.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, szp, szp, u32p]
        L.orc_rle_canonical.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, szp, szp]
        L.orc_bwt.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, u32p, C.c_void_p]
        L.orc_bwt_naive.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, u32p, C.c_void_p]
        L.orc_mtf_and_rle.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, szp, szp,
                                      C.c_void_p]
        L.orc_build_table.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p]
        L.orc_huffman_model.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, szp,
                                        C.c_void_p, C.c_void_p, szp]
        L.orc_huffman_encode.restype = C.c_size_t
        L.orc_huffman_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                         C.c_void_p, C.c_size_t]
        L.orc_encode_ex.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p), szp,
                                    szp, C.c_void_p, C.c_size_t, szp]
        L.orc_encode_mt.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                    szp, szp]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_bw_new.restype = C.c_void_p
        L.orc_bw_write_bits.argtypes = [C.c_void_p, C.c_uint32, C.c_size_t]
        L.orc_bw_write_bits_u32.argtypes = [C.c_void_p, C.c_uint32, C.c_size_t]
        L.orc_bw_write_byte.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_bw_write_bytes.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_bw_close.restype = C.c_size_t
        L.orc_bw_close.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _lib = L
    return _lib


def _as_u8(data):
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def crc32(data):
    a = _as_u8(data)
    return int(lib().orc_crc32(_ptr(a), a.size))


def _rle(fn, data, level, with_crc):
    a = _as_u8(data)
    out = np.empty(100000 * level, dtype=np.uint8)
    olen, cons, crc = C.c_size_t(), C.c_size_t(), C.c_uint32()
    if with_crc:
        rc = fn(_ptr(a), a.size, level, _ptr(out), C.byref(olen), C.byref(cons), C.byref(crc))
    else:
        rc = fn(_ptr(a), a.size, level, _ptr(out), C.byref(olen), C.byref(cons))
    assert rc == 0
    return out[:olen.value].copy(), cons.value, crc.value


def rle_one(data, level):
    """(rle1 bytes, consumed, block crc) — reference lib/rle.rs:102"""
    return _rle(lib().orc_rle_one, data, level, True)


def rle_canonical(data, level):
    out, cons, _ = _rle(lib().orc_rle_canonical, data, level, False)
    return out, cons


def _bwt(fn, data):
    a = _as_u8(data)
    out = np.empty(max(a.size, 1), dtype=np.uint8)
    ptr = C.c_uint32()
    has = np.zeros(256, dtype=np.uint8)
    rc = fn(_ptr(a), a.size, _ptr(out), C.byref(ptr), _ptr(has))
    if rc != 0:
        return np.empty(0, dtype=np.uint8), None, has
    return out[:a.size].copy(), ptr.value, has


def bwt(data):
    """(bwt bytes, ptr, has_byte[256]) — reference lib/bwt.rs:526"""
    return _bwt(lib().orc_bwt, data)


def bwt_naive(data):
    return _bwt(lib().orc_bwt_naive, data)


def mtf_and_rle(bwt_bytes, has_byte):
    """(symbols u16[m], num_syms, freqs[258]) — reference lib/mtf.rs:14"""
    a = _as_u8(bwt_bytes)
    has = np.ascontiguousarray(has_byte, dtype=np.uint8)
    out = np.empty(a.size + 1, dtype=np.uint16)
    m, ns = C.c_size_t(), C.c_size_t()
    freqs = np.zeros(258, dtype=np.uint64)
    rc = lib().orc_mtf_and_rle(_ptr(a), a.size, _ptr(has), _ptr(out), C.byref(m), C.byref(ns),
                               _ptr(freqs))
    assert rc == 0
    return out[:m.value].copy(), ns.value, freqs


def build_table(freqs, num_syms=None):
    """code lengths — reference lib/huffman.rs:271"""
    f = np.ascontiguousarray(freqs, dtype=np.uint64)
    n = f.size if num_syms is None else num_syms
    out = np.zeros(n, dtype=np.uint8)
    rc = lib().orc_build_table(n, _ptr(f), _ptr(out))
    assert rc == 0
    return out


def huffman_model(syms, num_syms, freqs):
    """(num_tables, tables[num_tables, num_syms], selectors) — lib/huffman.rs:313-460"""
    s = np.ascontiguousarray(syms, dtype=np.uint16)
    f = np.ascontiguousarray(freqs, dtype=np.uint64)
    tables = np.zeros((6, 258), dtype=np.uint8)
    sel = np.zeros(s.size // 50 + 2, dtype=np.uint8)
    nt, ns = C.c_size_t(), C.c_size_t()
    rc = lib().orc_huffman_model(_ptr(s), s.size, num_syms, _ptr(f), C.byref(nt), _ptr(tables),
                                 _ptr(sel), C.byref(ns))
    assert rc == 0
    return nt.value, tables[:nt.value, :num_syms].copy(), sel[:ns.value].copy()


def huffman_encode(syms, num_syms, freqs):
    """(bytes zero-padded, bit count) of huffman::encode on a fresh writer"""
    s = np.ascontiguousarray(syms, dtype=np.uint16)
    f = np.ascontiguousarray(freqs, dtype=np.uint64)
    cap = s.size * 3 + 4096
    out = np.zeros(cap, dtype=np.uint8)
    bits = lib().orc_huffman_encode(_ptr(s), s.size, num_syms, _ptr(f), _ptr(out), cap)
    return out[:(bits + 7) // 8].copy(), int(bits)


def encode(data, level, with_info=False):
    """banzai::encode restated — reference lib/lib.rs:84. Returns bytes (and block infos)."""
    a = _as_u8(data)
    out = C.c_void_p()
    olen, cons, nb = C.c_size_t(), C.c_size_t(), C.c_size_t()
    cap = a.size // (79999 * level) + 2 if with_info else 0
    infos = (BlockInfo * cap)() if with_info else None
    rc = lib().orc_encode_ex(_ptr(a), a.size, level, C.byref(out), C.byref(olen), C.byref(cons),
                             infos, cap, C.byref(nb))
    if rc != 0:
        raise ValueError("level out of range")
    res = C.string_at(out.value, olen.value)
    lib().orc_free(out)
    assert cons.value == a.size
    if with_info:
        assert nb.value <= cap
        return res, [infos[i] for i in range(nb.value)]
    return res


def encode_mt(data, level, threads=0, digest=False):
    """The same restatement with one worker thread per block (orc_encode_mt): sequential cut
    chain, blocks encoded in parallel, bit strings appended in order.  Byte-identical to
    encode(); affordable at the benchmark's full sizes.  threads=0: all host cores.
    digest=True returns (sha256 hex, length, n_blocks) without copying the stream into Python."""
    import hashlib
    a = _as_u8(data)
    if threads <= 0:
        threads = os.cpu_count() or 1
    out = C.c_void_p()
    olen, nb = C.c_size_t(), C.c_size_t()
    rc = lib().orc_encode_mt(_ptr(a), a.size, level, threads, C.byref(out), C.byref(olen), C.byref(nb))
    if rc != 0:
        raise ValueError("level out of range")
    try:
        if digest:
            view = (C.c_uint8 * olen.value).from_address(out.value)
            return hashlib.sha256(view).hexdigest(), olen.value, nb.value
        return C.string_at(out.value, olen.value)
    finally:
        lib().orc_free(out)


class BitWriter:
    """reference lib/out.rs OutputStream"""

    def __init__(self):
        self.h = lib().orc_bw_new()

    def write_bits(self, chunk, n):
        lib().orc_bw_write_bits(self.h, chunk, n)

    def write_bits_u32(self, chunk, n):
        lib().orc_bw_write_bits_u32(self.h, chunk, n)

    def write_byte(self, b):
        lib().orc_bw_write_byte(self.h, b)

    def write_bytes(self, bs):
        a = _as_u8(bs)
        lib().orc_bw_write_bytes(self.h, _ptr(a), a.size)

    def close(self):
        out = np.zeros(1 << 20, dtype=np.uint8)
        n = lib().orc_bw_close(self.h, _ptr(out), out.size)
        self.h = None
        return bytes(out[:n])
