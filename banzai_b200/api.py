"""Python host side over the C ABI: mirrors banzai's `encode` / `encode_file` and exposes the
stage seams (`rle::rle_one`, `bwt::bwt`, `mtf::mtf_and_rle`, `huffman::encode`) for parity tests."""
import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import lib


class BanzaiError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        msg = lib.bnz_strerror(code).decode()
        super().__init__(f"{msg}: {detail}" if detail else msg)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(data):
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data) if not isinstance(data, (bytes, bytearray, memoryview)) else data,
                         dtype=np.uint8)


class Context:
    """Owns device streams/arenas (include/banzai_b200.h: bnz_ctx). Not thread-safe."""

    def __init__(self, n_gpus=1, devices=None):
        h = C.c_void_p()
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            rc = lib.bnz_ctx_create_on(C.byref(h), arr, len(devices))
        else:
            rc = lib.bnz_ctx_create(C.byref(h), n_gpus)
        if rc != _ffi.OK:
            raise BanzaiError(rc, "bnz_ctx_create (is a B200 visible?)")
        self._h = h

    def close(self):
        if self._h:
            lib.bnz_ctx_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != _ffi.OK:
            raise BanzaiError(rc, lib.bnz_last_error(self._h).decode())

    def set(self, key, value):
        self._check(lib.bnz_ctx_set(self._h, key.encode(), int(value)))

    def stats(self):
        s = _ffi.Stats()
        self._check(lib.bnz_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    # ---- the hot path -------------------------------------------------------------
    def encode_bytes(self, data, level=9):
        """whole-buffer `banzai::encode` (lib/lib.rs:84): returns the .bz2 stream as bytes"""
        if not (1 <= level <= 9):
            # the reference asserts (lib/lib.rs:89); surface it as the ABI's EINVAL
            raise BanzaiError(_ffi.EINVAL, f"level {level}")
        a = _u8(data)
        out = C.c_void_p()
        olen, cons = C.c_size_t(), C.c_size_t()
        self._check(lib.bnz_encode(self._h, _ptr(a), a.size, level, C.byref(out), C.byref(olen),
                                   C.byref(cons)))
        try:
            if cons.value != a.size:
                raise BanzaiError(_ffi.EINTERNAL, "short encode")
            # (ctypes.string_at takes a C int length: it would truncate streams >= 2 GiB)
            return bytes((C.c_ubyte * olen.value).from_address(out.value))
        finally:
            lib.bnz_free(self._h, out)

    def encode_ptr(self, host_ptr, n, level):
        """encode n bytes at a raw host pointer (e.g. pinned); returns (out_ptr, out_len) —
        caller releases with free_out()."""
        out = C.c_void_p()
        olen, cons = C.c_size_t(), C.c_size_t()
        self._check(lib.bnz_encode(self._h, host_ptr, n, level, C.byref(out), C.byref(olen),
                                   C.byref(cons)))
        return out, olen.value

    def free_out(self, out):
        lib.bnz_free(self._h, out)

    def encode_device(self, d_in, h_in, n, level, d_out, d_out_cap):
        olen = C.c_size_t()
        self._check(lib.bnz_encode_device(self._h, d_in, h_in, n, level, d_out, d_out_cap,
                                          C.byref(olen)))
        return olen.value

    def encode_stream(self, reader, writer, level, read_size=8 << 20):
        """streaming `banzai::encode(reader, writer, level)` over bnz_stream_* : the reader fills
        pinned windows in place (readinto when it has it) while the previous window is on the
        GPU; stream bytes reach `writer.write` in order, on this thread.  Returns bytes read."""
        if not (1 <= level <= 9):
            raise BanzaiError(_ffi.EINVAL, f"level {level}")
        sink_exc = []

        def _sink(_user, data, n):
            try:
                writer.write(bytes((C.c_ubyte * n).from_address(data)))
                return 0
            except Exception as e:          # surfaces as BNZ_EIO; the Python exception is re-raised
                sink_exc.append(e)
                return 1

        cb = _ffi.SINK_FN(_sink)
        h = C.c_void_p()
        self._check(lib.bnz_stream_open(self._h, level, cb, None, C.byref(h)))
        try:
            buf, cap, used = C.c_void_p(), C.c_size_t(), C.c_size_t()
            readinto = getattr(reader, "readinto", None)
            try:
                while True:
                    self._check(lib.bnz_stream_reserve(h, C.byref(buf), C.byref(cap)))
                    want = min(cap.value, read_size)
                    if readinto is not None:
                        got = readinto(memoryview((C.c_ubyte * want).from_address(buf.value)).cast("B"))
                        got = 0 if got is None else got
                    else:
                        chunk = reader.read(want)
                        got = len(chunk)
                        if got:
                            C.memmove(buf, chunk, got)
                    if got == 0:
                        break
                    self._check(lib.bnz_stream_commit(h, got))
                self._check(lib.bnz_stream_finish(h, C.byref(used)))
            except BanzaiError:
                if sink_exc:
                    raise sink_exc[0]
                raise
            if hasattr(writer, "flush"):
                writer.flush()
            return used.value
        finally:
            lib.bnz_stream_close(h)

    # ---- stage seams ----------------------------------------------------------------
    def stage_bwt(self, blocks, level=9, with_stats=False):
        """`bwt::bwt` (lib/bwt.rs:526) on a list of byte blocks -> [(bwt, ptr, has_byte)]"""
        arrs = [_u8(b) for b in blocks]
        nb = len(arrs)
        lens = np.array([a.size for a in arrs], dtype=np.uint32)
        offs = np.zeros(nb, dtype=np.uint64)
        if nb:
            offs[1:] = np.cumsum(lens.astype(np.uint64))[:-1]
        cat = np.concatenate(arrs) if nb else np.zeros(0, np.uint8)
        out = np.zeros(max(cat.size, 1), dtype=np.uint8)
        ptr = np.zeros(max(nb, 1), dtype=np.uint32)
        has = np.zeros((max(nb, 1), 256), dtype=np.uint8)
        st = (_ffi.BwtBlockStats * max(nb, 1))()
        self._check(lib.bnz_stage_bwt(self._h, _ptr(cat), _ptr(offs), _ptr(lens), nb, level,
                                      _ptr(out), _ptr(ptr), _ptr(has), st))
        res = []
        for b in range(nb):
            o = int(offs[b])
            item = (out[o:o + int(lens[b])].copy(), int(ptr[b]), has[b].copy())
            if with_stats:
                item = item + ({"rounds": st[b].rounds, "tied": st[b].tied,
                                "sum_active": st[b].sum_active,
                                "sum_active_passes": st[b].sum_active_passes,
                                "cycles": st[b].cycles, "period": st[b].period},)
            res.append(item)
        return res

    def stage_rle1(self, data, level=9):
        """`rle::rle_one` driven as `encode` does -> list of dicts per block"""
        a = _u8(data)
        max_blocks = a.size // (79999 * level) + 2
        in_off = np.zeros(max_blocks, np.uint64)
        in_len = np.zeros(max_blocks, np.uint64)
        rle_off = np.zeros(max_blocks, np.uint64)
        rle_len = np.zeros(max_blocks, np.uint32)
        crc = np.zeros(max_blocks, np.uint32)
        cap = a.size + a.size // 4 + 16
        rle = np.zeros(cap, np.uint8)
        nb = C.c_size_t()
        self._check(lib.bnz_stage_rle1(self._h, _ptr(a), a.size, level, _ptr(in_off), _ptr(in_len),
                                       _ptr(rle_off), _ptr(rle_len), _ptr(crc), max_blocks,
                                       _ptr(rle), cap, C.byref(nb)))
        res = []
        for b in range(nb.value):
            o = int(rle_off[b])
            res.append({"in_off": int(in_off[b]), "consumed": int(in_len[b]),
                        "rle": rle[o:o + int(rle_len[b])].copy(), "crc": int(crc[b])})
        return res

    def stage_mtf(self, bwts, has_bytes):
        """`mtf::mtf_and_rle` (lib/mtf.rs:14) on a list of BWT blocks -> [(syms, num_syms, freqs)]"""
        arrs = [_u8(b) for b in bwts]
        nb = len(arrs)
        lens = np.array([a.size for a in arrs], dtype=np.uint32)
        offs = np.zeros(nb, dtype=np.uint64)
        if nb:
            offs[1:] = np.cumsum(lens.astype(np.uint64))[:-1]
        cat = np.concatenate(arrs)
        has = np.ascontiguousarray(np.stack([np.asarray(h, dtype=np.uint8) for h in has_bytes]))
        syms = np.zeros(cat.size + nb, dtype=np.uint16)
        sym_len = np.zeros(nb, np.uint32)
        num_syms = np.zeros(nb, np.uint32)
        freqs = np.zeros((nb, 258), np.uint32)
        self._check(lib.bnz_stage_mtf(self._h, _ptr(cat), _ptr(offs), _ptr(lens), _ptr(has), nb,
                                      _ptr(syms), _ptr(sym_len), _ptr(num_syms), _ptr(freqs)))
        res = []
        for b in range(nb):
            o = int(offs[b]) + b
            res.append((syms[o:o + int(sym_len[b])].copy(), int(num_syms[b]), freqs[b].copy()))
        return res

    def stage_huffman(self, sym_blocks, num_syms_list, freqs_list):
        """`huffman::encode` (lib/huffman.rs:313) -> [(bytes, bit_len, tables, num_tables)]"""
        arrs = [np.ascontiguousarray(s, dtype=np.uint16) for s in sym_blocks]
        nb = len(arrs)
        lens = np.array([a.size for a in arrs], dtype=np.uint32)
        offs = np.zeros(nb, dtype=np.uint64)
        if nb:
            offs[1:] = np.cumsum(lens.astype(np.uint64))[:-1]
        cat = np.concatenate(arrs)
        ns = np.array(num_syms_list, dtype=np.uint32)
        fr = np.ascontiguousarray(np.stack([np.asarray(f, dtype=np.uint32)[:258] for f in freqs_list]))
        stride = int(lens.max()) * 3 + 8192
        stride = (stride + 15) & ~15
        bits = np.zeros(nb * stride, np.uint8)
        bit_len = np.zeros(nb, np.uint64)
        tables = np.zeros((nb, 6, 258), np.uint8)
        nt = np.zeros(nb, np.uint32)
        self._check(lib.bnz_stage_huffman(self._h, _ptr(cat), _ptr(offs), _ptr(lens), _ptr(ns),
                                          _ptr(fr), nb, _ptr(bits), stride, _ptr(bit_len),
                                          _ptr(tables), _ptr(nt)))
        res = []
        for b in range(nb):
            nbytes = (int(bit_len[b]) + 7) // 8
            res.append((bits[b * stride:b * stride + nbytes].copy(), int(bit_len[b]),
                        tables[b, :int(nt[b]), :int(ns[b])].copy(), int(nt[b])))
        return res


_default_ctx = None


def _ctx():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(n_gpus=1)
    return _default_ctx


def encode_bytes(data, level=9):
    return _ctx().encode_bytes(data, level)


def encode(reader, writer, level):
    """banzai::encode(reader, BufWriter, level) -> usize  (reference lib/lib.rs:84-132).

    `reader` is any object with .readinto() or .read(n); `writer` any object with .write() (and
    optionally .flush()).  The input is streamed window by window (include/banzai_b200.h,
    bnz_stream_*), never held whole.  Like the reference it returns the number of input bytes encoded, and it
    rejects level outside 1..=9 (the reference asserts at lib/lib.rs:89)."""
    if not (1 <= level <= 9):
        raise BanzaiError(_ffi.EINVAL, f"level {level}")
    return _ctx().encode_stream(reader, writer, level)


def encode_file(in_path, out_path):
    """banzai::encode_file(in_path, out_path) -> usize at level 9 (reference lib/lib.rs:141-153)"""
    with open(in_path, "rb") as inf, open(out_path, "wb") as outf:
        return encode(inf, outf, 9)


def stage_rle1(data, level=9):
    return _ctx().stage_rle1(data, level)


def stage_bwt(blocks, level=9, with_stats=False):
    return _ctx().stage_bwt(blocks, level, with_stats)


def stage_mtf(bwts, has_bytes):
    return _ctx().stage_mtf(bwts, has_bytes)


def stage_huffman(sym_blocks, num_syms_list, freqs_list):
    return _ctx().stage_huffman(sym_blocks, num_syms_list, freqs_list)
