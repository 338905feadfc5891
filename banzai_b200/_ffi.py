"""ctypes binding of libbanzai_b200.so (the C ABI declared in include/banzai_b200.h).

There is no CPU fallback: importing this module fails loudly when the CUDA library has not
been built (run `python -c "import __graft_entry__ as g; g.build()"` or `make -C
banzai_b200/csrc`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BANZAI_B200_LIB") or os.path.join(_HERE, "libbanzai_b200.so")   # override: kernel experiments

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build the CUDA extension first (make -C banzai_b200/csrc). "
        "banzai_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

OK, EINVAL, ECUDA, ENOMEM, EINTERNAL, EIO, EVERIFY = 0, 1, 2, 3, 4, 5, 6


class Stats(C.Structure):
    _fields_ = [
        ("in_bytes", C.c_uint64), ("out_bytes", C.c_uint64),
        ("n_blocks", C.c_uint32), ("n_devices", C.c_uint32),
        ("kernel_launches", C.c_uint32), ("bwt_radix_bits", C.c_uint32),
        ("total_ms", C.c_float), ("h2d_ms", C.c_float), ("d2h_ms", C.c_float),
        ("rle_ms", C.c_float), ("crc_ms", C.c_float), ("bwt_ms", C.c_float),
        ("mtf_ms", C.c_float), ("huff_ms", C.c_float), ("pack_ms", C.c_float),
        ("bwt_n", C.c_uint64), ("bwt_sum_active", C.c_uint64),
        ("bwt_sum_active_passes", C.c_uint64),
        ("bwt_max_rounds", C.c_uint32), ("bwt_tied_blocks", C.c_uint32),
        ("bwt_rounds_total", C.c_uint64), ("bwt_algorithmic_bytes", C.c_uint64),
        ("bwt_cyc_build", C.c_uint64), ("bwt_cyc_radix", C.c_uint64), ("bwt_cyc_rerank", C.c_uint64),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
        ("bwt_sum_tile", C.c_uint64), ("bwt_cyc_tile", C.c_uint64), ("bwt_cyc_final", C.c_uint64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class BwtBlockStats(C.Structure):
    _fields_ = [("n", C.c_uint32), ("rounds", C.c_uint32), ("tied", C.c_uint32), ("period", C.c_uint32),
                ("sum_active", C.c_uint64), ("sum_active_passes", C.c_uint64), ("cycles", C.c_uint64),
                ("sum_tile", C.c_uint64)]


_vp, _sz = C.c_void_p, C.c_size_t
_szp = C.POINTER(C.c_size_t)

lib.bnz_ctx_create.argtypes = [C.POINTER(_vp), C.c_int]
lib.bnz_ctx_create_on.argtypes = [C.POINTER(_vp), C.POINTER(C.c_int), C.c_int]
lib.bnz_ctx_destroy.argtypes = [_vp]
lib.bnz_ctx_destroy.restype = None
lib.bnz_strerror.argtypes = [C.c_int]
lib.bnz_strerror.restype = C.c_char_p
lib.bnz_last_error.argtypes = [_vp]
lib.bnz_last_error.restype = C.c_char_p
lib.bnz_ctx_set.argtypes = [_vp, C.c_char_p, C.c_long]
lib.bnz_encode.argtypes = [_vp, _vp, _sz, C.c_int, C.POINTER(_vp), _szp, _szp]
lib.bnz_free.argtypes = [_vp, _vp]
lib.bnz_free.restype = None
lib.bnz_encode_device.argtypes = [_vp, _vp, _vp, _sz, C.c_int, _vp, _sz, _szp]
lib.bnz_max_compressed_size.argtypes = [_sz]
lib.bnz_max_compressed_size.restype = _sz
lib.bnz_encode_file.argtypes = [_vp, C.c_char_p, C.c_char_p, _szp]
SINK_FN = C.CFUNCTYPE(C.c_int, _vp, _vp, _sz)
lib.bnz_stream_open.argtypes = [_vp, C.c_int, SINK_FN, _vp, C.POINTER(_vp)]
lib.bnz_stream_reserve.argtypes = [_vp, C.POINTER(_vp), _szp]
lib.bnz_stream_commit.argtypes = [_vp, _sz]
lib.bnz_stream_write.argtypes = [_vp, _vp, _sz]
lib.bnz_stream_finish.argtypes = [_vp, _szp]
lib.bnz_stream_close.argtypes = [_vp]
lib.bnz_stream_close.restype = None
lib.bnz_host_alloc.argtypes = [_sz]
lib.bnz_host_alloc.restype = _vp
lib.bnz_host_free.argtypes = [_vp]
lib.bnz_host_free.restype = None
lib.bnz_device_alloc.argtypes = [_vp, _sz]
lib.bnz_device_alloc.restype = _vp
lib.bnz_device_free.argtypes = [_vp, _vp]
lib.bnz_device_free.restype = None
lib.bnz_memcpy_h2d.argtypes = [_vp, _vp, _vp, _sz]
lib.bnz_memcpy_d2h.argtypes = [_vp, _vp, _vp, _sz]
lib.bnz_get_stats.argtypes = [_vp, C.POINTER(Stats)]
lib.bnz_stage_rle1.argtypes = [_vp, _vp, _sz, C.c_int, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _szp]
lib.bnz_host_cut_chain.argtypes = [_vp, _sz, C.c_int, _vp, _vp, _sz, C.c_int, _vp, _vp, _vp, _sz, _szp, _szp]
lib.bnz_stage_bwt.argtypes = [_vp, _vp, _vp, _vp, _sz, C.c_int, _vp, _vp, _vp, _vp]
lib.bnz_stage_mtf.argtypes = [_vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp]
lib.bnz_stage_huffman.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp, _vp, _vp]

EXPORTS = [
    "bnz_ctx_create", "bnz_ctx_create_on", "bnz_ctx_destroy", "bnz_strerror", "bnz_last_error",
    "bnz_ctx_set", "bnz_encode", "bnz_free", "bnz_encode_device", "bnz_max_compressed_size",
    "bnz_encode_file", "bnz_host_alloc", "bnz_host_free", "bnz_device_alloc", "bnz_device_free",
    "bnz_memcpy_h2d", "bnz_memcpy_d2h", "bnz_get_stats", "bnz_stage_rle1", "bnz_stage_bwt",
    "bnz_stage_mtf", "bnz_stage_huffman", "bnz_stream_open", "bnz_stream_reserve", "bnz_stream_commit",
    "bnz_stream_write", "bnz_stream_finish", "bnz_stream_close", "bnz_host_cut_chain",
]
