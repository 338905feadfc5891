"""banzai_b200 — B200-native bzip2 encoder core behind banzai's `encode` API.

Host-side mirror of the reference's public surface (jgbyrne/banzai lib/lib.rs):
    encode(reader, writer, level) -> int      lib/lib.rs:84
    encode_file(in_path, out_path) -> int     lib/lib.rs:141
plus `Context` (device resources) and stage-level functions used by the parity tests.
All compute runs in hand-written sm_100a CUDA kernels through the C ABI in
include/banzai_b200.h; there is no CPU fallback.
"""
from .api import (BanzaiError, Context, encode, encode_bytes, encode_file, stage_bwt,  # noqa: F401
                  stage_huffman, stage_mtf, stage_rle1)

__all__ = ["BanzaiError", "Context", "encode", "encode_bytes", "encode_file", "stage_rle1",
           "stage_bwt", "stage_mtf", "stage_huffman"]
