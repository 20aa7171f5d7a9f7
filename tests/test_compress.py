"""utils/compress.py: the native LZ4 block codec (csrc/core/exb_core.cpp) and the zlib codec behind
server.message_compress. Reference: pico-core Compress (snappy / lz4 / zlib)."""
import os
import struct

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from openembedding_b200.utils import compress as C


def _payloads():
    rng = np.random.default_rng(0)
    yield b""
    yield b"a"
    yield b"abcabcabcabcabcabcabcabcabcabcabcabc" * 50
    yield bytes(100000)                                            # one long run: length bytes chain past 255
    yield rng.integers(0, 256, 70000, dtype=np.uint8).tobytes()   # incompressible
    ids = (rng.integers(0, 1 << 20, 20000).astype(np.uint64) * 7919)
    yield ids.tobytes()                                            # what a pull request carries
    rows = np.zeros((4096, 16), dtype=np.float32); rows[::7] = rng.random((586, 16), dtype=np.float32)
    yield rows.tobytes()                                           # sparse rows: zeros initializer + a few trained


@pytest.mark.parametrize("enc", ["lz4", "deflate"])
def test_roundtrip(enc):
    for p in _payloads():
        z = C.compress(p, enc)
        assert C.decompress(z, enc) == p
    big = bytes(100000)
    assert len(C.compress(big, enc)) < 2000


def test_lz4_is_the_block_format():
    """a hand-assembled LZ4 block decodes: literals 'abcd', then a match of 8 at offset 4, then literals 'xyz12'"""
    block = bytes([0x44]) + b"abcd" + struct.pack("<H", 4) + bytes([0x50]) + b"xyz12"
    frame = struct.pack("<Q", 4 + 8 + 5) + block
    assert C.decompress(frame, "lz4") == b"abcd" + b"abcdabcd" + b"xyz12"


@settings(max_examples=200, deadline=None)
@given(st.binary(max_size=3000), st.integers(0, 40))
def test_lz4_fuzz_roundtrip_and_corruption(data, rep):
    payload = data * (1 + rep)
    z = C.compress(payload, "lz4")
    assert C.decompress(z, "lz4") == payload
    if len(z) > 9:                       # flip a byte of the block: an error or different bytes, never a crash
        bad = bytearray(z); bad[8 + (len(z) - 9) // 2] ^= 0x5A
        try:
            out = C.decompress(bytes(bad), "lz4")
            assert len(out) == len(payload)
        except ValueError:
            pass
    with pytest.raises(ValueError):
        C.decompress(z[:7], "lz4")


def test_config_values_map_to_encodings():
    assert C.encoding_of("") is None and C.encoding_of("zlib") == "deflate"
    assert C.encoding_of("lz4") == "lz4" and C.encoding_of("snappy") == "lz4"
