"""Payload codecs for ``server.message_compress`` (reference: pico-core ``Compress`` -- snappy / lz4 / zlib behind
``RpcView`` and the file layer, pico-core/include/pico-core/Compress.h).

* ``zlib``  -- deflate level 1 (python's zlib), HTTP ``Content-Encoding: deflate``
* ``lz4``   -- the LZ4 block format, native codec in ``csrc/core/exb_core.cpp`` (``exb_lz4_*``), framed with the
  8-byte little-endian uncompressed size; ``Content-Encoding: lz4``
* ``snappy`` is accepted as a configuration value and served by the lz4 codec (same speed class; there is no
  snappy library in the image and the wire format is private to this framework's own nodes and clients)
"""
import ctypes
import struct
import zlib

from .. import _native

ENCODINGS = {"zlib": "deflate", "lz4": "lz4", "snappy": "lz4"}


def encoding_of(method):
    """HTTP Content-Encoding token of a ``message_compress`` value ("" -> None)"""
    return ENCODINGS.get(method or "")


def compress(data, encoding):
    if encoding == "deflate":
        return zlib.compress(data, 1)
    if encoding == "lz4":
        lib = _native.core()
        n = len(data)
        cap = lib.exb_lz4_bound(n)
        out = ctypes.create_string_buffer(cap)
        m = lib.exb_lz4_compress(bytes(data), n, out, cap)
        if m < 0:
            raise ValueError("lz4: compression failed")
        return struct.pack("<Q", n) + out.raw[:m]
    raise ValueError("unknown content encoding %r" % (encoding,))


def decompress(data, encoding):
    if encoding == "deflate":
        return zlib.decompress(data)
    if encoding == "lz4":
        if len(data) < 8:
            raise ValueError("lz4: truncated frame")
        (n,) = struct.unpack("<Q", data[:8])
        if n > (1 << 40):
            raise ValueError("lz4: implausible frame size")
        lib = _native.core()
        out = ctypes.create_string_buffer(max(1, n))
        body = bytes(data[8:])
        m = lib.exb_lz4_decompress(body, len(body), out, n)
        if m != n:
            raise ValueError("lz4: corrupt frame")
        return out.raw[:n]
    raise ValueError("unknown content encoding %r" % (encoding,))
