"""Hash functions and the key partitioner of the generic parameter server.

Reference: pico-core ``MurmurHash3`` / ``HashFunction`` and pico-ps ``Partitioner``
(murmur + jump-consistent-hash; pico-ps/pico-ps/common/Partitioner.h). OpenEmbedding's embedding
tables do NOT use it (they route ``id % shard_num``, EmbeddingPullOperator.cpp:74-76) and its serving
placement is a rotating cursor (Model.cpp:153-186 -> ``serving/controller.py``); the functions are the
generic-PS utilities of the inventory, used by ``ServingClient(policy="hash")`` to keep a key range on
one replica (cache affinity) without reshuffling when replicas are added.
"""

_M64 = (1 << 64) - 1


def murmur3_fmix64(k):
    """64-bit finalizer of MurmurHash3 (bijective mix of an integer key)."""
    k &= _M64
    k ^= k >> 33
    k = (k * 0xFF51AFD7ED558CCD) & _M64
    k ^= k >> 33
    k = (k * 0xC4CEB9FE1A85EC53) & _M64
    k ^= k >> 33
    return k


def murmur3_32(data, seed=0):
    """MurmurHash3_x86_32 of a bytes object."""
    c1, c2 = 0xCC9E2D51, 0x1B873593
    h = seed & 0xFFFFFFFF
    n = len(data)
    for i in range(0, n - n % 4, 4):
        k = int.from_bytes(data[i:i + 4], "little")
        k = (k * c1) & 0xFFFFFFFF
        k = ((k << 15) | (k >> 17)) & 0xFFFFFFFF
        k = (k * c2) & 0xFFFFFFFF
        h ^= k
        h = ((h << 13) | (h >> 19)) & 0xFFFFFFFF
        h = (h * 5 + 0xE6546B64) & 0xFFFFFFFF
    tail = data[n - n % 4:]
    if tail:
        k = int.from_bytes(tail, "little")
        k = (k * c1) & 0xFFFFFFFF
        k = ((k << 15) | (k >> 17)) & 0xFFFFFFFF
        k = (k * c2) & 0xFFFFFFFF
        h ^= k
    h ^= n
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & 0xFFFFFFFF
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & 0xFFFFFFFF
    h ^= h >> 16
    return h


def jump_consistent_hash(key, num_buckets):
    """Lamping & Veach: bucket in [0, num_buckets); growing the bucket count moves 1/n of the keys."""
    if num_buckets <= 0:
        raise ValueError("num_buckets must be positive")
    key &= _M64
    b, j = -1, 0
    while j < num_buckets:
        b = j
        key = (key * 2862933555777941757 + 1) & _M64
        j = int((b + 1) * (float(1 << 31) / float((key >> 33) + 1)))
    return b


class Partitioner:
    """key -> partition: murmur mix, then jump-consistent-hash."""

    def __init__(self, num_partitions):
        self.n = int(num_partitions)

    def __call__(self, key):
        if isinstance(key, (bytes, str)):
            key = murmur3_32(key.encode() if isinstance(key, str) else key)
        return jump_consistent_hash(murmur3_fmix64(int(key)), self.n)
