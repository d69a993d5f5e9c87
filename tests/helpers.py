"""Pure-python restatements of the product's counter-based streams (Philox4x32-10, ring offsets) as
documented in include/ols_b200.h / DESIGN.md — shared by CPU and GPU tests."""

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF


def philox4x32_10(ctr, k0, k1):
    c = list(ctr)
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c[3] ^ k1) & M32, p0 & M32]
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c


def philox_uniform(seed, stream, which, d):
    blk = d >> 1
    c = philox4x32_10([blk & M32, (blk >> 32) & M32, stream & M32, ((stream >> 32) & M32) ^ ((which << 28) & M32)],
                      seed & M32, (seed >> 32) & M32)
    v = ((c[3] << 32) | c[2]) if (d & 1) else ((c[1] << 32) | c[0])
    return (v >> 11) * (1.0 / 9007199254740992.0)


def fmix32(h):
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & M32
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & M32
    h ^= h >> 16
    return h


def ring_replica_hash(ring_seed, replica):
    h = fmix32((replica >> 32) & M32)
    h = fmix32((replica & M32) ^ h)
    h = fmix32(((ring_seed >> 32) & M32) ^ h)
    return fmix32((ring_seed & M32) ^ h)


def ring_start(ring_seed, replica, bucket, length):
    h = fmix32(ring_replica_hash(ring_seed, replica) ^ (((bucket + 1) * 0x9E3779B1) & M32))
    return (h * length) >> 32


def init_ring_start(ring_seed, replica, n_init):
    return (fmix32(ring_replica_hash(ring_seed, replica) ^ 0x85EBCA6B) * n_init) >> 32
