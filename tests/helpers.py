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


def mix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def ring_start(ring_seed, replica, bucket, length):
    h = mix64((mix64(ring_seed ^ ((replica * 0xD1342543DE82EF95) & M64)) + bucket) & M64)
    return ((h >> 32) * length) >> 32


def init_ring_start(ring_seed, replica, n_init):
    h = mix64(ring_seed ^ mix64((replica + 0x51ED270B1) & M64))
    return ((h >> 32) * n_init) >> 32
