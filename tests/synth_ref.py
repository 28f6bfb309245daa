"""Independent numpy restatement of the device-side synthetic graph (smore_b200/csrc/device_graph.cu: synth_line,
synth_endpoint, synth_perm) -- test infrastructure only. Line i of the edge list is a pure function of (seed, i):
Philox4x32-10 block i of stream 2^62 + 2^61 -> two endpoints with inverse CDF rank = V * u^3, relabelled by a 4-round
Feistel bijection cycle-walked into [0, V), and an integer weight 1 + umulhi(word, 5)."""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & M32
        hi1, lo1 = p1 >> np.uint64(32), p1 & M32
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & M32
        k1 = (k1 + np.uint64(0xBB67AE85)) & M32
    return c0, c1, c2, c3


def _mul32(a, b):
    return (a * np.uint64(b)) & M32


def synth_perm(x, V, seed):
    bits = 2
    while (1 << bits) < V:
        bits += 1
    hb = bits >> 1
    lb = bits - hb
    hmask, lmask = np.uint64((1 << hb) - 1), np.uint64((1 << lb) - 1)
    s_lo, s_hi = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    x = np.asarray(x, dtype=np.uint64).copy()
    todo = np.ones(len(x), dtype=bool)
    while todo.any():
        y = x[todo]
        hi, lo = y >> np.uint64(lb), y & lmask
        for r in range(4):
            f = _mul32((lo + _mul32(np.uint64(0x9E3779B9), r + 1) + s_lo) & M32, 0x85EBCA6B)
            f ^= f >> np.uint64(13)
            f = _mul32(f, 0xC2B2AE35)
            f ^= f >> np.uint64(16)
            hi = (hi ^ f) & hmask
            g = _mul32((hi + _mul32(np.uint64(0x7F4A7C15), r + 1) + s_hi) & M32, 0xCC9E2D51)
            g ^= g >> np.uint64(15)
            g = _mul32(g, 0x1B873593)
            g ^= g >> np.uint64(16)
            lo = (lo ^ g) & lmask
        y = (hi << np.uint64(lb)) | lo
        x[todo] = y
        todo[todo] = y >= np.uint64(V)
    return x.astype(np.int64)


def synth_endpoint(k, V, seed):
    u = (k.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)
    r = (((float(V) * u) * u) * u).astype(np.int64)
    r = np.minimum(r, V - 1)
    return synth_perm(r, V, seed)


def synth_edges(V, E_lines, seed):
    """(u, v, w) of every kept line (self loops dropped), in line order."""
    i = np.arange(E_lines, dtype=np.uint64)
    stream = (1 << 62) + (1 << 61)
    r0, r1, r2, _ = philox4x32_10(i & M32, i >> np.uint64(32), np.uint64(stream & 0xFFFFFFFF), np.uint64(stream >> 32),
                                  seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = synth_endpoint(r0, V, seed)
    v = synth_endpoint(r1, V, seed)
    w = 1 + ((r2 * np.uint64(5)) >> np.uint64(32)).astype(np.int64)
    keep = u != v
    return u[keep], v[keep], w[keep].astype(np.float64)
