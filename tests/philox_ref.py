"""numpy restatement of Philox4x32-10 (Salmon et al. SC'11) for the RNG tests."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c = [np.asarray(x, dtype=np.uint32).copy() for x in (c0, c1, c2, c3)]
    k0, k1 = int(k0), int(k1)
    for _ in range(10):
        p0 = M0 * c[0].astype(np.uint64)
        p1 = M1 * c[2].astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
        c = [hi1 ^ c[1] ^ np.uint32(k0), lo1, hi0 ^ c[3] ^ np.uint32(k1), lo0]
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, axis=-1)


# Known-answer vectors from the Random123 distribution (kat_vectors)
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def pair_words(seed, ids, pair):
    """The two Philox words (wa, wb) of Box-Muller pair `pair` of draws `ids`
    (csrc/philox.cuh: philox_pair_words -- call c2 = pair // 2 of stream 0, words
    (x, y) for even pairs, (z, w) for odd ones)."""
    ids = np.asarray(ids, dtype=np.uint64)
    n = ids.size
    w = philox4x32_10(ids.astype(np.uint32), (ids >> np.uint64(32)).astype(np.uint32),
                      np.full(n, pair // 2, np.uint32), np.zeros(n, np.uint32),
                      int(seed) & 0xFFFFFFFF, int(seed) >> 32)
    return (w[:, 2], w[:, 3]) if pair & 1 else (w[:, 0], w[:, 1])


def normals_from_words(wa, wb):
    """numpy restatement of the device's word -> N(0,1) pair mapping (csrc/philox.cuh):
    u = ((wa : wb[13:0]) + 1/2) 2^-46, angle = 2 pi ((wb >> 14) + 1/2) / 2^18."""
    wa = wa.astype(np.uint64)
    wb = wb.astype(np.uint64)
    x = (wa << np.uint64(14)) | (wb & np.uint64(0x3FFF))          # 46 random bits
    u = (x.astype(np.float64) + 0.5) * 2.0 ** -46
    rad = np.sqrt(-2.0 * np.log(u))
    ang = 2.0 * np.pi * ((wb >> np.uint64(14)).astype(np.float64) + 0.5) / 2.0 ** 18
    return rad * np.cos(ang), rad * np.sin(ang)
