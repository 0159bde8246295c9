"""oracle/points.py -- TEST INFRASTRUCTURE: numpy restatement of the reference's Halton point sampler.

  pointsampler()       src/pointsampler.d/halton.c:69-84   dimension = rand_beg + i, index clipped to 32 bits,
                                                           dimensions >= 256 fall back to the Mersenne twister
  halton_sample        ext/halton/halton.h:291-305 (base 2 = bit reversal), :309-2416 (one function per prime:
                       acc = sum perm_b[digit_k] * b^(D-1-k) over the D lowest base-b digits, times
                       (float)(0x1.fffffcp-1 / b^D)), dispatcher :2418
  halton_init_random   ext/halton/halton.h:3244-3274: Fisher-Yates per base from lrand48() after srand48(frame),
                       bases 1..3 identity; _halton_init_tables expands the digit permutation of base b to blocks of
                       b^digits entries (`perm<P>[i]` = permuted digits of i, most significant first reversed, see
                       ext/halton/halton_gen.py) -- restated here digit by digit, which is the same number.

Pinned by tests/golden/halton.npz (values from the compiled reference, oracle/_ref/libref_halton.so)."""
import numpy as np

NUM_DIMS = 256


def primes(n):
    out, c = [], 1
    while len(out) < n:
        c += 1
        if all(c % p for p in out if p * p <= c):
            out.append(c)
    return out


class Lrand48:
    """glibc srand48/lrand48: X <- (0x5DEECE66D X + 0xB) mod 2^48, X0 = seed<<16 | 0x330E, result X >> 17"""

    def __init__(self, seed):
        self.x = ((int(seed) & 0xFFFFFFFF) << 16) | 0x330E

    def __call__(self):
        self.x = (0x5DEECE66D * self.x + 0xB) & ((1 << 48) - 1)
        return self.x >> 17


def digit_permutations(frame):
    """perms[base] for base 1..1619 (halton.h:3244-3270)"""
    r = Lrand48(frame)
    perms = {}
    for base in range(1, 1620):
        p = list(range(base))
        if base >= 4:
            for i in range(base - 1):
                j = i + r() // ((1 << 31) // (base - i) + 1)
                p[i], p[j] = p[j], p[i]
        perms[base] = p
    return perms


class Halton:
    def __init__(self, frame):
        self.perms = digit_permutations(frame)
        self.bases = primes(NUM_DIMS)
        self.ndigits, self.scale = [], []
        for b in self.bases:
            # halton_gen.py: digits per table lookup = largest power of b <= 500, lookups = as many as fit in 32 bits
            dg, pw = 1, b
            while pw * b <= 500:
                pw *= b
                dg += 1
            mx, lk = pw, 1
            while mx * pw < (1 << 32):
                mx *= pw
                lk += 1
            self.ndigits.append(dg * lk)
            self.scale.append(np.float32(float.fromhex("0x1.fffffcp-1") / mx))

    def sample(self, dim, index):
        """halton_sample(h, dim, (unsigned)index) for arrays"""
        dim = np.asarray(dim, np.int64)
        index = np.asarray(index, np.uint64) & np.uint64(0xFFFFFFFF)
        out = np.zeros(len(dim), np.float32)
        for d in np.unique(dim):
            m = dim == d
            idx = index[m].astype(np.uint64)
            if d == 0:
                v = idx.astype(np.uint32)
                v = (v << np.uint32(16)) | (v >> np.uint32(16))
                v = ((v & np.uint32(0x00ff00ff)) << np.uint32(8)) | ((v & np.uint32(0xff00ff00)) >> np.uint32(8))
                v = ((v & np.uint32(0x0f0f0f0f)) << np.uint32(4)) | ((v & np.uint32(0xf0f0f0f0)) >> np.uint32(4))
                v = ((v & np.uint32(0x33333333)) << np.uint32(2)) | ((v & np.uint32(0xcccccccc)) >> np.uint32(2))
                v = ((v & np.uint32(0x55555555)) << np.uint32(1)) | ((v & np.uint32(0xaaaaaaaa)) >> np.uint32(1))
                out[m] = (np.uint32(0x3f800000) | (v >> np.uint32(9))).view(np.float32) - np.float32(1.0)
                continue
            b = self.bases[d]
            perm = np.asarray(self.perms[b], np.uint64)
            acc = np.zeros(len(idx), np.uint64)
            for _ in range(self.ndigits[d]):
                acc = acc * np.uint64(b) + perm[idx % np.uint64(b)]
                idx = idx // np.uint64(b)
            # unsigned -> float conversion (the C expression's integer sum is `unsigned`), then one float multiply
            out[m] = (acc & np.uint64(0xFFFFFFFF)).astype(np.uint32).astype(np.float32) * self.scale[d]
        return out
