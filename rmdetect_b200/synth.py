"""Deterministic synthetic inputs: random genomes and stand-in bpp matrices.

ViennaRNA is not part of this path (it stays the reference's external
dependency, reference lib/scanner.py:130) and is absent from the build image,
so benchmarks and parity tests need a base-pair-probability source that
(i) depends only on the window's letters, as a fold does, (ii) is cheap, and
(iii) is bit-identical on the host (here, numpy) and on the device
(`rmd_synth_bpp`, csrc/synth_kernels.cu).  Everything below is integer
hashing; the only floating-point steps are one division and one multiplication
per entry, both IEEE-exact, mirroring how the reference turns a `ubox` line
into a probability (`float(sqrt_p) ** 2`, reference lib/rnatools.py:189-195).

Model of a matrix: cells (i, j), j - i >= 4, whose letters can pair
(AU, GC, GU) may start a helix with probability Q; a helix runs inward
(i+1, j-1), … while letters stay complementary, for a hashed length 1..8; its
members share a base level drawn from three bands (p in [1e-5,1e-3),
[1e-3,1e-1), [1e-1,1]) with per-cell jitter.  With Q = 31/256 a random 150-nt
window carries about 4 entries per nt, three quarters of them below 1e-3 — the
regime SURVEY.md §6 measured on the reference's real `dot.ps` files.
"""
from __future__ import annotations

import numpy as np

from .bpprobs import BPProbs

U64 = np.uint64
GOLDEN = U64(0x9E3779B97F4A7C15)
MIN_SPAN = 4
LOOKBACK = 8
SEED_Q24 = 31 << 16                  # threshold on a 24-bit uniform: Q = 31/256
SQRT_MIN = 3162278                  # sqrt(1e-5) in units of 1e-9
BAND_LO = (3162278, 31622777, 316227767)
BAND_HI = (31622777, 316227767, 1000000001)
LEN_TAB = np.array([1, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 5, 6, 7, 8], dtype=np.int64)


def mix64(x):
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    x = np.asarray(x, dtype=U64)
    with np.errstate(over="ignore"):
        x = (x ^ (x >> U64(30))) * U64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> U64(27))) * U64(0x94D049BB133111EB)
        x = x ^ (x >> U64(31))
    return x


def synth_codes(n, seed, start=0, reverse_total=0):
    """iid uniform letter codes 0..3 for positions start..start+n-1 of the stream, or, with `reverse_total`,
    of the reverse-complement strand of a stream of that many letters (reference lib/sequences.py:138-143:
    reversed, A<->U, C<->G; with A,C,G,U = 0..3 the complement of c is 3 - c)."""
    k = np.arange(start, start + n, dtype=np.int64)
    pos = (reverse_total - 1 - k if reverse_total else k).astype(U64)
    with np.errstate(over="ignore"):
        h = mix64(U64(seed) + GOLDEN * (pos + U64(1)))
    c = (h >> U64(62)).astype(np.uint8)
    return (3 - c).astype(np.uint8) if reverse_total else c


def synth_sequence(n, seed, start=0, reverse_total=0) -> str:
    return np.array([65, 67, 71, 85], dtype=np.uint8)[synth_codes(n, seed, start, reverse_total)].tobytes().decode()


def synth_calibration_samples(seed, first, n, L):
    """Host twin of the draw inside rmd_calibrate_synth (csrc/calib_kernels.cuh): sample s is the low 2L bits of
    mix64(seed + GOLDEN * (s + 1)), letter j its bits 2j and 2j + 1, s in [first, first + n); L <= 32.  Returns codes [n][L]."""
    g = np.arange(first, first + n, dtype=np.int64).astype(U64)
    with np.errstate(over="ignore"):
        h = mix64(U64(seed) + GOLDEN * (g + U64(1)))
    return ((h[:, None] >> (U64(2) * np.arange(L, dtype=U64))[None, :]) & U64(3)).astype(np.uint8)


def _hash_bytes(seq: str) -> np.ndarray:
    """Per-letter byte used for the window key: 0..3 for ACGU, else the ASCII code."""
    raw = np.frombuffer(seq.encode("latin-1"), dtype=np.uint8).copy()
    out = raw.copy()
    for k, b in enumerate("ACGU"):
        out[raw == ord(b)] = k
    return out


def window_key(seq: str, seed) -> int:
    """64-bit key of a window: chained mix over 8-letter little-endian words, then the length."""
    hb = _hash_bytes(seq)
    n = len(hb)
    pad = (-n) % 8
    words = np.concatenate([hb, np.zeros(pad, dtype=np.uint8)]).view("<u8")
    key = U64(seed)
    for w in words:
        key = mix64(key ^ U64(w))
    return int(mix64(key ^ U64(n)))


def _cell_hash(key, i, j, salt):
    with np.errstate(over="ignore"):
        v = (i.astype(U64) << U64(34)) | (j.astype(U64) << U64(4)) | U64(salt)
        return mix64(U64(key) + GOLDEN * (v + U64(1)))


def _can_pair(a, b):
    """AU, GC, GU in either order on codes A=0 C=1 G=2 U=3; anything else cannot."""
    s = a.astype(np.int64) + b.astype(np.int64)
    ok = (a < 4) & (b < 4)
    return ok & (((s == 3) & (a != b)) | ((s == 5) & (a != b) & (np.minimum(a, b) == 2)))


def synth_bpp_coo(seq: str, seed):
    """Upper-triangle coordinate list (i, j, p) of the window's stand-in matrix, sorted by (i, j)."""
    n = len(seq)
    if n < MIN_SPAN + 1:
        z = np.zeros(0, dtype=np.uint16)
        return z, z.copy(), np.zeros(0, dtype=np.float64)
    key = window_key(seq, seed)
    codes = _hash_bytes(seq)
    ii, jj = np.triu_indices(n, k=MIN_SPAN)
    best = np.zeros(len(ii), dtype=np.int64)
    alive = _can_pair(codes[ii], codes[jj])
    for t in range(LOOKBACK):
        si, sj = ii - t, jj + t
        inb = (si >= 0) & (sj < n)
        alive = alive & inb
        if not alive.any():
            break
        ci = np.where(inb, si, 0)
        cj = np.where(inb, sj, 0)
        alive = alive & _can_pair(codes[ci], codes[cj])
        h0 = _cell_hash(key, ci, cj, 0)
        seeded = alive & ((h0 >> U64(40)) < U64(SEED_Q24))
        length = LEN_TAB[((h0 >> U64(8)) & U64(15)).astype(np.int64)]
        member = seeded & (length > t)
        if member.any():
            band_u = ((h0 >> U64(16)) & U64(0xFFFF)).astype(np.int64)
            band = np.where(band_u < 49152, 0, np.where(band_u < 60000, 1, 2))
            lo = np.array(BAND_LO, dtype=np.int64)[band]
            hi = np.array(BAND_HI, dtype=np.int64)[band]
            h1 = _cell_hash(key, ci, cj, 1)
            base = lo + (h1 % (hi - lo).astype(U64)).astype(np.int64)
            h2 = _cell_hash(key, ii, jj, 2 + t)
            half = base >> 1
            val = half + ((half * (h2 & U64(1023)).astype(np.int64)) >> 10)
            val = np.maximum(val, SQRT_MIN)
            best = np.where(member, np.maximum(best, val), best)
    on = best > 0
    i = ii[on].astype(np.uint16)
    j = jj[on].astype(np.uint16)
    root = best[on].astype(np.float64) / 1e9
    return i, j, root * root


class SynthFold:
    """Stand-in for the reference's `RNATools` at the one call the scan path makes.

    `bpprobs(seq, constraint)` has the reference signature (lib/rnatools.py:31-39)
    and returns a `BPProbs`; `coo(seq)` returns the arrays directly.
    """

    def __init__(self, seed=20261019, factory=BPProbs):
        self.seed = seed
        self.factory = factory

    def coo(self, seq):
        return synth_bpp_coo(seq, self.seed)

    def bpprobs(self, seq, constraint=""):
        i, j, p = synth_bpp_coo(seq, self.seed)
        out = self.factory()
        for a, b, v in zip(i.tolist(), j.tolist(), p.tolist()):
            out.add(a, b, v)
        return out
