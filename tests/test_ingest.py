"""Host-side ingest in the C-ABI library: dot.ps parser and the compact bpp transport (no GPU needed)."""
import os
import random

import numpy as np

from helpers import REF_DATA
from rmdetect_b200._native import compact_bpp, parse_dot_ps
from rmdetect_b200.dotps import read_dot_ps, read_dot_ps_native

DOTS = [os.path.join(REF_DATA, "lys_dot.ps"), os.path.join(REF_DATA, "arich", "dot.ps")]


def expand(ij, q):
    """numpy twin of expand_bpp_kernel (rmdetect_b200/csrc/post_kernels.cuh)."""
    x = (q & 0x3FFFFFFF).astype(np.float64) / 1e9
    bits = (x * x).view(np.int64) + (q >> 30).astype(np.int64) - 1
    return (ij & 0xFF).astype(np.uint16), (ij >> 8).astype(np.uint16), bits.view(np.float64)


def test_native_dot_ps_parser_matches_python_reader():
    for path in DOTS:
        seq, i, j, p = read_dot_ps(path)
        nseq, ni, nj, np_, nq = read_dot_ps_native(path)
        assert nseq == seq
        assert np.array_equal(ni, i) and np.array_equal(nj, j)
        assert np.array_equal(np_, p)                       # bit for bit: float(tok) ** 2.0
        assert len(p) > 200 and (nq != 0).all()
        ei, ej, ep = expand((ni | (nj << 8)).astype(np.uint16), nq)
        assert np.array_equal(ep, p)


def test_parser_line_rules():
    text = ("%!PS\n/sequence { (\\\nACGUACGU\\\n) } def\n"
            "1 8 0.500000000 ubox\n2 7 0.000000000 ubox\n8 1 0.250000000 ubox\n3 6 0.9 lbox\n"
            "% 4 5 0.3 ubox\n4 5 0.100000000 ubox\n4 5 0.200000000 ubox\n1 2 3 4 ubox\n")
    seq, i, j, p, q = parse_dot_ps(text)
    assert seq == "ACGUACGU"
    # zero dropped, lbox and comment ignored, five-token line ignored, (8, 1) replaces (1, 8), later (4, 5) wins
    assert i.tolist() == [0, 3] and j.tolist() == [7, 4]
    assert p.tolist() == [0.25 ** 2.0, 0.2 ** 2.0]


def test_compact_round_trip_and_rejects():
    rng = random.Random(5)
    ks = [1, 2, 999999999, 10 ** 9] + [rng.randrange(1, 10 ** 9) for _ in range(200000)]
    p = np.array([float("%.9f" % (k / 1e9)) ** 2.0 for k in ks])
    i = np.array([rng.randrange(0, 100) for _ in ks], dtype=np.uint16)
    j = (i + 1 + np.array([rng.randrange(0, 100) for _ in ks])).astype(np.uint16)
    ij, q = compact_bpp(i, j, p)
    assert ((q & 0x3FFFFFFF) == np.array(ks)).all()
    assert set((q >> 30).tolist()) <= {0, 1, 2} and (q >> 30 != 1).sum() > 50     # libm pow is not always x * x
    ei, ej, ep = expand(ij, q)
    assert np.array_equal(ei, i) and np.array_equal(ej, j) and np.array_equal(ep, p)
    assert compact_bpp(i[:3], j[:3], np.array([0.1234567891234, 0.5, 0.5])) is None          # not a squared 9-decimal
    assert compact_bpp(np.array([300], dtype=np.uint16), np.array([301], dtype=np.uint16), p[:1]) is None
    assert compact_bpp(i[:1], j[:1], np.array([-0.25])) is None


def test_transport_root_is_the_ieee_quotient_for_every_word():
    """The device expands a transport word q as q * 1e-9 plus one residual correction (csrc/device_common.cuh:
    transport_root) instead of dividing by 1e9; the same three IEEE operations on the host give q / 1e9 for all 2^30 words."""
    import ctypes as C
    from oracle import pyoracle
    L = pyoracle.lib()
    L.orc_check_div1e9.restype = C.c_int64
    assert L.orc_check_div1e9() == 0
