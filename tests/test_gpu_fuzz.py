"""Seeded random configurations of the whole scan (models, window geometry, thresholds, letters, matrix densities,
screens forced on) against the CPU oracle: hits, positions, gap configurations, fp64 sums, tries."""
import math
import random
import zlib

import numpy as np
import pytest

from helpers import LN_UNKNOWN, MODEL_FILES, evalues_path, load_evalues, load_model
from rmdetect_b200.synth import SynthFold

pytestmark = pytest.mark.gpu


class _Mixed:
    """Stand-in matrices of varying density: the synthetic fold, thinned or boosted per window."""

    def __init__(self, seed, keep, boost):
        self.fold, self.keep, self.boost, self.seed = SynthFold(seed), keep, boost, seed

    def coo(self, win):
        i, j, p = self.fold.coo(win)
        rng = np.random.default_rng(len(win) * 7919 + self.seed + zlib.crc32(win.encode()))
        sel = rng.random(len(p)) < self.keep
        i, j, p = i[sel], j[sel], p[sel].copy()
        hot = rng.random(len(p)) < self.boost
        p[hot] = np.minimum(p[hot] * rng.choice([3.0, 10.0, 40.0], size=int(hot.sum())), 1.0)
        return i, j, p


@pytest.mark.parametrize("case", range(14))
def test_random_configuration_matches_oracle(case):
    from oracle import pyoracle
    from rmdetect_b200.scanner import ScanEngine
    rng = random.Random(1000 + case)
    names = rng.sample(sorted(MODEL_FILES), rng.randint(1, 5))
    win_len = rng.choice([150, 150, 120, 100, 160, 97])
    win_step = rng.choice([75, 75, 50, 33, 150, 10])
    min_bpp = rng.choice([0.001, 0.001, 0.01, 0.0003, 0.05])
    cutoff = math.log(rng.choice([0.1, 0.1, 0.5, 0.01]))
    letters = rng.choice(["ACGU", "ACGU", "AACGGU", "ACGUN", "GGCCAU"])
    seqs = [("s%d" % k, "".join(rng.choice(letters) for _ in range(rng.choice([40, 151, 333, 700, 1200]))), None)
            for k in range(rng.randint(1, 3))]
    fold = _Mixed(case, rng.choice([1.0, 0.6, 0.25]), rng.choice([0.0, 0.05, 0.3]))
    eng = ScanEngine([load_model(n) for n in names], [load_evalues(n) for n in names], LN_UNKNOWN)
    eng.ctx.set_screen_policy(0)
    res = eng.scan_job(eng.build_job(seqs, win_len, win_step, fold, "", min_bpp, cutoff))
    eng.close()
    oms = [pyoracle.OracleModel(load_model(n)) for n in names]
    oes = [pyoracle.OracleEValues(evalues_path(n)) for n in names]
    at, tests = 0, [0, 0]
    for si, (_, seq, _) in enumerate(seqs):
        starts = pyoracle.windows(len(seq), win_len, win_step)
        want, tries, _ = pyoracle.scan_sequence(oms, oes, si, seq, win_len, win_step, [fold.coo(seq[s:s + win_len]) for s in starts],
                                                min_bpp, LN_UNKNOWN, cutoff)
        tests[0] += tries[0]; tests[1] += tries[1]
        for w in want:
            assert at < len(res["hits"]), (case, "missing hits")
            g = res["hits"][at]; at += 1
            assert (g["seq"], g["model"], g["start"], g["p0"], g["p1"], g["leaf"]) == (si, w["model"], w["coords"][0], w["p"][0], w["p"][1], w["cfg"]), case
            assert (g["prob_model"], g["prob_rand"], g["score"], g["evalue"]) == (w["prob_model"], w["prob_rand"], w["score"], w["evalue"]), case
    assert at == len(res["hits"]) and res["tries"] == tests, (case, res["tries"], tests)
