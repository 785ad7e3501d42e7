"""Shared fixtures: golden files written by tests/golden/make_golden.py, model loading."""
import functools
import gzip
import json
import math
import os

from rmdetect_b200.evalues import parse_evalues
from rmdetect_b200.modelfile import parse_model

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_DATA = os.path.join(GOLDEN, "ref_data")
MODEL_FILES = {
    "KT_1.0": "models/kt_1", "GB_1.0": "models/gb_1", "CL_1.0": "models/cl_1",
    "TGA_1.0": "models/tga_1", "ARICH_1.0": "arich/arich_1.0",
}
MODEL_ORDER = ["CL_1.0", "TGA_1.0", "GB_1.0", "KT_1.0"]
LN_CUTOFF = math.log(0.1)
LN_UNKNOWN = math.log(0.0001)


@functools.lru_cache(maxsize=None)
def golden(name):
    with gzip.open(os.path.join(GOLDEN, name + ".json.gz"), "rt") as fh:
        return json.load(fh)


def model_path(full_name):
    return os.path.join(REF_DATA, MODEL_FILES[full_name] + ".model")


def evalues_path(full_name):
    return os.path.join(REF_DATA, MODEL_FILES[full_name] + ".evalues")


@functools.lru_cache(maxsize=None)
def load_model(full_name):
    return parse_model(model_path(full_name))


@functools.lru_cache(maxsize=None)
def load_evalues(full_name):
    return parse_evalues(evalues_path(full_name))


class Cfg:
    """Stand-in for the reference's Config object: just the attributes Scanner reads."""

    def __init__(self, models, evalues, win_len=150, win_step=75, min_bpp=0.001, cutoff=0.1, unknown=1e-4):
        self.models, self.evalues = models, evalues
        self.win_len, self.win_step = win_len, win_step
        self.constraint = ""
        self.min_bpp_mean = min_bpp
        self.cutoff = math.log(cutoff)
        self.unknown_prob = math.log(unknown)


class Seq:
    def __init__(self, key, seq, col_index):
        self.key, self.seq, self.col_index = key, seq, col_index


class SeqList:
    def __init__(self, triples):
        self.sequence_list = [Seq(*t) for t in triples]


def fresh_models(names):
    """New model/evalues objects (the Scanner caches its engine by object identity)."""
    return [parse_model(model_path(n)) for n in names], [parse_evalues(evalues_path(n)) for n in names]
