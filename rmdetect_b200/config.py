"""The configuration object `Scanner(config)` reads, for callers without a reference checkout.

Restates lib/config.py:19-140 (reference root) — defaults, the search order for the model directory, models and their
`.evalues` tables loaded in `os.listdir` order with the `--inc` / `--exc` lists — and the lines of rmdetect.py:196-222 that
turn the command line's numbers into the fields the scan uses (`cutoff` and `unknown_prob` as natural logs, `win_step`
half the window unless given).  The command line itself stays the reference's.
"""
import math
import os
import sys

from .evalues import parse_evalues
from .modelfile import parse_model


class Config:
    VERSION = "0.0.6"
    ENV_DATA_PATH = "RMDETECT_DATA"
    MODEL_SUFFIX = "model"
    EVALUES_SUFFIX = "evalues"

    def __init__(self, win_len=150, win_step=None, cutoff=0.1, unknown_prob=0.0001, min_bpp_mean=0.001, data_path=None,
                 include=(), exclude=(), both_strands=False, constraint="", log=sys.stderr):
        """The keyword defaults are rmdetect.py's option defaults (:171-190); lib/config.py's own (:27-47) differ only in
        win_len = win_step = 0, which rmdetect.py always overwrites."""
        self.verbose = False
        self.cutoff_correction = math.log(0.1)
        self.cutoff = math.log(cutoff)                      # rmdetect.py:220
        self.unknown_prob = math.log(unknown_prob)          # :221
        self.min_bpp_mean = min_bpp_mean
        self.win_len = win_len
        self.win_step = int(win_len / 2) if win_step is None else win_step      # :199-202
        self.data_path = data_path
        self.both_strands = both_strands
        self.models = []
        self.evalues = []
        self.exclude = list(exclude)
        self.include = list(include)
        self.constraint = constraint
        self.log = log

    def _say(self, text):
        if self.log is not None:
            self.log.write(text)

    def configure(self):
        self._set_data_path()
        self._load_models()
        return self

    def _set_data_path(self):
        """lib/config.py:58-93: --data-path, then $RMDETECT_DATA, then <script directory>/models, then the current directory."""
        env = os.getenv(self.ENV_DATA_PATH)
        script = "%s/models" % os.path.dirname(sys.argv[0])
        if not (self.data_path is not None and os.path.isdir(self.data_path)):
            if env is not None and os.path.isdir(env):
                self.data_path = env
            elif os.path.isdir(script):
                self.data_path = script
            else:
                self.data_path = os.curdir
        self.data_path = os.path.dirname(self.data_path + "/")
        self._say("Searching for models: OK! -> '%s'\n" % self.data_path)

    def _load_models(self):
        """lib/config.py:95-140: every file that ends in "model", in directory order; a model is taken when it is named by
        `include` (or `include` is empty) and not by `exclude`; its table is the file with the ending swapped."""
        if not os.path.isdir(self.data_path):
            sys.stderr.write("ERROR: defined data_path directory '%s' doesn't exist!\n" % self.data_path)
            raise SystemExit
        included, excluded = [], []
        for f in os.listdir(self.data_path):
            if not f.endswith(self.MODEL_SUFFIX):
                continue
            model = parse_model("%s/%s" % (self.data_path, f))
            if (model.full_name() in self.include or len(self.include) == 0) and model.full_name() not in self.exclude:
                self.models.append(model)
                self.evalues.append(parse_evalues("%s/%s" % (self.data_path, f.replace(self.MODEL_SUFFIX, self.EVALUES_SUFFIX))))
                included.append(model.full_name())
            else:
                excluded.append(model.full_name())
        self._say("Included models:  %s\nExcluded models:  %s\n" % (", ".join(included or ["none"]), ", ".join(excluded or ["none"])))
        if len(self.models) == 0:
            print("WARNING: No models read from '%s'\n" % self.data_path)
