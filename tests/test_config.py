"""rmdetect_b200/config.py: the fields Scanner reads, with rmdetect.py's defaults, and the model directory loader
(lib/config.py:95-140)."""
import math
import os

import pytest

from rmdetect_b200.config import Config

HERE = os.path.dirname(os.path.abspath(__file__))
MODELS = os.path.join(HERE, "golden", "ref_data", "models")


def test_defaults_are_the_command_lines():
    c = Config(log=None)
    assert (c.win_len, c.win_step, c.min_bpp_mean, c.constraint) == (150, 75, 0.001, "")
    assert c.cutoff == math.log(0.1) and c.unknown_prob == math.log(0.0001)
    assert Config(win_len=101, log=None).win_step == 50 and Config(win_len=100, win_step=10, log=None).win_step == 10


def test_directory_order_include_and_exclude():
    c = Config(data_path=MODELS, log=None).configure()
    order = [f for f in os.listdir(MODELS) if f.endswith("model")]
    assert [m.full_name() for m in c.models] == [n.split(".")[0].rsplit("_", 1)[0].upper() + "_1.0" for n in order]
    assert len(c.evalues) == len(c.models) == 4 and c.data_path == MODELS
    one = Config(data_path=MODELS + "/", include=["KT_1.0"], log=None).configure()
    assert [m.full_name() for m in one.models] == ["KT_1.0"] and len(one.evalues) == 1
    rest = Config(data_path=MODELS, exclude=["KT_1.0", "CL_1.0"], log=None).configure()
    assert sorted(m.full_name() for m in rest.models) == ["GB_1.0", "TGA_1.0"]


def test_search_order(tmp_path, monkeypatch, capsys):
    monkeypatch.setenv("RMDETECT_DATA", MODELS)
    c = Config(data_path=str(tmp_path / "missing"), log=None).configure()      # --data-path is not a directory: the environment
    assert c.data_path == MODELS and len(c.models) == 4
    monkeypatch.delenv("RMDETECT_DATA")
    monkeypatch.setattr("sys.argv", [str(tmp_path / "script.py")])
    monkeypatch.chdir(tmp_path)
    c = Config(log=None).configure()                                           # nothing found: the current directory
    assert c.data_path == os.curdir and c.models == []
    assert "WARNING: No models read" in capsys.readouterr().out
