"""The C ABI (include/rmdetect_b200.h): the library loads, exports every function the header declares, the ctypes
layer binds the same set, struct sizes agree, and without a CUDA device the library fails loudly instead of falling
back to anything (no compute call is made here)."""
import ctypes as C
import os
import re

import numpy as np

from rmdetect_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rmdetect_b200.h")


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"^[A-Za-z_][\w \*]*?\b(rmd_\w+)\s*\(", text, flags=re.M)))


def test_header_declares_the_entry_points():
    names = declared_functions()
    assert len(names) >= 25 and "rmd_scan" in names and "rmd_create" in names and "rmd_parse_dot_ps" in names


def test_library_exports_every_declared_function():
    lib = C.CDLL(_native.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_layer_binds_exactly_the_header():
    assert sorted(_native.EXPORTS) == declared_functions()
    lib = _native.load_library()
    for name in _native.EXPORTS:
        getattr(lib, name)


def test_struct_layouts_match_the_header(tmp_path):
    """The header compiles as plain C, and its struct sizes are the ones the ctypes / numpy mirrors use."""
    import subprocess
    from rmdetect_b200.modelflat import PAIR_DTYPE, TREE_DTYPE
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "%s"\nint main(void) { printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu\\n", sizeof(rmd_scan_result), '
                   'sizeof(rmd_scan_input), sizeof(rmd_model_desc), sizeof(rmd_hit), sizeof(rmd_tree_node), sizeof(rmd_gate_pair), sizeof(rmd_hit32)); return 0; }\n' % HEADER)
    exe = tmp_path / "sizes"
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", str(src), "-o", str(exe)], env=env)
    sizes = [int(t) for t in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(_native.ScanResult), C.sizeof(_native.ScanInput), C.sizeof(_native.ModelDesc),
                     _native.HIT_DTYPE.itemsize, TREE_DTYPE.itemsize, PAIR_DTYPE.itemsize, _native.HIT32_DTYPE.itemsize]


def test_version_and_error_text_without_a_context():
    lib = _native.load_library()
    assert b"sm_100a" in lib.rmd_version()
    assert isinstance(lib.rmd_last_error(None), bytes)


def test_no_device_means_failure_not_fallback():
    import torch
    lib = _native.load_library()
    h = C.c_void_p()
    rc = lib.rmd_create(0, C.byref(h))
    if torch.cuda.is_available():
        assert rc == 0
        lib.rmd_destroy(h)
    else:
        assert rc == -2 and b"no CPU path" in lib.rmd_last_error(None)
        try:
            _native.Context(0)
        except _native.NativeError as exc:
            assert "no CPU path" in str(exc)
        else:
            raise AssertionError("a context without a CUDA device")


def test_host_only_entry_points_work_without_a_device():
    """rmd_compact_bpp and rmd_parse_dot_ps are host code: exercised in test_ingest.py; here only the round trip of
    one entry through the 6-byte transport word."""
    i = np.array([3], dtype=np.uint16); j = np.array([40], dtype=np.uint16)
    p = np.array([0.031622780 ** 2.0])
    out = _native.compact_bpp(i, j, p)
    assert out is not None and int(out[0][0]) == 3 | (40 << 8) and int(out[1][0]) & 0x3fffffff == 31622780


def test_release_library_reads_no_tuning_switch():
    """RMD_DEBUG, RMD_TEAMS, RMD_CTAS_PER_SM, RMD_CARVEOUT, RMD_LOCAL_STACK, RMD_NO_SCREEN and RMD_NO_FUSED_EXPAND exist only in
    -DRMD_TUNING builds: the shipped library must not even contain their names (nor import getenv for them)."""
    blob = open(_native.LIB_PATH, "rb").read()
    for name in (b"RMD_DEBUG", b"RMD_TEAMS", b"RMD_CTAS_PER_SM", b"RMD_CARVEOUT", b"RMD_LOCAL_STACK", b"RMD_NO_SCREEN", b"RMD_NO_FUSED_EXPAND"):
        assert name not in blob, name
