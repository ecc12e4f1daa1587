"""The C-ABI library loads, exports every symbol include/sdvpcm.h declares, and refuses to compute without a GPU."""
import ctypes as C
import os
import re

import pytest

from sdvpcmdecoder_b200 import _build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def library():
    _build.build_library()
    return capi.lib()


def test_exports_match_header(library):
    header = open(os.path.join(ROOT, "include", "sdvpcm.h")).read()
    declared = sorted(set(re.findall(r"SDV_API\s+[\w\s\*]+?\b(sdv_\w+)\s*\(", header)))
    assert declared and set(declared) == set(capi.EXPORTS)
    for name in declared:
        assert hasattr(library, name), name


def test_struct_sizes(library):
    assert C.sizeof(capi.BinConfig) == 16 and C.sizeof(capi.DeintConfig) == 16 and C.sizeof(capi.Geometry) == 16
    assert C.sizeof(capi.BinStats) == 40
    assert library.sdv_version() == 100
    g = capi.Geometry(lines_per_field=294, lead_in=80)
    assert library.sdv_stc007_block_count(C.byref(g), 10) == 80 + 10 * 588


def test_no_cpu_fallback(library):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert library.sdv_create(C.byref(h), 0) == capi.SDV_ERR_CUDA
    with pytest.raises(capi.SdvError):
        capi.Handle(0)


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "sdvpcmdecoder_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("against the oracle", "") or f in ("sdv_common.cuh", "stc007_line.cuh", "__init__.py"), f


def test_plain_c_client_compiles_links_and_runs(library, tmp_path):
    """include/sdvpcm.h is valid C99, the layouts are the documented ones, and a C program linked against the library gets
    SDV_ERR_CUDA from sdv_create on a machine without a GPU (or passes the argument checks on one with)."""
    import subprocess
    exe = str(tmp_path / "cabi_client")
    src = os.path.join(ROOT, "tests", "cabi", "cabi_client.c")
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                    "-L", libdir, "-lsdvpcm_b200", "-Wl,-rpath," + libdir], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert ("no-gpu" in res.stdout) or ("gpu: ok" in res.stdout)
