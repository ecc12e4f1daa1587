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


@pytest.mark.gpu
def test_c_client_decodes_golden_frames(library, tmp_path):
    """A C99 program (tests/cabi/cabi_decode_client.c), linked against the library, decodes tapes through the host-buffer entry
    points and compares with the reference's golden records / PCMSamplePair streams (tests/golden, written by make_golden.py from
    oracle/_ref) byte for byte: STC-007 (the golden pipeline tape), PCM-1 and PCM-16x0 (the damaged golden tapes: line records;
    PCM-1 also its sample stream)."""
    import subprocess
    import numpy as np
    from sdvpcmdecoder_b200 import synth
    from tests.test_pcm1_line import pcm1_cases
    from tests.test_pcm16x0_line import pcm16x0_cases
    exe = str(tmp_path / "cabi_decode_client")
    src = os.path.join(ROOT, "tests", "cabi", "cabi_decode_client.c")
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                    "-L", libdir, "-lsdvpcm_b200", "-Wl,-rpath," + libdir], check=True)
    gold = os.path.join(ROOT, "tests", "golden")

    def run(fmt, luma, recs, samples, flags):
        luma = np.ascontiguousarray(luma)
        paths = []
        for name, arr in (("luma", luma), ("recs", recs), ("smp", samples), ("fl", flags)):
            if arr is None:
                paths.append("-")
                continue
            pth = str(tmp_path / f"{fmt}_{name}.bin")
            np.ascontiguousarray(arr).tofile(pth)
            paths.append(pth)
        res = subprocess.run([exe, fmt, str(luma.shape[0]), str(luma.shape[1]), str(luma.shape[2])] + paths, capture_output=True, text=True)
        assert res.returncode == 0, res.stdout + res.stderr
        assert "equal the expected files" in res.stdout

    g = np.load(os.path.join(gold, "stc007_pipeline_pal.npz"))
    tape = synth.make_stc007(int(g["n_frames"]), seed=int(g["seed"]))
    smp = np.stack([g["l"], g["r"]], axis=1).astype(np.int16)
    fl = np.stack([g["flags_l"] & 7, g["flags_r"] & 7], axis=1).astype(np.uint8)
    run("stc007", tape["luma"], None, smp, fl)
    g1 = np.load(os.path.join(gold, "pcm1_lines.npz"))
    run("pcm1", pcm1_cases()["damaged"], g1["damaged_recs"], g1["damaged_samples"], np.zeros(0, np.uint8))
    gx = np.load(os.path.join(gold, "pcm16x0_lines.npz"))
    run("pcm16x0", pcm16x0_cases()["damaged"], gx["damaged_recs"], np.zeros(0, np.int16), np.zeros(0, np.uint8))
