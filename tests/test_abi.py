"""CPU tier: the C-ABI library loads and exports exactly what include/psnw_klt.h declares."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT
from mcmtt_opticalflow_b200 import klt

HEADER = os.path.join(ROOT, "include", "psnw_klt.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"PSNW_API[^;(]*?\b(psnw_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_three_hot_path_exports():
    syms = declared_symbols()
    for s in ("psnw_klt_detect", "psnw_klt_track", "psnw_box_update", "psnw_push_frame", "psnw_create", "psnw_destroy"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    lib = klt.load()
    for s in declared_symbols():
        assert hasattr(lib, s), "libpsnw_klt.so does not export %s" % s
    assert sorted(klt.EXPORTS) == declared_symbols()


def test_no_undeclared_exports():
    out = subprocess.check_output(["nm", "-D", "--defined-only", klt.LIB_PATH]).decode()
    exported = sorted(set(re.findall(r" T (\w+)", out)))
    assert [s for s in exported if not s.startswith("psnw_")] == []
    assert exported == declared_symbols()


def test_params_layout_matches_header():
    # int x9, double, float x2, int x2, pointer, int  (natural alignment)
    assert C.sizeof(klt.Params) == 80
    p = klt.default_params()
    assert p.apron == 0
    assert (p.levels, p.ring_slots, p.win_w, p.win_h, p.max_iters) == (4, 4, 21, 21, 30)
    assert p.epsilon == 0.01 and abs(p.min_eig_thresh - 1e-4) < 1e-10 and p.fb_thresh == 1.0
    assert C.sizeof(klt.TrackJob) == 80


def test_python_constants_match_header():
    """The ctypes stub restates the header's flags and error codes: compile a tiny C program that prints the header's values."""
    import tempfile
    names = ["PSNW_MEM_HOST", "PSNW_MEM_PINNED", "PSNW_MEM_DEVICE", "PSNW_MEM_OVERLAP", "PSNW_PUSH_EIG",
             "PSNW_OK", "PSNW_EARG", "PSNW_ECUDA", "PSNW_ENOMEM", "PSNW_ESTATE", "(int)sizeof(psnw_params)",
             "(int)sizeof(psnw_track_job)", "(int)sizeof(psnw_detect_job)"]
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "c.c")
        open(src, "w").write('#include <stdio.h>\n#include "psnw_klt.h"\nint main(void){%s return 0;}\n'
                             % "".join('printf("%%d\\n", (int)(%s));' % n for n in names))
        exe = os.path.join(d, "c")
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.dirname(HEADER), src, "-o", exe])
        vals = [int(v) for v in subprocess.check_output([exe]).decode().split()]
    want = [klt.MEM_HOST, klt.MEM_PINNED, klt.MEM_DEVICE, klt.MEM_OVERLAP, klt.PUSH_EIG, 0, klt.EARG, klt.ECUDA, klt.ENOMEM,
            klt.ESTATE, C.sizeof(klt.Params), C.sizeof(klt.TrackJob), C.sizeof(klt.DetectJob)]
    assert vals == want


def test_header_compiles_as_c():
    subprocess.check_call(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", HEADER])


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(klt.PsnwError) as e:
        klt.KLTContext(768, 576)
    assert e.value.code == klt.ECUDA


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mcmtt_opticalflow_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert "psnw_oracle" not in text and "libpsnw_oracle" not in text and "cv2_ref" not in text, f
                if f.endswith(".py") and f != "synth.py":
                    assert "import cv2" not in text, f
