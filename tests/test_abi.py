"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol that
include/misob200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "misob200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    from miso_b200 import build
    return build.build()


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    for must in ("mb_nms", "mb_multiscale_roi_align", "mb_rpn_proposals", "mb_det_postprocess", "mb_crop_plan",
                 "mb_crop_gather", "mb_box_decode", "mb_box_convert"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in header_symbols():
        assert hasattr(lib, name), f"{name} declared in misob200.h but not exported"


def test_binding_table_matches_header(lib_path):
    from miso_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()
    lib = _lib.load()
    assert lib.mb_abi_version() == 1
    info = lib.mb_build_info().decode()
    assert "sm_100a" in info


def test_struct_sizes_match_c_layout(lib_path):
    """ctypes mirrors of the parameter structs must have the C compiler's layout."""
    import subprocess, tempfile, textwrap
    from miso_b200 import _lib
    src = textwrap.dedent("""
        #include <stdio.h>
        #include "misob200.h"
        int main(void) {
            printf("%zu %zu %zu %zu\\n", sizeof(mb_roi_align_params), sizeof(mb_rpn_params),
                   sizeof(mb_det_params), sizeof(mb_crop_params));
            return 0;
        }""")
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, c])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(_lib.RoiAlignParams), ctypes.sizeof(_lib.RpnParams),
                     ctypes.sizeof(_lib.DetParams), ctypes.sizeof(_lib.CropParams)]


def test_cpu_tensor_raises_instead_of_falling_back():
    import torch
    from miso_b200 import MisoB200Error, ops
    with pytest.raises(MisoB200Error):
        ops.nms(torch.zeros(3, 4), torch.zeros(3), 0.5)
    with pytest.raises(MisoB200Error):
        ops.roi_align(torch.zeros(1, 1, 4, 4), torch.zeros(1, 5), 2)
    with pytest.raises(ValueError):
        ops.box_convert(torch.zeros(1, 4), "xyxy", "bogus")
