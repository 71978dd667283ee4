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
    # the kernel list it reports is the set of __global__ functions in the sources
    import glob, re
    src = os.path.join(os.path.dirname(os.path.abspath(_lib.__file__)), "csrc")
    kernels = set()
    for f in glob.glob(os.path.join(src, "*.cu")) + glob.glob(os.path.join(src, "*.cuh")):
        with open(f) as fh:
            kernels |= set(re.findall(r"__global__\s+void\s+(?:__launch_bounds__\([^)]*\)\s+)?(k_\w+)", fh.read()))
    assert set(info.split("kernels=")[1].split(",")) == kernels


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


def test_workspace_queries_need_no_device_pointers(lib_path):
    """Workspace sizing is host arithmetic: it must work before any input is bound."""
    import ctypes as C
    from miso_b200 import _lib
    lib = _lib.load()
    p = _lib.RpnParams()
    p.num_images, p.num_levels = 4, 5
    for l, g in enumerate((200, 100, 50, 25, 13)):
        p.feat_h[l] = p.feat_w[l] = g
        p.anchors_per_loc[l] = 3
        p.stride_h[l] = p.stride_w[l] = 800 // g
    p.pre_nms_top_n = p.post_nms_top_n = 1000
    assert lib.mb_rpn_workspace_bytes(C.byref(p)) > 4 * 159882 * 8
    p.pre_nms_top_n = 5000                      # outside the envelope -> 0, the host raises
    assert lib.mb_rpn_workspace_bytes(C.byref(p)) == 0
    d = _lib.DetParams()
    d.num_images, d.num_classes, d.max_props_per_image, d.detections_per_img = 4, 3, 1000, 300
    assert lib.mb_det_workspace_bytes(C.byref(d)) > 0
    assert lib.mb_nms_workspace_bytes(200000, 80) > 200000 * 60
