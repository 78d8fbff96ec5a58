"""CPU suite, part 2: the C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol
include/fi_b200.h declares (and nothing in the product package touches the oracle)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "fi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:unsigned long long|size_t|int|void|char)\s*\*?\s*(\w+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_header_declares_the_reference_launchers():
    names = _declared_symbols()
    for n in ("CropAndResizeLaucher", "CropAndResizeBackpropImageLaucher", "ROIPoolForwardLaucher", "ROIPoolBackwardLaucher", "_nms"):
        assert n in names          # crop_and_resize_kernel.h:8-18, roi_pooling_kernel.h:8-18, nms_kernel.h:11-12
    assert len(names) >= 20


def test_library_builds_loads_and_exports_every_declared_symbol():
    from feature_intertwiner_b200 import _lib, build
    path = build.build_library()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    for name in _declared_symbols():
        assert hasattr(handle, name), "libfi_b200.so does not export %s" % name
    assert set(_lib.SIGNATURES) == set(_declared_symbols())
    L = _lib.lib()
    assert L.fi_abi_version() == 1 and L.fi_last_error() is not None


def test_library_is_sm_100a_and_uses_vector_reductions():
    from feature_intertwiner_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.build_library()], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
    sass = subprocess.run(["cuobjdump", "-sass", build.build_library()], capture_output=True, text=True).stdout
    assert "REDG.E.ADD.F32x4" in sass          # 128-bit vector reduction in the NHWC backward
    assert "SYNCS.ARRIVE.TRANS64" in sass      # mbarrier expect_tx
    assert "UBLKCP.S.G" in sass                # cp.async.bulk global -> shared: gradient rows of the NHWC backward (roi_align_bwd_pix.cu)


def test_option_table_matches_the_header():
    """Every FI_OPT_* of the header is bound by name in _lib.OPTIONS (or is the count), fi_set_option accepts 0 for each, returns
    the previous value, and refuses values past the documented range -- no GPU needed; and the packed-lerp forward is in the SASS
    un-fused: FFMA2 with the opaque -0.0 addend between two FADD2 (a contraction of the reference's `top + (bottom - top) * w`
    would change bits)."""
    from feature_intertwiner_b200 import _lib, build
    src = open(os.path.join(ROOT, "include", "fi_b200.h")).read()
    keys = {name: int(v) for name, v in re.findall(r"#define\s+(FI_OPT_\w+)\s+(\d+)", src)}
    count = keys.pop("FI_OPT_COUNT")
    assert sorted(keys.values()) == list(range(count))
    assert sorted(_lib.OPTIONS.values()) == sorted(keys.values())
    assert {"FI_OPT_" + k.upper() for k in _lib.OPTIONS} == set(keys)
    L = _lib.lib()
    for name, key in _lib.OPTIONS.items():
        old = L.fi_get_option(key)
        assert L.fi_set_option(key, 0) == old
        assert L.fi_set_option(key, 99) == -1 and L.fi_get_option(key) == 0
        assert L.fi_set_option(key, old) == 0
    assert L.fi_set_option(count, 0) == -1
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN2fi30crop_fwd_nhwc_sets_lean_kernelILi2ELi2ELi8ELi2ELb1EEEvNS_7FwdSetsENS_7FwdPlanEy",
                           build.build_library()], capture_output=True, text=True).stdout
    assert sass.count("FADD2") >= 48 and sass.count("FFMA2") >= 24 and "FMUL2" not in sass, "packed lerps of the lean forward changed shape"
    assert "ATOMG.E.ADD" in sass               # the ticket draw


def test_argument_validation_needs_no_gpu():
    from feature_intertwiner_b200 import _lib
    L = _lib.lib()
    assert L.fi_sinkhorn(None, None, 1, 300, 1, 1.0, 5, None, None, None, None) == -1     # N > 256
    assert b"N in [1,256]" in L.fi_last_error()
    assert L.fi_split_levels(None, 70000, None, None, None, None, None, None) == -1
    assert L.fi_crop_and_resize_forward(None, 0, None, None, None, 4, 1, 8, 8, 7, 7, 16, 0.0, None, 0, None) == -1
    assert L.fi_last_status() == -1


def test_product_refuses_cpu_tensors_and_never_imports_the_oracle():
    import torch
    import feature_intertwiner_b200 as fi
    with pytest.raises(fi.FiError):
        fi.crop_and_resize(torch.randn(1, 4, 8, 8), torch.zeros(1, 4), torch.zeros(1, dtype=torch.int32), 7, 7)
    with pytest.raises(fi.FiError):
        fi.sinkhorn_loss(torch.rand(1, 8, 1), torch.rand(1, 8, 1))
    pkg = os.path.join(ROOT, "feature_intertwiner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "_ref/" not in text, f
