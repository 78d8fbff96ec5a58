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
