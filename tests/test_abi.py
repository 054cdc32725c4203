"""CPU suite: the C-ABI library loads and exports every symbol include/sbod.h declares
(no compute calls here — there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(REPO, "include", "sbod.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sbod_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as G
    from shape_based_object_detection_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        G.build()
    return _lib


def test_header_symbols_are_exported(built):
    handle = ctypes.CDLL(built.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, missing
    assert handle.sbod_abi_version() == 2


def test_binding_table_matches_header(built):
    assert sorted(built.EXPORTS) == declared_symbols()


def test_validation_codes_without_a_gpu(built):
    lib = built.lib()
    assert lib.sbod_error_string(0) == b"ok"
    assert b"invalid" in lib.sbod_error_string(-1)
    # argument validation happens before any CUDA call
    assert lib.sbod_iou_matrix(None, -1, None, 4, 0, None, None) == -1
    assert lib.sbod_box_convert(None, None, 4, 7, None) == -1
    d = built.LossDesc()
    assert lib.sbod_loss_forward(ctypes.byref(d), None) == -1
    dd = built.DetectDesc()
    assert lib.sbod_detect(ctypes.byref(dd), None) == -1
    assert lib.sbod_nms_workspace_bytes(1000) >= 1000 * 16 * 8


def test_product_has_no_cpu_path(built):
    """The operators refuse CPU tensors instead of silently computing on the host."""
    import torch
    from shape_based_object_detection_b200 import metrics
    with pytest.raises(built.SbodError):
        metrics.find_jaccard_overlap(torch.zeros((2, 4)), torch.zeros((3, 4)))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(REPO, "shape_based_object_detection_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(root, f)
                assert "oracle." not in src and "box_pipeline" not in src, os.path.join(root, f)
