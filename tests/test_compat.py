"""CPU suite: compat.install() rebinds the reference's hot-path names (build container only — the
reference tree does not exist on the GPU box)."""
import os
import sys

import pytest

REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_install_patches_reference_modules():
    import shape_based_object_detection_b200 as S
    patched = S.install(REF)
    try:
        assert "models.SSD512.MultiBoxLoss512" in patched
        assert "models.utils.detect" in patched
        assert "operators.iou_utils.jaccard" in patched
        assert "detect_scripts.detect_tools.detect_refine" in patched
        assert "metrics.calculate_mAP" in patched
        import models  # noqa: F401
        ref_ssd512 = sys.modules["models.SSD512"]  # (models.SSD512 the attribute is the class, models/__init__.py:1)
        from shape_based_object_detection_b200.models.SSD512 import MultiBoxLoss512
        assert ref_ssd512.MultiBoxLoss512 is MultiBoxLoss512
        import models as ref_models
        assert ref_models.MultiBoxLoss512 is MultiBoxLoss512          # what model_entry() hands to the drivers
        import metrics as ref_metrics
        assert ref_metrics.find_jaccard_overlap.__module__ == "metrics"  # CPU data-loader path left alone
    finally:
        for name in [m for m in sys.modules if m.split(".")[0] in ("models", "operators", "detect_scripts", "dataset",
                                                                   "metrics")]:
            del sys.modules[name]
        if REF in sys.path:
            sys.path.remove(REF)
