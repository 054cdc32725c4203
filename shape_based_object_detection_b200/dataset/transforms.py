"""Box format converters of dataset/transforms.py:26-83, as CUDA kernels (differentiable with respect to
the boxes, like the reference's torch expressions)."""
from .. import _boxops as B
from .. import _lib as L


def xy_to_cxcy(xy):
    """(x_min, y_min, x_max, y_max) -> (c_x, c_y, w, h); transforms.py:26-34."""
    return B.convert(xy, L.BOX_XY_TO_CXCY)


def cxcy_to_xy(cxcy):
    """(c_x, c_y, w, h) -> (x_min, y_min, x_max, y_max); transforms.py:37-45."""
    return B.convert(cxcy, L.BOX_CXCY_TO_XY)


def cxcy_to_gcxgcy(cxcy, priors_cxcy):
    """Encode centre-size boxes w.r.t. priors (variances 10 and 5); transforms.py:48-66."""
    return B.encode(cxcy, priors_cxcy, L.CODEC_TRANSFORMS, 0.1, 0.2)


def gcxgcy_to_cxcy(gcxgcy, priors_cxcy):
    """Decode model offsets into centre-size boxes; transforms.py:69-83."""
    return B.decode(gcxgcy, priors_cxcy, L.CODEC_TRANSFORMS, 0.1, 0.2)
