"""Box format converters of dataset/transforms.py:26-83, as CUDA kernels."""
import torch

from .. import _lib as L


def _unary(t, op):
    L.need_cuda(t)
    x = L.f32c(t)
    out = torch.empty_like(x)
    L.check(L.lib().sbod_box_convert(L.ptr(x), L.ptr(out), x.size(0), op, L.stream_ptr()))
    return out


def xy_to_cxcy(xy):
    """(x_min, y_min, x_max, y_max) -> (c_x, c_y, w, h); transforms.py:26-34."""
    return _unary(xy, L.BOX_XY_TO_CXCY)


def cxcy_to_xy(cxcy):
    """(c_x, c_y, w, h) -> (x_min, y_min, x_max, y_max); transforms.py:37-45."""
    return _unary(cxcy, L.BOX_CXCY_TO_XY)


def cxcy_to_gcxgcy(cxcy, priors_cxcy):
    """Encode centre-size boxes w.r.t. priors (variances 10 and 5); transforms.py:48-66."""
    L.need_cuda(cxcy, priors_cxcy)
    a, p = L.f32c(cxcy), L.f32c(priors_cxcy)
    out = torch.empty_like(a)
    L.check(L.lib().sbod_box_encode(L.ptr(a), L.ptr(p), L.ptr(out), a.size(0), L.CODEC_TRANSFORMS, 0.1, 0.2,
                                    L.stream_ptr()))
    return out


def gcxgcy_to_cxcy(gcxgcy, priors_cxcy):
    """Decode model offsets into centre-size boxes; transforms.py:69-83."""
    L.need_cuda(gcxgcy, priors_cxcy)
    a, p = L.f32c(gcxgcy), L.f32c(priors_cxcy)
    out = torch.empty_like(a)
    L.check(L.lib().sbod_box_decode(L.ptr(a), L.ptr(p), L.ptr(out), a.size(0), L.CODEC_TRANSFORMS, 0.1, 0.2,
                                    L.stream_ptr()))
    return out
