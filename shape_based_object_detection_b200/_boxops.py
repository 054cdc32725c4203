"""Box converters / gcxgcy codec as differentiable operators (CUDA forward and backward kernels).

In the reference these are plain torch expressions (dataset/transforms.py:26-83, operators/iou_utils.py:
167-189, 324-368), so gradients flow through them - IouLoss(pred_mode='Center') decodes its predictions
first (operators/Loss.py:176-178). The priors are constants on the path: a prior tensor that requires
grad is rejected instead of being silently detached.
"""
import torch

from . import _lib as L

# backward op codes of sbod_box_op_bwd
BWD_XY_TO_CXCY, BWD_CXCY_TO_XY, BWD_ENC_T, BWD_ENC_U, BWD_DEC_T, BWD_DEC_U = range(6)


class _BoxOp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pri, fwd, bwd_op, v0, v1):
        a = L.f32c(x.detach())
        p = L.f32c(pri.detach()) if pri is not None else None
        out = torch.empty_like(a)
        with torch.cuda.device(a.device):
            fwd(a, p, out)
        ctx.save_for_backward(a, p if p is not None else a)
        ctx.meta = (bwd_op, v0, v1, p is not None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        a, p = ctx.saved_tensors
        bwd_op, v0, v1, has_pri = ctx.meta
        go = L.f32c(grad_out)
        gi = torch.empty_like(a)
        with torch.cuda.device(a.device):
            L.check(L.lib().sbod_box_op_bwd(bwd_op, L.ptr(a), L.ptr(p) if has_pri else None, L.ptr(go), L.ptr(gi),
                                            a.size(0), v0, v1, L.stream_ptr()))
        return gi, None, None, None, None, None


def _apply(x, pri, fwd, bwd_op, v0=0.0, v1=0.0):
    L.need_cuda(x, pri)
    L.device_of(x, pri)
    if pri is not None and pri.requires_grad:
        raise L.SbodError("gradients with respect to the priors are not implemented (they are constants on the path)")
    return _BoxOp.apply(x, pri, fwd, bwd_op, float(v0), float(v1))


def convert(x, op):
    def fwd(a, p, out):
        L.check(L.lib().sbod_box_convert(L.ptr(a), L.ptr(out), a.size(0), op, L.stream_ptr()))
    return _apply(x, None, fwd, BWD_XY_TO_CXCY if op == L.BOX_XY_TO_CXCY else BWD_CXCY_TO_XY)


def encode(x, pri, flavour, v0, v1):
    def fwd(a, p, out):
        L.check(L.lib().sbod_box_encode(L.ptr(a), L.ptr(p), L.ptr(out), a.size(0), flavour, float(v0), float(v1),
                                        L.stream_ptr()))
    return _apply(x, pri, fwd, BWD_ENC_T if flavour == L.CODEC_TRANSFORMS else BWD_ENC_U, v0, v1)


def decode(x, pri, flavour, v0, v1):
    def fwd(a, p, out):
        L.check(L.lib().sbod_box_decode(L.ptr(a), L.ptr(p), L.ptr(out), a.size(0), flavour, float(v0), float(v1),
                                        L.stream_ptr()))
    return _apply(x, pri, fwd, BWD_DEC_T if flavour == L.CODEC_TRANSFORMS else BWD_DEC_U, v0, v1)
