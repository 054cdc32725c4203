// pair_iou.cuh — paired (elementwise) IoU / GIoU / DIoU / CIoU, forward and hand-derived backward.
// Semantics: operators/iou_utils.py:6-164 (bbox_overlaps_iou/giou/diou/ciou), equal-length inputs.
// Gradients follow torch autograd conventions: maximum/minimum split ties evenly, clamp passes
// the gradient on the closed interval, and CIoU's alpha / arctan / w_temp are constants
// (the reference computes them under torch.no_grad(), iou_utils.py:86-92).
#pragma once

#include "common.cuh"

namespace sbod {

struct PairGrad {
  float4 d1;  // d out / d b1 (x1,y1,x2,y2)
  float4 d2;  // d out / d b2
};

// d max(a,b) -> (wa, wb)
SBOD_DEVINL void max_w(float a, float b, float& wa, float& wb) {
  wa = a > b ? 1.f : (a == b ? 0.5f : 0.f);
  wb = 1.f - wa;
}
SBOD_DEVINL void min_w(float a, float b, float& wa, float& wb) {
  wa = a < b ? 1.f : (a == b ? 0.5f : 0.f);
  wb = 1.f - wa;
}

// Returns the clamped overlap value; if G != nullptr also the gradient of that value.
template <bool WITH_GRAD>
SBOD_DEVINL float pair_overlap(const float4 p, const float4 q, int kind, PairGrad* G) {
  const float w1 = p.z - p.x, h1 = p.w - p.y;
  const float w2 = q.z - q.x, h2 = q.w - q.y;
  const float area1 = w1 * h1, area2 = w2 * h2;
  const float ix1 = fmaxf(p.x, q.x), iy1 = fmaxf(p.y, q.y);
  const float ix2 = fminf(p.z, q.z), iy2 = fminf(p.w, q.w);
  const float iwr = ix2 - ix1, ihr = iy2 - iy1;
  const float iw = fmaxf(iwr, 0.f), ih = fmaxf(ihr, 0.f);
  const float inter = iw * ih;
  const float uni = area1 + area2 - inter;
  const float iou = inter / uni;

  float val, lo = -1.f, hi = 1.f;
  // pieces for the penalty terms
  float ox1 = 0, oy1 = 0, ox2 = 0, oy2 = 0, owr = 0, ohr = 0, ow = 0, oh = 0;
  float dx = 0, dy = 0, d2 = 0, c2 = 0, closure = 0, arct = 0, alpha = 0;
  constexpr float kPi = 3.14159265358979323846f;
  if (kind != SBOD_PAIR_IOU) {
    ox1 = fminf(p.x, q.x); oy1 = fminf(p.y, q.y);
    ox2 = fmaxf(p.z, q.z); oy2 = fmaxf(p.w, q.w);
    owr = ox2 - ox1; ohr = oy2 - oy1;
    ow = fmaxf(owr, 0.f); oh = fmaxf(ohr, 0.f);
  }
  if (kind == SBOD_PAIR_IOU) {
    val = iou; lo = 0.f;
  } else if (kind == SBOD_PAIR_GIOU) {
    closure = ow * oh;
    val = iou - (closure - uni) / closure;
  } else {
    dx = (q.z + q.x) / 2.f - (p.z + p.x) / 2.f;
    dy = (q.w + q.y) / 2.f - (p.w + p.y) / 2.f;
    d2 = dx * dx + dy * dy;
    c2 = ow * ow + oh * oh;
    if (kind == SBOD_PAIR_DIOU) {
      val = iou - d2 / c2;
    } else {  // CIOU
      arct = atanf(w2 / h2) - atanf(w1 / h1);
      const float v = (4.f / (kPi * kPi)) * arct * arct;
      const float S = 1.f - iou;
      alpha = v / (S + v);
      const float ar = (8.f / (kPi * kPi)) * arct * ((w1 - 2.f * w1) * h1);
      val = iou - (d2 / c2 + alpha * ar);
    }
  }
  const float out = (val != val) ? val : fminf(fmaxf(val, lo), hi);  // torch.clamp keeps NaN
  if (!WITH_GRAD) return out;

  // ---- reverse mode ----
  float g_val = (val >= lo && val <= hi) ? 1.f : 0.f;
  float g_iou = g_val;
  float g_uni = 0.f, g_ow = 0.f, g_oh = 0.f, g_dx = 0.f, g_dy = 0.f;
  float g_w1 = 0.f, g_h1 = 0.f;  // extra direct paths (CIoU aspect term)
  if (kind == SBOD_PAIR_GIOU) {
    // val = iou - (closure - uni)/closure = iou - 1 + uni/closure
    g_uni += g_val / closure;
    const float g_cl = -g_val * uni / (closure * closure);
    g_ow += g_cl * oh;
    g_oh += g_cl * ow;
  } else if (kind == SBOD_PAIR_DIOU || kind == SBOD_PAIR_CIOU) {
    const float g_d2 = -g_val / c2;
    const float g_c2 = g_val * d2 / (c2 * c2);
    g_dx += g_d2 * 2.f * dx;
    g_dy += g_d2 * 2.f * dy;
    g_ow += g_c2 * 2.f * ow;
    g_oh += g_c2 * 2.f * oh;
    if (kind == SBOD_PAIR_CIOU) {
      // ar = K*arct*((w1 - w_temp)*h1), w_temp constant = 2*w1
      const float K = (8.f / (kPi * kPi)) * arct;
      const float g_ar = -g_val * alpha;
      g_w1 += g_ar * K * h1;
      g_h1 += g_ar * K * (w1 - 2.f * w1);
    }
  }
  // iou = inter / uni
  float g_inter = g_iou / uni;
  g_uni += -g_iou * inter / (uni * uni);
  // uni = area1 + area2 - inter
  const float g_a1 = g_uni, g_a2 = g_uni;
  g_inter -= g_uni;
  // inter = iw*ih ; iw = clamp(iwr, min=0)
  const float g_iwr = (iwr >= 0.f) ? g_inter * ih : 0.f;
  const float g_ihr = (ihr >= 0.f) ? g_inter * iw : 0.f;
  const float g_owr = (owr >= 0.f) ? g_ow : 0.f;
  const float g_ohr = (ohr >= 0.f) ? g_oh : 0.f;

  float4 d1 = make_float4(0, 0, 0, 0), d2g = make_float4(0, 0, 0, 0);
  float wa, wb;
  // ix1 = max(p.x,q.x) (iwr = ix2 - ix1)
  max_w(p.x, q.x, wa, wb); d1.x += -g_iwr * wa; d2g.x += -g_iwr * wb;
  max_w(p.y, q.y, wa, wb); d1.y += -g_ihr * wa; d2g.y += -g_ihr * wb;
  min_w(p.z, q.z, wa, wb); d1.z += g_iwr * wa;  d2g.z += g_iwr * wb;
  min_w(p.w, q.w, wa, wb); d1.w += g_ihr * wa;  d2g.w += g_ihr * wb;
  if (kind != SBOD_PAIR_IOU) {
    min_w(p.x, q.x, wa, wb); d1.x += -g_owr * wa; d2g.x += -g_owr * wb;
    min_w(p.y, q.y, wa, wb); d1.y += -g_ohr * wa; d2g.y += -g_ohr * wb;
    max_w(p.z, q.z, wa, wb); d1.z += g_owr * wa;  d2g.z += g_owr * wb;
    max_w(p.w, q.w, wa, wb); d1.w += g_ohr * wa;  d2g.w += g_ohr * wb;
    // dx = (q.z+q.x)/2 - (p.z+p.x)/2
    d1.x += -0.5f * g_dx; d1.z += -0.5f * g_dx; d2g.x += 0.5f * g_dx; d2g.z += 0.5f * g_dx;
    d1.y += -0.5f * g_dy; d1.w += -0.5f * g_dy; d2g.y += 0.5f * g_dy; d2g.w += 0.5f * g_dy;
  }
  // area1 = w1*h1
  const float t_w1 = g_a1 * h1 + g_w1, t_h1 = g_a1 * w1 + g_h1;
  d1.x -= t_w1; d1.z += t_w1; d1.y -= t_h1; d1.w += t_h1;
  const float t_w2 = g_a2 * h2, t_h2 = g_a2 * w2;
  d2g.x -= t_w2; d2g.z += t_w2; d2g.y -= t_h2; d2g.w += t_h2;
  G->d1 = d1;
  G->d2 = d2g;
  return out;
}

}  // namespace sbod
