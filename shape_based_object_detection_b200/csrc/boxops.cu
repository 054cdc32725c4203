// boxops.cu — the stand-alone operators of the path (sm_100a): dense IoU matrix, box format
// converters and the gcxgcy codec, paired IoU family (+backward), row losses of operators/Loss.py,
// stand-alone greedy NMS (bitmask + single-CTA ballot reduction) and the SSD-pytorch style match.
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "pair_iou.cuh"

namespace sbod {

static inline int grid_for(size_t n, int block) {
  size_t g = (n + block - 1) / block;
  const size_t cap = size_t(sm_count()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return int(g);
}

// ---- dense IoU matrix: a [A,4] staged in smem by chunks, b streamed with 128-bit loads --------
constexpr int kIouChunk = 64;
__global__ void __launch_bounds__(256) iou_matrix_kernel(const float4* __restrict__ a, int A,
                                                         const float4* __restrict__ b, int B,
                                                         int mode, float* __restrict__ out) {
  __shared__ float4 s_a[kIouChunk];
  __shared__ float s_area[kIouChunk];
  __shared__ uint8_t s_zero[kIouChunk];
  const int a0 = blockIdx.y * kIouChunk;
  const int na = min(kIouChunk, A - a0);
  for (int i = threadIdx.x; i < na; i += blockDim.x) {
    const float4 g = a[a0 + i];
    s_a[i] = g;
    s_area[i] = box_area_rn(g);
    s_zero[i] = gt_is_zero(g) ? 1 : 0;
  }
  __syncthreads();
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < B; j += gridDim.x * blockDim.x) {
    const float4 bb = ld_stream_f4(b + j);
    const float ab = box_area_rn(bb);
    const bool bz = anchor_is_zero(bb);
    for (int i = 0; i < na; ++i) {
      float v;
      if (mode == SBOD_IOU_METRICS) {
        v = iou_metrics_rn(s_a[i], s_area[i], bb, ab);
        if (s_zero[i]) v = 0.f;   // metrics.py:249
        if (bz) v = -1.f;         // metrics.py:250 (applied last)
      } else if (mode == SBOD_IOU_JACCARD) {
        v = iou_plain_rn(s_a[i], s_area[i], bb, ab);
      } else {
        v = inter_rn(s_a[i], bb);
      }
      out[size_t(a0 + i) * B + j] = v;  // coalesced over j
    }
  }
}

// ---- converters / codec ------------------------------------------------------------------------
__global__ void box_convert_kernel(const float4* __restrict__ in, float4* __restrict__ out, int n,
                                   int op) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 v = in[i];
    float4 r;
    if (op == SBOD_BOX_XY_TO_CXCY) {  // transforms.py:26-34
      r = make_float4(__fdiv_rn(__fadd_rn(v.z, v.x), 2.f), __fdiv_rn(__fadd_rn(v.w, v.y), 2.f),
                      __fsub_rn(v.z, v.x), __fsub_rn(v.w, v.y));
    } else {  // transforms.py:37-45, iou_utils.py:167-177
      const float hw = __fdiv_rn(v.z, 2.f), hh = __fdiv_rn(v.w, 2.f);
      r = make_float4(__fsub_rn(v.x, hw), __fsub_rn(v.y, hh), __fadd_rn(v.x, hw), __fadd_rn(v.y, hh));
    }
    out[i] = r;
  }
}

__global__ void box_encode_kernel(const float4* __restrict__ boxes, const float4* __restrict__ pri,
                                  float4* __restrict__ out, int n, int flavour, float v0, float v1) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 b = boxes[i], p = pri[i];
    float4 r;
    if (flavour == SBOD_CODEC_TRANSFORMS) {  // input cxcy, transforms.py:48-66
      r.x = __fdiv_rn(__fsub_rn(b.x, p.x), __fdiv_rn(p.z, 10.f));
      r.y = __fdiv_rn(__fsub_rn(b.y, p.y), __fdiv_rn(p.w, 10.f));
      r.z = __fmul_rn(logf(__fdiv_rn(b.z, p.z)), 5.f);
      r.w = __fmul_rn(logf(__fdiv_rn(b.w, p.w)), 5.f);
    } else {  // input xyxy "matched", iou_utils.py:324-345
      const float gx = __fsub_rn(__fdiv_rn(__fadd_rn(b.x, b.z), 2.f), p.x);
      const float gy = __fsub_rn(__fdiv_rn(__fadd_rn(b.y, b.w), 2.f), p.y);
      r.x = __fdiv_rn(gx, __fmul_rn(v0, p.z));
      r.y = __fdiv_rn(gy, __fmul_rn(v0, p.w));
      r.z = __fdiv_rn(logf(__fdiv_rn(__fsub_rn(b.z, b.x), p.z)), v1);
      r.w = __fdiv_rn(logf(__fdiv_rn(__fsub_rn(b.w, b.y), p.w)), v1);
    }
    out[i] = r;
  }
}

__global__ void box_decode_kernel(const float4* __restrict__ locs, const float4* __restrict__ pri,
                                  float4* __restrict__ out, int n, int flavour, float v0, float v1) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 l = locs[i], p = pri[i];
    float4 r;
    if (flavour == SBOD_CODEC_TRANSFORMS) {  // -> cxcy, transforms.py:69-83
      r.x = __fadd_rn(__fdiv_rn(__fmul_rn(l.x, p.z), 10.f), p.x);
      r.y = __fadd_rn(__fdiv_rn(__fmul_rn(l.y, p.w), 10.f), p.y);
      r.z = __fmul_rn(expf(__fdiv_rn(l.z, 5.f)), p.z);
      r.w = __fmul_rn(expf(__fdiv_rn(l.w, 5.f)), p.w);
    } else {  // -> xyxy, iou_utils.py:349-368
      const float cx = __fadd_rn(p.x, __fmul_rn(__fmul_rn(l.x, v0), p.z));
      const float cy = __fadd_rn(p.y, __fmul_rn(__fmul_rn(l.y, v0), p.w));
      const float w = __fmul_rn(p.z, expf(__fmul_rn(l.z, v1)));
      const float h = __fmul_rn(p.w, expf(__fmul_rn(l.w, v1)));
      const float x1 = __fsub_rn(cx, __fdiv_rn(w, 2.f)), y1 = __fsub_rn(cy, __fdiv_rn(h, 2.f));
      r = make_float4(x1, y1, __fadd_rn(w, x1), __fadd_rn(h, y1));
    }
    out[i] = r;
  }
}

// Backward of the converters / codec with respect to their FIRST argument (the boxes or offsets; the
// priors are constants on the path). op: 0 xy_to_cxcy, 1 cxcy_to_xy, 2 encode (transforms), 3 encode
// (iou_utils), 4 decode (transforms), 5 decode (iou_utils). `in` is the forward input.
__global__ void box_op_bwd_kernel(int op, const float4* __restrict__ in, const float4* __restrict__ pri,
                                  const float4* __restrict__ go, float4* __restrict__ gi, int n, float v0,
                                  float v1) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 g = go[i];
    float4 r;
    if (op == 0) {  // out = ((x2+x1)/2, (y2+y1)/2, x2-x1, y2-y1)
      r = make_float4(0.5f * g.x - g.z, 0.5f * g.y - g.w, 0.5f * g.x + g.z, 0.5f * g.y + g.w);
    } else if (op == 1) {  // out = (cx - w/2, cy - h/2, cx + w/2, cy + h/2)
      r = make_float4(g.x + g.z, g.y + g.w, 0.5f * (g.z - g.x), 0.5f * (g.w - g.y));
    } else {
      const float4 a = in[i], p = pri[i];
      if (op == 2) {  // ((c - pc) / (pwh / 10), log(wh / pwh) * 5)
        r = make_float4(g.x * 10.f / p.z, g.y * 10.f / p.w, g.z * 5.f / a.z, g.w * 5.f / a.w);
      } else if (op == 3) {  // (((x1+x2)/2 - pc) / (v0 pwh), log((x2-x1) / pwh) / v1)
        const float cx = g.x / (2.f * v0 * p.z), cy = g.y / (2.f * v0 * p.w);
        const float sw = g.z / (v1 * (a.z - a.x)), sh = g.w / (v1 * (a.w - a.y));
        r = make_float4(cx - sw, cy - sh, cx + sw, cy + sh);
      } else if (op == 4) {  // (l_c * pwh / 10 + pc, exp(l_wh / 5) * pwh)
        r = make_float4(g.x * p.z / 10.f, g.y * p.w / 10.f, g.z * expf(a.z / 5.f) * p.z / 5.f,
                        g.w * expf(a.w / 5.f) * p.w / 5.f);
      } else {  // c = pc + l_c v0 pwh; wh = pwh exp(l_wh v1); out = (c - wh/2, c + wh/2)
        const float w = p.z * expf(a.z * v1), h = p.w * expf(a.w * v1);
        r = make_float4((g.x + g.z) * v0 * p.z, (g.y + g.w) * v0 * p.w, 0.5f * (g.z - g.x) * w * v1,
                        0.5f * (g.w - g.y) * h * v1);
      }
    }
    gi[i] = r;
  }
}

__global__ void offset2bbox_kernel(const float4* __restrict__ arm, const float4* __restrict__ odm,
                                   const float4* __restrict__ pri, float4* __restrict__ out, int N,
                                   int P) {
  const size_t total = size_t(N) * P;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total;
       i += size_t(gridDim.x) * blockDim.x) {
    const float4 p = pri[i % P], a = arm[i], o = odm[i];
    // init = gcxgcy_to_cxcy(arm, priors); out = cxcy_to_xy(gcxgcy_to_cxcy(odm, init))  RefineDet512.py:650-651
    const float icx = a.x * p.z / 10.f + p.x, icy = a.y * p.w / 10.f + p.y;
    const float iw = expf(a.z / 5.f) * p.z, ih = expf(a.w / 5.f) * p.w;
    const float cx = o.x * iw / 10.f + icx, cy = o.y * ih / 10.f + icy;
    const float w = expf(o.z / 5.f) * iw, h = expf(o.w / 5.f) * ih;
    out[i] = make_float4(cx - w / 2.f, cy - h / 2.f, cx + w / 2.f, cy + h / 2.f);
  }
}

// ---- prior / anchor / location tables generated on the device ---------------------------------------
// models/SSD300.py:389-443, SSD512.py:417-474, RetinaNet.py:261-300, RefineDet512.py:655-695 (python triple
// loops: level, row i, column j, box shape) and FCOSDet.py:235-251 (cell centres). One thread per prior; the
// arithmetic is the reference's: python floats (float64), rounded once to fp32, then clamp_(0, 1).
constexpr int kGridMaxLevels = 8;
constexpr int kGridMaxShapes = 64;
struct GridSpec {
  int n_levels;
  int rows[kGridMaxLevels], cols[kGridMaxLevels], n_shapes[kGridMaxLevels], shape0[kGridMaxLevels];
  long long offset[kGridMaxLevels + 1];  // first output row of each level
  double mul_x[kGridMaxLevels], div_x[kGridMaxLevels], mul_y[kGridMaxLevels], div_y[kGridMaxLevels];
  double shape_w[kGridMaxShapes], shape_h[kGridMaxShapes];
  int clamp01, centres_only;
};

__global__ void prior_grid_kernel(const GridSpec g, float* __restrict__ out, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int lv = 0;
    while (lv + 1 < g.n_levels && idx >= g.offset[lv + 1]) ++lv;
    const long long local = idx - g.offset[lv];
    const int ns = g.centres_only ? 1 : g.n_shapes[lv];
    const long long cell = local / ns;
    const int sidx = int(local - cell * ns);
    const int i = int(cell / g.cols[lv]), j = int(cell - (long long)i * g.cols[lv]);
    float cx = float(__ddiv_rn(__dmul_rn(double(j) + 0.5, g.mul_x[lv]), g.div_x[lv]));
    float cy = float(__ddiv_rn(__dmul_rn(double(i) + 0.5, g.mul_y[lv]), g.div_y[lv]));
    if (g.centres_only) {
      reinterpret_cast<float2*>(out)[idx] = make_float2(cx, cy);
    } else {
      float w = float(g.shape_w[g.shape0[lv] + sidx]), h = float(g.shape_h[g.shape0[lv] + sidx]);
      if (g.clamp01) {
        cx = fminf(fmaxf(cx, 0.f), 1.f); cy = fminf(fmaxf(cy, 0.f), 1.f);
        w = fminf(fmaxf(w, 0.f), 1.f); h = fminf(fmaxf(h, 0.f), 1.f);
      }
      reinterpret_cast<float4*>(out)[idx] = make_float4(cx, cy, w, h);
    }
  }
}

// ---- RefineDet: ARM easy-negative mask  softmax(arm_scores)[:, 1] < theta  (RefineDet512.py:894-895)
__global__ void arm_easy_negative_kernel(const float2* __restrict__ arm_scores, size_t n, float theta,
                                         uint8_t* __restrict__ out) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const float2 x = arm_scores[i];
    const float m = fmaxf(x.x, x.y);
    const float e0 = expf(x.x - m), e1 = expf(x.y - m);
    out[i] = (e1 / (e0 + e1) < theta) ? 1 : 0;
  }
}

// ---- paired IoU family -------------------------------------------------------------------------
__global__ void pair_iou_fwd_kernel(const float4* __restrict__ b1, const float4* __restrict__ b2,
                                    int M, int kind, float* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x)
    out[i] = pair_overlap<false>(b1[i], b2[i], kind, nullptr);
}
__global__ void pair_iou_bwd_kernel(const float4* __restrict__ b1, const float4* __restrict__ b2,
                                    const float* __restrict__ go, int M, int kind,
                                    float4* __restrict__ g1, float4* __restrict__ g2) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    PairGrad pg;
    pair_overlap<true>(b1[i], b2[i], kind, &pg);
    const float g = go[i];
    if (g1) g1[i] = make_float4(pg.d1.x * g, pg.d1.y * g, pg.d1.z * g, pg.d1.w * g);
    if (g2) g2[i] = make_float4(pg.d2.x * g, pg.d2.y * g, pg.d2.z * g, pg.d2.w * g);
  }
}

// ---- row losses --------------------------------------------------------------------------------
__global__ void smooth_l1_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                 int n, float beta, float* __restrict__ out,
                                 float* __restrict__ grad) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float d = pred[i] - tgt[i];
    const float x = fabsf(d);
    const bool lin = x >= beta;  // Loss.py:214-217
    out[i] = lin ? x - 0.5f * beta : 0.5f * x * x / beta;
    if (grad) grad[i] = lin ? (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) : d / beta;
  }
}

// one warp per row
__global__ void softmax_focal_kernel(const float* __restrict__ x, const int64_t* __restrict__ tgt,
                                     int M, int C, float afg, float abg, float gamma,
                                     float* __restrict__ row_out, float* __restrict__ grad) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < M; r += warps) {
    const float* row = x + size_t(r) * C;
    float mx = -__int_as_float(0x7f800000);
    for (int k = lane; k < C; k += 32) mx = fmaxf(mx, row[k]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int k = lane; k < C; k += 32) s += __expf(row[k] - mx);
    s = warp_sum(s);
    const float lg = logf(s);
    int t = int(tgt[r]);
    t = min(max(t, 0), C - 1);
    const float ce = (mx - row[t]) + lg;
    const float pt = __expf(-ce);
    const bool fg = t != 0;
    const float A = fg ? afg : abg;
    const float w = fg ? 1.f - pt : pt;  // Loss.py:32: the background weight is p_0 itself
    const float L = A * powf(w, gamma) * ce;
    if (lane == 0) row_out[r] = L;
    if (grad) {
      const float dw = fg ? pt : -pt;  // d w / d ce
      const float dL = A * (powf(w, gamma) + ce * gamma * powf(w, gamma - 1.f) * dw);
      const float lse = mx + lg;
      for (int k = lane; k < C; k += 32)
        grad[size_t(r) * C + k] = dL * (__expf(row[k] - lse) - (k == t ? 1.f : 0.f));
    }
  }
}

__global__ void sigmoid_focal_kernel(const float* __restrict__ x, const int64_t* __restrict__ tgt,
                                     int M, int C, float alpha, float gamma,
                                     float* __restrict__ row_out, float* __restrict__ grad) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < M; r += warps) {
    const float* row = x + size_t(r) * C;
    const int64_t t = tgt[r];
    float acc = 0.f;
    for (int k = lane; k < C; k += 32) {
      float g = 0.f;
      if (k >= 1) {  // out[:, 1:] against class ids 1..C-1, Loss.py:51-58
        const float z = row[k];
        const float p = 1.f / (1.f + __expf(-z));
        if (t == k) {  // -alpha * (1-p)^g * log p
          const float q1 = 1.f - p;
          acc += -alpha * powf(q1, gamma) * logf(p);
          // d/dz: p' = p(1-p)
          g = -alpha * (-gamma * powf(q1, gamma - 1.f) * logf(p) + powf(q1, gamma) / p) * p * q1;
        } else if (t > 0) {  // -(1-alpha) * p^g * log(1-p), only for foreground rows (Loss.py:70-73)
          const float q1 = 1.f - p;
          acc += -(1.f - alpha) * powf(p, gamma) * logf(q1);
          g = -(1.f - alpha) * (gamma * powf(p, gamma - 1.f) * logf(q1) - powf(p, gamma) / q1) * p * q1;
        }
      }
      if (grad) grad[size_t(r) * C + k] = g;
    }
    acc = warp_sum(acc);
    if (lane == 0) row_out[r] = acc;
  }
}

// Loss.py:83-103 FocalLoss: one-hot targets over ALL columns (class ids 0..C-1),
//   p = clamp(sigmoid(z), 1e-4, 1 - 1e-4); ce = BCE-with-logits(z, y); a = y ? alpha : 1 - alpha;
//   pt = y ? p : 1 - p;  loss = a * (1 - pt)^gamma * ce, summed. Gradient through the clamp is zero
//   outside (1e-4, 1 - 1e-4) as in torch; d ce / d z = sigmoid(z) - y.
__global__ void bce_focal_kernel(const float* __restrict__ x, const int64_t* __restrict__ tgt, int M, int C,
                                 float alpha, float gamma, float* __restrict__ row_out,
                                 float* __restrict__ grad) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < M; r += warps) {
    const float* row = x + size_t(r) * C;
    const int64_t t = tgt[r];
    float acc = 0.f;
    for (int k = lane; k < C; k += 32) {
      const float z = row[k];
      const bool y = t == k;
      const float ps = 1.f / (1.f + expf(-z));
      const float p = fminf(fmaxf(ps, 1e-4f), 1.f - 1e-4f);
      const float ce = fmaxf(z, 0.f) - (y ? z : 0.f) + log1pf(expf(-fabsf(z)));
      const float a = y ? alpha : 1.f - alpha;
      const float om = y ? 1.f - p : p;  // 1 - pt
      const float wgt = powf(om, gamma);
      acc += a * wgt * ce;
      if (grad) {
        const float dp = (ps >= 1e-4f && ps <= 1.f - 1e-4f) ? ps * (1.f - ps) : 0.f;  // d clamp(sigmoid) / dz
        const float dom = y ? -dp : dp;                                               // d (1 - pt) / dz
        grad[size_t(r) * C + k] = a * (gamma * powf(om, gamma - 1.f) * dom * ce + wgt * (ps - (y ? 1.f : 0.f)));
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) row_out[r] = acc;
  }
}

// ---- stand-alone NMS ---------------------------------------------------------------------------
// order key: score descending, index ascending (stable), NaN-free scores assumed
SBOD_DEVINL unsigned long long nms_key(float s, int i) {
  uint32_t b = __float_as_uint(s);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // monotone map float -> uint
  return (static_cast<unsigned long long>(~b) << 32) | uint32_t(i);
}

// rank sort: position of box i in the sorted order = number of smaller keys
__global__ void __launch_bounds__(256) nms_rank_kernel(const float* __restrict__ scores, int n,
                                                       int* __restrict__ order) {
  __shared__ unsigned long long s_k[256];
  const int i = blockIdx.x * 256 + threadIdx.x;
  const unsigned long long mine = i < n ? nms_key(scores[i], i) : ~0ull;
  int rank = 0;
  for (int base = 0; base < n; base += 256) {
    const int j = base + threadIdx.x;
    __syncthreads();
    s_k[threadIdx.x] = j < n ? nms_key(scores[j], j) : ~0ull;
    __syncthreads();
    const int lim = min(256, n - base);
    for (int t = 0; t < lim; ++t) rank += s_k[t] < mine ? 1 : 0;
  }
  if (i < n) order[rank] = i;
}

// iou_utils.diounms (iou_utils.py:453-530) as written: box i is kept, box j a lower-scored candidate.
//   value = inter / ((area_j - inter) + area_i) - (d / c) ** beta1,  suppress j unless value <= thr
//   d = (cx_i - cx_j)^2 + (cy_i - y2_j)^2   (:507 takes y2 of the candidate where its centre was meant)
//   c = squared diagonal of the enclosing box.
SBOD_DEVINL bool diou_suppresses(const float4 bi, float ai, const float4 bj, float thr, float beta1) {
  const float w = fmaxf(__fsub_rn(fminf(bj.z, bi.z), fmaxf(bj.x, bi.x)), 0.f);
  const float h = fmaxf(__fsub_rn(fminf(bj.w, bi.w), fmaxf(bj.y, bi.y)), 0.f);
  const float inter = __fmul_rn(w, h);
  const float dx = __fsub_rn(__fdiv_rn(__fadd_rn(bi.x, bi.z), 2.f), __fdiv_rn(__fadd_rn(bj.x, bj.z), 2.f));
  const float dy = __fsub_rn(__fdiv_rn(__fadd_rn(bi.y, bi.w), 2.f), __fdiv_rn(__fadd_rn(bj.w, bj.w), 2.f));
  const float d = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
  const float ex = __fsub_rn(fmaxf(bj.z, bi.z), fminf(bj.x, bi.x));
  const float ey = __fsub_rn(fmaxf(bj.w, bi.w), fminf(bj.y, bi.y));
  const float c = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
  const float u = __fdiv_rn(d, c);
  const float pen = beta1 == 1.f ? u : powf(u, beta1);
  const float aj = box_area_rn(bj);
  const float v = __fsub_rn(__fdiv_rn(inter, __fadd_rn(__fsub_rn(aj, inter), ai)), pen);
  return !(v <= thr);  // IoU.le(overlap) keeps; NaN is dropped like in the reference
}

template <bool kDiou>
__global__ void __launch_bounds__(64) nms_mask_kernel(const float4* __restrict__ boxes,
                                                      const int* __restrict__ order, int m,
                                                      float thr, unsigned long long* __restrict__ mask,
                                                      int words, float beta1) {
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (cb < rb) return;  // only j > i matters
  __shared__ float4 s_b[64];
  const int cn = min(64, m - cb * 64);
  if (threadIdx.x < cn) s_b[threadIdx.x] = boxes[order[cb * 64 + threadIdx.x]];
  __syncthreads();
  const int i = rb * 64 + threadIdx.x;
  if (i < m) {
    const float4 bi = boxes[order[i]];
    const float ai = box_area_rn(bi);
    unsigned long long bits = 0ull;
    const int start = (rb == cb) ? threadIdx.x + 1 : 0;
    for (int t = start; t < cn; ++t) {
      const bool sup = kDiou ? diou_suppresses(bi, ai, s_b[t], thr, beta1)
                             : iou_plain_rn(bi, ai, s_b[t], box_area_rn(s_b[t])) > thr;
      if (sup) bits |= 1ull << t;
    }
    mask[size_t(i) * words + cb] = bits;
  }
}

// single CTA: walk 64-box blocks in order; warp 0 resolves the block's diagonal word serially with
// shuffles, then every thread ORs the rows of the newly kept boxes into the removed set.
__global__ void __launch_bounds__(1024) nms_reduce_kernel(const unsigned long long* __restrict__ mask,
                                                          const int* __restrict__ order, int m,
                                                          int words, int64_t* __restrict__ keep,
                                                          int32_t* __restrict__ count,
                                                          unsigned long long* __restrict__ removed /*[words] global scratch*/) {
  __shared__ unsigned long long s_keepbits;
  __shared__ int s_count;
  const int tid = threadIdx.x;
  for (int w = tid; w < words; w += blockDim.x) removed[w] = 0ull;
  if (tid == 0) s_count = 0;
  __syncthreads();
  for (int b = 0; b < words; ++b) {
    if (tid < 32) {
      // lane handles rows 2*lane and 2*lane+1 of this block
      unsigned long long rem = removed[b];
      const int r0 = b * 64 + 2 * tid, r1 = r0 + 1;
      const unsigned long long d0 = r0 < m ? mask[size_t(r0) * words + b] : 0ull;
      const unsigned long long d1 = r1 < m ? mask[size_t(r1) * words + b] : 0ull;
      unsigned long long keepbits = 0ull;
      for (int t = 0; t < 64; ++t) {
        const int src = t >> 1;
        const unsigned long long dt =
            __shfl_sync(0xffffffffu, (t & 1) ? d1 : d0, src);  // row t's diagonal word
        if (b * 64 + t < m && !((rem >> t) & 1ull)) {
          keepbits |= 1ull << t;
          rem |= dt;
        }
      }
      if (tid == 0) s_keepbits = keepbits;
    }
    __syncthreads();
    const unsigned long long kb = s_keepbits;
    const int base_count = s_count;
    // emit kept indices in order
    if (tid < 64 && ((kb >> tid) & 1ull)) {
      const int pos = base_count + __popcll(kb & ((1ull << tid) - 1ull));
      keep[pos] = order[b * 64 + tid];
    }
    // OR the kept rows into removed for later blocks
    for (int w = b + 1 + tid; w < words; w += blockDim.x) {
      unsigned long long acc = removed[w];
      unsigned long long bits = kb;
      while (bits) {
        const int t = __ffsll((long long)bits) - 1;
        bits &= bits - 1;
        acc |= mask[size_t(b * 64 + t) * words + w];
      }
      removed[w] = acc;
    }
    __syncthreads();
    if (tid == 0) s_count = base_count + __popcll(kb);
    __syncthreads();
  }
  if (tid == 0) *count = s_count;
}

// ---- SSD-pytorch style match (iou_utils.py:236-321), one image -----------------------------------
__global__ void __launch_bounds__(256) match_best_kernel(const float4* __restrict__ truths, int G,
                                                         const float4* __restrict__ pri_cxcy, int P,
                                                         float* __restrict__ best_ov,
                                                         int* __restrict__ best_idx,
                                                         unsigned long long* __restrict__ gkey) {
  extern __shared__ float4 s_t[];
  for (int i = threadIdx.x; i < G; i += blockDim.x) s_t[i] = truths[i];
  __syncthreads();
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
    const float4 c = pri_cxcy[p];
    const float hw = __fdiv_rn(c.z, 2.f), hh = __fdiv_rn(c.w, 2.f);
    const float4 a = make_float4(__fsub_rn(c.x, hw), __fsub_rn(c.y, hh), __fadd_rn(c.x, hw),
                                 __fadd_rn(c.y, hh));  // point_form
    const float aa = box_area_rn(a);
    float best = -__int_as_float(0x7f800000);
    int bi = 0;
    for (int g = 0; g < G; ++g) {
      const float4 t = s_t[g];
      const float v = iou_plain_rn(t, box_area_rn(t), a, aa);
      if (v > best) { best = v; bi = g; }
      if (v >= 0.f) {
        const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(v)) << 32) |
                                       (0xffffffffu - unsigned(p));
        if (key > gkey[g]) atomicMax(&gkey[g], key);
      }
    }
    best_ov[p] = best;
    best_idx[p] = bi;
  }
}

__global__ void match_force_kernel(int G, float* __restrict__ best_ov, int* __restrict__ best_idx,
                                   unsigned long long* __restrict__ gkey) {
  // for j in range(G): best_truth_idx[best_prior_idx[j]] = j  -> last j wins; overlap := 2
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const uint32_t p = 0xffffffffu - uint32_t(gkey[g] & 0xffffffffull);
    bool winner = true;
    for (int h = g + 1; h < G; ++h)
      if (0xffffffffu - uint32_t(gkey[h] & 0xffffffffull) == p) { winner = false; break; }
    best_ov[p] = 2.0f;
    if (winner) best_idx[p] = g;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x) gkey[g] = 0ull;
}

__global__ void match_emit_kernel(float threshold, const float4* __restrict__ truths,
                                  const int64_t* __restrict__ labels,
                                  const float4* __restrict__ pri_cxcy, int P, float v0, float v1,
                                  int encode_loc, const float* __restrict__ best_ov,
                                  const int* __restrict__ best_idx, float4* __restrict__ loc_out,
                                  int64_t* __restrict__ conf_out) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
    const int g = best_idx[p];
    const float4 b = truths[g];
    int64_t conf = labels[g] + 1;
    if (best_ov[p] < threshold) conf = 0;
    conf_out[p] = conf;
    if (encode_loc) {
      const float4 pr = pri_cxcy[p];
      float4 r;
      const float gx = __fsub_rn(__fdiv_rn(__fadd_rn(b.x, b.z), 2.f), pr.x);
      const float gy = __fsub_rn(__fdiv_rn(__fadd_rn(b.y, b.w), 2.f), pr.y);
      r.x = __fdiv_rn(gx, __fmul_rn(v0, pr.z));
      r.y = __fdiv_rn(gy, __fmul_rn(v0, pr.w));
      r.z = __fdiv_rn(logf(__fdiv_rn(__fsub_rn(b.z, b.x), pr.z)), v1);
      r.w = __fdiv_rn(logf(__fdiv_rn(__fsub_rn(b.w, b.y), pr.w)), v1);
      loc_out[p] = r;
    } else {
      loc_out[p] = b;
    }
  }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace sbod

using namespace sbod;

extern "C" int sbod_abi_version(void) { return SBOD_ABI_VERSION; }

extern "C" const char* sbod_error_string(int code) {
  switch (code) {
    case SBOD_OK: return "ok";
    case SBOD_ERR_INVALID: return "sbod: invalid argument";
    case SBOD_ERR_WORKSPACE: return "sbod: workspace too small or misaligned";
    case SBOD_ERR_UNSUPPORTED: return "sbod: shape outside the supported range";
    case SBOD_ERR_ALIGNMENT: return "sbod: streamed tensor is not 16-byte aligned";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "sbod: unknown error";
}

extern "C" int sbod_iou_matrix(const float* a, int A, const float* b, int B, int mode, float* out,
                               sbod_stream_t stream) {
  if (A < 0 || B < 0 || mode < SBOD_IOU_METRICS || mode > SBOD_IOU_INTERSECT) return SBOD_ERR_INVALID;
  if (A == 0 || B == 0) return SBOD_OK;
  if (!a || !b || !out) return SBOD_ERR_INVALID;
  dim3 grid((B + 255) / 256, (A + kIouChunk - 1) / kIouChunk);
  if (grid.x > 4096) grid.x = 4096;
  iou_matrix_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(a), A, reinterpret_cast<const float4*>(b), B, mode, out);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" int sbod_box_convert(const float* in, float* out, int n, int op, sbod_stream_t stream) {
  if (n < 0 || (op != SBOD_BOX_XY_TO_CXCY && op != SBOD_BOX_CXCY_TO_XY)) return SBOD_ERR_INVALID;
  if (n == 0) return SBOD_OK;
  if (!in || !out) return SBOD_ERR_INVALID;
  box_convert_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), n, op);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" int sbod_box_encode(const float* boxes, const float* priors_cxcy, float* out, int n,
                               int flavour, float v0, float v1, sbod_stream_t stream) {
  if (n < 0 || flavour < 0 || flavour > 1) return SBOD_ERR_INVALID;
  if (n == 0) return SBOD_OK;
  if (!boxes || !priors_cxcy || !out) return SBOD_ERR_INVALID;
  box_encode_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(boxes), reinterpret_cast<const float4*>(priors_cxcy),
      reinterpret_cast<float4*>(out), n, flavour, v0, v1);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" int sbod_box_decode(const float* locs, const float* priors_cxcy, float* out, int n,
                               int flavour, float v0, float v1, sbod_stream_t stream) {
  if (n < 0 || flavour < 0 || flavour > 1) return SBOD_ERR_INVALID;
  if (n == 0) return SBOD_OK;
  if (!locs || !priors_cxcy || !out) return SBOD_ERR_INVALID;
  box_decode_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(locs), reinterpret_cast<const float4*>(priors_cxcy),
      reinterpret_cast<float4*>(out), n, flavour, v0, v1);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" int sbod_prior_grid(int n_levels, const int32_t* rows, const int32_t* cols, const int32_t* n_shapes,
                               const double* scale, const double* shapes, int clamp01, float* out, long long n_out,
                               sbod_stream_t stream) {
  if (n_levels < 1 || n_levels > kGridMaxLevels || !rows || !cols || !n_shapes || !scale || !out) return SBOD_ERR_INVALID;
  GridSpec g;
  memset(&g, 0, sizeof(g));
  g.n_levels = n_levels;
  g.clamp01 = clamp01;
  int total_shapes = 0;
  bool centres = true;
  for (int l = 0; l < n_levels; ++l) centres = centres && n_shapes[l] == 0;
  g.centres_only = centres ? 1 : 0;
  long long off = 0;
  for (int l = 0; l < n_levels; ++l) {
    if (rows[l] <= 0 || cols[l] <= 0 || n_shapes[l] < 0 || (!centres && n_shapes[l] == 0)) return SBOD_ERR_INVALID;
    g.rows[l] = rows[l]; g.cols[l] = cols[l]; g.n_shapes[l] = n_shapes[l]; g.shape0[l] = total_shapes;
    g.mul_x[l] = scale[4 * l]; g.div_x[l] = scale[4 * l + 1]; g.mul_y[l] = scale[4 * l + 2]; g.div_y[l] = scale[4 * l + 3];
    g.offset[l] = off;
    off += (long long)rows[l] * cols[l] * (centres ? 1 : n_shapes[l]);
    if (total_shapes + n_shapes[l] > kGridMaxShapes) return SBOD_ERR_UNSUPPORTED;
    for (int k = 0; k < n_shapes[l]; ++k) {
      if (!shapes) return SBOD_ERR_INVALID;
      g.shape_w[total_shapes + k] = shapes[2 * (total_shapes + k)];
      g.shape_h[total_shapes + k] = shapes[2 * (total_shapes + k) + 1];
    }
    total_shapes += n_shapes[l];
  }
  g.offset[n_levels] = off;
  if (off != n_out) return SBOD_ERR_INVALID;
  if (reinterpret_cast<uintptr_t>(out) & 15) return SBOD_ERR_ALIGNMENT;
  prior_grid_kernel<<<grid_for(size_t(off), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(g, out, off);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" int sbod_box_op_bwd(int op, const float* in, const float* priors_cxcy, const float* grad_out,
                               float* grad_in, int n, float v0, float v1, sbod_stream_t stream) {
  if (n < 0 || !grad_out || !grad_in || op < 0 || op > 5) return SBOD_ERR_INVALID;
  if (op >= 2 && (!in || !priors_cxcy)) return SBOD_ERR_INVALID;
  if (n == 0) return SBOD_OK;
  box_op_bwd_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      op, reinterpret_cast<const float4*>(in), reinterpret_cast<const float4*>(priors_cxcy),
      reinterpret_cast<const float4*>(grad_out), reinterpret_cast<float4*>(grad_in), n, v0, v1);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" int sbod_offset2bbox(const float* arm_locs, const float* odm_locs,
                                const float* priors_cxcy, float* out, int N, int P,
                                sbod_stream_t stream) {
  if (N < 0 || P < 0) return SBOD_ERR_INVALID;
  if (N == 0 || P == 0) return SBOD_OK;
  if (!arm_locs || !odm_locs || !priors_cxcy || !out) return SBOD_ERR_INVALID;
  offset2bbox_kernel<<<grid_for(size_t(N) * P, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(arm_locs), reinterpret_cast<const float4*>(odm_locs),
      reinterpret_cast<const float4*>(priors_cxcy), reinterpret_cast<float4*>(out), N, P);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

namespace sbod {
__global__ void selftest_div_kernel(const float* a, const float* b, long long n, float* out, float* ref,
                                    uint8_t* fast_ok) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = div_rn_fast(a[i], b[i]);
  ref[i] = __fdiv_rn(a[i], b[i]);
  fast_ok[i] = div_fast_ok(a[i], b[i]) ? 1 : 0;
}
}  // namespace sbod

extern "C" int sbod_selftest_div(const float* a, const float* b, long long n, float* out, float* ref,
                                 uint8_t* fast_ok, sbod_stream_t stream) {
  if (!a || !b || !out || !ref || !fast_ok || n < 0) return SBOD_ERR_INVALID;
  if (n == 0) return SBOD_OK;
  sbod::selftest_div_kernel<<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, n, out, ref,
                                                                                                  fast_ok);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" int sbod_arm_easy_negative(const float* arm_scores, long long n_rows, float theta,
                                      uint8_t* out, sbod_stream_t stream) {
  if (n_rows < 0) return SBOD_ERR_INVALID;
  if (n_rows == 0) return SBOD_OK;
  if (!arm_scores || !out) return SBOD_ERR_INVALID;
  arm_easy_negative_kernel<<<grid_for(size_t(n_rows), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(arm_scores), size_t(n_rows), theta, out);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" int sbod_pair_iou_fwd(const float* b1, const float* b2, int M, int kind, float* out,
                                 sbod_stream_t stream) {
  if (M < 0 || kind < 0 || kind > SBOD_PAIR_CIOU) return SBOD_ERR_INVALID;
  if (M == 0) return SBOD_OK;
  if (!b1 || !b2 || !out) return SBOD_ERR_INVALID;
  pair_iou_fwd_kernel<<<grid_for(M, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(b1), reinterpret_cast<const float4*>(b2), M, kind, out);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" int sbod_pair_iou_bwd(const float* b1, const float* b2, const float* grad_out, int M,
                                 int kind, float* grad_b1, float* grad_b2, sbod_stream_t stream) {
  if (M < 0 || kind < 0 || kind > SBOD_PAIR_CIOU) return SBOD_ERR_INVALID;
  if (M == 0) return SBOD_OK;
  if (!b1 || !b2 || !grad_out) return SBOD_ERR_INVALID;
  pair_iou_bwd_kernel<<<grid_for(M, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(b1), reinterpret_cast<const float4*>(b2), grad_out, M, kind,
      reinterpret_cast<float4*>(grad_b1), reinterpret_cast<float4*>(grad_b2));
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" int sbod_smooth_l1(const float* pred, const float* target, int n_elem, float beta,
                              float* out, float* grad_pred, sbod_stream_t stream) {
  if (n_elem < 0) return SBOD_ERR_INVALID;
  if (n_elem == 0) return SBOD_OK;
  if (!pred || !target || !out) return SBOD_ERR_INVALID;
  smooth_l1_kernel<<<grid_for(n_elem, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, target, n_elem, beta, out, grad_pred);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" int sbod_softmax_focal(const float* logits, const int64_t* target, int M, int C,
                                  float alpha_fg, float alpha_bg, float gamma, float* row_out,
                                  float* grad_logits, sbod_stream_t stream) {
  if (M < 0 || C <= 0) return SBOD_ERR_INVALID;
  if (M == 0) return SBOD_OK;
  if (!logits || !target || !row_out) return SBOD_ERR_INVALID;
  softmax_focal_kernel<<<grid_for(size_t(M) * 32, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, target, M, C, alpha_fg, alpha_bg, gamma, row_out, grad_logits);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" int sbod_sigmoid_focal(const float* logits, const int64_t* target, int M, int C,
                                  float alpha, float gamma, float* row_out, float* grad_logits,
                                  sbod_stream_t stream) {
  if (M < 0 || C <= 0) return SBOD_ERR_INVALID;
  if (M == 0) return SBOD_OK;
  if (!logits || !target || !row_out) return SBOD_ERR_INVALID;
  sigmoid_focal_kernel<<<grid_for(size_t(M) * 32, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, target, M, C, alpha, gamma, row_out, grad_logits);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" int sbod_bce_focal(const float* logits, const int64_t* target, int M, int C, float alpha,
                              float gamma, float* row_out, float* grad_logits, sbod_stream_t stream) {
  if (M < 0 || C <= 0) return SBOD_ERR_INVALID;
  if (M == 0) return SBOD_OK;
  if (!logits || !target || !row_out) return SBOD_ERR_INVALID;
  bce_focal_kernel<<<grid_for(size_t(M) * 32, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, target, M, C, alpha, gamma, row_out, grad_logits);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" size_t sbod_nms_workspace_bytes(int n) {
  if (n <= 0) return 256;
  const size_t words = (size_t(n) + 63) / 64;
  return align_up(size_t(n) * 4, 256) + align_up(words * 8, 256) + align_up(size_t(n) * words * 8, 256);
}

static int nms_run(const float* boxes, const float* scores, int n, float iou_thr, int top_k, int64_t* keep_out,
                   int32_t* count_out, void* workspace, size_t workspace_bytes, sbod_stream_t stream, bool diou,
                   float beta1);

extern "C" int sbod_nms(const float* boxes, const float* scores, int n, float iou_thr, int top_k,
                        int64_t* keep_out, int32_t* count_out, void* workspace,
                        size_t workspace_bytes, sbod_stream_t stream) {
  return nms_run(boxes, scores, n, iou_thr, top_k, keep_out, count_out, workspace, workspace_bytes, stream, false,
                 1.f);
}

extern "C" int sbod_diou_nms(const float* boxes, const float* scores, int n, float thr, int top_k, float beta1,
                             int64_t* keep_out, int32_t* count_out, void* workspace, size_t workspace_bytes,
                             sbod_stream_t stream) {
  return nms_run(boxes, scores, n, thr, top_k, keep_out, count_out, workspace, workspace_bytes, stream, true, beta1);
}

static int nms_run(const float* boxes, const float* scores, int n, float iou_thr, int top_k, int64_t* keep_out,
                   int32_t* count_out, void* workspace, size_t workspace_bytes, sbod_stream_t stream, bool diou,
                   float beta1) {
  if (n < 0 || !count_out) return SBOD_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    SBOD_CUDA_TRY(cudaMemsetAsync(count_out, 0, 4, st));
    return SBOD_OK;
  }
  if (!boxes || !scores || !keep_out) return SBOD_ERR_INVALID;
  if (!workspace || workspace_bytes < sbod_nms_workspace_bytes(n)) return SBOD_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return SBOD_ERR_WORKSPACE;
  unsigned char* w = static_cast<unsigned char*>(workspace);
  int* order = reinterpret_cast<int*>(w);
  w += align_up(size_t(n) * 4, 256);
  const int m = (top_k > 0 && top_k < n) ? top_k : n;  // iou_utils.nms considers the top_k best only
  const int words = (m + 63) / 64;
  unsigned long long* removed = reinterpret_cast<unsigned long long*>(w);
  w += align_up((size_t(n) + 63) / 64 * 8, 256);
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(w);
  nms_rank_kernel<<<(n + 255) / 256, 256, 0, st>>>(scores, n, order);
  SBOD_LAUNCH_CHECK();
  dim3 grid(words, words);
  if (diou)
    nms_mask_kernel<true><<<grid, 64, 0, st>>>(reinterpret_cast<const float4*>(boxes), order, m, iou_thr, mask,
                                               words, beta1);
  else
    nms_mask_kernel<false><<<grid, 64, 0, st>>>(reinterpret_cast<const float4*>(boxes), order, m, iou_thr, mask,
                                                words, beta1);
  SBOD_LAUNCH_CHECK();
  nms_reduce_kernel<<<1, 1024, 0, st>>>(mask, order, m, words, keep_out, count_out, removed);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" size_t sbod_match_workspace_bytes(int G, int P) {
  if (G < 1) G = 1;
  if (P < 1) P = 1;
  return align_up(size_t(G) * 8, 256) + align_up(size_t(P) * 4, 256) * 2;
}

extern "C" int sbod_match(float threshold, const float* truths, int G, const float* priors_cxcy,
                          int P, float v0, float v1, const int64_t* labels, int encode_loc,
                          float* loc_out, int64_t* conf_out, void* workspace,
                          size_t workspace_bytes, sbod_stream_t stream) {
  // workspace: G keys (zero before the first call, left zero), then P floats + P ints of scratch
  if (G <= 0 || P <= 0 || !truths || !priors_cxcy || !labels || !loc_out || !conf_out)
    return SBOD_ERR_INVALID;
  if (G > 8192) return SBOD_ERR_UNSUPPORTED;
  const size_t need = sbod_match_workspace_bytes(G, P);
  if (!workspace || workspace_bytes < need) return SBOD_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return SBOD_ERR_WORKSPACE;
  unsigned char* w = static_cast<unsigned char*>(workspace);
  unsigned long long* gkey = reinterpret_cast<unsigned long long*>(w);
  w += align_up(size_t(G) * 8, 256);
  float* best_ov = reinterpret_cast<float*>(w);
  w += align_up(size_t(P) * 4, 256);
  int* best_idx = reinterpret_cast<int*>(w);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    SBOD_CUDA_TRY(cudaFuncSetAttribute(match_best_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_once.mark();
  }
  match_best_kernel<<<grid_for(P, 256), 256, size_t(G) * 16, st>>>(
      reinterpret_cast<const float4*>(truths), G, reinterpret_cast<const float4*>(priors_cxcy), P,
      best_ov, best_idx, gkey);
  SBOD_LAUNCH_CHECK();
  match_force_kernel<<<1, 256, 0, st>>>(G, best_ov, best_idx, gkey);
  SBOD_LAUNCH_CHECK();
  match_emit_kernel<<<grid_for(P, 256), 256, 0, st>>>(
      threshold, reinterpret_cast<const float4*>(truths), labels,
      reinterpret_cast<const float4*>(priors_cxcy), P, v0, v1, encode_loc, best_ov, best_idx,
      reinterpret_cast<float4*>(loc_out), conf_out);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}
