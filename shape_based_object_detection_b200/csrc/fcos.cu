// fcos.cu — anchor-free (FCOS) targets, loss (forward / backward) and post-processing (sm_100a).
//
// The reference's FCOSLoss / FCOS.postprocess (models/FCOSDet.py:253-270, 311-544) do not run
// (SURVEY.md §8 a-F); these kernels implement the semantics that file spells out — constants and
// structure from FCOSDet.py:333-335 (strides, size-of-interest ranges, radius 1.5), :343-486
// (centre sampling, min-area assignment, centerness target) and :527-544 (loss composition), with
// the canonical FCOS meaning where a line cannot execute. Parity is pinned to the oracle's
// restatement only ("parity unpinned" with respect to the reference itself).
#include <math.h>

#include "comm.cuh"
#include "common.cuh"
#include "pair_iou.cuh"

namespace sbod {

constexpr float kFcosInf = 1e6f;  // FCOSDet.py:326
constexpr int kFcosChunk = 256;

struct FcosParams {
  const float4* locs;        // [N,P] predicted l,t,r,b
  const float* scores;       // [N,P,C] logits
  const float* centerness;   // [N,P] logits
  const float2* locations;   // [P] cell centres
  const float* aux;          // [P,4]: centre-sampling radius in x and in y, size-of-interest lo, hi
  const float4* gt_boxes;
  const int64_t* gt_labels;
  const int32_t* gt_offsets;
  int N, P, C;
  int center_sample;
  float reg_weight, falpha, fgamma;
  int32_t* lab;              // [N,P]
  float4* tgt;               // [N,P] ltrb target
  double* blockpart;         // [N, nblk, 5]
  double* sums;              // [6]: focal, sum(l*w), sum(w), bce, n_pos, n_images
  float* loss;               // [4]: total, conf, loc, center
  const CommDev* comm;       // in-kernel all-reduce of the six sums over the ranks (or null)
};

// ---- assignment: one thread per location, objects staged in shared memory ----------------------
__global__ void __launch_bounds__(256) fcos_assign_kernel(const FcosParams q) {
  __shared__ float4 s_box[kFcosChunk];
  __shared__ float s_area[kFcosChunk];
  const int n = blockIdx.y, p = blockIdx.x * blockDim.x + threadIdx.x;
  const int g0 = q.gt_offsets[n];
  const int G = q.gt_offsets[n + 1] - g0;
  const bool valid = p < q.P;
  float2 xy = make_float2(0.f, 0.f);
  float rad_x = 0.f, rad_y = 0.f, lo = 0.f, hi = 0.f;
  if (valid) {
    xy = q.locations[p];
    const float4 a = reinterpret_cast<const float4*>(q.aux)[p];
    rad_x = a.x;
    rad_y = a.y;
    lo = a.z;
    hi = a.w;
  }
  float best_area = kFcosInf;  // locs_gt_area.min(dim=1): first index among ties (FCOSDet.py:416)
  int best = 0;
  float4 best_t = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c0 = 0; c0 < G; c0 += kFcosChunk) {
    const int gc = min(kFcosChunk, G - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < gc; i += blockDim.x) {
      const float4 b = q.gt_boxes[g0 + c0 + i];
      s_box[i] = b;
      s_area[i] = box_area_rn(b);
    }
    __syncthreads();
    if (!valid) continue;
    for (int j = 0; j < gc; ++j) {
      const float4 b = s_box[j];
      const float l = __fsub_rn(xy.x, b.x), t = __fsub_rn(xy.y, b.y);
      const float r = __fsub_rn(b.z, xy.x), bt = __fsub_rn(b.w, xy.y);
      bool inside;
      if (q.center_sample) {  // FCOSDet.py:424-474: the object's centre box of half-size stride*radius, clipped to the object
        const float cx = __fdiv_rn(__fadd_rn(b.x, b.z), 2.f), cy = __fdiv_rn(__fadd_rn(b.y, b.w), 2.f);
        const float x0 = __fsub_rn(cx, rad_x), y0 = __fsub_rn(cy, rad_y);
        const float x1 = __fadd_rn(cx, rad_x), y1 = __fadd_rn(cy, rad_y);
        const float bx0 = x0 > b.x ? x0 : b.x, by0 = y0 > b.y ? y0 : b.y;
        const float bx1 = x1 < b.z ? x1 : b.z, by1 = y1 < b.w ? y1 : b.w;
        const float m = fminf(fminf(__fsub_rn(xy.x, bx0), __fsub_rn(xy.y, by0)),
                              fminf(__fsub_rn(bx1, xy.x), __fsub_rn(by1, xy.y)));
        inside = m > 0.f;
      } else {
        inside = fminf(fminf(l, t), fminf(r, bt)) > 0.f;
      }
      const float mx = fmaxf(fmaxf(l, t), fmaxf(r, bt));
      const bool in_level = mx >= lo && mx <= hi;  // size of interest of this pyramid level
      const float area = (inside && in_level) ? s_area[j] : kFcosInf;
      if (area < best_area) {
        best_area = area;
        best = c0 + j;
        best_t = make_float4(l, t, r, bt);
      } else if (c0 + j == 0) {
        best_t = make_float4(l, t, r, bt);  // argmin of an all-INF row is object 0
      }
    }
  }
  if (valid) {
    const size_t np = size_t(n) * q.P + p;
    int lab = 0;
    if (G > 0 && best_area != kFcosInf) lab = int(q.gt_labels[g0 + best]);
    q.lab[np] = lab;
    q.tgt[np] = best_t;
  }
}

// ---- per-positive terms ------------------------------------------------------------------------
SBOD_DEVINL float centerness_target(const float4 t) {  // FCOSDet.py:479-486
  const float lr = fminf(t.x, t.z) / fmaxf(t.x, t.z);
  const float tb = fminf(t.y, t.w) / fmaxf(t.y, t.w);
  return sqrtf(lr * tb);
}
SBOD_DEVINL float4 anchored_box(const float2 xy, const float4 d) {  // FCOSDet.py:264-267
  return make_float4(xy.x - d.x, xy.y - d.y, xy.x + d.z, xy.y + d.w);
}

// one warp per 32 locations; the class row of each positive is walked by the whole warp
template <bool BACKWARD>
__global__ void __launch_bounds__(128) fcos_terms_kernel(const FcosParams q, const float* __restrict__ grad_loss,
                                                         float4* __restrict__ grad_locs,
                                                         float* __restrict__ grad_scores,
                                                         float* __restrict__ grad_center) {
  __shared__ double s_red[34];
  const int n = blockIdx.y, lane = threadIdx.x & 31;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = p < q.P;
  const size_t np = size_t(n) * q.P + (valid ? p : 0);
  const int lab = valid ? q.lab[np] : 0;
  const bool pos = lab > 0;
  double a_focal = 0.0, a_lw = 0.0, a_w = 0.0, a_bce = 0.0, a_np = 0.0;

  // scales of the backward pass (FCOSDet.py:527-544)
  float s_conf = 0.f, s_loc = 0.f, s_ctr = 0.f;
  bool weighted = true;
  if (BACKWARD) {
    const double npos = q.sums[4], sw = q.sums[2];
    const float gout = grad_loss ? *grad_loss : 1.f;
    s_conf = float(double(gout) / (npos + q.sums[5]));  // sums[5]: images of the whole (possibly sharded) batch
    weighted = sw > 1e-6;  // IouLoss: weights branch only if their sum is > 1e-6 (Loss.py:192-199)
    s_loc = npos > 0.0 ? float(double(gout) * double(q.reg_weight) / (weighted ? sw : npos)) : 0.f;
    s_ctr = npos > 0.0 ? float(double(gout) / npos) : 0.f;
  }

  if (pos) {
    const float4 t = q.tgt[np];
    const float w = centerness_target(t);
    const float2 xy = q.locations[p];
    const float4 pl = q.locs[np];
    PairGrad pg;
    const float v = pair_overlap<BACKWARD>(anchored_box(xy, pl), anchored_box(xy, t), SBOD_PAIR_DIOU, &pg);
    const float x = q.centerness[np];
    if (!BACKWARD) {
      a_lw = double((1.f - v) * w);
      a_w = double(w);
      a_bce = double(fmaxf(x, 0.f) - x * w + log1pf(__expf(-fabsf(x))));  // BCEWithLogits
      a_np = 1.0;
    } else {
      const float k = s_loc * (weighted ? w : 1.f);
      // loss = 1 - diou; box = (x - l, y - t, x + r, y + b)
      if (grad_locs) grad_locs[np] = make_float4(k * pg.d1.x, k * pg.d1.y, -k * pg.d1.z, -k * pg.d1.w);
      if (grad_center) grad_center[np] = s_ctr * (1.f / (1.f + __expf(-x)) - w);
    }
  } else if (BACKWARD && valid) {
    if (grad_locs) grad_locs[np] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (grad_center) grad_center[np] = 0.f;
  }

  // sigmoid focal over columns 1..C-1 of the positive rows only: rows with target 0 contribute
  // nothing because the negative term is masked by (t > 0) (Loss.py:70-73)
  unsigned m = __ballot_sync(0xffffffffu, pos);
  while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    const int t = __shfl_sync(0xffffffffu, lab, src);
    const size_t row = (size_t(n) * q.P + size_t(blockIdx.x) * blockDim.x + (threadIdx.x & ~31) + src) * q.C;
    float acc = 0.f;
    for (int k = lane; k < q.C; k += 32) {
      float g = 0.f;
      if (k >= 1) {
        const float z = q.scores[row + k];
        const float pr = 1.f / (1.f + __expf(-z));
        const float q1 = 1.f - pr;
        if (k == t) {
          acc += -q.falpha * powf(q1, q.fgamma) * logf(pr);
          g = -q.falpha * (-q.fgamma * powf(q1, q.fgamma - 1.f) * logf(pr) + powf(q1, q.fgamma) / pr) * pr * q1;
        } else {
          acc += -(1.f - q.falpha) * powf(pr, q.fgamma) * logf(q1);
          g = -(1.f - q.falpha) * (q.fgamma * powf(pr, q.fgamma - 1.f) * logf(q1) - powf(pr, q.fgamma) / q1) * pr * q1;
        }
      }
      if (BACKWARD && grad_scores) grad_scores[row + k] = s_conf * g;
    }
    if (!BACKWARD) {
      acc = warp_sum(acc);
      if (lane == src) a_focal += double(acc);
    }
  }
  if (!BACKWARD) {
    const double t0 = block_sum(a_focal, s_red), t1 = block_sum(a_lw, s_red), t2 = block_sum(a_w, s_red);
    const double t3 = block_sum(a_bce, s_red), t4 = block_sum(a_np, s_red);
    if (threadIdx.x == 0) {
      double* bp = q.blockpart + (size_t(n) * gridDim.x + blockIdx.x) * 5;
      bp[0] = t0; bp[1] = t1; bp[2] = t2; bp[3] = t3; bp[4] = t4;
    }
  }
}

// loss from the six sums (FCOSDet.py:527-544); the sums may have been all-reduced over the ranks
SBOD_DEVINL void fcos_finalize(const FcosParams& q, const double* tot) {
  const double npos = tot[4];
  const double conf = tot[0] / (npos + tot[5]);
  double loc = 0.0, ctr = 0.0;
  if (npos > 0.0) {
    loc = tot[2] > 1e-6 ? tot[1] / tot[2] : tot[1] / npos;
    ctr = tot[3] / npos;
  }
  q.loss[0] = float(conf + double(q.reg_weight) * loc + ctr);
  q.loss[1] = float(conf);
  q.loss[2] = float(loc);
  q.loss[3] = float(ctr);
}

// fold the block partials in a fixed order and form the loss
__global__ void __launch_bounds__(256) fcos_finalize_kernel(const FcosParams q, int n_parts) {
  __shared__ double s_red[34];
  double acc[5] = {0, 0, 0, 0, 0};
  // each thread sums a strided subset, then a block sum: the order is fixed by the launch shape
  for (int i = threadIdx.x; i < n_parts; i += blockDim.x)
    for (int k = 0; k < 5; ++k) acc[k] += q.blockpart[size_t(i) * 5 + k];
  double tot[6];
  for (int k = 0; k < 5; ++k) tot[k] = block_sum(acc[k], s_red);
  tot[5] = double(q.N);
  if (q.comm && threadIdx.x < 32) comm_allreduce_sum(q.comm, tot, 6);  // sharded batch: sums of all ranks, rank order
  if (threadIdx.x == 0) {
    for (int k = 0; k < 6; ++k) q.sums[k] = tot[k];
    fcos_finalize(q, tot);
  }
}

__global__ void fcos_refinalize_kernel(const FcosParams q) {
  if (threadIdx.x == 0 && blockIdx.x == 0) fcos_finalize(q, q.sums);
}

// ---- FCOS.postprocess (FCOSDet.py:253-270) -------------------------------------------------------
// A warp per location row: the C class logits of a row are contiguous, so lanes walk them (coalesced; one thread per
// location with a stride of C floats between lanes cost 32 sectors per load and ran at 0.15 of the HBM roofline).
__global__ void __launch_bounds__(256) fcos_postprocess_kernel(const float4* __restrict__ box_pred,
                                                               const float* __restrict__ cls_pred,
                                                               const float* __restrict__ center_pred,
                                                               const float2* __restrict__ locations, int N, int P, int C,
                                                               float4* __restrict__ out_locs,
                                                               float* __restrict__ out_scores) {
  const size_t total = size_t(N) * P;
  const int lane = threadIdx.x & 31;
  const size_t warp0 = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = (size_t(gridDim.x) * blockDim.x) >> 5;
  // boxes: one thread per location (float4 in, float4 out: coalesced)
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x)
    out_locs[i] = anchored_box(locations[i % P], box_pred[i]);
  // scores: eight rows per warp and step, their loads in flight together
  for (size_t r0 = warp0 * 8; r0 < total; r0 += n_warps * 8) {
    float ctr[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const size_t i = r0 + u;
      ctr[u] = i < total ? 1.f / (1.f + expf(-center_pred[i])) : 0.f;
    }
    for (int k0 = 0; k0 < C; k0 += 32) {
      const int k = k0 + lane;
      float x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const size_t i = r0 + u;
        x[u] = (i < total && k < C) ? cls_pred[i * C + k] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const size_t i = r0 + u;
        if (i < total && k < C) out_scores[i * C + k] = (1.f / (1.f + expf(-x[u]))) * ctr[u];
      }
    }
  }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int fill_fcos(const sbod_fcos_desc* d, FcosParams& q, int* nblk) {
  if (!d || d->N <= 0 || d->P <= 0 || d->C <= 1) return SBOD_ERR_INVALID;
  if (!d->locs || !d->scores || !d->centerness || !d->locations || !d->loc_aux || !d->gt_boxes ||
      !d->gt_labels || !d->gt_offsets || !d->lab || !d->tgt || !d->sums || !d->loss)
    return SBOD_ERR_INVALID;
  *nblk = (d->P + 127) / 128;
  const size_t need = align_up(size_t(d->N) * *nblk * 5 * 8, 256);
  if (!d->workspace || d->workspace_bytes < need) return SBOD_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(d->workspace) & 255) return SBOD_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(d->loc_aux) & 15) return SBOD_ERR_ALIGNMENT;
  q.locs = reinterpret_cast<const float4*>(d->locs);
  q.scores = d->scores;
  q.centerness = d->centerness;
  q.locations = reinterpret_cast<const float2*>(d->locations);
  q.aux = d->loc_aux;
  q.gt_boxes = reinterpret_cast<const float4*>(d->gt_boxes);
  q.gt_labels = d->gt_labels;
  q.gt_offsets = d->gt_offsets;
  q.N = d->N; q.P = d->P; q.C = d->C;
  q.center_sample = d->center_sample;
  q.reg_weight = d->reg_weight; q.falpha = d->focal_alpha; q.fgamma = d->focal_gamma;
  q.lab = d->lab;
  q.tgt = reinterpret_cast<float4*>(d->tgt);
  q.blockpart = static_cast<double*>(d->workspace);
  q.sums = d->sums;
  q.loss = d->loss;
  q.comm = static_cast<const CommDev*>(d->comm);
  return SBOD_OK;
}

}  // namespace sbod

using namespace sbod;

extern "C" size_t sbod_fcos_workspace_bytes(const sbod_fcos_desc* d) {
  if (!d) return 0;
  return align_up(size_t(d->N) * ((d->P + 127) / 128) * 5 * 8, 256);
}

extern "C" int sbod_fcos_forward(const sbod_fcos_desc* d, sbod_stream_t stream) {
  FcosParams q;
  int nblk = 0;
  int rc = fill_fcos(d, q, &nblk);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 ga((q.P + 255) / 256, q.N);
  fcos_assign_kernel<<<ga, 256, 0, st>>>(q);
  SBOD_LAUNCH_CHECK();
  dim3 gt(nblk, q.N);
  fcos_terms_kernel<false><<<gt, 128, 0, st>>>(q, nullptr, nullptr, nullptr, nullptr);
  SBOD_LAUNCH_CHECK();
  fcos_finalize_kernel<<<1, 256, 0, st>>>(q, nblk * q.N);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" int sbod_fcos_finalize(const sbod_fcos_desc* d, sbod_stream_t stream) {
  FcosParams q;
  int nblk = 0;
  int rc = fill_fcos(d, q, &nblk);
  if (rc) return rc;
  fcos_refinalize_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(q);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" int sbod_fcos_backward(const sbod_fcos_desc* d, const float* grad_loss, float* grad_locs,
                                  float* grad_scores, float* grad_center, sbod_stream_t stream) {
  FcosParams q;
  int nblk = 0;
  int rc = fill_fcos(d, q, &nblk);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (grad_scores)  // zero everywhere except the positive rows patched below
    SBOD_CUDA_TRY(cudaMemsetAsync(grad_scores, 0, size_t(q.N) * q.P * size_t(q.C) * 4, st));
  dim3 gt(nblk, q.N);
  fcos_terms_kernel<true><<<gt, 128, 0, st>>>(q, grad_loss, reinterpret_cast<float4*>(grad_locs), grad_scores,
                                              grad_center);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" int sbod_fcos_postprocess(const float* box_pred, const float* cls_pred, const float* center_pred,
                                     const float* locations, int N, int P, int C, float* out_locs,
                                     float* out_scores, sbod_stream_t stream) {
  if (N <= 0 || P <= 0 || C <= 0) return SBOD_ERR_INVALID;
  if (!box_pred || !cls_pred || !center_pred || !locations || !out_locs || !out_scores) return SBOD_ERR_INVALID;
  const size_t total = size_t(N) * P;
  int grid = int((total + 63) / 64);  // eight warps of eight rows per CTA
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  fcos_postprocess_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(box_pred), cls_pred, center_pred,
      reinterpret_cast<const float2*>(locations), N, P, C, reinterpret_cast<float4*>(out_locs), out_scores);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}
