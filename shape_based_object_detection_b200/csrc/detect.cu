// detect.cu — the eval path (sm_100a): activation + decode + clamp + score threshold +
// per-class NMS + top-k, batched, no host round trips.
//
//   detect_bound_kernel  the streaming pass: logits HBM -> smem by bulk TMA exactly once (L2 evict-first), two
//                        threads per prior row, ONE pass: best foreground probability of the row, inflated
//                        to an upper bound of what the exact evaluation can produce (+ per-image histogram of
//                        the bounds). HBM-bound: five instructions per logit, nothing else in the kernel.
//   detect_refine_kernel exact activation (as torch: max shift, accurate exp, division) of the rows whose
//                        bound lies above the image's cutoff: an octet of lanes per row, the row in
//                        registers, class count a template parameter, classes above a logit threshold queued
//                        per warp and evaluated by consecutive lanes; every (class, prior) above
//                        min_score and the cutoff is emitted as one 64-bit key
//                        [0x3F800000 - score_bits : 32][class : 12][prior : 20]  so that
//                        ascending key order == (score desc, class asc, prior asc) == the order in
//                        which the reference's per-class NMS + stable top-k sort consumes them.
//   detect_nms_kernel    one CTA per image. The reference runs NMS over every candidate of every
//                        class and only then keeps the top_k best (models/utils.py:245-290). Greedy
//                        NMS is prefix-stable, so the first top_k+1 survivors in global key order
//                        are all that is ever needed: candidates are pulled in key-ordered chunks
//                        (histogram-guided, radix descent for oversized bins), sorted in smem,
//                        suppressed per class (pairwise bit masks; large class segments and the
//                        class-agnostic second stage of detect_tools as bitmask NMS over all warps),
//                        and the loop stops as soon as top_k+1 boxes survive. If the candidates above
//                        the cutoff run out first the CTA evaluates the image's remaining rows itself
//                        (second band); if everything runs out, everything is kept and emitted
//                        class-major, exactly as the reference does.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "row_stream.cuh"

namespace sbod {

constexpr int kChunk = 1024;       // candidates per NMS round (== threads of detect_nms_kernel)
constexpr int kNmsThreads = 1024;
constexpr int kMaxBins = 2048;
constexpr int kBigSeg = 64;        // class segments longer than this use the bitmask path
constexpr int kClassBits = 12, kPriorBits = 20;
constexpr uint32_t kOneBits = 0x3F800000u;

struct DetParams {
  float* locs;
  const float* scores;
  const float4* priors_cxcy;
  const uint8_t* prior_keep;
  int N, P, C;
  int act_kind, box_kind, clamp_inplace;
  float min_score, max_overlap;
  int top_k;
  float second_thr;
  int pre_nms_topk;
  int agnostic;                 // detect_objects: one candidate per prior (best foreground class), class-agnostic NMS
  float* out_boxes;
  int64_t* out_labels;
  float* out_scores;
  int32_t* out_prior;
  int32_t* out_counts;
  int out_cap;
  // workspace
  unsigned int* cand_count;     // [N]
  unsigned int* hist;           // [N, kMaxBins] level-0 digit histogram of the emitted candidate keys
  unsigned long long* cand;     // [N, cand_cap]
  unsigned int* class_seen;     // [N, C]  (pre-NMS per-class rank counters)
  long long cand_cap;
  int n_bins, shift0;           // level-0 digit = key >> shift0
  int kcap;                     // kept-list capacity in smem
  // row bounds (exact two-step candidate generation, see detect_bound_kernel)
  float* pbound;                // [N, P] upper bound of the best foreground probability of each prior
  unsigned int* rhist;          // [N, kMaxBins] histogram of the row bounds (same bins as the candidate keys)
  unsigned int* cutoff;         // [N] first-band cutoff of each image (written by the refine pass, read by the NMS)
  int rows_target;              // rows evaluated per image in the first band
  int32_t* agn_label;           // [N, P] arg-max class of each prior (class-agnostic mode)
  unsigned int* nms_mask;       // [N, kChunk, kChunk/32] suppression bits of large class segments
  // second-stage kept list spill (detect_tools: the first stage may keep more boxes than fit in smem)
  unsigned long long* spill_key;  // [N, spill_cap]
  float4* spill_box;              // [N, spill_cap]
  uint8_t* spill_st2;             // [N, spill_cap]
  int spill_cap;
  // tiling of the bound pass
  int rows_per_tile, tiles_per_image, n_tiles, n_stages;
  uint32_t stage_floats;
};

SBOD_DEVINL float4 decode_box(const DetParams& q, int n, int p) {
  const float4 l = reinterpret_cast<const float4*>(q.locs)[size_t(n) * q.P + p];
  float4 b;
  if (q.box_kind == SBOD_BOX_OFFSET) {  // cxcy_to_xy(gcxgcy_to_cxcy(l, prior)), transforms.py:37-45,69-83
    const float4 pr = q.priors_cxcy[p];
    const float cx = l.x * pr.z / 10.f + pr.x, cy = l.y * pr.w / 10.f + pr.y;
    const float w = expf(l.z / 5.f) * pr.z, h = expf(l.w / 5.f) * pr.w;
    b = make_float4(cx - w / 2.f, cy - h / 2.f, cx + w / 2.f, cy + h / 2.f);
  } else if (q.box_kind == SBOD_BOX_CENTER) {
    b = make_float4(l.x - l.z / 2.f, l.y - l.w / 2.f, l.x + l.z / 2.f, l.y + l.w / 2.f);
  } else {
    b = l;
  }
  // clamp_(0, 1): torch.clamp maps NaN to NaN; fminf/fmaxf would not — NaN parity is out of scope
  b.x = fminf(fmaxf(b.x, 0.f), 1.f);
  b.y = fminf(fmaxf(b.y, 0.f), 1.f);
  b.z = fminf(fmaxf(b.z, 0.f), 1.f);
  b.w = fminf(fmaxf(b.w, 0.f), 1.f);
  return b;
}

// score -> upper 32 bits of the candidate key (ascending = better score first)
SBOD_DEVINL uint32_t score_k32(float pr) {
  const uint32_t bits = __float_as_uint(pr);
  return bits > kOneBits ? 0u : kOneBits - bits;
}
SBOD_DEVINL int k32_bin(const DetParams& q, uint32_t k32) {
  const int bin = int(k32 >> (q.shift0 - 32));
  return bin < q.n_bins ? bin : q.n_bins - 1;
}

// ------------------------------------------------------------------------------------------
// detect_bound_kernel — the streaming pass of the eval path. The logits go HBM -> shared memory exactly
// once (bulk TMA through an mbarrier ring, same 128-row tiles and two-threads-per-row layout as the train
// kernel). Per prior row it computes the row's BEST FOREGROUND PROBABILITY in one pass — softmax:
// max_k exp(x_k - x_0) / sum_j exp(x_j - x_0), shifted by the row's own background logit so that neither a
// separate max pass nor a per-class compare / mask is needed (5 instructions per logit instead of ~8);
// sigmoid: sigmoid(max_k x_k); none: max_k x_k — inflated by a small slack so that it is an UPPER BOUND of
// every probability the exact evaluation can produce for the row. It writes the bound (4 bytes per prior)
// and a per-image histogram of the bounds in the bins of the candidate keys. The exact probabilities are
// then computed only for the rows whose bound can matter (detect_refine_kernel) — exact, because a row's
// candidates all lie below its bound.
// ------------------------------------------------------------------------------------------
constexpr float kBoundSlack = 1.00002f;  // covers ex2.approx / summation-order differences to the exact evaluation

template <int kC>  // kC > 0: compile-time class count (unrolled); 0: run-time
__global__ void __launch_bounds__(kStreamThreads, 2) detect_bound_kernel(const DetParams q) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stages = reinterpret_cast<float*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + size_t(q.n_stages) * q.stage_floats * 4);
  __shared__ unsigned int s_hist[kMaxBins];

  const int tid = threadIdx.x;
  const int C = kC ? kC : q.C;
  int t0, t1;
  tile_range(q.n_tiles, blockIdx.x, gridDim.x, t0, t1);
  const int n_my = t1 - t0;
  if (n_my <= 0) return;
  if (tid == 0) {
    for (int s = 0; s < q.n_stages; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  for (int b = tid; b < q.n_bins; b += kStreamThreads) s_hist[b] = 0u;
  __syncthreads();
  if (tid == 0) {
    for (int s = 0; s < q.n_stages && s < n_my; ++s)
      stream_issue(q.scores, q.N, q.P, C, stream_tile(t0 + s, q.tiles_per_image, kTileRows, q.P),
                   stages + size_t(s) * q.stage_floats, &bars[s]);
  }
  int row, h;
  stream_map(tid, row, h);
  const int nh = (C + 1 - h) >> 1;
  const float INF = __int_as_float(0x7f800000);

  auto flush_hist = [&](int n) {
    __syncthreads();
    unsigned int* gh = q.rhist + size_t(n) * kMaxBins;
    for (int b = tid; b < q.n_bins; b += kStreamThreads) {
      const unsigned int v = s_hist[b];
      if (v) {
        atomicAdd(&gh[b], v);
        s_hist[b] = 0u;
      }
    }
    __syncthreads();
  };

  int cur_n = t0 / q.tiles_per_image;  // tile coordinates advance incrementally
  int cur_t = t0 - cur_n * q.tiles_per_image;
  int hist_n = -1;
  int ring_s = 0;
  uint32_t ring_parity = 0;
  for (int it = 0; it < n_my; ++it) {
    const int n = cur_n;
    const int p0 = cur_t * kTileRows;
    const int rows = min(kTileRows, q.P - p0);
    if (++cur_t == q.tiles_per_image) {
      cur_t = 0;
      ++cur_n;
    }
    if (n != hist_n) {
      if (hist_n >= 0) flush_hist(hist_n);
      hist_n = n;
    }
    const bool valid = row < rows;
    const int r = min(row, rows - 1);  // keep every lane in the shuffles
    const size_t np = size_t(n) * q.P + p0 + r;
    if (valid && h == 0 && q.clamp_inplace) {  // clamp_ on the caller's tensor (models/utils.py:224, detect_tools.py:264)
      float4* lp = reinterpret_cast<float4*>(q.locs) + np;
      float4 b = *lp;
      b.x = fminf(fmaxf(b.x, 0.f), 1.f); b.y = fminf(fmaxf(b.y, 0.f), 1.f);
      b.z = fminf(fmaxf(b.z, 0.f), 1.f); b.w = fminf(fmaxf(b.w, 0.f), 1.f);
      *lp = b;
    }
    float* stage = stages + size_t(ring_s) * q.stage_floats;
    mbar_wait(&bars[ring_s], ring_parity);
    const uint32_t head = ((uint32_t(n) * uint32_t(q.P) + uint32_t(p0)) * uint32_t(C)) & 3u;
    const float* rbase = stage + head + r * C;
    float pb;
    if (q.act_kind == SBOD_ACT_SOFTMAX) {
      const float x0 = rbase[0];
      float sum, fgmax;
      pair_row_sum_fgmax<kC>(rbase, h, C, -x0 * kLog2e, sum, fgmax);
      pb = fgmax * __frcp_rn(sum);
      if (!(sum < INF) || !(pb == pb)) pb = 2.f;  // overflow (a logit > background + 88) or NaN: always evaluate the row
    } else {
      // maximum foreground logit of this thread's elements rp[0], rp[2], ... (element 0 of parity 0 = background)
      const float* rp = rbase + h;
      float m = -INF;
      for (int j = (h == 0 ? 1 : 0); j < nh; ++j) m = fmaxf(m, rp[2 * j]);
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      pb = q.act_kind == SBOD_ACT_SIGMOID ? __frcp_rn(1.f + __expf(-m)) : m;
      if (!(pb == pb)) pb = 2.f;
    }
    pb = pb * kBoundSlack;
    if (valid && h == 0) {
      if (q.prior_keep && q.prior_keep[np] == 0) pb = -1.f;  // filtered prior: never a candidate
      q.pbound[np] = pb;
      if (pb > q.min_score) atomicAdd(&s_hist[k32_bin(q, score_k32(pb))], 1u);
    }
    __syncthreads();  // every thread is done with the stage
    if (tid == 0 && it + q.n_stages < n_my)
      stream_issue(q.scores, q.N, q.P, C, stream_tile(t0 + it + q.n_stages, q.tiles_per_image, kTileRows, q.P), stage,
                   &bars[ring_s]);
    if (++ring_s == q.n_stages) {
      ring_s = 0;
      ring_parity ^= 1u;
    }
  }
  if (hist_n >= 0) flush_hist(hist_n);
}

// Generic bound pass (C > 128): one warp per prior row straight from global memory (coalesced over classes).
__global__ void __launch_bounds__(256) detect_bound_generic_kernel(const DetParams q) {
  const int lane = threadIdx.x & 31;
  const size_t total = size_t(q.N) * q.P;
  const float INF = __int_as_float(0x7f800000);
  for (size_t np = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; np < total;
       np += (size_t(gridDim.x) * blockDim.x) >> 5) {
    const float* x = q.scores + np * size_t(q.C);
    if (lane == 0 && q.clamp_inplace) {
      float4* lp = reinterpret_cast<float4*>(q.locs) + np;
      float4 b = *lp;
      b.x = fminf(fmaxf(b.x, 0.f), 1.f); b.y = fminf(fmaxf(b.y, 0.f), 1.f);
      b.z = fminf(fmaxf(b.z, 0.f), 1.f); b.w = fminf(fmaxf(b.w, 0.f), 1.f);
      *lp = b;
    }
    float m = -INF, mall = -INF;
    for (int k = lane; k < q.C; k += 32) {
      const float v = x[k];
      mall = fmaxf(mall, v);
      if (k >= 1) m = fmaxf(m, v);
    }
    m = warp_max(m);
    mall = warp_max(mall);
    float pb;
    if (q.act_kind == SBOD_ACT_SOFTMAX) {
      float s = 0.f;
      for (int k = lane; k < q.C; k += 32) s += __expf(x[k] - mall);
      s = warp_sum(s);
      pb = __expf(m - mall) / s;
    } else {
      pb = q.act_kind == SBOD_ACT_SIGMOID ? __frcp_rn(1.f + __expf(-m)) : m;
    }
    if (!(pb == pb)) pb = 2.f;
    pb *= kBoundSlack;
    if (lane == 0) {
      if (q.prior_keep && q.prior_keep[np] == 0) pb = -1.f;
      q.pbound[np] = pb;
      if (pb > q.min_score) atomicAdd(&q.rhist[size_t(np / q.P) * kMaxBins + k32_bin(q, score_k32(pb))], 1u);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Exact evaluation of prior rows, FOUR rows per warp: an octet (8 lanes) owns a row, lane s of the octet
// handles classes s, s + 8, ... (one 32-byte sector per row and step). A whole warp per row cost ~500 warp
// instructions per row, nearly all of them control / shuffles replicated over 32 lanes (measured: the refine
// pass was issue-bound at 30 M warp instructions); an octet per row needs about a quarter of that.
// Activation as torch computes it (max shift, accurate exp, true division); a class is evaluated exactly only
// if a cheap estimate says it can pass. Every (class, prior) whose probability exceeds min_score and whose key
// lies in [k_lo, k_hi) is emitted through `emit(key)`. Class-agnostic mode emits the row's best foreground
// class once, with class field 1, and records the arg-max class.
// All 32 lanes must call this; `p` < 0 marks an idle octet.
// ------------------------------------------------------------------------------------------
SBOD_DEVINL float octet_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
}
SBOD_DEVINL float octet_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v + __shfl_xor_sync(0xffffffffu, v, 4);
}

constexpr int kRowRegs = 16;  // logits per lane of an octet kept in registers: rows of up to 128 classes
constexpr int kRowQueue = 64;  // per-warp queue of (row, class) pairs waiting for their exact evaluation

// exp(x - mx) of the softmax denominator: the accurate expf, as torch evaluates it - with ex2.approx (2e-7
// relative) one more image in 32 of the config-2 batch swaps two near-tied candidates against the CPU reference.
// detect_probabilities_kernel uses the same function.
SBOD_DEVINL float exp_term(float x, float mx) { return expf(x - mx); }

// kC > 0: compile-time class count (immediate offsets, no bounds tests). kRegs: the row lives in registers.
// wq: this warp's queue in shared memory (kRowQueue entries) or null. With a queue the classes that pass the
// threshold test are first collected and then evaluated exactly by consecutive lanes, so the expensive part
// (accurate exp, division, key, emit) runs once per ~32 classes instead of divergently.
template <int kC, bool kRegs, typename Emit>
SBOD_DEVINL void eval_rows_impl(const DetParams& q, int n, int p, int lane, uint32_t k_lo, uint32_t k_hi, Emit&& emit,
                                uint2* wq) {
  const int C = kC ? kC : q.C;
  const int s = lane & 7;
  const bool on = p >= 0;
  const float* x = q.scores + (size_t(n) * q.P + (on ? p : 0)) * size_t(C) + s;
  const float NEG = -__int_as_float(0x7f800000);
  constexpr int kNj = kC ? (kC + 7) / 8 : kRowRegs;
  const int nj = kRegs ? kNj : (C + 7) / 8;
  // class s + 8 j exists for this lane? (compile-time for all but the last step when kC is known)
  auto has = [&](int j) -> bool { return kC ? (8 * j + 7 < kC || s < kC - 8 * j) : (s + 8 * j < C); };
  // the lane's logits x[s], x[s + 8], ...: every load of the row in flight at once (kRegs), or re-read from
  // memory in each of the three passes (rows of more than 128 classes)
  float xv[kRegs ? kNj : 1];
  if (kRegs) {
#pragma unroll
    for (int j = 0; j < kNj; ++j) xv[j] = (on && has(j)) ? __ldg(x + 8 * j) : NEG;
  }
  auto at = [&](int j) -> float { return kRegs ? xv[kRegs ? j : 0] : ((on && has(j)) ? __ldg(x + 8 * j) : NEG); };
  float mx = 0.f, sum = 1.f;
  if (q.act_kind == SBOD_ACT_SOFTMAX) {
    float m = NEG;
#pragma unroll
    for (int j = 0; j < nj; ++j) m = fmaxf(m, at(j));
    mx = octet_max(m);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < nj; ++j)
      if (on && has(j)) acc += exp_term(at(j), mx);
    sum = octet_sum(acc);
  }
  // lowest probability that can still be emitted, with a margin, as a threshold on the logit
  const float floor_p = fmaxf(q.min_score, k_hi > kOneBits ? 0.f : __uint_as_float(kOneBits - k_hi)) * 0.999f;
  float thr_logit;
  if (q.act_kind == SBOD_ACT_SOFTMAX) thr_logit = mx + __logf(floor_p * sum) - 1e-3f;
  else if (q.act_kind == SBOD_ACT_SIGMOID) thr_logit = floor_p < 1.f ? __logf(floor_p / (1.f - floor_p)) - 1e-3f : -NEG;
  else thr_logit = floor_p;
  if (!(thr_logit == thr_logit)) thr_logit = NEG;  // (NaN sums: evaluate everything)
  auto prob = [&](float v, float mx_, float sum_) -> float {
    if (q.act_kind == SBOD_ACT_SOFTMAX) return __fdiv_rn(expf(v - mx_), sum_);
    if (q.act_kind == SBOD_ACT_SIGMOID) return __fdiv_rn(1.f, 1.f + expf(-v));
    return v;
  };
  if (q.agnostic) {
    // best foreground class of the row, first index among ties (torch.max, models/utils.py:135)
    unsigned long long best = 0ull;  // (probability bits : ~class): max = best probability, then lowest class
#pragma unroll
    for (int j = 0; j < nj; ++j) {
      const int k = s + 8 * j;
      if (!on || k == 0 || !has(j)) continue;
      const float v = at(j);
      if (!(v > thr_logit)) continue;
      const float pr = prob(v, mx, sum);
      if (pr > q.min_score) {
        const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(pr)) << 32) | (0xffffffffu - unsigned(k));
        if (key > best) best = key;
      }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      if (other > best) best = other;
    }
    if (on && s == 0 && best != 0ull) {
      const float pr = __uint_as_float(uint32_t(best >> 32));
      const int cls = int(0xffffffffu - uint32_t(best & 0xffffffffull));
      const uint32_t k32 = score_k32(pr);
      if (k32 >= k_lo && k32 < k_hi) {
        emit((static_cast<unsigned long long>(k32) << 32) | (1ull << kPriorBits) | static_cast<unsigned long long>(p));
        q.agn_label[size_t(n) * q.P + p] = cls;
      }
    }
    return;
  }
  auto finish = [&](float v, int k, int pp, float mx_, float sum_) {
    const float pr = prob(v, mx_, sum_);
    const uint32_t k32 = score_k32(pr);
    if (pr > q.min_score && k32 >= k_lo && k32 < k_hi)
      emit((static_cast<unsigned long long>(k32) << 32) | (static_cast<unsigned long long>(k) << kPriorBits) |
           static_cast<unsigned long long>(pp));
  };
  if (kRegs && wq) {
    // which of the lane's classes pass (bit j = class s + 8 j), and where they go in the warp's queue
    unsigned pm = 0;
#pragma unroll
    for (int j = 0; j < kNj; ++j) {
      const bool pass = on && (j > 0 || s != 0) && has(j) && xv[kRegs ? j : 0] > thr_logit;
      pm |= pass ? (1u << j) : 0u;
    }
    const int c = __popc(pm);
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    if (total == 0) return;        // (warp-uniform)
    if (total <= kRowQueue) {
      int pos = inc - c;
#pragma unroll
      for (int j = 0; j < kNj; ++j)
        if (pm & (1u << j))
          wq[pos++] = make_uint2(__float_as_uint(xv[kRegs ? j : 0]), unsigned(s + 8 * j) | (unsigned(lane >> 3) << 16));
      __syncwarp();
      for (int base = 0; base < total; base += 32) {
        const int e = base + lane;
        const uint2 ent = e < total ? wq[e] : make_uint2(0u, 0u);
        const int src = int(ent.y >> 16) * 8;  // the octet that owns the row
        const float mx_ = __shfl_sync(0xffffffffu, mx, src), sum_ = __shfl_sync(0xffffffffu, sum, src);
        const int pp = __shfl_sync(0xffffffffu, p, src);
        if (e < total) finish(__uint_as_float(ent.x), int(ent.y & 0xffffu), pp, mx_, sum_);
      }
      __syncwarp();
      return;
    }
    // more passing classes than the queue holds (rows of many near-equal scores): in place
#pragma unroll
    for (int j = 0; j < kNj; ++j)
      if (pm & (1u << j)) finish(xv[kRegs ? j : 0], s + 8 * j, p, mx, sum);
    return;
  }
#pragma unroll
  for (int j = 0; j < nj; ++j) {
    const int k = s + 8 * j;
    if (!on || k == 0 || !has(j)) continue;
    const float v = at(j);
    if (v > thr_logit) finish(v, k, p, mx, sum);
  }
}

template <int kC, typename Emit>
SBOD_DEVINL void eval_rows(const DetParams& q, int n, int p, int lane, uint32_t k_lo, uint32_t k_hi, Emit&& emit,
                           uint2* wq = nullptr) {
  if (kC || q.C <= 8 * kRowRegs) eval_rows_impl<kC, true>(q, n, p, lane, k_lo, k_hi, emit, wq);
  else eval_rows_impl<0, false>(q, n, p, lane, k_lo, k_hi, emit, wq);
}

// ------------------------------------------------------------------------------------------
// detect_refine_kernel: grid (ceil(P / kRefRows), N). Rows whose bound lies above the image's cutoff are
// compacted and evaluated exactly, a warp per row; their candidates above the cutoff go to the image's
// candidate list as sortable 64-bit keys  [0x3F800000 - score_bits : 32][class : 12][prior : 20]
// (ascending key = score descending, class ascending, prior ascending = the order in which the reference's
// per-class NMS + stable top-k sort consumes them) through a CTA-wide shared buffer: one shared-memory
// atomic per warp and row, one global atomic per CTA. The histogram of the keys' leading digit is
// accumulated in shared memory and flushed once.
// ------------------------------------------------------------------------------------------
#ifdef SBOD_DEBUG_HOOKS  // phase time stamps (profiling builds only)
__device__ unsigned long long g_det_times[16];
SBOD_DEVINL unsigned long long det_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define DET_STAMP_MIN(i) do { if (threadIdx.x == 0) atomicMin(&g_det_times[i], det_now()); } while (0)
#define DET_STAMP_MAX(i) do { if (threadIdx.x == 0) atomicMax(&g_det_times[i], det_now()); } while (0)
#else
#define DET_STAMP_MIN(i)
#define DET_STAMP_MAX(i)
#endif

constexpr int kRefThreads = 512;
constexpr int kRefMaxRows = 2048;  // prior rows per CTA (host picks 512 .. 2048 so that the grid is a few waves)
constexpr int kRefKeyBuf = kMaxBins / 2;  // keys buffered per CTA (shares its storage with the histogram copy)

template <int kC>  // kC > 0: compile-time class count
__global__ void __launch_bounds__(kRefThreads, 3) detect_refine_kernel(const DetParams q, int rows_per_cta) {
  __shared__ __align__(16) unsigned long long s_keys[kRefKeyBuf];
  __shared__ int s_rows[kRefMaxRows];
  __shared__ unsigned int s_nrows, s_ncand, s_gbase;
  __shared__ uint2 s_wq[kRefThreads / 32][kRowQueue];
  const int n = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // The CTAs of an image share its rows in interleaved blocks of 32 (block i belongs to CTA i mod gridDim.x):
  // rows with a high bound cluster (neighbouring priors look at the same object), contiguous slices would give a
  // few CTAs most of the work. The bounds of this CTA's rows are requested before anything else.
  constexpr int kRefIters = kRefMaxRows / kRefThreads;
  float pbv[kRefIters];
#pragma unroll
  for (int it = 0; it < kRefIters; ++it) {
    const int p = ((it * (kRefThreads / 32) + wid) * int(gridDim.x) + int(blockIdx.x)) * 32 + lane;
    pbv[it] = (it * kRefThreads < rows_per_cta && p < q.P) ? __ldcg(q.pbound + size_t(n) * q.P + p) : -1.f;
  }
  DET_STAMP_MIN(0);
  DET_STAMP_MAX(1);
  // The image's cutoff from the histogram of the row bounds (every CTA of the image computes the same value):
  // the smallest bin count cb such that at least rows_target rows have their bound in bins [0, cb). Four bins
  // per thread straight from L2 (rows of the histogram are kMaxBins long, unused bins stay zero).
  __shared__ unsigned int s_wt[kRefThreads / 32], s_cb;
  static_assert(kRefThreads * 4 == kMaxBins, "four bins per thread");
  const uint4 hc = __ldcg(reinterpret_cast<const uint4*>(q.rhist + size_t(n) * kMaxBins) + tid);
  if (tid == 0) {
    s_nrows = 0u;
    s_ncand = 0u;
    s_cb = unsigned(q.n_bins);
  }
  unsigned int cut;
  {
    const unsigned int mine = hc.x + hc.y + hc.z + hc.w;
    unsigned int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_wt[wid] = inc;
    __syncthreads();
    unsigned int before = inc - mine;
#pragma unroll
    for (int w = 0; w < kRefThreads / 32; ++w)
      if (w < wid) before += s_wt[w];
    const unsigned int need = unsigned(q.rows_target);
    if (before < need && before + mine >= need) {
      const unsigned int c4[4] = {hc.x, hc.y, hc.z, hc.w};
      unsigned int acc = before;
      bool found = false;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc += c4[j];
        if (!found && acc >= need) {
          s_cb = unsigned(4 * tid + j + 1);  // bins [0, 4 tid + j]
          found = true;
        }
      }
    }
    __syncthreads();
    const int cb = int(s_cb);
    cut = cb >= q.n_bins ? 0xffffffffu : unsigned(cb) << (q.shift0 - 32);
  }
  if (blockIdx.x == 0 && tid == 0) q.cutoff[n] = cut;  // the NMS kernel reads it instead of recomputing it
  DET_STAMP_MAX(2);
  // rows of this CTA whose bound is above the cutoff
#pragma unroll
  for (int it = 0; it < kRefIters; ++it) {
    if (it * kRefThreads >= rows_per_cta) break;  // (CTA-uniform)
    const int p = ((it * (kRefThreads / 32) + wid) * int(gridDim.x) + int(blockIdx.x)) * 32 + lane;
    const float b = pbv[it];
    const bool take = b > q.min_score && score_k32(b) < cut;  // (rows past the end carry -1)
    const unsigned bal = __ballot_sync(0xffffffffu, take);
    unsigned int base = 0;
    if (lane == 0 && bal) base = atomicAdd(&s_nrows, unsigned(__popc(bal)));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (take) s_rows[base + __popc(bal & ((1u << lane) - 1u))] = p;
  }
  __syncthreads();
  const int n_rows = int(s_nrows);
  DET_STAMP_MAX(3);
  unsigned long long* g_list = q.cand + size_t(n) * q.cand_cap;
  unsigned int* g_hist = q.hist + size_t(n) * kMaxBins;
  auto emit = [&](unsigned long long key) {
    const unsigned int at = atomicAdd(&s_ncand, 1u);
    if (at < unsigned(kRefKeyBuf)) {
      s_keys[at] = key;
    } else {  // the shared buffer is full (rare): straight to the global list
      const unsigned int slot = atomicAdd(&q.cand_count[n], 1u);
      if ((long long)slot < q.cand_cap) g_list[slot] = key;
    }
    atomicAdd(&g_hist[k32_bin(q, uint32_t(key >> 32))], 1u);  // a few dozen keys per CTA: straight to L2
  };
  // four rows per warp (an octet each)
  for (int i = wid * 4; i < n_rows; i += (kRefThreads / 32) * 4) {
    const int r = i + (lane >> 3);
    eval_rows<kC>(q, n, r < n_rows ? s_rows[r] : -1, lane, 0u, cut, emit, s_wq[wid]);
  }
  __syncthreads();
  DET_STAMP_MAX(4);
  // hand the buffered keys over: one global atomic for the CTA's keys
  const unsigned int have = min(s_ncand, unsigned(kRefKeyBuf));
  if (have) {  // (CTA-uniform)
    if (tid == 0) s_gbase = atomicAdd(&q.cand_count[n], have);
    __syncthreads();
    for (unsigned int i = tid; i < have; i += kRefThreads) {
      const long long slot = (long long)s_gbase + i;
      if (slot < q.cand_cap) g_list[slot] = s_keys[i];
    }
  }
  DET_STAMP_MAX(5);
}

// ------------------------------------------------------------------------------------------
// detect_nms_kernel
// ------------------------------------------------------------------------------------------
SBOD_DEVINL int key_class(unsigned long long k) { return int((k >> kPriorBits) & ((1u << kClassBits) - 1u)); }
SBOD_DEVINL int key_prior(unsigned long long k) { return int(k & ((1u << kPriorBits) - 1u)); }
SBOD_DEVINL float key_score(unsigned long long k) { return __uint_as_float(kOneBits - uint32_t(k >> 32)); }

SBOD_DEVINL int64_t out_label(const DetParams& q, int n, unsigned long long k) {
  return q.agnostic ? int64_t(q.agn_label[size_t(n) * q.P + key_prior(k)]) : int64_t(key_class(k));
}

// torchvision nms criterion: inter / (area_i + area_j - inter) > thr
SBOD_DEVINL bool overlaps(const float4 a, const float4 b, float thr) {
  const float aa = box_area_rn(a), ab = box_area_rn(b);
  return iou_plain_rn(a, aa, b, ab) > thr;
}

constexpr int kClsGrouped = 128;  // up to this many classes the chunk is grouped by class in one pass
constexpr int kKeptGrouped = 1024; // ... and the kept list too, while it is not longer than this

struct NmsSmem {
  unsigned int hist[kMaxBins];
  unsigned long long ckey[kChunk];
  float4 cbox[kChunk];
  uint16_t cidx[kChunk];   // chunk positions grouped by class (segments in key order)
  uint16_t cnew[kChunk];   // per class segment: positions kept in this round
  uint8_t cflag[kChunk];   // 1 = survives stage 1, 2 = survives stage 2 as well
  unsigned int wscan[40];
  uint16_t big_seg0[kChunk / kBigSeg + 1];  // large class segments of the round (start, length)
  uint16_t big_len[kChunk / kBigSeg + 1];
  int misc[16];
  uint16_t wcnt[(kChunk / 32) * (kClsGrouped + 1)];  // per warp and class: entries of the class in the warp's 32 chunk positions
  unsigned int kc_off[kClsGrouped + 3];  // kept list grouped by class: segment bounds ...
  uint16_t kc_idx[kKeptGrouped];         // ... and kept-list positions
};

// radix levels below the level-0 digit: 11 bits each
SBOD_DEVINL int level_shift(const DetParams& q, int level) {
  const int s = q.shift0 - 11 * level;
  return s > 0 ? s : 0;
}
SBOD_DEVINL int level_bins(const DetParams& q, int level) {
  if (level == 0) return q.n_bins;
  const int hi = level_shift(q, level - 1), lo = level_shift(q, level);
  return 1 << (hi - lo);
}

__global__ void __launch_bounds__(kNmsThreads) detect_nms_kernel(const DetParams q) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  NmsSmem& S = *reinterpret_cast<NmsSmem*>(smem_raw);
  unsigned char* dyn = smem_raw + ((sizeof(NmsSmem) + 127) & ~size_t(127));
  // kept list: in shared memory; moved to the global spill buffer if the first stage of detect_tools keeps more
  unsigned long long* kkey = reinterpret_cast<unsigned long long*>(dyn);          // [kcap]
  float4* kbox = reinterpret_cast<float4*>(dyn + size_t(q.kcap) * 8);             // [kcap]
  uint8_t* kst2 = reinterpret_cast<uint8_t*>(dyn + size_t(q.kcap) * 24);          // [kcap]
  unsigned int* cls_off = reinterpret_cast<unsigned int*>(dyn + size_t(q.kcap) * 25 + 128 - (size_t(q.kcap) * 25) % 128);  // [C+1]

  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  DET_STAMP_MIN(8);
  // same function of the same histogram as detect_refine_kernel: candidates below `cut` are complete
  {
    unsigned int* rh = q.rhist + size_t(n) * kMaxBins;
    for (int b = tid; b < q.n_bins; b += kNmsThreads) rh[b] = 0u;  // clean for the next call
  }
  unsigned int cut = q.cutoff[n];  // candidates below `cut` are complete (detect_refine_kernel)
  const bool two_stage = q.second_thr >= 0.f;
  unsigned int* g_hist = q.hist + size_t(n) * kMaxBins;
  unsigned int* g_seen = q.class_seen + size_t(n) * q.C;
  unsigned long long* g_cand = q.cand + size_t(n) * q.cand_cap;
  unsigned int raw_total = q.cand_count[n];
  long long total = raw_total < (unsigned long long)q.cand_cap ? raw_total : q.cand_cap;
  bool overflow = raw_total > (unsigned long long)q.cand_cap;
  int kcap_now = q.kcap;  // capacity of the kept list (shared memory; the spill buffer once it has been switched to)

  int kept_n = 0;   // stage-1 survivors so far (kept list length)
  int kept2_n = 0;  // stage-2 survivors so far
  int status = 0;   // 1 = kept-list capacity exceeded
  unsigned long long lo_key = 0ull;
  bool exhausted = total == 0;
  const int stop_at = q.top_k + 1;

  // The first round only takes about twice as many candidates as boxes are wanted (a power of two):
  // when suppression is moderate that is enough, and sorting / decoding / grouping a short chunk is
  // cheaper; if it is not enough the following rounds take full chunks.
  int chunk_cap = 64;
  if (!two_stage) {
    while (4 * chunk_cap < 5 * stop_at && chunk_cap < kChunk) chunk_cap <<= 1;  // >= 1.25 x the boxes wanted
  } else {
    // two suppression stages thin the chunk twice, and their cost grows with the square of the survivors
    while (2 * chunk_cap < 5 * stop_at && chunk_cap < kChunk) chunk_cap <<= 1;  // >= 2.5 x the boxes wanted
  }
  for (;;) {  // bands: first the candidates above the cutoff, then (rarely) everything else
  while (!exhausted && (two_stage ? kept2_n : kept_n) < stop_at && !status) {
    // ---- choose [lo_key, hi_key) holding at most chunk_cap candidates ----------------------
    int level = 0;
    while (level_shift(q, level) > 0 && (lo_key & ((1ull << level_shift(q, level)) - 1ull)) != 0ull)
      ++level;
    unsigned long long hi_key = 0ull;
    bool have_hi = false;
    while (!have_hi) {
      const int sh = level_shift(q, level);
      const int nb = level_bins(q, level);
      const int dlo = (level == 0) ? int(lo_key >> sh) : int((lo_key >> sh) & (unsigned(nb) - 1u));
      // node = keys sharing lo_key's bits above this level's digit
      const int up = (level == 0) ? 64 : level_shift(q, level - 1);
      const unsigned long long node_base = (up >= 64) ? 0ull : (lo_key >> up) << up;
      if (level == 0) {
        for (int b = tid; b < nb; b += kNmsThreads) S.hist[b] = __ldcg(&g_hist[b]);  // (L2: the second band adds to it)
      } else {
        for (int b = tid; b < nb; b += kNmsThreads) S.hist[b] = 0u;
        __syncthreads();
        for (long long i = tid; i < total; i += kNmsThreads) {
          const unsigned long long k = g_cand[i];
          if (k >= lo_key && (up >= 64 || (k >> up) == (lo_key >> up)))
            atomicAdd(&S.hist[(k >> sh) & (unsigned(nb) - 1u)], 1u);
        }
      }
      __syncthreads();
      // d_end = first digit >= dlo whose inclusive running count exceeds kChunk (block-wide scan,
      // two bins per thread); first_nonempty = first digit >= dlo with a non-zero count.
      {
        if (tid == 0) {
          S.misc[0] = nb;  // d_end
          S.misc[4] = nb;  // first non-empty
        }
        const int b0 = 2 * tid, b1 = 2 * tid + 1;
        const unsigned int h0 = (b0 >= dlo && b0 < nb) ? S.hist[b0] : 0u;
        const unsigned int h1 = (b1 >= dlo && b1 < nb) ? S.hist[b1] : 0u;
        unsigned int inc = h0 + h1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        if (lane == 31) S.wscan[wid] = inc;
        __syncthreads();
        unsigned int before = inc - (h0 + h1);
        for (int w = 0; w < wid; ++w) before += S.wscan[w];
        if (h0 && before + h0 > unsigned(chunk_cap)) atomicMin(&S.misc[0], b0);
        else if (h1 && before + h0 + h1 > unsigned(chunk_cap)) atomicMin(&S.misc[0], b1);
        if (h0) atomicMin(&S.misc[4], b0);
        else if (h1) atomicMin(&S.misc[4], b1);
        __syncthreads();
        if (tid == 0) S.misc[1] = (S.misc[0] == S.misc[4] && S.misc[0] < nb) ? 1 : 0;  // descend
      }
      __syncthreads();
      const int d_end = S.misc[0];
      const bool descend = S.misc[1] != 0;
      __syncthreads();
      if (descend) {
        // restart at the oversized digit, one level down (its keys all share this digit)
        lo_key = lo_key > (node_base + (static_cast<unsigned long long>(d_end) << sh))
                     ? lo_key
                     : node_base + (static_cast<unsigned long long>(d_end) << sh);
        ++level;
        continue;
      }
      if (d_end >= nb) {
        if (level == 0) { hi_key = ~0ull; exhausted = true; }   // everything that is left fits
        else hi_key = node_base + (1ull << up);                  // end of this node
        if (level != 0 && hi_key == 0ull) { hi_key = ~0ull; exhausted = true; }  // wrapped
      } else {
        hi_key = node_base + (static_cast<unsigned long long>(d_end) << sh);
      }
      have_hi = true;
    }

    DET_STAMP_MAX(9);
    // ---- gather the chunk -----------------------------------------------------------------
    if (tid == 0) S.misc[3] = 0;
    __syncthreads();
    // eight independent 8-byte loads in flight per thread (the list lives in L2 / HBM)
    for (long long r0 = 0; r0 < total; r0 += 8LL * kNmsThreads) {  // CTA-uniform trip count (ballots inside)
      const long long i0 = r0 + tid;
      unsigned long long kk[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long i = i0 + (long long)u * kNmsThreads;
        kk[u] = i < total ? g_cand[i] : ~0ull;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const unsigned long long k = kk[u];
        const bool in = k != ~0ull && k >= lo_key && (k < hi_key || hi_key == ~0ull);
        const unsigned bal = __ballot_sync(0xffffffffu, in);  // (the loop bounds are warp-uniform)
        if (bal) {
          int base = 0;
          if (lane == __ffs(bal) - 1) base = atomicAdd(&S.misc[3], __popc(bal));
          base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
          const int slot = base + __popc(bal & ((1u << lane) - 1u));
          if (in && slot < chunk_cap) S.ckey[slot] = k;
        }
      }
    }
    __syncthreads();
    const int m = min(S.misc[3], chunk_cap);
    lo_key = hi_key;
    chunk_cap = kChunk;  // later rounds take full chunks
    if (m == 0) {
      __syncthreads();
      continue;
    }
    DET_STAMP_MAX(10);
    // ---- sort by key (bitonic, padded with ~0) ----------------------------------------------
    // Each thread keeps its element in a register; partners less than 32 apart are reached by
    // shuffles (40 of the 55 steps, no block barrier), the others through shared memory, alternating
    // between ckey and the not yet used cbox array so that one barrier per step suffices.
    if (m <= 256) {
      // Short chunks (the usual first round): rank sort. Four threads per key count the keys below it (keys are
      // unique: one per (class, prior)), interleaved so that a warp's reads fall into different banks; the key then
      // goes straight to its rank. Two barriers instead of the bitonic network's six plus thirty shuffle steps.
      unsigned long long* bufB = reinterpret_cast<unsigned long long*>(S.cbox);
      __syncthreads();
      const int i = tid >> 2, part = tid & 3;
      const unsigned long long my = i < m ? S.ckey[i] : ~0ull;
      int below = 0;
      for (int j = part; j < m; j += 4) below += S.ckey[j] < my ? 1 : 0;
      below += __shfl_xor_sync(0xffffffffu, below, 1);
      below += __shfl_xor_sync(0xffffffffu, below, 2);
      if (part == 0 && i < m) bufB[below] = my;
      __syncthreads();
      const unsigned long long v = tid < m ? bufB[tid] : ~0ull;
      __syncthreads();  // (bufB is the box array of the next phase)
      S.ckey[tid] = v;
      __syncthreads();
    } else {
      unsigned long long v = tid < m ? S.ckey[tid] : ~0ull;
      unsigned long long* bufA = S.ckey;
      unsigned long long* bufB = reinterpret_cast<unsigned long long*>(S.cbox);
      __syncthreads();
      int sort_n = 2;
      while (sort_n < m) sort_n <<= 1;
      const bool in_net = (tid & ~31) < sort_n;  // warps beyond the network only keep the barriers company
      for (int k2 = 2; k2 <= sort_n; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
          unsigned long long o = 0ull;
          if (j >= 32) {
            if (in_net) bufA[tid] = v;
            __syncthreads();
            if (in_net) o = bufA[tid ^ j];
            unsigned long long* t = bufA;
            bufA = bufB;
            bufB = t;
          } else if (in_net) {
            o = __shfl_xor_sync(0xffffffffu, v, j);
          }
          if (in_net) {
            const bool up_dir = (tid & k2) == 0, lower = (tid & j) == 0;
            const unsigned long long mn = v < o ? v : o, mx = v < o ? o : v;
            v = (lower == up_dir) ? mn : mx;
          }
        }
      }
      __syncthreads();  // every read of the exchange buffers is done
      S.ckey[tid] = v;
      __syncthreads();
    }
    DET_STAMP_MAX(11);
    // ---- decode boxes, group by class -------------------------------------------------------
    for (int c = tid; c <= q.C; c += kNmsThreads) cls_off[c] = 0u;
    if (tid == 0) {
      S.misc[6] = 1;  // next class to process (dynamic assignment to warps)
      S.misc[8] = 0;  // number of large class segments of this round
    }
    __syncthreads();
    if (tid < m) {
      const unsigned long long k = S.ckey[tid];
      S.cbox[tid] = decode_box(q, n, key_prior(k));
      S.cflag[tid] = 0;
      atomicAdd(&cls_off[key_class(k) + 1], 1u);
    }
    __syncthreads();
    // exclusive scan of the class histogram (C+1 entries), by warp 0
    if (wid == 0) {
      unsigned int carry = 0;
      for (int base = 0; base <= q.C; base += 32) {
        const int c = base + lane;
        const unsigned int v = c <= q.C ? cls_off[c] : 0u;
        unsigned int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        if (c <= q.C) cls_off[c] = carry + inc;  // inclusive over the shifted histogram == start of class c+... see below
        carry += __shfl_sync(0xffffffffu, inc, 31);
      }
    }
    __syncthreads();
    // cls_off[c] now holds the number of chunk entries with class < c ... (entry c+1 counted class c),
    // i.e. segment of class c is [cls_off[c], cls_off[c+1]).
    // Chunk positions grouped by class, key order inside a class: position = segment start + entries
    // of the class in earlier warps + rank among the warp's own peers (match_any), one pass for all
    // classes instead of one scan of the chunk per class.
    const bool grouped = q.C <= kClsGrouped;
    if (grouped) {
      const int cw = q.C + 1;
      for (int i = tid; i < (kChunk / 32) * cw; i += kNmsThreads) S.wcnt[i] = 0;
      __syncthreads();
      const int c = tid < m ? key_class(S.ckey[tid]) : q.C;  // padding joins a dummy class
      const unsigned peers = __match_any_sync(0xffffffffu, c);
      const int r = __popc(peers & ((1u << lane) - 1u));
      if (r == 0) S.wcnt[wid * cw + c] = uint16_t(__popc(peers));
      __syncthreads();
      if (tid < m) {
        unsigned int before = 0;
        for (int w2 = 0; w2 < wid; ++w2) before += S.wcnt[w2 * cw + c];
        S.cidx[cls_off[c] + before + r] = uint16_t(tid);
      }
      // (visible to the class warps after the barrier that ends stage 1a)
    }

    DET_STAMP_MAX(12);
    // ---- stage 1a: every candidate against the boxes kept in EARLIER rounds, one thread each ----
    // (parallel over the whole CTA whatever the number of classes; flag 3 = already suppressed).
    // With few classes the kept list is first grouped by class (counting sort in shared memory), so a
    // candidate only meets the kept boxes of its own class instead of walking the whole list.
    if (grouped && kept_n > 0 && kept_n <= kKeptGrouped) {
      for (int c = tid; c <= q.C + 1; c += kNmsThreads) S.kc_off[c] = 0;
      __syncthreads();
      for (int i = tid; i < kept_n; i += kNmsThreads) atomicAdd(&S.kc_off[key_class(kkey[i]) + 2], 1u);
      __syncthreads();
      if (wid == 0) {  // kc_off[c + 1] = kept boxes with class < c (filled as a cursor below), C + 2 entries
        unsigned int carry = 0;
        for (int base = 0; base <= q.C + 1; base += 32) {
          const int c = base + lane;
          const unsigned int v = c <= q.C + 1 ? S.kc_off[c] : 0u;
          unsigned int inc = v;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
          }
          if (c <= q.C + 1) S.kc_off[c] = carry + inc;
          carry += __shfl_sync(0xffffffffu, inc, 31);
        }
      }
      __syncthreads();
      // now kc_off[c + 1] = start of class c; scatter advances it to the end of class c = start of c + 1,
      // so afterwards class c occupies [kc_off[c], kc_off[c + 1])
      for (int i = tid; i < kept_n; i += kNmsThreads) {
        const unsigned int at = atomicAdd(&S.kc_off[key_class(kkey[i]) + 1], 1u);
        S.kc_idx[at] = uint16_t(i);
      }
      __syncthreads();
      if (tid < m) {
        const int c = key_class(S.ckey[tid]);
        const float4 be = S.cbox[tid];
        bool sup = false;
        for (unsigned int r = S.kc_off[c]; r < S.kc_off[c + 1] && !sup; ++r)
          if (overlaps(kbox[S.kc_idx[r]], be, q.max_overlap)) sup = true;
        if (sup) S.cflag[tid] = 3;
      }
    } else if (tid < m && kept_n > 0) {
      const int c = key_class(S.ckey[tid]);
      const float4 be = S.cbox[tid];
      bool sup = false;
      for (int i = 0; i < kept_n && !sup; ++i)
        if (key_class(kkey[i]) == c && overlaps(kbox[i], be, q.max_overlap)) sup = true;
      if (sup) S.cflag[tid] = 3;
    }
    __syncthreads();
    DET_STAMP_MAX(13);
    // ---- stage 1b: greedy suppression inside the chunk -------------------------------------------
    // Small class segments (<= kBigSeg candidates): one warp per class walks its segment and tests
    // each candidate against the boxes it kept so far (warp ballot).
    // Large segments (few classes, dense scenes): the pairwise suppression bits of the segment are
    // built by ALL warps (one warp per (row, 32-column word), upper triangle only) and one warp then
    // resolves the greedy order by scanning the bit rows — the classic bitmask NMS, per class.
    for (;;) {
      int c = 0;
      if (lane == 0) c = atomicAdd(&S.misc[6], 1);
      c = __shfl_sync(0xffffffffu, c, 0);
      if (c >= q.C) break;
      const unsigned int seg0 = cls_off[c], seg1 = cls_off[c + 1];
      if (seg1 == seg0) continue;
      if (!grouped) {  // many classes: fill the segment with this class's chunk positions in key order
        unsigned int w = seg0;
        for (int base = 0; base < m; base += 32) {
          const int i = base + lane;
          const bool mine = i < m && key_class(S.ckey[i]) == c;
          const unsigned bal = __ballot_sync(0xffffffffu, mine);
          if (mine) S.cidx[w + __popc(bal & ((1u << lane) - 1u))] = uint16_t(i);
          w += __popc(bal);
        }
        __syncwarp();
      }
      unsigned int lim = seg1;               // pre-NMS per-class cap: later candidates are dropped
      if (q.pre_nms_topk > 0) {
        const unsigned int seen = g_seen[c];  // candidates of this class consumed by earlier rounds (global
        if (lane == 0) g_seen[c] = seen + (seg1 - seg0);  // memory: only touched when the cap is in use)
        const unsigned int room = seen >= unsigned(q.pre_nms_topk) ? 0u : unsigned(q.pre_nms_topk) - seen;
        lim = min(seg1, seg0 + room);
      }
      if (lim - seg0 > unsigned(kBigSeg)) {  // leave it to the bitmask path below
        if (lane == 0) {
          const int slot = atomicAdd(&S.misc[8], 1);
          S.big_seg0[slot] = uint16_t(seg0);
          S.big_len[slot] = uint16_t(lim - seg0);
        }
        continue;
      }
      if (lim - seg0 <= 32u) {
        // up to 32 candidates (the usual case): lane i owns candidate i. All pairwise tests run in
        // parallel - bit j of sup_by says "candidate j (earlier in key order) overlaps me" - and the
        // greedy order is then resolved on the bit masks alone.
        const int L = int(lim - seg0);
        const int e = lane < L ? int(S.cidx[seg0 + lane]) : -1;
        const bool alive = e >= 0 && S.cflag[e] != 3;  // 3: suppressed by a box of an earlier round
        const float4 b = e >= 0 ? S.cbox[e] : make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned int sup_by = 0u;
        for (int j = 0; j + 1 < L; ++j) {
          float4 bj;
          bj.x = __shfl_sync(0xffffffffu, b.x, j);
          bj.y = __shfl_sync(0xffffffffu, b.y, j);
          bj.z = __shfl_sync(0xffffffffu, b.z, j);
          bj.w = __shfl_sync(0xffffffffu, b.w, j);
          if (lane > j && lane < L && overlaps(bj, b, q.max_overlap)) sup_by |= 1u << j;
        }
        const unsigned int alive_mask = __ballot_sync(0xffffffffu, alive);
        unsigned int kept = 0u;
        for (int i = 0; i < L; ++i) {
          const unsigned int mi = __shfl_sync(0xffffffffu, sup_by, i);
          if (((alive_mask >> i) & 1u) && !(mi & kept)) kept |= 1u << i;
        }
        if (alive && ((kept >> lane) & 1u)) S.cflag[e] = 1;
        continue;
      }
      unsigned int nnew = 0;
      for (unsigned int r = seg0; r < lim; ++r) {
        const int e = S.cidx[r];
        if (S.cflag[e] == 3) continue;  // suppressed by a box of an earlier round (stage 1a)
        const float4 be = S.cbox[e];
        bool sup = false;
        for (unsigned int i = lane; i < nnew && !sup; i += 32)
          if (overlaps(S.cbox[S.cnew[seg0 + i]], be, q.max_overlap)) sup = true;
        if (!__any_sync(0xffffffffu, sup)) {
          if (lane == 0) {
            S.cnew[seg0 + nnew] = uint16_t(e);
            S.cflag[e] = 1;
          }
          ++nnew;
          __syncwarp();
        }
      }
    }
    __syncthreads();
    const int n_big = S.misc[8];
    if (n_big > 0) {
      unsigned int* gmask = q.nms_mask + size_t(n) * (kChunk * (kChunk / 32));  // row r: words [r*32, r*32+32)
      // (1) suppression bits, upper triangle: task = (segment, row, word)
      for (int b = 0; b < n_big; ++b) {
        const int seg0 = S.big_seg0[b], L = S.big_len[b];
        const int W = (L + 31) >> 5;
        for (int task = wid; task < L * W; task += kNmsThreads / 32) {
          const int row = task / W, word = task - row * W;
          if (word < (row >> 5)) continue;  // below the diagonal
          const int col = word * 32 + lane;
          const int er = S.cidx[seg0 + row];
          bool bit = false;
          if (col > row && col < L && S.cflag[er] != 3) {
            const int ec = S.cidx[seg0 + col];
            bit = S.cflag[ec] != 3 && overlaps(S.cbox[er], S.cbox[ec], q.max_overlap);
          }
          const unsigned bal = __ballot_sync(0xffffffffu, bit);
          if (lane == 0) gmask[size_t(seg0 + row) * 32 + word] = bal;
        }
      }
      __threadfence_block();
      __syncthreads();
      // (2) greedy order: one warp per big segment; lane w owns word w of the "removed" set
      for (int b = wid; b < n_big; b += kNmsThreads / 32) {
        const int seg0 = S.big_seg0[b], L = S.big_len[b];
        const int W = (L + 31) >> 5;
        unsigned int removed = 0u;
        for (int r0 = 0; r0 < L; r0 += 8) {  // eight bit rows in flight
          unsigned int rows[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int r = r0 + u;
            rows[u] = (r < L && lane < W && lane >= (r >> 5)) ? gmask[size_t(seg0 + r) * 32 + lane] : 0u;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int r = r0 + u;
            if (r >= L) break;
            const int e = S.cidx[seg0 + r];
            const unsigned int remw = __shfl_sync(0xffffffffu, removed, r >> 5);
            const bool dead = ((remw >> (r & 31)) & 1u) || S.cflag[e] == 3;
            if (!dead) {
              if (lane == 0) S.cflag[e] = 1;
              removed |= rows[u];
            }
          }
        }
      }
      __syncthreads();
    }

    DET_STAMP_MAX(15);
    // ---- stage 2 (detect_tools): class-agnostic NMS over the stage-1 survivors, key order ----
    // (a) every stage-1 survivor against the stage-2 survivors of EARLIER rounds, one thread each;
    // (b) the remaining ones are compacted in key order; (c) their pairwise suppression bits (upper triangle)
    // are built by all warps; (d) one warp resolves the greedy order on the bit rows - the same decisions as
    // walking the candidates one by one, without the walk (one warp testing ~700 survivors against each other
    // took 3 ms on a dense scene).
    if (two_stage) {
      {
        const int e = tid;
        if (e < m && S.cflag[e] == 1) {
          const float4 be = S.cbox[e];
          bool sup = false;
          for (int i = 0; i < kept_n && !sup; ++i)
            if (kst2[i] && overlaps(kbox[i], be, q.second_thr)) sup = true;
          if (sup) S.cflag[e] = 4;  // stage-1 survivor that an earlier round's box suppresses in stage 2
        }
      }
      __syncthreads();
      int s2 = 0;
      {
        const bool live = tid < m && S.cflag[tid] == 1;
        const unsigned bal = __ballot_sync(0xffffffffu, live);
        if (lane == 0) S.wscan[wid] = unsigned(__popc(bal));
        __syncthreads();
        int before = 0;
        for (int w = 0; w < kNmsThreads / 32; ++w) {
          const int c = int(S.wscan[w]);
          if (w < wid) before += c;
          s2 += c;
        }
        if (live) S.cidx[before + __popc(bal & ((1u << lane) - 1u))] = uint16_t(tid);
        if (tid < m && S.cflag[tid] == 4) S.cflag[tid] = 1;  // still a stage-1 survivor for the kept list
        __syncthreads();
      }
      if (s2 > 0) {  // (CTA-uniform)
        unsigned int* gmask = q.nms_mask + size_t(n) * (kChunk * (kChunk / 32));  // row r: words [r*32, r*32+32)
        const int W = (s2 + 31) >> 5;
        for (int task = wid; task < s2 * W; task += kNmsThreads / 32) {
          const int row = task / W, word = task - row * W;
          if (word < (row >> 5)) continue;  // below the diagonal
          const int col = word * 32 + lane;
          bool bit = false;
          if (col > row && col < s2) bit = overlaps(S.cbox[S.cidx[row]], S.cbox[S.cidx[col]], q.second_thr);
          const unsigned bal = __ballot_sync(0xffffffffu, bit);
          if (lane == 0) gmask[size_t(row) * 32 + word] = bal;
        }
        __threadfence_block();
        __syncthreads();
        if (wid == 0) {  // lane w owns word w of the "removed" set
          unsigned int removed = 0u;
          int nnew2 = 0;
          for (int r0 = 0; r0 < s2; r0 += 8) {  // eight bit rows in flight
            unsigned int rows[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int r = r0 + u;
              rows[u] = (r < s2 && lane < W && lane >= (r >> 5)) ? gmask[size_t(r) * 32 + lane] : 0u;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int r = r0 + u;
              if (r >= s2) break;
              const unsigned int remw = __shfl_sync(0xffffffffu, removed, r >> 5);
              if (!((remw >> (r & 31)) & 1u)) {
                if (lane == 0) S.cflag[S.cidx[r]] = 2;
                ++nnew2;
                removed |= rows[u];
              }
            }
            if (kept2_n + nnew2 >= stop_at) break;  // enough survivors: the rest of the chunk cannot reach the output
          }
          if (lane == 0) S.misc[5] = nnew2;
        }
      } else if (tid == 0) {
        S.misc[5] = 0;
      }
      __syncthreads();
    }

    DET_STAMP_MAX(14);
    // ---- append the survivors to the kept list, in key order ---------------------------------
    const bool keep_me = tid < m && (S.cflag[tid] == 1 || S.cflag[tid] == 2);
    const unsigned bal = __ballot_sync(0xffffffffu, keep_me);
    if (lane == 0) S.wscan[wid] = __popc(bal);
    __syncthreads();
    int pos = kept_n + __popc(bal & ((1u << lane) - 1u));
    int add = 0;
    for (int w = 0; w < kNmsThreads / 32; ++w) {
      if (w < wid) pos += S.wscan[w];
      add += S.wscan[w];
    }
    if (kept_n + add > kcap_now && kcap_now == q.kcap && q.spill_key && kept_n + add <= q.spill_cap) {
      // detect_tools only: the first stage kept more boxes than fit in shared memory -> the kept list moves to
      // the global spill buffer (slower, exact; the reference has no limit here)
      unsigned long long* gk = q.spill_key + size_t(n) * q.spill_cap;
      float4* gb = q.spill_box + size_t(n) * q.spill_cap;
      uint8_t* gs = q.spill_st2 + size_t(n) * q.spill_cap;
      for (int i = tid; i < kept_n; i += kNmsThreads) {
        gk[i] = kkey[i];
        gb[i] = kbox[i];
        gs[i] = kst2[i];
      }
      __syncthreads();
      kkey = gk;
      kbox = gb;
      kst2 = gs;
      kcap_now = q.spill_cap;
    }
    if (kept_n + add > kcap_now) {
      status = 1;
    } else {
      if (keep_me) {
        kkey[pos] = S.ckey[tid];
        kbox[pos] = S.cbox[tid];
        kst2[pos] = S.cflag[tid] == 2 ? 1 : 0;
      }
      kept_n += add;
      if (two_stage) kept2_n += S.misc[5];
    }
    __syncthreads();
  }

    // The candidates above the cutoff ran out before top_k+1 boxes survived: the answer needs lower-scored
    // candidates. Second band, exact: this CTA evaluates every row of its image and emits what lies below the
    // cutoff (rare; the bound pass sized the first band for ~4x the boxes wanted).
    if (exhausted && !status && !overflow && cut != 0xffffffffu && (two_stage ? kept2_n : kept_n) < stop_at) {
      auto emit = [&](unsigned long long key) {
        const unsigned int slot = atomicAdd(&q.cand_count[n], 1u);
        if ((long long)slot < q.cand_cap) g_cand[slot] = key;
        atomicAdd(&g_hist[k32_bin(q, uint32_t(key >> 32))], 1u);
      };
      for (int p0 = wid * 4; p0 < q.P; p0 += (kNmsThreads / 32) * 4) {  // four rows per warp (an octet each)
        const int p = p0 + (lane >> 3);
        const bool take = p < q.P && q.pbound[size_t(n) * q.P + p] > q.min_score;
        eval_rows<0>(q, n, take ? p : -1, lane, cut, 0xffffffffu, emit);
      }
      __threadfence();
      __syncthreads();
      raw_total = *reinterpret_cast<volatile unsigned int*>(&q.cand_count[n]);
      total = raw_total < (unsigned long long)q.cand_cap ? raw_total : q.cand_cap;
      overflow = raw_total > (unsigned long long)q.cand_cap;
      lo_key = static_cast<unsigned long long>(cut) << 32;  // every key of the second band is >= this
      cut = 0xffffffffu;
      exhausted = false;
      __syncthreads();
      continue;
    }
    break;
  }

  // ---- leave the workspace clean for the next call ----------------------------------------
  __syncthreads();
  for (int b = tid; b < q.n_bins; b += kNmsThreads) g_hist[b] = 0u;
  for (int c = tid; c < q.C; c += kNmsThreads) g_seen[c] = 0u;
  if (tid == 0) q.cand_count[n] = 0u;

  // ---- emit --------------------------------------------------------------------------------
  float4* ob = reinterpret_cast<float4*>(q.out_boxes) + size_t(n) * q.out_cap;
  int64_t* ol = q.out_labels + size_t(n) * q.out_cap;
  float* os = q.out_scores + size_t(n) * q.out_cap;
  int32_t* op = q.out_prior + size_t(n) * q.out_cap;
  if (status || overflow) {
    if (tid == 0) q.out_counts[n] = overflow ? -2 : -1;  // host raises: capacity exceeded
    return;
  }
  if (kept_n == 0) {  // placeholder for 'background', models/utils.py:274-277
    if (tid == 0) {
      ob[0] = make_float4(0.f, 0.f, 1.f, 1.f);
      ol[0] = 0;
      os[0] = 0.f;
      op[0] = -1;
      q.out_counts[n] = 1;
    }
    return;
  }
  if (!two_stage) {
    if (kept_n > q.top_k) {
      // more than top_k survivors: the reference sorts by score (stable) and keeps the first top_k
      for (int i = tid; i < q.top_k; i += kNmsThreads) {
        const unsigned long long k = kkey[i];
        ob[i] = kbox[i]; ol[i] = out_label(q, n, k); os[i] = key_score(k); op[i] = key_prior(k);
      }
      if (tid == 0) q.out_counts[n] = q.top_k;
    } else {
      // everything survives: class-major, NMS order inside a class (models/utils.py:245-281)
      for (int i = tid; i < kept_n; i += kNmsThreads) {
        const unsigned long long k = kkey[i];
        const int c = key_class(k);
        int dst = 0;
        for (int j = 0; j < kept_n; ++j) {
          const int cj = key_class(kkey[j]);
          dst += (cj < c || (cj == c && j < i)) ? 1 : 0;
        }
        ob[dst] = kbox[i]; ol[dst] = out_label(q, n, k); os[dst] = key_score(k); op[dst] = key_prior(k);
      }
      if (tid == 0) q.out_counts[n] = kept_n;
    }
  } else {
    // detect_tools.py:200-212: survivors of the second NMS in score order; truncated to top_k only
    // when the FIRST stage produced more than top_k boxes.
    const bool truncate = kept_n > q.top_k;
    if (wid == 0) {
      int w = 0;
      for (int base = 0; base < kept_n; base += 32) {
        const int i = base + lane;
        const bool s2 = i < kept_n && kst2[i];
        const unsigned bal = __ballot_sync(0xffffffffu, s2);
        const int dst = w + __popc(bal & ((1u << lane) - 1u));
        if (s2 && (!truncate || dst < q.top_k) && dst < q.out_cap) {
          const unsigned long long k = kkey[i];
          ob[dst] = kbox[i]; ol[dst] = out_label(q, n, k); os[dst] = key_score(k); op[dst] = key_prior(k);
        }
        w += __popc(bal);
      }
      if (lane == 0) q.out_counts[n] = truncate ? min(w, q.top_k) : min(w, q.out_cap);
    }
  }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static void level0_layout(float min_score, int* shift0, int* n_bins) {
  // keys: (0x3F800000 - score_bits) << 32 | ...; score in (min_score, 1]
  uint32_t minbits = 0;
  float ms = min_score > 0.f ? min_score : 0.f;
  memcpy(&minbits, &ms, 4);
  if (minbits > kOneBits) minbits = kOneBits;
  const uint32_t range = kOneBits - minbits;  // max k32
  int sh = 0;
  while (((unsigned long long)range >> sh) + 1ull > (unsigned long long)kMaxBins) ++sh;
  *shift0 = 32 + sh;
  *n_bins = int(((unsigned long long)range >> sh) + 1ull);
}

static long long cand_capacity(const sbod_detect_desc* d) {
  if (d->class_agnostic) return (long long)d->P;
  return (long long)d->P * (long long)(d->C > 1 ? d->C - 1 : 1);
}

constexpr int kSpillCap = 65536;  // kept boxes per image of the detect_tools first stage before the call gives up

static int kept_capacity(const sbod_detect_desc* d) {
  // multiple of 16 so that the float4 / u64 arrays carved after each other stay aligned
  if (d->second_nms_thr >= 0.f) return 4096 + kChunk;
  return (d->top_k + 1 + kChunk + 15) & ~15;
}

// workspace carve-up (the leading block carries the zero contract)
struct DetLayout {
  size_t cand_count, hist, class_seen, rhist, zero_end, cutoff, pbound, agn_label, nms_mask, spill_key, spill_box,
      spill_st2, cand, total;
};
static DetLayout det_layout(const sbod_detect_desc* d) {
  DetLayout l;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    const size_t at = o;
    o += align_up(bytes, 256);
    return at;
  };
  const size_t N = size_t(d->N), P = size_t(d->P);
  l.cand_count = take(N * 4);
  l.hist = take(N * kMaxBins * 4);
  l.class_seen = take(N * size_t(d->C) * 4);
  l.rhist = take(N * kMaxBins * 4);
  l.zero_end = o;
  l.cutoff = take(N * 4);
  l.pbound = take(N * P * 4);
  l.agn_label = take(d->class_agnostic ? N * P * 4 : 0);
  l.nms_mask = take(N * kChunk * (kChunk / 32) * 4);
  const bool spill = d->second_nms_thr >= 0.f;
  l.spill_key = take(spill ? N * kSpillCap * 8 : 0);
  l.spill_box = take(spill ? N * kSpillCap * 16 : 0);
  l.spill_st2 = take(spill ? N * kSpillCap : 0);
  l.cand = take(N * size_t(cand_capacity(d)) * 8);
  l.total = o;
  return l;
}

}  // namespace sbod

using namespace sbod;

extern "C" size_t sbod_detect_workspace_bytes(const sbod_detect_desc* d) {
  if (!d) return 0;
  return det_layout(d).total;
}

// bytes at the start of the workspace that must be zero before the first call
extern "C" size_t sbod_detect_workspace_zero_bytes(const sbod_detect_desc* d) {
  if (!d) return 0;
  return det_layout(d).zero_end;
}

static int detect_run(const sbod_detect_desc* d, sbod_stream_t stream, int stage_mask) {
  if (!d || !d->locs || !d->scores || !d->out_boxes || !d->out_labels || !d->out_scores ||
      !d->out_prior || !d->out_counts)
    return SBOD_ERR_INVALID;
  if (d->N <= 0 || d->P <= 0 || d->C <= 1 || d->top_k <= 0) return SBOD_ERR_INVALID;
  if (d->box_kind == SBOD_BOX_OFFSET && !d->priors_cxcy) return SBOD_ERR_INVALID;
  if (d->box_kind < 0 || d->box_kind > SBOD_BOX_CORNER) return SBOD_ERR_INVALID;
  if (d->act_kind < SBOD_ACT_SOFTMAX || d->act_kind > SBOD_ACT_NONE) return SBOD_ERR_INVALID;
  if (d->P > (1 << kPriorBits) || d->C > (1 << kClassBits)) return SBOD_ERR_UNSUPPORTED;
  if (d->top_k > 4096) return SBOD_ERR_UNSUPPORTED;
  if (d->out_cap < (d->top_k > 1 ? d->top_k : 1)) return SBOD_ERR_INVALID;
  if (d->class_agnostic && (d->second_nms_thr >= 0.f || d->pre_nms_topk > 0)) return SBOD_ERR_INVALID;
  if (reinterpret_cast<uintptr_t>(d->scores) & 15) return SBOD_ERR_ALIGNMENT;
  if (reinterpret_cast<uintptr_t>(d->locs) & 15) return SBOD_ERR_ALIGNMENT;
  const DetLayout lay = det_layout(d);
  if (!d->workspace || d->workspace_bytes < lay.total) return SBOD_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(d->workspace) & 255) return SBOD_ERR_WORKSPACE;

  DetParams q;
  q.locs = d->locs; q.scores = d->scores;
  q.priors_cxcy = reinterpret_cast<const float4*>(d->priors_cxcy);
  q.prior_keep = d->prior_keep;
  q.N = d->N; q.P = d->P; q.C = d->C;
  q.act_kind = d->act_kind; q.box_kind = d->box_kind; q.clamp_inplace = d->clamp_inplace;
  q.min_score = d->min_score; q.max_overlap = d->max_overlap; q.top_k = d->top_k;
  q.second_thr = d->second_nms_thr; q.pre_nms_topk = d->pre_nms_topk;
  q.agnostic = d->class_agnostic ? 1 : 0;
  q.out_boxes = d->out_boxes; q.out_labels = d->out_labels; q.out_scores = d->out_scores;
  q.out_prior = d->out_prior; q.out_counts = d->out_counts; q.out_cap = d->out_cap;
  level0_layout(d->min_score, &q.shift0, &q.n_bins);
  q.cand_cap = cand_capacity(d);
  q.kcap = kept_capacity(d);
  unsigned char* w = static_cast<unsigned char*>(d->workspace);
  q.cand_count = reinterpret_cast<unsigned int*>(w + lay.cand_count);
  q.hist = reinterpret_cast<unsigned int*>(w + lay.hist);
  q.class_seen = reinterpret_cast<unsigned int*>(w + lay.class_seen);
  q.rhist = reinterpret_cast<unsigned int*>(w + lay.rhist);
  q.cutoff = reinterpret_cast<unsigned int*>(w + lay.cutoff);
  q.pbound = reinterpret_cast<float*>(w + lay.pbound);
  q.agn_label = reinterpret_cast<int32_t*>(w + lay.agn_label);
  q.nms_mask = reinterpret_cast<unsigned int*>(w + lay.nms_mask);
  const bool spill = d->second_nms_thr >= 0.f;
  q.spill_key = spill ? reinterpret_cast<unsigned long long*>(w + lay.spill_key) : nullptr;
  q.spill_box = spill ? reinterpret_cast<float4*>(w + lay.spill_box) : nullptr;
  q.spill_st2 = spill ? reinterpret_cast<uint8_t*>(w + lay.spill_st2) : nullptr;
  q.spill_cap = spill ? kSpillCap : 0;
  q.cand = reinterpret_cast<unsigned long long*>(w + lay.cand);
  // first band: about four times the boxes wanted plus a chunk of slack (rows; each holds >= 0 candidates)
  q.rows_target = 4 * (q.top_k + 1) + 1024;
  if (q.pre_nms_topk > 0 || q.second_thr >= 0.f) q.rows_target *= 2;

  // bound pass tiling: 128-row tiles, two threads per row; as many ring stages as fit in ~100 KB (two CTAs per SM)
  const bool fast = q.C <= 128;
  q.rows_per_tile = kTileRows;
  q.stage_floats = uint32_t(align_up(size_t(kTileRows) * q.C + 8, 32));
  {
    const size_t sb = size_t(q.stage_floats) * 4;
    q.n_stages = int((100 * 1024) / sb);
    if (q.n_stages > 4) q.n_stages = 4;
    if (q.n_stages < 1) q.n_stages = 1;
  }
  q.tiles_per_image = (q.P + kTileRows - 1) / kTileRows;
  q.n_tiles = q.tiles_per_image * q.N;

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    SBOD_CUDA_TRY(cudaFuncSetAttribute(detect_bound_kernel<81>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SBOD_CUDA_TRY(cudaFuncSetAttribute(detect_bound_kernel<21>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SBOD_CUDA_TRY(cudaFuncSetAttribute(detect_bound_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SBOD_CUDA_TRY(cudaFuncSetAttribute(detect_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    // without a preference the driver picks the smallest shared-memory carve-out that fits ONE CTA
    SBOD_CUDA_TRY(cudaFuncSetAttribute(detect_bound_kernel<81>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    SBOD_CUDA_TRY(cudaFuncSetAttribute(detect_bound_kernel<21>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    SBOD_CUDA_TRY(cudaFuncSetAttribute(detect_bound_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    SBOD_CUDA_TRY(cudaFuncSetAttribute(detect_refine_kernel<81>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    SBOD_CUDA_TRY(cudaFuncSetAttribute(detect_refine_kernel<21>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    SBOD_CUDA_TRY(cudaFuncSetAttribute(detect_refine_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_once.mark();
  }
  if (stage_mask & 1) {  // bound pass
    if (fast) {
      const size_t smem = size_t(q.n_stages) * q.stage_floats * 4 + 4 * 8;
      int grid = sm_count() * 2;
      if (grid > q.n_tiles) grid = q.n_tiles;
      if (q.C == 81) detect_bound_kernel<81><<<grid, kStreamThreads, smem, st>>>(q);       // COCO
      else if (q.C == 21) detect_bound_kernel<21><<<grid, kStreamThreads, smem, st>>>(q);  // VOC
      else detect_bound_kernel<0><<<grid, kStreamThreads, smem, st>>>(q);
    } else {
      detect_bound_generic_kernel<<<sm_count() * 8, 256, 0, st>>>(q);
    }
    SBOD_LAUNCH_CHECK();
  }
  if (stage_mask & 4) {  // exact evaluation of the rows above the cutoff
    // rows per CTA: 512 .. 2048, so that the whole grid is resident (three CTAs per SM) whenever it can be
    int rows_per_cta = 512;
    while (rows_per_cta < kRefMaxRows &&
           (long long)q.N * ((q.P + rows_per_cta - 1) / rows_per_cta) > (long long)sm_count() * 3)
      rows_per_cta *= 2;
    dim3 grid((q.P + rows_per_cta - 1) / rows_per_cta, q.N);
    if (q.C == 81) detect_refine_kernel<81><<<grid, kRefThreads, 0, st>>>(q, rows_per_cta);
    else if (q.C == 21) detect_refine_kernel<21><<<grid, kRefThreads, 0, st>>>(q, rows_per_cta);
    else detect_refine_kernel<0><<<grid, kRefThreads, 0, st>>>(q, rows_per_cta);
    SBOD_LAUNCH_CHECK();
  }
  if (stage_mask & 2) {  // NMS (+ in-kernel second band for the images that need it)
    const size_t nms_smem = ((sizeof(NmsSmem) + 127) & ~size_t(127)) + size_t(q.kcap) * 25 + 256 + size_t(q.C + 1) * 4;
    if (nms_smem > 220 * 1024) return SBOD_ERR_UNSUPPORTED;
    detect_nms_kernel<<<q.N, kNmsThreads, nms_smem, st>>>(q);
    SBOD_LAUNCH_CHECK();
  }
  return SBOD_OK;
}

// The activation exactly as the refine pass evaluates it (same device functions, a warp per row), for every
// (image, prior, class): lets a test separate "scores within 1e-5 of torch" from "kept indices bit-exact given
// the scores".
__global__ void __launch_bounds__(256) detect_probabilities_kernel(const DetParams q, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, s = lane & 7;
  const size_t total = size_t(q.N) * q.P;
  const size_t n_octets = (size_t(gridDim.x) * blockDim.x) >> 3;
  const float NEG = -__int_as_float(0x7f800000);
  for (size_t np0 = ((size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5) * 4; np0 < total; np0 += (n_octets >> 2) * 4) {
    const size_t np = np0 + (lane >> 3);
    const bool on = np < total;
    const float* x = q.scores + (on ? np : 0) * size_t(q.C);
    float mx = 0.f, sum = 1.f;
    if (q.act_kind == SBOD_ACT_SOFTMAX) {  // the same sequence of operations as eval_rows
      float m = NEG;
      if (on)
        for (int k = s; k < q.C; k += 8) m = fmaxf(m, __ldg(x + k));
      mx = octet_max(m);
      float acc = 0.f;
      if (on)
        for (int k = s; k < q.C; k += 8) acc += exp_term(__ldg(x + k), mx);
      sum = octet_sum(acc);
    }
    if (on)
      for (int k = s; k < q.C; k += 8) {
        const float v = __ldg(x + k);
        float pr = v;
        if (q.act_kind == SBOD_ACT_SOFTMAX) pr = __fdiv_rn(expf(v - mx), sum);
        else if (q.act_kind == SBOD_ACT_SIGMOID) pr = __fdiv_rn(1.f, 1.f + expf(-v));
        out[np * size_t(q.C) + k] = pr;
      }
  }
}

extern "C" int sbod_detect_probabilities(const float* scores, int N, int P, int C, int act_kind, float* out,
                                         sbod_stream_t stream) {
  if (!scores || !out || N <= 0 || P <= 0 || C <= 1) return SBOD_ERR_INVALID;
  if (act_kind < SBOD_ACT_SOFTMAX || act_kind > SBOD_ACT_NONE) return SBOD_ERR_INVALID;
  DetParams q;
  memset(&q, 0, sizeof(q));
  q.scores = scores;
  q.N = N; q.P = P; q.C = C;
  q.act_kind = act_kind;
  detect_probabilities_kernel<<<sm_count() * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(q, out);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

#ifdef SBOD_DEBUG_HOOKS
extern "C" __attribute__((visibility("default"))) int sbod_debug_det_times(unsigned long long* out, int reset) {
  if (out) cudaMemcpyFromSymbol(out, g_det_times, sizeof(g_det_times));
  if (reset) {
    unsigned long long init[16];
    for (int i = 0; i < 16; ++i) init[i] = (i == 0 || i == 8) ? ~0ull : 0ull;
    cudaMemcpyToSymbol(g_det_times, init, sizeof(init));
  }
  return 0;
}
#endif

extern "C" int sbod_detect(const sbod_detect_desc* d, sbod_stream_t stream) {
  return detect_run(d, stream, 7);
}

// Profiling / bench hook. stage 0 = bound pass + refine, 1 = NMS kernel, 2 = bound pass only (the streaming
// kernel), 3 = refine only, 4 = refine + NMS (everything after stage 2). A bound pass must be followed by the
// later stages before the next full sbod_detect (the NMS kernel consumes and cleans the workspace).
extern "C" int sbod_detect_stage(const sbod_detect_desc* d, int stage, sbod_stream_t stream) {
  static const int masks[5] = {1 | 4, 2, 1, 4, 4 | 2};
  if (stage < 0 || stage > 4) return SBOD_ERR_INVALID;
  return detect_run(d, stream, masks[stage]);
}
