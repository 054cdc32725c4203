// detect.cu — the eval path (sm_100a): activation + decode + clamp + score threshold +
// per-class NMS + top-k, batched, no host round trips.
//
//   detect_score_fast_kernel<C>  2 <= C <= 128 (C = 81 / 21 compile-time): two threads per row, a
//                        sampling pass fixes a per-image score cutoff, the main pass emits only the
//                        candidates above it (exact: flagged images are redone without cutoff).
//   detect_score_kernel  generic (C > 128); persistent; logits tiles streamed by bulk TMA (same ring as
//                        the train kernel); one thread per prior row turns the row into probabilities in
//                        place, and every (class, prior) above min_score is emitted as one 64-bit
//                        key  [0x3F800000 - score_bits : 32][class : 12][prior : 20]  so that
//                        ascending key order == (score desc, class asc, prior asc) == the order in
//                        which the reference's per-class NMS + stable top-k sort consumes them.
//                        A per-image histogram of the key's top digit is built on the fly.
//   detect_nms_kernel    one CTA per image. The reference runs NMS over every candidate of every
//                        class and only then keeps the top_k best (models/utils.py:245-290). Greedy
//                        NMS is prefix-stable, so the first top_k+1 survivors in global key order
//                        are all that is ever needed: candidates are pulled in key-ordered chunks
//                        (histogram-guided, radix descent for oversized bins), sorted in smem,
//                        suppressed per class by one warp per class (warp-ballot over the kept
//                        list), and the loop stops as soon as top_k+1 boxes survive. If the
//                        candidates run out first, everything is kept and emitted class-major,
//                        exactly as the reference does.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "row_stream.cuh"

namespace sbod {

constexpr int kDRows = 128;
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
  float* out_boxes;
  int64_t* out_labels;
  float* out_scores;
  int32_t* out_prior;
  int32_t* out_counts;
  int out_cap;
  // workspace
  unsigned int* cand_count;     // [N]
  unsigned int* hist;           // [N, n_bins]
  unsigned long long* cand;     // [N, cand_cap]
  unsigned int* class_seen;     // [N, C]  (pre-NMS per-class rank counters)
  long long cand_cap;
  int n_bins, shift0;           // level-0 digit = key >> shift0
  int kcap;                     // kept-list capacity in smem
  int debug_skip;               // SBOD_DEBUG_SKIP (profiling): bit0 no emission, bit1 no exact refine, bit2 no cutoff
  // speculative per-image score cutoff (exact: images whose candidates run out are redone in full)
  int mode;                     // 0 = main pass, 1 = sampling pass, 2 = fallback pass (flagged images only)
  unsigned int* shist;          // [N, n_bins] histogram of the sampled tiles
  unsigned int* cutoff_k32;     // [N] emit only keys with k32 < cutoff (0xffffffff = no cutoff)
  float* cutoff_floor;          // [N] score just below the cutoff (mask floor)
  unsigned int* flags;          // [N] 1 = the cutoff was too strict, redo this image in full
  unsigned int* nms_mask;       // [N, kChunk, kChunk/32] suppression bits of large class segments
  int sample_stride, sample_target;
  int speculate;                // the sampling pass ran: cutoffs come from shist
  // tiling
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

struct DTile {
  int n, p0, rows;
};
SBOD_DEVINL DTile dtile(const DetParams& q, int tile) {
  DTile t;
  t.n = tile / q.tiles_per_image;
  t.p0 = (tile - t.n * q.tiles_per_image) * q.rows_per_tile;
  t.rows = min(q.rows_per_tile, q.P - t.p0);
  return t;
}

SBOD_DEVINL void issue_dtile(const DetParams& q, int tile, float* stage, uint64_t* bar) {
  const DTile t = dtile(q, tile);
  const size_t first = (size_t(t.n) * q.P + t.p0) * size_t(q.C);
  const size_t total = size_t(q.N) * q.P * size_t(q.C);
  const TileSpan s = make_tile_span(q.scores, first, size_t(t.rows) * q.C, total);
  for (uint32_t i = 0; i < s.tail_floats; ++i)
    stage[s.bulk_bytes / 4 + i] = q.scores[(s.src16 - q.scores) + s.bulk_bytes / 4 + i];
  if (s.bulk_bytes) {
    mbar_arrive_expect_tx(bar, s.bulk_bytes);
    tma_load_1d(stage, s.src16, s.bulk_bytes, bar);
  } else {
    mbar_arrive(bar);
  }
}

__global__ void __launch_bounds__(kDRows) detect_score_kernel(const DetParams q) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stages = reinterpret_cast<float*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + size_t(q.n_stages) * q.stage_floats * 4);
  __shared__ unsigned int s_wtot[kDRows / 32];
  __shared__ unsigned int s_base;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n_my = (q.n_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  if (tid == 0) {
    for (int s = 0; s < q.n_stages; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (int s = 0; s < q.n_stages && s < n_my; ++s)
      issue_dtile(q, blockIdx.x + s * gridDim.x, stages + size_t(s) * q.stage_floats, &bars[s]);
  }
  int gcd = 1;
  while (gcd < 32 && (q.C % (gcd * 2)) == 0) gcd *= 2;
  const int rot = (lane * gcd) >> 5;
  const int C = q.C;
  const float INF = __int_as_float(0x7f800000);

  for (int it = 0; it < n_my; ++it) {
    const int tile = blockIdx.x + it * gridDim.x;
    const DTile tc = dtile(q, tile);
    const int n = tc.n;
    const bool valid = tid < tc.rows;
    const int p = tc.p0 + tid;
    const size_t np = size_t(n) * q.P + (valid ? p : tc.p0);
    const bool keep_row = valid && (q.prior_keep ? q.prior_keep[np] != 0 : true);

    if (valid && q.clamp_inplace) {  // models/utils.py:224, detect_tools.py:264: clamp_ on the caller's tensor
      float4* lp = reinterpret_cast<float4*>(q.locs) + np;
      float4 b = *lp;
      b.x = fminf(fmaxf(b.x, 0.f), 1.f); b.y = fminf(fmaxf(b.y, 0.f), 1.f);
      b.z = fminf(fmaxf(b.z, 0.f), 1.f); b.w = fminf(fmaxf(b.w, 0.f), 1.f);
      *lp = b;
    }

    const int s = it % q.n_stages;
    float* stage = stages + size_t(s) * q.stage_floats;
    mbar_wait(&bars[s], (it / q.n_stages) & 1);
    const size_t first = (size_t(n) * q.P + tc.p0) * size_t(C);
    float* row = stage + (first & 3) + size_t(tid) * C;
    unsigned int cnt = 0;
    if (keep_row) {
      if (q.act_kind == SBOD_ACT_SOFTMAX) {
        float m0 = -INF, m1 = -INF, m2 = -INF, m3 = -INF;
        int k = rot;
        for (; k + 3 < C; k += 4) {
          m0 = fmaxf(m0, row[k]); m1 = fmaxf(m1, row[k + 1]);
          m2 = fmaxf(m2, row[k + 2]); m3 = fmaxf(m3, row[k + 3]);
        }
        for (; k < C; ++k) m0 = fmaxf(m0, row[k]);
        for (k = 0; k < rot; ++k) m1 = fmaxf(m1, row[k]);
        const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        k = rot;
        for (; k + 3 < C; k += 4) {
          const float e0 = __expf(row[k] - mx), e1 = __expf(row[k + 1] - mx);
          const float e2 = __expf(row[k + 2] - mx), e3 = __expf(row[k + 3] - mx);
          row[k] = e0; row[k + 1] = e1; row[k + 2] = e2; row[k + 3] = e3;
          s0 += e0; s1 += e1; s2 += e2; s3 += e3;
        }
        for (; k < C; ++k) { const float e = __expf(row[k] - mx); row[k] = e; s0 += e; }
        for (k = 0; k < rot; ++k) { const float e = __expf(row[k] - mx); row[k] = e; s1 += e; }
        const float inv = __frcp_rn((s0 + s1) + (s2 + s3));
        for (int kk = 0; kk < C; ++kk) {
          k = kk + rot; if (k >= C) k -= C;
          const float pr = row[k] * inv;
          row[k] = pr;
          cnt += (k > 0 && pr > q.min_score) ? 1u : 0u;
        }
      } else {
        for (int kk = 0; kk < C; ++kk) {
          int k = kk + rot; if (k >= C) k -= C;
          const float pr = q.act_kind == SBOD_ACT_NONE ? row[k] : __frcp_rn(1.f + __expf(-row[k]));
          row[k] = pr;
          cnt += (k > 0 && pr > q.min_score) ? 1u : 0u;
        }
      }
    }
    // block-exclusive scan of the per-row candidate counts, one atomic per tile
    unsigned int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_wtot[wid] = inc;
    __syncthreads();
    unsigned int off = inc - cnt, tot = 0;
    for (int w = 0; w < kDRows / 32; ++w) {
      if (w < wid) off += s_wtot[w];
      tot += s_wtot[w];
    }
    if (tid == 0) s_base = tot ? atomicAdd(&q.cand_count[n], tot) : 0u;
    __syncthreads();
    if (cnt) {
      long long slot = (long long)s_base + off;
      unsigned long long* dst = q.cand + size_t(n) * q.cand_cap;
      unsigned int* hist = q.hist + size_t(n) * q.n_bins;
      for (int k = 1; k < C; ++k) {  // ascending class order inside the row
        const float pr = row[k];
        if (pr > q.min_score) {
          uint32_t bits = __float_as_uint(pr);
          const uint32_t k32 = bits > kOneBits ? 0u : kOneBits - bits;
          const unsigned long long key = (static_cast<unsigned long long>(k32) << 32) |
                                         (static_cast<unsigned long long>(k) << kPriorBits) |
                                         static_cast<unsigned long long>(p);
          if (slot < q.cand_cap) dst[slot] = key;
          ++slot;
          int bin = int(key >> q.shift0);
          if (bin >= q.n_bins) bin = q.n_bins - 1;
          atomicAdd(&hist[bin], 1u);
        }
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0 && it + q.n_stages < n_my)
      issue_dtile(q, blockIdx.x + (it + q.n_stages) * gridDim.x, stage, &bars[s]);
  }
}

// Per-image emission cutoff from the histogram of the sampled tiles: the score bin above which about
// sample_target candidates are expected. Block-wide (any block size that is a multiple of 32);
// every thread returns the same (k32 cutoff, score floor). scratch: >= 34 unsigned ints.
__device__ void compute_cutoff(const DetParams& q, int n, unsigned int* scratch, unsigned int& k32_out,
                               float& floor_out) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nt = blockDim.x;
  const unsigned int* sh = q.shist + size_t(n) * q.n_bins;
  const int per = (q.n_bins + nt - 1) / nt;  // contiguous bins per thread, best scores first
  unsigned int mine = 0;
  for (int j = 0; j < per; ++j) {
    const int b = tid * per + j;
    if (b < q.n_bins) mine += sh[b];
  }
  unsigned int inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (tid == 0) scratch[33] = unsigned(q.n_bins);
  if (lane == 31) scratch[wid] = inc;
  __syncthreads();
  unsigned int before = inc - mine;
  for (int w = 0; w < wid; ++w) before += scratch[w];
  const unsigned int need = unsigned((q.sample_target + q.sample_stride - 1) / q.sample_stride);
  if (before < need && before + mine >= need) {
    unsigned int acc = before;
    for (int j = 0; j < per; ++j) {
      const int b = tid * per + j;
      if (b >= q.n_bins) break;
      acc += sh[b];
      if (acc >= need) {
        scratch[33] = unsigned(b + 1);  // emit bins [0, b]
        break;
      }
    }
  }
  __syncthreads();
  const int cb = int(scratch[33]);
  if (cb >= q.n_bins) {
    k32_out = 0xffffffffu;
    floor_out = q.min_score;
  } else {
    k32_out = unsigned(cb) << (q.shift0 - 32);
    floor_out = fmaxf(q.min_score, __uint_as_float(kOneBits - min(k32_out, kOneBits)));
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------
// detect_score_fast_kernel — odd C <= 128, softmax or sigmoid. Two threads per row (row_stream.cuh),
// contiguous tile ranges per CTA. Pass 2 of the softmax also builds, per thread, a bit mask of the
// classes whose un-normalised exp already exceeds min_score (a superset of the candidates, since the
// row sum is >= 1); only those few are re-evaluated exactly and emitted. The per-image histogram
// of the key's top digit is accumulated in shared memory and flushed once per image.
// ------------------------------------------------------------------------------------------
template <int kC>  // kC > 0: compile-time class count (unrolled softmax passes)
__global__ void __launch_bounds__(kStreamThreads, 2) detect_score_fast_kernel(const DetParams q) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stages = reinterpret_cast<float*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + size_t(q.n_stages) * q.stage_floats * 4);
  __shared__ unsigned int s_hist[kMaxBins + 40];  // + scratch of compute_cutoff
  constexpr int kKeyBuf = 1024, kKeyFlush = 512;  // shared key buffer (see the slot reservation below)
  __shared__ unsigned long long s_keys[kKeyBuf];
  __shared__ unsigned int s_ncand, s_gbase;

  const int tid = threadIdx.x, lane = tid & 31;
  const int C = kC ? kC : q.C;
  // tiles of this CTA: all of a contiguous range (modes 0, 2) or every sample_stride-th tile (mode 1)
  const int stride = q.mode == 1 ? q.sample_stride : 1;
  const int n_units = q.mode == 1 ? (q.n_tiles + stride - 1 - stride / 2) / stride : q.n_tiles;
  int t0, t1;
  tile_range(n_units, blockIdx.x, gridDim.x, t0, t1);
  const int n_my = t1 - t0;
  auto tile_of = [&](int idx) { return (t0 + idx) * stride + (q.mode == 1 ? stride / 2 : 0); };
  auto image_active = [&](int n) { return q.mode != 2 || q.flags[n] != 0u; };
  if (n_my <= 0) return;
  if (q.mode == 2) {  // nothing flagged in my range: leave at once
    bool any = false;
    for (int n = tile_of(0) / q.tiles_per_image; n <= tile_of(n_my - 1) / q.tiles_per_image; ++n)
      any = any || q.flags[n] != 0u;
    if (!any) return;
  }
  auto issue = [&](int idx, float* stage, uint64_t* bar) {
    const StreamTile t = stream_tile(tile_of(idx), q.tiles_per_image, kTileRows, q.P);
    if (image_active(t.n)) stream_issue(q.scores, q.N, q.P, C, t, stage, bar);
    else mbar_arrive(bar);  // keep the ring's phase bookkeeping, move no data
  };
  if (tid == 0) {
    for (int s = 0; s < q.n_stages; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  for (int b = tid; b < q.n_bins; b += kStreamThreads) s_hist[b] = 0u;
  if (tid == 0) s_ncand = 0u;
  __syncthreads();
  if (tid == 0) {
    for (int s = 0; s < q.n_stages && s < n_my; ++s) issue(s, stages + size_t(s) * q.stage_floats, &bars[s]);
  }
  int row, h;
  stream_map(tid, row, h);
  const int nh = (C + 1 - h) >> 1;
  int hist_n = -1;
  unsigned int cut_k32 = 0xffffffffu;
  float cut_floor = q.min_score;

  auto flush_hist = [&](int n) {
    __syncthreads();
    unsigned int* gh = (q.mode == 1 ? q.shist : q.hist) + size_t(n) * q.n_bins;
    for (int b = tid; b < q.n_bins; b += kStreamThreads) {
      const unsigned int v = s_hist[b];
      if (v) {
        atomicAdd(&gh[b], v);
        s_hist[b] = 0u;
      }
    }
    __syncthreads();
  };

  // tile coordinates advance incrementally (no integer division in the loop)
  int cur_n = tile_of(0) / q.tiles_per_image;
  int cur_t = tile_of(0) - cur_n * q.tiles_per_image;
  int ring_s = 0;
  uint32_t ring_parity = 0;
  for (int it = 0; it < n_my; ++it) {
    StreamTile tc;
    tc.n = cur_n;
    tc.p0 = cur_t * kTileRows;
    tc.rows = min(kTileRows, q.P - tc.p0);
    cur_t += stride;
    while (cur_t >= q.tiles_per_image) {
      cur_t -= q.tiles_per_image;
      ++cur_n;
    }
    const int n = tc.n;
    if (n != hist_n) {
      if (hist_n >= 0) flush_hist(hist_n);
      hist_n = n;
      if (q.mode == 0 && q.speculate) compute_cutoff(q, n, s_hist + kMaxBins, cut_k32, cut_floor);
    }
    const bool act = image_active(n);
    const bool valid = row < tc.rows && act;
    const int r = min(row, tc.rows - 1);
    const int p = tc.p0 + r;
    const size_t np = size_t(n) * q.P + p;
    if (valid && h == 0 && q.clamp_inplace && q.mode == 0) {  // clamp_ on the caller's tensor (models/utils.py:224)
      float4* lp = reinterpret_cast<float4*>(q.locs) + np;
      float4 b = *lp;
      b.x = fminf(fmaxf(b.x, 0.f), 1.f); b.y = fminf(fmaxf(b.y, 0.f), 1.f);
      b.z = fminf(fmaxf(b.z, 0.f), 1.f); b.w = fminf(fmaxf(b.w, 0.f), 1.f);
      *lp = b;
    }
    const bool keep_row = valid && (q.prior_keep ? q.prior_keep[np] != 0 : true);

    const int s = ring_s;  // stage / phase advance incrementally (no integer division per tile)
    float* stage = stages + size_t(s) * q.stage_floats;
    mbar_wait(&bars[s], ring_parity);
    if (++ring_s == q.n_stages) {
      ring_s = 0;
      ring_parity ^= 1u;
    }
    const uint32_t head = ((uint32_t(n) * uint32_t(q.P) + uint32_t(tc.p0)) * uint32_t(C)) & 3u;
    const float* rp = stage + head + r * C + h;

    uint32_t m0 = 0u, m1 = 0u;
    float nmx2 = 0.f, inv = 1.f;
    if (q.act_kind == SBOD_ACT_SOFTMAX) {
      const float mx = kC ? pair_row_max_fixed<kC>(rp - h, h) : half_row_max(rp, nh);
      nmx2 = -mx * kLog2e;
      const float sum = kC ? pair_row_sumexp_mask_fixed<kC>(rp - h, h, nmx2, cut_floor, m0, m1)
                           : half_row_sumexp_mask(rp, nh, nmx2, cut_floor, m0, m1);
      inv = __frcp_rn(sum);
    } else {
      // sigmoid(x) > t  <=>  x > logit(t); keep a small margin, the exact test follows
      const float t = fminf(fmaxf(cut_floor, 1e-30f), 1.f - 1e-7f);
      const float lim = q.act_kind == SBOD_ACT_NONE ? cut_floor * (1.f - 1e-6f) : logf(t / (1.f - t)) - 1e-3f;
      for (int j = 0; j < nh; ++j)
        if (rp[2 * j] > lim) {
          if (j < 32) m0 |= 1u << j;
          else m1 |= 1u << (j - 32);
        }
    }
    if (h == 0) m0 &= ~1u;  // class 0 (background) never yields a detection
    if (q.debug_skip & 2) m0 = m1 = 0u;
    if (!keep_row) m0 = m1 = 0u;
    // exact test of the few flagged classes: prob = e * (1/sum) > min_score, as the emitted score
    unsigned int cnt = 0;
    {
      uint32_t mm = m0;
      while (mm) {
        const int j = __ffs(mm) - 1;
        mm &= mm - 1;
        const float x = rp[2 * j];
        const float pr = q.act_kind == SBOD_ACT_SOFTMAX ? ex2_approx(fmaf(x, kLog2e, nmx2)) * inv
                         : (q.act_kind == SBOD_ACT_NONE ? x : __frcp_rn(1.f + __expf(-x)));
        if (pr > q.min_score && (kOneBits - min(__float_as_uint(pr), kOneBits)) < cut_k32) ++cnt; else m0 &= ~(1u << j);
      }
      mm = m1;
      while (mm) {
        const int j = __ffs(mm) - 1;
        mm &= mm - 1;
        const float x = rp[2 * (j + 32)];
        const float pr = q.act_kind == SBOD_ACT_SOFTMAX ? ex2_approx(fmaf(x, kLog2e, nmx2)) * inv
                         : (q.act_kind == SBOD_ACT_NONE ? x : __frcp_rn(1.f + __expf(-x)));
        if (pr > q.min_score && (kOneBits - min(__float_as_uint(pr), kOneBits)) < cut_k32) ++cnt; else m1 &= ~(1u << j);
      }
    }
    // warp-exclusive scan of the per-thread counts; one global atomic per warp that has candidates
    // (no CTA-wide barrier: while this warp waits for its slot the other warps keep streaming)
    unsigned int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    const unsigned int wtot = __shfl_sync(0xffffffffu, inc, 31);
    // Slots: the warp's keys go to a CTA-wide shared buffer (one shared-memory atomic per warp and
    // tile) that is handed to global memory with ONE global atomic per flush; a warp never waits for
    // a global atomic's round trip unless the buffer overflows (then the excess goes out directly).
    unsigned int wbase = 0u, fits = 0u, gbase = 0u;
    if (wtot && q.mode != 1) {
      if (lane == 31) wbase = atomicAdd(&s_ncand, wtot);
      wbase = __shfl_sync(0xffffffffu, wbase, 31);
      fits = wbase >= unsigned(kKeyBuf) ? 0u : min(wtot, unsigned(kKeyBuf) - wbase);
      if (fits < wtot) {
        if (lane == 31) gbase = atomicAdd(&q.cand_count[n], wtot - fits);
        gbase = __shfl_sync(0xffffffffu, gbase, 31);
      }
    }
    const unsigned int off = inc - cnt;
    if (cnt && !(q.debug_skip & 1)) {
      unsigned int rank = off;  // position inside the warp's block of wtot keys
      unsigned long long* dst = q.cand + size_t(n) * q.cand_cap;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t mm = half ? m1 : m0;
        while (mm) {
          const int j = __ffs(mm) - 1 + 32 * half;
          mm &= mm - 1;
          const int k = 2 * j + h;
          const float x = rp[2 * j];
          const float pr = q.act_kind == SBOD_ACT_SOFTMAX ? ex2_approx(fmaf(x, kLog2e, nmx2)) * inv
                           : (q.act_kind == SBOD_ACT_NONE ? x : __frcp_rn(1.f + __expf(-x)));
          const uint32_t bits = __float_as_uint(pr);
          const uint32_t k32 = bits > kOneBits ? 0u : kOneBits - bits;
          const unsigned long long key = (static_cast<unsigned long long>(k32) << 32) |
                                         (static_cast<unsigned long long>(k) << kPriorBits) |
                                         static_cast<unsigned long long>(p);
          if (q.mode != 1) {
            if (rank < fits) {
              s_keys[wbase + rank] = key;
            } else {
              const long long slot = (long long)gbase + (rank - fits);
              if (slot < q.cand_cap) dst[slot] = key;
            }
          }
          ++rank;
          int bin = int(key >> q.shift0);
          if (bin >= q.n_bins) bin = q.n_bins - 1;
          atomicAdd(&s_hist[bin], 1u);
        }
      }
    }
    __syncthreads();  // every thread is done with stage s
    if (tid == 0 && it + q.n_stages < n_my) issue(it + q.n_stages, stage, &bars[s]);
    // hand the buffered keys over when the buffer fills up, the image changes or the CTA is done
    // (s_ncand is stable here: the barrier above is behind every append of this tile)
    const unsigned int have = min(s_ncand, unsigned(kKeyBuf));
    if (have && (have >= unsigned(kKeyFlush) || it == n_my - 1 || cur_n != n)) {
      if (tid == 0) s_gbase = atomicAdd(&q.cand_count[n], have);
      __syncthreads();
      unsigned long long* dst = q.cand + size_t(n) * q.cand_cap;
      for (unsigned int i = tid; i < have; i += kStreamThreads) {
        const long long slot = (long long)s_gbase + i;
        if (slot < q.cand_cap) dst[slot] = s_keys[i];
      }
      __syncthreads();
      if (tid == 0) s_ncand = 0u;
      __syncthreads();
    }
  }
  if (hist_n >= 0) flush_hist(hist_n);
}

// ------------------------------------------------------------------------------------------
// detect_nms_kernel
// ------------------------------------------------------------------------------------------
SBOD_DEVINL int key_class(unsigned long long k) { return int((k >> kPriorBits) & ((1u << kClassBits) - 1u)); }
SBOD_DEVINL int key_prior(unsigned long long k) { return int(k & ((1u << kPriorBits) - 1u)); }
SBOD_DEVINL float key_score(unsigned long long k) { return __uint_as_float(kOneBits - uint32_t(k >> 32)); }

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
  unsigned long long* kkey = reinterpret_cast<unsigned long long*>(dyn);          // [kcap]
  float4* kbox = reinterpret_cast<float4*>(dyn + size_t(q.kcap) * 8);             // [kcap]
  uint8_t* kst2 = reinterpret_cast<uint8_t*>(dyn + size_t(q.kcap) * 24);          // [kcap]
  unsigned int* cls_off = reinterpret_cast<unsigned int*>(dyn + size_t(q.kcap) * 25 + 128 - (size_t(q.kcap) * 25) % 128);  // [C+1]

  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (q.mode == 2 && q.flags[n] == 0u) return;  // fallback pass: only the flagged images
  bool cutoff_active = false;
  if (q.mode == 0 && q.speculate) {
    unsigned int ck = 0xffffffffu;
    float cf = 0.f;
    compute_cutoff(q, n, S.wscan, ck, cf);  // same function of the same histogram as the score pass
    cutoff_active = ck != 0xffffffffu;
    unsigned int* sh = q.shist + size_t(n) * q.n_bins;
    for (int b = tid; b < q.n_bins; b += kNmsThreads) sh[b] = 0u;  // clean for the next call
  }
  const bool two_stage = q.second_thr >= 0.f;
  unsigned int* g_hist = q.hist + size_t(n) * q.n_bins;
  unsigned int* g_seen = q.class_seen + size_t(n) * q.C;
  const unsigned long long* g_cand = q.cand + size_t(n) * q.cand_cap;
  const unsigned int raw_total = q.cand_count[n];
  const long long total = raw_total < (unsigned long long)q.cand_cap ? raw_total : q.cand_cap;
  const bool overflow = raw_total > (unsigned long long)q.cand_cap;

  int kept_n = 0;   // stage-1 survivors so far (kept list length)
  int kept2_n = 0;  // stage-2 survivors so far
  int status = 0;   // 1 = kept-list capacity exceeded
  unsigned long long lo_key = 0ull;
  bool exhausted = total == 0;
  const int stop_at = q.top_k + 1;

  // The first round only takes about twice as many candidates as boxes are wanted (a power of two):
  // when suppression is moderate that is enough, and sorting / decoding / grouping a short chunk is
  // cheaper; if it is not enough the following rounds take full chunks.
  int chunk_cap = kChunk;
  if (!two_stage) {
    chunk_cap = 256;
    while (chunk_cap < 2 * stop_at && chunk_cap < kChunk) chunk_cap <<= 1;
  }
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
        for (int b = tid; b < nb; b += kNmsThreads) S.hist[b] = g_hist[b];
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
    // ---- sort by key (bitonic, padded with ~0) ----------------------------------------------
    // Each thread keeps its element in a register; partners less than 32 apart are reached by
    // shuffles (40 of the 55 steps, no block barrier), the others through shared memory, alternating
    // between ckey and the not yet used cbox array so that one barrier per step suffices.
    {
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
      const unsigned int seen = g_seen[c];  // candidates of this class consumed by earlier rounds
      if (lane == 0) g_seen[c] = seen + (seg1 - seg0);
      unsigned int lim = seg1;               // pre-NMS per-class cap: later candidates are dropped
      if (q.pre_nms_topk > 0) {
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

    // ---- stage 2 (detect_tools): class-agnostic NMS over the stage-1 survivors, key order ----
    if (two_stage) {
      if (wid == 0) {
        int nnew2 = 0;  // new stage-2 survivors, their chunk positions reuse S.cidx[0..)
        for (int e = 0; e < m; ++e) {
          if (S.cflag[e] != 1) continue;
          const float4 be = S.cbox[e];
          bool sup = false;
          for (int i = lane; i < kept_n && !sup; i += 32)
            if (kst2[i] && overlaps(kbox[i], be, q.second_thr)) sup = true;
          for (int i = lane; i < nnew2 && !sup; i += 32)
            if (overlaps(S.cbox[S.cidx[i]], be, q.second_thr)) sup = true;
          if (!__any_sync(0xffffffffu, sup)) {
            if (lane == 0) {
              S.cidx[nnew2] = uint16_t(e);
              S.cflag[e] = 2;
            }
            ++nnew2;
            __syncwarp();
          }
        }
        if (lane == 0) S.misc[5] = nnew2;
      }
      __syncthreads();
    }

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
    if (kept_n + add > q.kcap) {
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

  // ---- leave the workspace clean for the next call ----------------------------------------
  for (int b = tid; b < q.n_bins; b += kNmsThreads) g_hist[b] = 0u;
  for (int c = tid; c < q.C; c += kNmsThreads) g_seen[c] = 0u;
  if (tid == 0) q.cand_count[n] = 0u;
  if (q.mode == 2 && tid == 0) q.flags[n] = 0u;
  // The candidates above the speculative cutoff ran out before top_k+1 boxes survived: the answer
  // needs lower-scored candidates that were not emitted. Flag the image; the fallback pass redoes it.
  if (cutoff_active && !status && !overflow && (two_stage ? kept2_n : kept_n) < stop_at) {
    if (tid == 0) {
      q.flags[n] = 1u;
      q.out_counts[n] = -3;
    }
    return;
  }

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
        ob[i] = kbox[i]; ol[i] = key_class(k); os[i] = key_score(k); op[i] = key_prior(k);
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
        ob[dst] = kbox[i]; ol[dst] = c; os[dst] = key_score(k); op[dst] = key_prior(k);
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
          ob[dst] = kbox[i]; ol[dst] = key_class(k); os[dst] = key_score(k); op[dst] = key_prior(k);
        }
        w += __popc(bal);
      }
      if (lane == 0) q.out_counts[n] = truncate ? min(w, q.top_k) : min(w, q.out_cap);
    }
  }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct DTiling {
  int rows, stages;
  uint32_t stage_floats;
  size_t smem;
};
static DTiling choose_dtiling(int C) {
  DTiling t;
  const size_t budget = 100 * 1024;
  t.rows = kDRows;
  while (t.rows > 4 && (size_t(t.rows) * C + 8) * 4 > budget) t.rows /= 2;
  t.stage_floats = uint32_t(align_up(size_t(t.rows) * C + 8, 32));
  t.stages = int(budget / (size_t(t.stage_floats) * 4));
  if (t.stages > 4) t.stages = 4;
  if (t.stages < 1) t.stages = 1;
  t.smem = size_t(t.stages) * t.stage_floats * 4 + 4 * 8;
  return t;
}

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
  return (long long)d->P * (long long)(d->C > 1 ? d->C - 1 : 1);
}

static int kept_capacity(const sbod_detect_desc* d) {
  // multiple of 16 so that the float4 / u64 arrays carved after each other stay aligned
  if (d->second_nms_thr >= 0.f) return 4096 + kChunk;
  return (d->top_k + 1 + kChunk + 15) & ~15;
}

}  // namespace sbod

using namespace sbod;

extern "C" size_t sbod_detect_workspace_bytes(const sbod_detect_desc* d) {
  if (!d) return 0;
  int sh, nb;
  level0_layout(d->min_score, &sh, &nb);
  size_t b = 0;
  b += align_up(size_t(d->N) * 4, 256);                       // cand_count
  b += align_up(size_t(d->N) * kMaxBins * 4, 256);            // hist
  b += align_up(size_t(d->N) * size_t(d->C) * 4, 256);        // class_seen
  b += align_up(size_t(d->N) * kMaxBins * 4, 256);            // shist
  b += align_up(size_t(d->N) * 4, 256);                       // flags
  b += align_up(size_t(d->N) * 4, 256) * 2;                   // cutoff_k32, cutoff_floor
  b += align_up(size_t(d->N) * kChunk * (kChunk / 32) * 4, 256);  // nms_mask
  b += align_up(size_t(d->N) * size_t(cand_capacity(d)) * 8, 256);
  return b;
}

// bytes at the start of the workspace that must be zero before the first call
extern "C" size_t sbod_detect_workspace_zero_bytes(const sbod_detect_desc* d) {
  if (!d) return 0;
  return align_up(size_t(d->N) * 4, 256) + align_up(size_t(d->N) * kMaxBins * 4, 256) +
         align_up(size_t(d->N) * size_t(d->C) * 4, 256) + align_up(size_t(d->N) * kMaxBins * 4, 256) +
         align_up(size_t(d->N) * 4, 256);
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
  if (reinterpret_cast<uintptr_t>(d->scores) & 15) return SBOD_ERR_ALIGNMENT;
  if (reinterpret_cast<uintptr_t>(d->locs) & 15) return SBOD_ERR_ALIGNMENT;
  if (!d->workspace || d->workspace_bytes < sbod_detect_workspace_bytes(d)) return SBOD_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(d->workspace) & 255) return SBOD_ERR_WORKSPACE;

  DetParams q;
  q.locs = d->locs; q.scores = d->scores;
  q.priors_cxcy = reinterpret_cast<const float4*>(d->priors_cxcy);
  q.prior_keep = d->prior_keep;
  q.N = d->N; q.P = d->P; q.C = d->C;
  q.act_kind = d->act_kind; q.box_kind = d->box_kind; q.clamp_inplace = d->clamp_inplace;
  q.min_score = d->min_score; q.max_overlap = d->max_overlap; q.top_k = d->top_k;
  q.second_thr = d->second_nms_thr; q.pre_nms_topk = d->pre_nms_topk;
  q.out_boxes = d->out_boxes; q.out_labels = d->out_labels; q.out_scores = d->out_scores;
  q.out_prior = d->out_prior; q.out_counts = d->out_counts; q.out_cap = d->out_cap;
  level0_layout(d->min_score, &q.shift0, &q.n_bins);
  q.cand_cap = cand_capacity(d);
  q.kcap = kept_capacity(d);
  q.debug_skip = 0;
#ifdef SBOD_DEBUG_HOOKS  // profiling builds only: release builds never read the environment
  {
    const char* e = getenv("SBOD_DEBUG_SKIP");
    q.debug_skip = e ? atoi(e) : 0;
  }
#endif
  unsigned char* w = static_cast<unsigned char*>(d->workspace);
  q.cand_count = reinterpret_cast<unsigned int*>(w); w += align_up(size_t(q.N) * 4, 256);
  q.hist = reinterpret_cast<unsigned int*>(w);       w += align_up(size_t(q.N) * kMaxBins * 4, 256);
  q.class_seen = reinterpret_cast<unsigned int*>(w); w += align_up(size_t(q.N) * size_t(q.C) * 4, 256);
  q.shist = reinterpret_cast<unsigned int*>(w);      w += align_up(size_t(q.N) * kMaxBins * 4, 256);
  q.flags = reinterpret_cast<unsigned int*>(w);      w += align_up(size_t(q.N) * 4, 256);
  q.cutoff_k32 = reinterpret_cast<unsigned int*>(w); w += align_up(size_t(q.N) * 4, 256);
  q.cutoff_floor = reinterpret_cast<float*>(w);      w += align_up(size_t(q.N) * 4, 256);
  q.nms_mask = reinterpret_cast<unsigned int*>(w);   w += align_up(size_t(q.N) * kChunk * (kChunk / 32) * 4, 256);
  q.cand = reinterpret_cast<unsigned long long*>(w);
  q.mode = 0;
  q.speculate = 0;
  q.sample_stride = 26;
  q.sample_target = 4 * (q.top_k + 1) + 1024;
#ifdef SBOD_DEBUG_HOOKS
  if (getenv("SBOD_TARGET")) q.sample_target = atoi(getenv("SBOD_TARGET"));
#endif
  // hist rows are n_bins wide inside the kMaxBins-strided allocation
  const DTiling t = choose_dtiling(q.C);
  q.rows_per_tile = t.rows;
  q.tiles_per_image = (q.P + t.rows - 1) / t.rows;
  q.n_tiles = q.tiles_per_image * q.N;
  q.n_stages = t.stages;
  q.stage_floats = t.stage_floats;

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    SBOD_CUDA_TRY(cudaFuncSetAttribute(detect_score_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SBOD_CUDA_TRY(cudaFuncSetAttribute(detect_score_fast_kernel<81>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SBOD_CUDA_TRY(cudaFuncSetAttribute(detect_score_fast_kernel<21>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SBOD_CUDA_TRY(cudaFuncSetAttribute(detect_score_fast_kernel<0>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    SBOD_CUDA_TRY(cudaFuncSetAttribute(detect_nms_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_once.mark();
  }
  int ctas_per_sm = int((220 * 1024) / (t.smem + 2 * 1024));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm > 8) ctas_per_sm = 8;
  int grid = sm_count() * ctas_per_sm;
  if (grid > q.n_tiles) grid = q.n_tiles;
  const bool fast = q.C >= 2 && q.C <= 128;  // odd C is bank-conflict free, even C only slower in smem
  if (fast) {
    q.rows_per_tile = kTileRows;
    q.stage_floats = uint32_t(align_up(size_t(kTileRows) * q.C + 8, 32));
    const size_t sb = size_t(q.stage_floats) * 4;
    int ctas = 2;
    q.n_stages = int((96 * 1024) / sb);
    if (q.n_stages < 2) {
      q.n_stages = int((200 * 1024) / sb);
      ctas = 1;
    }
    if (q.n_stages > 4) q.n_stages = 4;
    q.tiles_per_image = (q.P + kTileRows - 1) / kTileRows;
    q.n_tiles = q.tiles_per_image * q.N;
    const size_t fsmem = size_t(q.n_stages) * sb + 4 * 8;
    int fgrid = sm_count() * ctas;
    if (fgrid > q.n_tiles) fgrid = q.n_tiles;
    auto launch_score = [&](int g, const DetParams& qq) {
      if (qq.C == 81) detect_score_fast_kernel<81><<<g, kStreamThreads, fsmem, st>>>(qq);       // COCO
      else if (qq.C == 21) detect_score_fast_kernel<21><<<g, kStreamThreads, fsmem, st>>>(qq);  // VOC
      else detect_score_fast_kernel<0><<<g, kStreamThreads, fsmem, st>>>(qq);
    };
    // Speculation pays when the sampled histogram is cheap relative to the main pass.
    const bool speculate = q.tiles_per_image >= 2 * q.sample_stride && !(q.debug_skip & 4);
    q.speculate = speculate ? 1 : 0;
    if ((stage_mask & 1) && speculate) {
      {
        DetParams qs = q;
        qs.mode = 1;
        int sgrid = (q.n_tiles / q.sample_stride + 1);
        if (sgrid > fgrid) sgrid = fgrid;
        launch_score(sgrid, qs);
        SBOD_LAUNCH_CHECK();
      }
    }
    if (stage_mask & 4) {
      launch_score(fgrid, q);
      SBOD_LAUNCH_CHECK();
    }
    if (stage_mask & 2) {
      const size_t nms_smem_f = ((sizeof(NmsSmem) + 127) & ~size_t(127)) + size_t(q.kcap) * 25 + 256 +
                                size_t(q.C + 1) * 4;
      if (nms_smem_f > 220 * 1024) return SBOD_ERR_UNSUPPORTED;
      detect_nms_kernel<<<q.N, kNmsThreads, nms_smem_f, st>>>(q);
      SBOD_LAUNCH_CHECK();
      if (speculate) {  // exact fallback for the images whose candidates ran out (usually none)
        DetParams qf = q;
        qf.mode = 2;
        launch_score(fgrid, qf);
        SBOD_LAUNCH_CHECK();
        detect_nms_kernel<<<q.N, kNmsThreads, nms_smem_f, st>>>(qf);
        SBOD_LAUNCH_CHECK();
      }
      return SBOD_OK;
    }
    if (!(stage_mask & 2)) return SBOD_OK;
  } else if (stage_mask & 4) {
    detect_score_kernel<<<grid, kDRows, t.smem, st>>>(q);
    SBOD_LAUNCH_CHECK();
  }
  if (!(stage_mask & 2)) return SBOD_OK;
  const size_t nms_smem = ((sizeof(NmsSmem) + 127) & ~size_t(127)) + size_t(q.kcap) * 25 + 256 +
                          size_t(q.C + 1) * 4;
  if (nms_smem > 220 * 1024) return SBOD_ERR_UNSUPPORTED;
  detect_nms_kernel<<<q.N, kNmsThreads, nms_smem, st>>>(q);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" int sbod_detect(const sbod_detect_desc* d, sbod_stream_t stream) {
  return detect_run(d, stream, 7);
}

// Profiling / bench hook. stage 0 = sampling pass + main score pass, 1 = NMS kernel (+ fallback
// passes), 2 = sampling pass only, 3 = main score pass only. Score passes must be followed by a
// stage-1 launch before the next full sbod_detect (it consumes and cleans the workspace).
extern "C" int sbod_detect_stage(const sbod_detect_desc* d, int stage, sbod_stream_t stream) {
  static const int masks[5] = {1 | 4, 2, 1, 4, 4 | 2};
  if (stage < 0 || stage > 4) return SBOD_ERR_INVALID;
  return detect_run(d, stream, masks[stage]);
}
