// loss.cu — the fused train path: assignment + log-sum-exp + loc/conf losses + hard-negative
// mining, forward and backward (sm_100a).
//
// Kernels (forward = match_lse* -> classify -> mine; backward = bwd_patch, or loss_bwd for focal)
//   match_lse_fast_kernel<C>  2 <= C <= 128 (C = 81 / 21 compile-time): warp-specialised, persistent.
//                      Stream warps: logits tiles HBM->smem by 1-D bulk TMA through an mbarrier ring,
//                      one-pass log-sum-exp + background CE, optional zero-fill of the logits gradient.
//                      Match warps (match role): prior<->GT IoU with slice-bbox culling, both arg-maxes,
//                      ticket queues per image.
//   match_lse_kernel   generic fallback (C > 128, or no logits: sbod_assign): one thread per prior row.
//   forced_match_kernel (sbod_assign) per image: the reference's "every object keeps its best prior"
//                      override, including its filtered-index quirk (SSD512.py:546-553).
//   classify_kernel    2048 priors per CTA: forced-match list rebuilt from the per-object keys and applied,
//                      classes, true-class CE + loc term of the foreground rows, mining candidates counted in
//                      a fine per-image histogram, partial sums of the slice.
//   mine_kernel        one CTA per image: sums its 3 * n_pos largest candidate CEs (histogram bin + one pass
//                      over the candidates, no sort) and leaves the threshold for the backward; the last image
//                      folds the batch (and mines it for SSD300's batch-global rule, SSD300.py:580-588),
//                      exchanges the sums with the other ranks and finalises the loss.
//                      (programmatic dependent launches: match -> classify -> mine)
//   bwd_patch_kernel   sparse backward (CE + mining): softmax - onehot on the selected rows of a
//                      zero-filled gradient (zero_fill_kernel when the forward did not fill it), grad wrt locs.
//   loss_bwd_kernel    dense backward (focal): tile in by TMA, transformed in place, out by TMA store.
#include <math.h>
#include <stdlib.h>

#include "comm.cuh"
#include "common.cuh"
#include "pair_iou.cuh"
#include "row_stream.cuh"

namespace sbod {

constexpr int kRows = 128;      // threads per CTA == max rows per streamed tile
constexpr int kGtChunk = 256;   // GT boxes staged in smem at a time
constexpr int kMaxStages = 4;
constexpr int kBins = 4096;     // bins of the mining-candidate histogram

struct LossParams {
  const float* locs;
  const float* scores;
  const float4* priors_cxcy;
  const float4* priors_xy;
  const float4* anchors_xy;  // per image, or null
  const float4* gt_boxes;
  const int64_t* gt_labels;
  const int32_t* gt_offsets;
  const uint8_t* exclude;
  int N, P, C, gmax;
  float thr_pos, thr_neg;
  int reg_kind, cls_kind, binarize, ratio;
  float reg_weight, beta, falpha, fgamma;
  float* ov;
  int32_t* obj;
  float* lse;
  float* ce;
  uint8_t* sel;
  float* sel_thr;  // [N,2] per image: CE threshold of the mined negatives, 1 if every tie at the threshold is taken
  double* partials;
  double* sums;
  float* loss;
  // workspace
  unsigned long long* gtkey;  // [N, gmax]
  float* cand;                // [N, P]
  unsigned int* counters;     // [4]
  unsigned int* match_q;      // [N] per-image chunk queues of the match role
  unsigned int* sel_hist;     // [N, kBins] per image: histogram of the mining candidates (sel_bin; zero between calls)
  double* blockpart;          // [N, slices, 4] partial sums of the slices
  float* sel_seg;             // [N, kSelCap] batch-global mining: each image's candidates inside the batch's threshold bin
  int* sel_segn;              // [N] ... their number (-1: the bin does not fit, whole-batch path)
  double* sel_above;          // [N] ... and each image's sum of the candidates above the bin
  // tiling
  int rows_per_tile, tiles_per_image, n_tiles, n_stages;
  uint32_t stage_floats;
  int with_scores;
  int fast;        // odd C <= 128: warp-specialised two-threads-per-row kernels
  int ctas_per_sm;
  int debug_skip;  // SBOD_DEBUG_SKIP: bit0 = no stream role, bit1 = no match role (profiling only)
  float* prefill;  // grad wrt logits to zero-fill while streaming (or null)
  const CommDev* comm;  // in-kernel all-reduce of the loss sums over the ranks (or null)
};

SBOD_DEVINL int64_t map_label(const LossParams& q, int64_t lab) {
  return q.binarize ? (lab > 0 ? 1 : 0) : lab;
}

// ------------------------------------------------------------------------------------------
// match_lse_kernel
// ------------------------------------------------------------------------------------------
struct TileCoord {
  int n, p0, rows;
};
SBOD_DEVINL TileCoord tile_coord(const LossParams& q, int tile) {
  TileCoord t;
  t.n = tile / q.tiles_per_image;
  t.p0 = (tile - t.n * q.tiles_per_image) * q.rows_per_tile;
  t.rows = min(q.rows_per_tile, q.P - t.p0);
  return t;
}

SBOD_DEVINL void issue_tile_load(const LossParams& q, int tile, float* stage, uint64_t* bar) {
  const TileCoord t = tile_coord(q, tile);
  const size_t first = (size_t(t.n) * q.P + t.p0) * size_t(q.C);
  const size_t total = size_t(q.N) * q.P * size_t(q.C);
  const TileSpan s = make_tile_span(q.scores, first, size_t(t.rows) * q.C, total);
  for (uint32_t i = 0; i < s.tail_floats; ++i)
    stage[s.bulk_bytes / 4 + i] = q.scores[(s.src16 - q.scores) + s.bulk_bytes / 4 + i];
  if (s.bulk_bytes) {
    mbar_arrive_expect_tx(bar, s.bulk_bytes);
    tma_load_1d_hint(stage, s.src16, s.bulk_bytes, bar, l2_policy_evict_first());
  } else {
    mbar_arrive(bar);
  }
}

__global__ void __launch_bounds__(kRows) match_lse_kernel(const LossParams q) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stages = reinterpret_cast<float*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + size_t(q.n_stages) * q.stage_floats * 4);

  __shared__ float4 s_gbox[kGtChunk];
  __shared__ float s_garea[kGtChunk];
  __shared__ unsigned long long s_gkey[kGtChunk];
  __shared__ uint8_t s_gskip[kGtChunk];

  const int tid = threadIdx.x, lane = tid & 31;
  const int n_my = (q.n_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);

  if (q.with_scores) {
    if (tid == 0) {
      for (int s = 0; s < q.n_stages; ++s) mbar_init(&bars[s], 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
      for (int s = 0; s < q.n_stages && s < n_my; ++s)
        issue_tile_load(q, blockIdx.x + s * gridDim.x, stages + size_t(s) * q.stage_floats, &bars[s]);
    }
  }

  // rotation that makes the per-thread row walk bank-conflict free for even C
  int gcd = 1;
  while (gcd < 32 && (q.C % (gcd * 2)) == 0) gcd *= 2;
  const int rot = (lane * gcd) >> 5;

  for (int it = 0; it < n_my; ++it) {
    const int tile = blockIdx.x + it * gridDim.x;
    const TileCoord tc = tile_coord(q, tile);
    const int n = tc.n;
    const int g0 = q.gt_offsets[n];
    const int G = q.gt_offsets[n + 1] - g0;
    const bool valid = tid < tc.rows;
    const int p = tc.p0 + tid;
    const size_t np = size_t(n) * q.P + (valid ? p : tc.p0);

    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) a = q.anchors_xy ? q.anchors_xy[np] : q.priors_xy[p];
    const float aa = box_area_rn(a);
    const bool azero = anchor_is_zero(a);
    const bool a_ok = (a.z >= a.x) && (a.w >= a.y);
    const bool active = valid && !azero;

    // warp bounding box of the active priors, for GT culling
    const float INF = __int_as_float(0x7f800000);
    const float bx1 = warp_min(active ? a.x : INF), by1 = warp_min(active ? a.y : INF);
    const float bx2 = warp_max(active ? a.z : -INF), by2 = warp_max(active ? a.w : -INF);

    // running arg-max over objects, first index wins (torch.max semantics, metrics/SSD512.py:538)
    float best = -INF;
    int bobj = 0;

    for (int c0 = 0; c0 < G; c0 += kGtChunk) {
      const int gc = min(kGtChunk, G - c0);
      __syncthreads();
      bool g_ok = true;
      for (int i = tid; i < gc; i += kRows) {
        const float4 g = q.gt_boxes[g0 + c0 + i];
        s_gbox[i] = g;
        s_garea[i] = box_area_rn(g);
        s_gskip[i] = gt_is_zero(g) ? 1 : 0;
        s_gkey[i] = 0ull;
        g_ok = g_ok && (g.z >= g.x) && (g.w >= g.y);
      }
      // With non-negative widths/heights everywhere a pair that does not intersect is exactly +0,
      // so objects outside the warp's bounding box can be skipped without changing any result.
      const bool chunk_ok = __syncthreads_and(int(g_ok && (a_ok || !valid))) != 0;

      if (chunk_ok) {
        float cbest = azero ? -1.f : 0.f;  // value of every culled / zero-size object
        int cidx = c0;
        for (int base = 0; base < gc; base += 32) {
          const int i = base + lane;
          bool hit = false;
          if (i < gc && !s_gskip[i]) {
            const float4 g = s_gbox[i];
            hit = (g.z > bx1) && (g.x < bx2) && (g.w > by1) && (g.y < by2);
          }
          unsigned m = __ballot_sync(0xffffffffu, hit);
          while (m) {
            const int j = base + (__ffs(m) - 1);
            m &= m - 1;
            if (active) {
              const float iou = iou_metrics_rn(s_gbox[j], s_garea[j], a, aa);
              if (iou > cbest) {
                cbest = iou;
                cidx = c0 + j;
              }
              if (iou > 0.f) {
                const unsigned long long key =
                    (static_cast<unsigned long long>(__float_as_uint(iou)) << 32) |
                    (0xffffffffu - unsigned(p));
                if (key > s_gkey[j]) atomicMax(&s_gkey[j], key);
              }
            }
          }
        }
        if (cbest > best) {
          best = cbest;
          bobj = cidx;
        }
      } else if (valid) {
        // general path: every pair, masks applied as the reference does (metrics.py:249-250)
        for (int j = 0; j < gc; ++j) {
          float iou = iou_metrics_rn(s_gbox[j], s_garea[j], a, aa);
          if (s_gskip[j]) iou = 0.f;
          if (azero) iou = -1.f;
          if (iou > best) {
            best = iou;
            bobj = c0 + j;
          }
          if (iou > 0.f) {
            const unsigned long long key =
                (static_cast<unsigned long long>(__float_as_uint(iou)) << 32) |
                (0xffffffffu - unsigned(p));
            if (key > s_gkey[j]) atomicMax(&s_gkey[j], key);
          }
        }
      }
      __syncthreads();
      for (int i = tid; i < gc; i += kRows) {
        const unsigned long long k = s_gkey[i];
        if (k) atomicMax(&q.gtkey[size_t(n) * q.gmax + c0 + i], k);
      }
    }
    if (G == 0) {  // the reference raises on an image without objects; we define "all background"
      best = 0.f;
      bobj = 0;
    }

    if (!q.with_scores) {
      if (valid) {
        q.ov[np] = best;
        q.obj[np] = bobj;
      }
      continue;
    }

    // ---- scores tile: log-sum-exp + CE against the provisional class ----
    const int s = it % q.n_stages;
    float* stage = stages + size_t(s) * q.stage_floats;
    mbar_wait(&bars[s], (it / q.n_stages) & 1);
    if (valid) {
      const size_t first = (size_t(n) * q.P + tc.p0) * size_t(q.C);
      const float* row = stage + (first & 3) + size_t(tid) * q.C;
      const int C = q.C;
      float m0 = -INF, m1 = -INF, m2 = -INF, m3 = -INF;
      int k = rot;
      for (; k + 3 < C; k += 4) {
        m0 = fmaxf(m0, row[k]);
        m1 = fmaxf(m1, row[k + 1]);
        m2 = fmaxf(m2, row[k + 2]);
        m3 = fmaxf(m3, row[k + 3]);
      }
      for (; k < C; ++k) m0 = fmaxf(m0, row[k]);
      for (k = 0; k < rot; ++k) m1 = fmaxf(m1, row[k]);
      const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      k = rot;
      for (; k + 3 < C; k += 4) {
        s0 += __expf(row[k] - mx);
        s1 += __expf(row[k + 1] - mx);
        s2 += __expf(row[k + 2] - mx);
        s3 += __expf(row[k + 3] - mx);
      }
      for (; k < C; ++k) s0 += __expf(row[k] - mx);
      for (k = 0; k < rot; ++k) s1 += __expf(row[k] - mx);
      const float lg = logf((s0 + s1) + (s2 + s3));
      q.ov[np] = best;
      q.obj[np] = bobj;
      q.lse[np] = mx + lg;
      // CE against class 0 (background) == -log_softmax[0] in torch's operation order; the few
      // positive rows are re-evaluated against their true class by classify_kernel / mine_kernel.
      q.ce[np] = (mx - row[0]) + lg;
    }
    __syncthreads();  // every thread is done with stage s
    if (tid == 0 && it + q.n_stages < n_my)
      issue_tile_load(q, blockIdx.x + (it + q.n_stages) * gridDim.x, stage, &bars[s]);
  }
}

// ------------------------------------------------------------------------------------------
// match_lse_fast_kernel — warp-specialised version for odd C <= 128 (C = 81, 21, ...).
//   warps 0-7  (256 threads) "stream" role: consume the TMA ring, two threads per row,
//              log-sum-exp + background CE per prior. Never touches the ground truth.
//   warps 8-15 (256 threads) "match" role: prior<->GT IoU with slice-bbox culling, both arg-maxes.
//              Never touches the logits. The object group is staged in a per-warp shared buffer and
//              consumed two objects at a time; an object's best prior inside a slice is one REDUX +
//              one ballot and is kept as a key in a register of the lane that owns the object.
// The two roles share nothing but the SM: ALU-bound matching hides under the memory-bound stream.
// ------------------------------------------------------------------------------------------
constexpr int kMatchWarps = 8;
constexpr int kKeyGroups = 4;  // object keys kept in registers: 32 * kKeyGroups objects per image
constexpr int kZeroFloats = 1024;  // 4 KB zero tile for the gradient prefill
constexpr int kMatchThreads = 32 * kMatchWarps;
constexpr int kFastThreads = kStreamThreads + kMatchThreads;

template <int kC>  // kC > 0: the class count is a compile-time constant (unrolled stream role)
__global__ void __launch_bounds__(kFastThreads, 2) match_lse_fast_kernel(const LossParams q) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stages = reinterpret_cast<float*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + size_t(q.n_stages) * q.stage_floats * 4);
  __shared__ __align__(16) float4 s_gt[kMatchWarps][32];
  __shared__ float s_ga[kMatchWarps][32];
  __shared__ __align__(128) float s_zero[kZeroFloats];  // source of the gradient zero-fill stores

  const int tid = threadIdx.x, lane = tid & 31;
  int t0, t1;
  tile_range(q.n_tiles, blockIdx.x, gridDim.x, t0, t1);
  const int n_my = t1 - t0;
  const int C = kC ? kC : q.C;
  if (tid == 0) {
    for (int s = 0; s < q.n_stages; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < kZeroFloats; i += kFastThreads) s_zero[i] = 0.f;
  fence_proxy_async();
  __syncthreads();

  if (tid < kStreamThreads) {
    // =========================== stream role ===========================
    if (q.debug_skip & 1) return;
    if (tid == 0) {
      for (int s = 0; s < q.n_stages && s < n_my; ++s)
        stream_issue(q.scores, q.N, q.P, C, stream_tile(t0 + s, q.tiles_per_image, kTileRows, q.P),
                     stages + size_t(s) * q.stage_floats, &bars[s]);
    }
    int row, h;
    stream_map(tid, row, h);
    const int nh = (C + 1 - h) >> 1;
    int cur_n = t0 / q.tiles_per_image;  // tile coordinates advance incrementally (no division in the loop)
    int cur_t = t0 - cur_n * q.tiles_per_image;
    int s = 0;
    uint32_t parity = 0;
    for (int it = 0; it < n_my; ++it) {
      StreamTile tc;
      tc.n = cur_n;
      tc.p0 = cur_t * kTileRows;
      tc.rows = min(kTileRows, q.P - tc.p0);
      if (++cur_t == q.tiles_per_image) {
        cur_t = 0;
        ++cur_n;
      }
      float* stage = stages + size_t(s) * q.stage_floats;
      mbar_wait(&bars[s], parity);
      const uint32_t head = ((uint32_t(tc.n) * uint32_t(q.P) + uint32_t(tc.p0)) * uint32_t(C)) & 3u;
      const int r = min(row, tc.rows - 1);  // keep every lane in the shuffles
      const float* rbase = stage + head + r * C;
      // One pass: shift by the row's own background logit, whose term is then 1, so the sum cannot
      // underflow; when it overflows (some logit > background + 88) the warp redoes the row with the
      // usual max shift.
      const float x0 = rbase[0];
      float sum = kC ? pair_row_sumexp_fixed<kC>(rbase, h, -x0 * kLog2e)
                     : half_row_sumexp(rbase + h, nh, -x0 * kLog2e);
      float shift = x0;
      if (__any_sync(0xffffffffu, !(sum < __int_as_float(0x7f800000)))) {
        shift = half_row_max(rbase + h, nh);
        sum = half_row_sumexp(rbase + h, nh, -shift * kLog2e);
      }
      if (h == 0 && row < tc.rows) {
        const float lg = logf(sum);
        const size_t np = size_t(tc.n) * q.P + tc.p0 + row;
        q.lse[np] = shift + lg;
        q.ce[np] = (shift - x0) + lg;  // background CE
      }
      named_bar_sync(1, kStreamThreads);  // the stream group is done with stage s
      if (tid == 0 && it + q.n_stages < n_my)
        stream_issue(q.scores, q.N, q.P, C,
                     stream_tile(t0 + it + q.n_stages, q.tiles_per_image, kTileRows, q.P), stage, &bars[s]);
      if (tid == 32 && q.prefill) {
        const uint64_t zero_policy = l2_policy_evict_first();
        // zero-fill this tile of the gradient buffer: 16-byte aligned interior by bulk stores from
        // the zero tile (fire and forget: the source never changes), ragged ends by plain stores
        const size_t e0 = (size_t(tc.n) * q.P + tc.p0) * size_t(C), e1 = e0 + size_t(tc.rows) * C;
        const size_t a0 = (e0 + 3) & ~size_t(3), a1 = e1 & ~size_t(3);
        for (size_t e = e0; e < (a0 < e1 ? a0 : e1); ++e) q.prefill[e] = 0.f;
        for (size_t e = (a1 > a0 ? a1 : (a0 < e1 ? a0 : e1)); e < e1; ++e) q.prefill[e] = 0.f;
        for (size_t a = a0; a < a1; a += kZeroFloats) {
          const size_t len = (a1 - a) < size_t(kZeroFloats) ? (a1 - a) : size_t(kZeroFloats);
          tma_store_1d_hint(q.prefill + a, s_zero, uint32_t(len * 4), zero_policy);
        }
        tma_store_commit();
      }
      if (++s == q.n_stages) {
        s = 0;
        parity ^= 1u;
      }
    }
    if (tid == 32 && q.prefill) tma_store_wait_all<0>();
    return;
  }

  // =========================== match role ===========================
  // Every match WARP works on its own: it pulls 32-prior slices from per-image queues (image affinity:
  // a warp starts at its CTA's home image and moves on when that queue runs dry; heaviest slices —
  // the large priors at the end of an image — first; the next slice index is requested before the
  // current one is processed). Ground truth is read through L1 (an image's boxes are ~2 KB) and
  // handed from lane to lane with shuffles, so the role needs no barrier at all.
  if (q.debug_skip & 2) return;
  const int wm = (tid - kStreamThreads) >> 5;
  const float INF = __int_as_float(0x7f800000);
  const int slices_per_image = (q.P + 31) / 32;
  float4* sgt = s_gt[wm];  // this warp's staging of one group of 32 objects (+ their areas)
  float* sga = s_ga[wm];
  int prev_n = -1, g0 = 0, G = 0;
  // Best prior of object 32*u + lane seen by this warp in the current image, as a sortable key
  // (IoU bits : ~prior) held in a register of lane `lane`; flushed with one atomicMax per object
  // when the warp moves to another image. Objects beyond 32*kKeyGroups go straight to global memory.
  unsigned long long kreg[kKeyGroups];
#pragma unroll
  for (int u = 0; u < kKeyGroups; ++u) kreg[u] = 0ull;

  auto flush_keys = [&](int n) {
#pragma unroll
    for (int u = 0; u < kKeyGroups; ++u) {
      if (kreg[u]) atomicMax(&q.gtkey[size_t(n) * q.gmax + 32 * u + lane], kreg[u]);
      kreg[u] = 0ull;
    }
  };

  // Tickets: the queue counter of an image is ONE address and claims on it serialise in L2 (~40 ns
  // each), so the light slices (the fine feature maps, two thirds of an image) go four, two or one
  // to a ticket depending on how many objects the image has (a light slice costs ~0.5 us plus
  // ~0.15 us per object it meets); the heavy ones (large priors that meet every object) stay single.
  const int heavy = slices_per_image / 3;
  const int light = slices_per_image - heavy;
  const int by_size = light > 1200 ? 4 : (light > 600 ? 2 : 1);  // very large prior sets: cap the ticket count
  auto chunk_of = [&](int Gn) { return max(by_size, Gn <= 24 ? 4 : (Gn <= 64 ? 2 : 1)); };
  auto tickets_of = [&](int Gn) {
    const int ch = chunk_of(Gn);
    return heavy + (slices_per_image - heavy + ch - 1) / ch;
  };
  int chunk = 1, n_tickets = 0;
  auto enter_image = [&](int n) {  // objects, ticket geometry; the previous image's keys are flushed
    if (prev_n >= 0) flush_keys(prev_n);
    g0 = q.gt_offsets[n];
    G = q.gt_offsets[n + 1] - g0;
    prev_n = n;
    chunk = chunk_of(G);
    n_tickets = tickets_of(G);
  };
  int img_i = 0;
  int cur_n = int(blockIdx.x % q.N);
  enter_image(cur_n);
  int nxt = 0;
  if (lane == 0) nxt = int(atomicAdd(&q.match_q[cur_n], 1u));
  for (;;) {
    int item = __shfl_sync(0xffffffffu, nxt, 0);
    while (item >= n_tickets) {
      // This image is drained. Look at 32 other queues at once (plain loads) and jump to the first
      // one that still has work, instead of probing them one atomic at a time.
      bool found = false;
      while (!found) {
        const int remaining = q.N - 1 - img_i;  // images after cur_n that were not visited yet
        if (remaining <= 0) break;
        const int span = min(32, remaining);
        const int cand = (cur_n + 1 + lane) % q.N;
        bool has = false;
        if (lane < span)
          has = *reinterpret_cast<volatile unsigned int*>(&q.match_q[cand]) <
                unsigned(tickets_of(q.gt_offsets[cand + 1] - q.gt_offsets[cand]));
        const unsigned bal = __ballot_sync(0xffffffffu, has);
        if (bal) {
          const int skip = __ffs(bal) - 1;
          cur_n = (cur_n + 1 + skip) % q.N;
          img_i += skip + 1;
          found = true;
        } else {
          cur_n = (cur_n + span) % q.N;
          img_i += span;
        }
      }
      if (!found) break;
      enter_image(cur_n);
      if (lane == 0) nxt = int(atomicAdd(&q.match_q[cur_n], 1u));
      item = __shfl_sync(0xffffffffu, nxt, 0);
    }
    if (item >= n_tickets) break;
    if (lane == 0) nxt = int(atomicAdd(&q.match_q[cur_n], 1u));  // prefetch the next ticket
    const int n = cur_n;
    const int slice0 = item < heavy ? item : heavy + chunk * (item - heavy);
    const int slice1 = item < heavy ? slice0 + 1 : min(slice0 + chunk, slices_per_image);
   for (int slice = slice0; slice < slice1; ++slice) {
    const int ps = (slices_per_image - 1 - slice) * 32;  // heaviest (last) slices first
    const int p = ps + lane;
    const bool valid = p < q.P;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) a = q.anchors_xy ? q.anchors_xy[size_t(n) * q.P + p] : q.priors_xy[p];
    const float aa = box_area_rn(a);
    const bool azero = anchor_is_zero(a);
    const bool active = valid && !azero;
    const bool a_ok = ((a.z >= a.x) && (a.w >= a.y)) || !valid;
    // (areas up to 2^59 keep every IoU denominator of the fast path below 2^60: see the division test below)
    const bool warp_a_ok = __all_sync(0xffffffffu, a_ok && aa <= 5.76460752303e17f) != 0;
    // bounding box of the warp's priors
    const float bx1 = warp_min_redux(active ? a.x : INF), by1 = warp_min_redux(active ? a.y : INF);
    const float bx2 = warp_max_redux(active ? a.z : -INF), by2 = warp_max_redux(active ? a.w : -INF);
    float best = -INF;  // running arg-max over objects, first index wins
    int bobj = 0;

    // one group of 32 objects; `kr` is the key register of the group (in_reg) or unused
    auto do_group = [&](int base, unsigned long long& kr, bool in_reg) {
      const int gi = base + lane;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gi < G) g = q.gt_boxes[g0 + gi];
      const bool gskip = gi >= G || gt_is_zero(g);
      const float ga_own = box_area_rn(g);
      const bool grp_ok = __all_sync(0xffffffffu, (g.z >= g.x) && (g.w >= g.y) && ga_own <= 5.76460752303e17f) != 0;
      if (grp_ok && warp_a_ok) {
        // fast path: a pair that does not intersect is exactly +0, so only objects touching the
        // warp's bounding box are evaluated, two at a time (independent dependency chains)
        const bool hit = !gskip && (g.z > bx1) && (g.x < bx2) && (g.w > by1) && (g.y < by2);
        unsigned m = __ballot_sync(0xffffffffu, hit);
        float cbest = azero ? -1.f : 0.f;
        int cidx = base;
        if (m) {
          sgt[lane] = g;
          sga[lane] = ga_own;
          __syncwarp();
          while (m) {
            const int sA = __ffs(m) - 1;
            m &= m - 1;
            const int sB = m ? __ffs(m) - 1 : sA;  // odd count: the last object twice (idempotent)
            m &= m - 1;
            const float4 gA = sgt[sA], gB = sgt[sB];
            const float gaA = sga[sA], gaB = sga[sB];
            // straight-line code (both boxes of a pair have w, h >= 0 here, so the denominator is
            // >= eps and an empty intersection gives exactly +0, as in the reference)
            const float inA = inter_rn(gA, a), inB = inter_rn(gB, a);
            const float dA = __fadd_rn(__fsub_rn(__fadd_rn(gaA, aa), inA), kEps);
            const float dB = __fadd_rn(__fsub_rn(__fadd_rn(gaB, aa), inB), kEps);
            float qA, qB;
            // div_fast_ok(in, d) wants in == 0 or 2^-60 <= in <= 2^60, and 2^-60 <= d <= 2^60. In this path every
            // width and height is >= 0 and every area <= 2^59, so  in <= min(area) <= d,  EPS <= d <= 2^60  hold by
            // construction (monotonic rounding): what is left per pair is "in is not a tiny positive number".
            const float tiny = 8.67361737988e-19f;  // 2^-60
            if (__all_sync(0xffffffffu, !(inA > 0.f && inA < tiny) && !(inB > 0.f && inB < tiny))) {
              qA = div_rn_fast(inA, dA);
              qB = div_rn_fast(inB, dB);
            } else {
              qA = __fdiv_rn(inA, dA);
              qB = __fdiv_rn(inB, dB);
            }
            const float idle = azero ? -1.f : 0.f;
            const float iouA = active ? qA : idle, iouB = active ? qB : idle;
            if (iouA > cbest) {
              cbest = iouA;
              cidx = base + sA;
            }
            if (iouB > cbest) {
              cbest = iouB;
              cidx = base + sB;
            }
            // best prior of each object inside this slice: max IoU bits, then the lowest lane
            const unsigned bA = active ? __float_as_uint(iouA) : 0u, bB = active ? __float_as_uint(iouB) : 0u;
            const unsigned mA = __reduce_max_sync(0xffffffffu, bA), mB = __reduce_max_sync(0xffffffffu, bB);
            const unsigned vA = __ballot_sync(0xffffffffu, bA == mA), vB = __ballot_sync(0xffffffffu, bB == mB);
            const unsigned long long kA =
                (static_cast<unsigned long long>(mA) << 32) | (0xffffffffu - unsigned(ps + __ffs(vA) - 1));
            const unsigned long long kB =
                (static_cast<unsigned long long>(mB) << 32) | (0xffffffffu - unsigned(ps + __ffs(vB) - 1));
            if (in_reg) {
              if (lane == sA && mA && kA > kr) kr = kA;
              if (lane == sB && mB && kB > kr) kr = kB;
            } else {
              if (lane == 0 && mA) atomicMax(&q.gtkey[size_t(n) * q.gmax + base + sA], kA);
              if (lane == 1 && mB) atomicMax(&q.gtkey[size_t(n) * q.gmax + base + sB], kB);
            }
          }
          __syncwarp();  // the staging is free for the next group
        }
        if (cbest > best) {
          best = cbest;
          bobj = cidx;
        }
      } else {
        // general path: every pair, masks applied as the reference does (metrics.py:249-250)
        const int gcnt = min(32, G - base);
        for (int src = 0; src < gcnt; ++src) {
          float4 gj;
          gj.x = __shfl_sync(0xffffffffu, g.x, src);
          gj.y = __shfl_sync(0xffffffffu, g.y, src);
          gj.z = __shfl_sync(0xffffffffu, g.z, src);
          gj.w = __shfl_sync(0xffffffffu, g.w, src);
          const bool sk = __shfl_sync(0xffffffffu, int(gskip), src) != 0;
          float iou = iou_metrics_rn(gj, box_area_rn(gj), a, aa);
          if (sk) iou = 0.f;
          if (azero) iou = -1.f;
          if (valid && iou > best) {
            best = iou;
            bobj = base + src;
          }
          if (valid && iou > 0.f)
            atomicMax(&q.gtkey[size_t(n) * q.gmax + base + src],
                      (static_cast<unsigned long long>(__float_as_uint(iou)) << 32) | (0xffffffffu - unsigned(p)));
        }
      }
    };

#pragma unroll
    for (int u = 0; u < kKeyGroups; ++u)
      if (32 * u < G) do_group(32 * u, kreg[u], true);
    for (int base = 32 * kKeyGroups; base < G; base += 32) {
      unsigned long long dummy = 0ull;
      do_group(base, dummy, false);
    }
    if (G == 0) {
      best = 0.f;
      bobj = 0;
    }
    if (valid) {
      q.ov[size_t(n) * q.P + p] = best;
      q.obj[size_t(n) * q.P + p] = bobj;
    }
   }
  }
  if (prev_n >= 0) flush_keys(prev_n);
  // the last match warp of the grid resets the queues for the next launch
  if (lane == 0) {
    __threadfence();
    if (atomicAdd(&q.counters[2], 1u) == gridDim.x * kMatchWarps - 1) {
      for (int i = 0; i < q.N; ++i) q.match_q[i] = 0u;
      q.counters[2] = 0u;
    }
  }
}

// ------------------------------------------------------------------------------------------
// forced_match_kernel: one CTA per image
// ------------------------------------------------------------------------------------------
// Works for any block size that is a multiple of 32 (<= 1024). s_prior/s_rank: [gmax] each.
__device__ void forced_match_phase(const LossParams& q, int n, uint32_t* s_prior, int32_t* s_rank,
                                   int* s_warp_tot /*[32]*/, int* s_carry) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nt = blockDim.x, nw = nt >> 5;
  const int g0 = q.gt_offsets[n];
  const int G = q.gt_offsets[n + 1] - g0;
  if (tid == 0) *s_carry = 0;
  __syncthreads();
  // rank j of each object inside the filtered list "objects whose best overlap is > 0"
  for (int base = 0; base < G; base += nt) {
    const int g = base + tid;
    unsigned long long key = 0ull;
    if (g < G) {
      key = q.gtkey[size_t(n) * q.gmax + g];
      q.gtkey[size_t(n) * q.gmax + g] = 0ull;  // leave the workspace clean for the next call
    }
    const bool f = key != 0ull;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    const int in_warp = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) s_warp_tot[wid] = __popc(bal);
    __syncthreads();
    int off = *s_carry;
    for (int w = 0; w < wid; ++w) off += s_warp_tot[w];
    if (g < G) {
      s_prior[g] = 0xffffffffu - uint32_t(key & 0xffffffffull);
      s_rank[g] = f ? off + in_warp : -1;
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < nw; ++w) t += s_warp_tot[w];
      *s_carry += t;
    }
    __syncthreads();
  }
  for (int g = tid; g < G; g += nt) {
    const int j = s_rank[g];
    if (j < 0) continue;
    const uint32_t p = s_prior[g];
    bool winner = true;  // "for j: obj[pr[j]] = j" -> the last j wins
    for (int h = g + 1; h < G; ++h)
      if (s_rank[h] >= 0 && s_prior[h] == p) {
        winner = false;
        break;
      }
    const size_t np = size_t(n) * q.P + p;
    q.ov[np] = 1.0f;
    if (winner) q.obj[np] = j;
  }
}

__global__ void __launch_bounds__(256) forced_match_kernel(const LossParams q) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* s_prior = reinterpret_cast<uint32_t*>(smem_raw);       // [gmax]
  int32_t* s_rank = reinterpret_cast<int32_t*>(s_prior + q.gmax);  // [gmax], -1 = filtered out
  __shared__ int s_warp_tot[32];
  __shared__ int s_carry;
  forced_match_phase(q, blockIdx.x, s_prior, s_rank, s_warp_tot, &s_carry);
}

// ------------------------------------------------------------------------------------------
// loc loss of one positive prior
// ------------------------------------------------------------------------------------------
struct LocTerm {
  float loss;
  float4 grad;  // d loss / d predicted loc (gcx, gcy, gw, gh)
};

template <bool WITH_GRAD>
SBOD_DEVINL LocTerm loc_term(const LossParams& q, const float4 pred, const float4 pri /*cxcy*/,
                             const float4 box /*xyxy*/) {
  LocTerm r;
  r.loss = 0.f;
  r.grad = make_float4(0.f, 0.f, 0.f, 0.f);
  if (q.reg_kind == SBOD_REG_L1_ELEM_MEAN || q.reg_kind == SBOD_REG_SMOOTH_L1) {
    // cxcy_to_gcxgcy(xy_to_cxcy(box), prior): transforms.py:26-34,48-66
    const float cx = (box.z + box.x) / 2.f, cy = (box.w + box.y) / 2.f;
    const float w = box.z - box.x, h = box.w - box.y;
    float t[4] = {(cx - pri.x) / (pri.z / 10.f), (cy - pri.y) / (pri.w / 10.f),
                  logf(w / pri.z) * 5.f, logf(h / pri.w) * 5.f};
    const float pv[4] = {pred.x, pred.y, pred.z, pred.w};
    float gv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float d = pv[i] - t[i];
      const float x = fabsf(d);
      const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
      if (q.reg_kind == SBOD_REG_L1_ELEM_MEAN) {
        r.loss += x;
        gv[i] = sgn;
      } else if (x >= q.beta) {  // Loss.py:214-217
        r.loss += x - 0.5f * q.beta;
        gv[i] = sgn;
      } else {
        r.loss += 0.5f * x * x / q.beta;
        gv[i] = d / q.beta;
      }
    }
    if (WITH_GRAD) r.grad = make_float4(gv[0], gv[1], gv[2], gv[3]);
  } else {
    // decoded = cxcy_to_xy(gcxgcy_to_cxcy(pred, prior)): transforms.py:37-45,69-83
    const float cx = pred.x * pri.z / 10.f + pri.x, cy = pred.y * pri.w / 10.f + pri.y;
    const float w = __expf(pred.z / 5.f) * pri.z, h = __expf(pred.w / 5.f) * pri.w;
    const float4 dec = make_float4(cx - w / 2.f, cy - h / 2.f, cx + w / 2.f, cy + h / 2.f);
    const int kind = q.reg_kind - SBOD_REG_IOU;
    PairGrad pg;
    const float v = pair_overlap<WITH_GRAD>(dec, box, kind, &pg);
    r.loss = 1.0f - v;
    if (WITH_GRAD) {
      const float4 d = make_float4(-pg.d1.x, -pg.d1.y, -pg.d1.z, -pg.d1.w);
      r.grad.x = (d.x + d.z) * pri.z / 10.f;
      r.grad.y = (d.y + d.w) * pri.w / 10.f;
      r.grad.z = (d.z - d.x) * 0.5f * w / 5.f;
      r.grad.w = (d.w - d.y) * 0.5f * h / 5.f;
    }
  }
  return r;
}

SBOD_DEVINL float4 prior_cxcy_of(const LossParams& q, int n, int p) {
  if (q.anchors_xy) {  // RefineDet ODM: xy_to_cxcy(decoded ARM box), RefineDet512.py:885-886
    const float4 a = q.anchors_xy[size_t(n) * q.P + p];
    return make_float4((a.z + a.x) / 2.f, (a.w + a.y) / 2.f, a.z - a.x, a.w - a.y);
  }
  return q.priors_cxcy[p];
}

// softmax focal terms as functions of ce = -log p_t (Loss.py:9-38)
SBOD_DEVINL float focal_fg(const LossParams& q, float ce) {  // alpha * (1-p)^gamma * ce
  const float pt = __expf(-ce);
  return q.falpha * pow_gamma(1.f - pt, q.fgamma) * ce;
}
SBOD_DEVINL float focal_bg(const LossParams& q, float ce) {  // (1-alpha) * p0^gamma * ce  (Loss.py:32)
  const float pt = __expf(-ce);
  return (1.f - q.falpha) * pow_gamma(pt, q.fgamma) * ce;
}

// ------------------------------------------------------------------------------------------
// Hard-negative mining without a sort: only the SUM of the k largest candidate cross entropies is consumed
// (SSD512.py:610-619), plus the k-th value itself as the backward's threshold.
//   * classify_kernel counts every candidate in a 4096-bin histogram of its image while it classifies the
//     priors (shared-memory histogram per CTA, the non-empty bins added to the image's histogram by L2
//     reductions). The bins are the leading 17 bits of the fp32 value over the 16 octaves [2^-10, 2^6) - 256
//     bins per octave, so the bin that holds the k-th value holds a few hundred candidates at most; bin 0
//     takes everything smaller, bin 4095 everything larger (order preserved).
//   * select_topk_sum (one CTA of 1024 threads): finds that bin from the histogram, then ONE pass over the
//     candidate values (L2-resident, 128-bit loads) adds up everything above the bin in double and collects the
//     bin's values in shared memory - in a fixed order, so the sum is reproducible; the k-th value inside the
//     bin is resolved there by 8-bit radix passes. A bin with more values than fit (long runs of equal values)
//     takes the same passes over global memory instead.
// ------------------------------------------------------------------------------------------
constexpr int kSelLo = 117 << 8;  // leading 17 bits (sign 0, exponent, 8 mantissa bits) of 2^-10
constexpr int kSelCap = 4096;     // values of the threshold bin resolved in shared memory

SBOD_DEVINL int sel_bin(float v) {  // v >= 0 (or +inf)
  const int b = int(__float_as_uint(v) >> 15) - kSelLo + 1;
  return b < 1 ? 0 : (b > kBins - 1 ? kBins - 1 : b);
}

SBOD_DEVINL void finalize_loss(const LossParams& q, const double* sums, float* loss) {
  const double s_loc = sums[0], s_pos = sums[1], s_neg = sums[2], npos = sums[3];
  double loc, conf;
  if (q.reg_kind == SBOD_REG_L1_ELEM_MEAN) loc = s_loc / (4.0 * npos);
  else loc = s_loc / npos;
  if (q.cls_kind == SBOD_CLS_FOCAL_SUM) conf = s_pos + s_neg;
  else conf = (s_pos + s_neg) / npos;
  loss[0] = float(conf + double(q.reg_weight) * loc);
  loss[1] = float(conf);
  loss[2] = float(loc);
  loss[3] = float(npos);
}

#ifdef SBOD_DEBUG_HOOKS  // phase time stamps of the classify / mine kernels (profiling builds only)
__device__ unsigned long long g_cm_times[8 + 4 * 64];
SBOD_DEVINL unsigned long long cm_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define CM_STAMP_MIN(i) do { if (threadIdx.x == 0) atomicMin(&g_cm_times[i], cm_now()); } while (0)
#define CM_STAMP_MAX(i) do { if (threadIdx.x == 0) atomicMax(&g_cm_times[i], cm_now()); } while (0)
#define CM_STAMP_SET(i) do { if (threadIdx.x == 0) g_cm_times[i] = cm_now(); } while (0)
#else
#define CM_STAMP_MIN(i)
#define CM_STAMP_MAX(i)
#define CM_STAMP_SET(i)
#endif

constexpr int kMineThreads = 1024;
constexpr int kMineWarps = kMineThreads / 32;

struct MineShared {
  float vals[kSelCap];    // the threshold bin's candidate values
  unsigned int dig[256];  // digit histogram of one radix pass
  double red4[4];
  double red[34];
  unsigned int wt[32];
  int warp_tot[kMineWarps];
  int misc[8];
  int last;
  int segoff[kMineThreads + 1];  // batch-global mining: where each image's segment goes in the merged list
};

struct TopkSum {
  double sum;   // sum of the k largest candidates
  float thr;    // the k-th largest (+inf: nothing selected)
  float ties;   // 1: every candidate equal to thr belongs to the selection
};

// exclusive prefix of one int per thread over the CTA (thread order); *total = sum over the CTA. Two barriers.
SBOD_DEVINL int mine_excl_scan(int v, int* warp_tot, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // (warp_tot may still be read from the previous call)
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  const int t = warp_tot[lane];  // kMineWarps == 32: one total per lane
  int wsum = t;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int x = __shfl_up_sync(0xffffffffu, wsum, o);
    if (lane >= o) wsum += x;
  }
  *total = __shfl_sync(0xffffffffu, wsum, 31);
  const int before_warp = __shfl_sync(0xffffffffu, wsum - t, wid);
  return before_warp + inc - v;
}

// Sum of the k largest candidates of vals[0, n) (candidates are >= 0, everything else is negative or NaN);
// ghist[kBins] = their sel_bin histogram, read and left zero. 0 < k <= number of candidates, or k == 0
// (nothing selected). sel (may be null): flag bytes of the values; only touched when some but not all values
// equal to the threshold are selected (the first ones by index get bit 1). All kMineThreads threads call this.
// --- the pieces of the top-k sum (all kMineThreads threads of one CTA call them) ---

// The bin that holds the k-th largest candidate, from the histogram ghist[kBins] (every thread owns 4 consecutive
// bins, highest bins first; `zero`: the bins are left clean for the next call). Results (CTA-uniform) in
// S.misc[0..2] = bin, rank of the k-th value inside the bin (from the top, 1-based), candidates in the bin.
// k > 0 and k <= number of candidates. Ends with a barrier.
__device__ void find_threshold_bin(MineShared& S, unsigned int* __restrict__ ghist, const long long k, const bool zero,
                                   const uint4* preloaded) {
  const int tid = threadIdx.x;
  constexpr int kPer = kBins / kMineThreads;
  static_assert(kPer == 4, "one 128-bit load per thread");
  unsigned int cnt[kPer];
  {
    uint4* gp = reinterpret_cast<uint4*>(ghist + (kBins - kPer * (tid + 1)));
    const uint4 v = preloaded ? *preloaded : __ldcg(gp);
    cnt[0] = v.w; cnt[1] = v.z; cnt[2] = v.y; cnt[3] = v.x;
    if (zero) *gp = make_uint4(0u, 0u, 0u, 0u);
  }
  if (k <= 0) return;  // (CTA-uniform)
  {
    unsigned int mine = 0;
#pragma unroll
    for (int j = 0; j < kPer; ++j) mine += cnt[j];
    int tot;
    const unsigned int before = unsigned(mine_excl_scan(int(mine), S.warp_tot, &tot));
    if (before < (unsigned long long)k && (unsigned long long)k <= (unsigned long long)before + mine) {
      unsigned long long acc = before;
      bool found = false;
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const unsigned int h = cnt[j];
        if (!found && acc < (unsigned long long)k && (unsigned long long)k <= acc + h) {
          S.misc[0] = kBins - kPer * tid - 1 - j;
          S.misc[1] = int((unsigned long long)k - acc);
          S.misc[2] = int(h);
          found = true;
        }
        acc += h;
      }
    }
    __syncthreads();
  }
}

// One pass over vals[0, n): everything above bin d0 is added to `above` (per thread, double); the values of the bin
// are collected in S.vals in a fixed order when `fits` (at most kSelCap of them). Returns the number of values of
// the bin it met (CTA-uniform). Ends with a barrier.
__device__ int scan_collect(MineShared& S, const float* __restrict__ vals, const long long n, const int d0,
                            const bool fits, double& above) {
  const int tid = threadIdx.x;
  int filled = 0;  // (CTA-uniform) values placed so far
  {
    // 128-bit loads over the 16-byte aligned interior, the ragged ends by the first threads
    const int head = int((4u - unsigned((reinterpret_cast<uintptr_t>(vals) >> 2) & 3u)) & 3u);
    const long long nh = head < n ? head : n;
    const long long n4 = (n - nh) >> 2;
    const long long tail0 = nh + 4 * n4;
    const float4* v4 = reinterpret_cast<const float4*>(vals + nh);
    constexpr int kB = 6;  // 128-bit loads in flight per thread
    bool ends_done = false;
    for (long long i0 = 0; i0 < n4 || !ends_done; i0 += kB * kMineThreads) {
      float4 v[kB];
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        const long long i = i0 + u * kMineThreads + tid;
        v[u] = i < n4 ? __ldcg(v4 + i) : make_float4(-1.f, -1.f, -1.f, -1.f);
      }
      float ve = -1.f;  // the ragged ends ride along with the first round
      if (!ends_done) {
        if (tid < nh) ve = __ldcg(vals + tid);
        else if (tid >= 32 && tail0 + (tid - 32) < n) ve = __ldcg(vals + tail0 + (tid - 32));
      }
      ends_done = true;
      unsigned hits = 0;
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int b = e[c] >= 0.f ? sel_bin(e[c]) : -1;
          if (b > d0) above += double(e[c]);
          if (b == d0) hits |= 1u << (4 * u + c);
        }
      }
      {
        const int b = ve >= 0.f ? sel_bin(ve) : -1;
        if (b > d0) above += double(ve);
        if (b == d0) hits |= 1u << (4 * kB);
      }
      int tot;
      int at = filled + mine_excl_scan(__popc(hits), S.warp_tot, &tot);
      if (fits && hits) {
#pragma unroll
        for (int u = 0; u < kB; ++u) {
          const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if ((hits & (1u << (4 * u + c))) && at < kSelCap) S.vals[at++] = e[c];  // (at < n0 by construction)
        }
        if ((hits & (1u << (4 * kB))) && at < kSelCap) S.vals[at++] = ve;
      }
      filled += tot;
    }
  }
  __syncthreads();
  return filled;
}

struct BinResolve {
  float thr;       // the k-th largest value
  int take_ties;   // how many values equal to thr belong to the selection
  bool all_ties;   // ... all of them
  double inbin;    // this thread's part of the sum of the bin's values above thr
};

// The k-th value inside bin d0 (n0 values, rank `remaining` from the top): 8-bit radix passes over the collected
// values S.vals[0, n0) (fits) or over the bin's values in vals[0, n) (global memory).
__device__ BinResolve resolve_in_bin(MineShared& S, const float* __restrict__ vals, const long long n, const int d0,
                                     const int n0, const bool fits, int remaining) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool regular = d0 > 0 && d0 < kBins - 1;  // every value of a regular bin shares its leading 17 bits
  uint32_t prefix = regular ? uint32_t(d0 - 1 + kSelLo) << 15 : 0u;
  uint32_t mask = regular ? 0xffff8000u : 0u;
  int n_ties = n0;
  for (int shift = regular ? 7 : 24; shift >= 0; shift -= 8) {
    const uint32_t dmask = shift == 7 ? 0xffu : (regular ? 0x7fu : 0xffu);
    if (tid < 256) S.dig[tid] = 0u;
    __syncthreads();
    if (fits) {
      for (int i = tid; i < n0; i += kMineThreads) {
        const uint32_t bits = __float_as_uint(S.vals[i]);
        if ((bits & mask) == prefix) atomicAdd(&S.dig[(bits >> shift) & dmask], 1u);
      }
    } else {
      for (long long i = tid; i < n; i += kMineThreads) {
        const float v = __ldcg(vals + i);
        const uint32_t bits = __float_as_uint(v);
        if (v >= 0.f && sel_bin(v) == d0 && (bits & mask) == prefix) atomicAdd(&S.dig[(bits >> shift) & dmask], 1u);
      }
    }
    __syncthreads();
    if (wid == 0) {  // lane l owns the digits 255 - 8 l ... 248 - 8 l, highest first
      unsigned int c8[8];
      unsigned int mine = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        c8[j] = S.dig[255 - 8 * lane - j];
        mine += c8[j];
      }
      unsigned int inc = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      unsigned int acc = inc - mine;
      if (acc < unsigned(remaining) && unsigned(remaining) <= inc) {
        bool found = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (!found && acc < unsigned(remaining) && unsigned(remaining) <= acc + c8[j]) {
            S.misc[0] = 255 - 8 * lane - j;
            S.misc[1] = int(unsigned(remaining) - acc);
            S.misc[2] = int(c8[j]);
            found = true;
          }
          acc += c8[j];
        }
      }
    }
    __syncthreads();
    prefix |= uint32_t(S.misc[0]) << shift;
    mask |= dmask << shift;
    remaining = S.misc[1];
    n_ties = S.misc[2];
    __syncthreads();
    if (shift == 7) shift = 8;  // (the last pass of a regular bin covers the low 7 bits: next shift = 0)
  }
  const float thr = __uint_as_float(prefix);
  const int take_ties = remaining;
  const bool all_ties = take_ties >= n_ties;
  double inbin = 0.0;
  if (fits) {
    for (int i = tid; i < n0; i += kMineThreads)
      if (S.vals[i] > thr) inbin += double(S.vals[i]);
  } else {
    for (long long i = tid; i < n; i += kMineThreads) {
      const float v = __ldcg(vals + i);
      if (v > thr && sel_bin(v) == d0) inbin += double(v);
    }
  }
  BinResolve br;
  br.thr = thr;
  br.take_ties = take_ties;
  br.all_ties = all_ties;
  br.inbin = inbin;
  return br;
}

// rare: only some of the values equal to the threshold belong to the selection - the first ones by index get flag
// bit 1. Every thread walks a contiguous run of the values; runs are ranked by an exclusive scan.
__device__ void mark_partial_ties(MineShared& S, const float* __restrict__ vals, const long long n, const float thr,
                                  const int take_ties, uint8_t* __restrict__ sel) {
  const int tid = threadIdx.x;
  const long long per = (n + kMineThreads - 1) / kMineThreads;
  const long long lo = per * tid < n ? per * tid : n, hi = lo + per < n ? lo + per : n;
  int mine = 0;
  for (long long i = lo; i < hi; ++i) mine += __ldcg(vals + i) == thr ? 1 : 0;
  int tot;
  int left = take_ties - mine_excl_scan(mine, S.warp_tot, &tot);
  for (long long i = lo; i < hi && left > 0; ++i)
    if (__ldcg(vals + i) == thr) {
      sel[i] = uint8_t(sel[i] | 2);
      --left;
    }
}

__device__ TopkSum select_topk_sum(MineShared& S, const float* __restrict__ vals, const long long n,
                                   unsigned int* __restrict__ ghist, const long long k, uint8_t* __restrict__ sel,
                                   const uint4* preloaded = nullptr /* this thread's four bins, already read */) {
  static_assert(kMineWarps == 32, "mine_excl_scan");
  TopkSum r;
  r.sum = 0.0;
  r.thr = __int_as_float(0x7f800000);
  r.ties = 0.f;
  find_threshold_bin(S, ghist, k, true, preloaded);
  if (k <= 0) return r;  // (CTA-uniform)
  const int d0 = S.misc[0];
  const int remaining = S.misc[1];
  const int n0 = S.misc[2];
  const bool fits = n0 <= kSelCap;
  CM_STAMP_SET(8 + 4 * (blockIdx.x & 63) + 1);
  double above = 0.0;
  scan_collect(S, vals, n, d0, fits, above);
  CM_STAMP_SET(8 + 4 * (blockIdx.x & 63) + 2);
  const BinResolve br = resolve_in_bin(S, vals, n, d0, n0, fits, remaining);
  r.sum = block_sum(above + br.inbin, S.red) + double(br.take_ties) * double(br.thr);
  r.thr = br.thr;
  r.ties = br.all_ties ? 1.f : 0.f;
  if (!br.all_ties && sel) mark_partial_ties(S, vals, n, br.thr, br.take_ties, sel);
  return r;
}

// ------------------------------------------------------------------------------------------
// classify_kernel: grid (slices per image, N), 256 threads, each CTA owns 2048 consecutive priors of an image
// (three CTAs per SM: the whole grid is resident at the SSD512 shape). Rebuilds the image's short forced-match
// list from the per-object keys and applies it to its slice (rank inside the FILTERED list, last write wins -
// SSD512.py:546-553); classes, selection bits, mining candidates (8 priors per thread, all loads in flight
// together) and their histogram; the rows with a foreground class are recorded and evaluated densely
// (true-class CE, loc term); partial sums of the slice to global memory. Programmatic dependent launch behind
// the match kernel, and it releases mine_kernel's launch at once.
// ------------------------------------------------------------------------------------------
constexpr int kCmThreads = 256;
constexpr int kCmWarps = kCmThreads / 32;
constexpr int kCmUnroll = 8;                      // priors per thread whose loads are in flight together
constexpr int kCmSlice = kCmThreads * kCmUnroll;  // priors per CTA

struct CmShared {
  unsigned int hist[kBins];   // the slice's candidate histogram
  double red4[4 * 32];
  int warp_tot[kCmWarps];
  int carry;
  int fg_n;
  uint16_t cnt[kCmUnroll][kCmWarps];
  uint16_t fg_idx[kCmSlice];  // foreground rows: prior index inside the slice
  uint16_t fg_cls[kCmSlice];  // ... class (bit 15: positive)
  uint16_t fg_obj[kCmSlice];  // ... object of the image
  int patch[kCmSlice];        // forced matches of the slice: 1 + rank of the object that claims the prior, 0 = none
};

__global__ void __launch_bounds__(kCmThreads, 3) classify_kernel(const LossParams q) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  CmShared& S = *reinterpret_cast<CmShared*>(smem_raw);
  int32_t* s_label = reinterpret_cast<int32_t*>(smem_raw + ((sizeof(CmShared) + 127) & ~size_t(127)));  // [gmax]

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // mine_kernel may take its place early
  const int n = blockIdx.y, slices = gridDim.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int p_lo = min(q.P, int(blockIdx.x) * kCmSlice), p_hi = min(q.P, p_lo + kCmSlice);
  const size_t base = size_t(n) * q.P;
  const bool focal = q.cls_kind == SBOD_CLS_FOCAL_SUM || q.cls_kind == SBOD_CLS_FOCAL_NORM;
  const int g0 = q.gt_offsets[n];  // (inputs of the call, not results of the match kernel)
  const int G = q.gt_offsets[n + 1] - g0;
  CM_STAMP_MIN(0);
  if (tid == 0) S.carry = 0;
  for (int g = tid; g < G; g += kCmThreads) s_label[g] = int(map_label(q, q.gt_labels[g0 + g]));
  for (int i = tid; i < kCmSlice; i += kCmThreads) S.patch[i] = 0;
  if (!focal)
    for (int i = tid; i < kBins; i += kCmThreads) S.hist[i] = 0u;
  uint8_t ex[kCmUnroll];
#pragma unroll
  for (int u = 0; u < kCmUnroll; ++u) {
    const int p = p_lo + u * kCmThreads + tid;
    ex[u] = (q.exclude && p < p_hi) ? q.exclude[base + p] : uint8_t(0);
  }
  // the image's candidate histogram (SSD300 mines the batch: one histogram for all images)
  unsigned int* gh = q.sel_hist + (q.cls_kind == SBOD_CLS_CE_MINE_BATCH ? size_t(0) : size_t(n) * kBins);

  // everything above is independent of the match kernel's results; from here on they are needed
  asm volatile("griddepcontrol.wait;" ::: "memory");
  CM_STAMP_MIN(1);
  // ---- every load of the slice is requested at once: the first per-object keys, then the per-prior state of
  // 8 priors per thread ----
  unsigned long long key0 = tid < G ? q.gtkey[size_t(n) * q.gmax + tid] : 0ull;
  float ov[kCmUnroll], ce[kCmUnroll];
  int ob[kCmUnroll];
#pragma unroll
  for (int u = 0; u < kCmUnroll; ++u) {
    const int p = p_lo + u * kCmThreads + tid;
    ov[u] = 0.f; ce[u] = 0.f; ob[u] = 0;
    if (p < p_hi) {
      ov[u] = q.ov[base + p];
      ob[u] = q.obj[base + p];
      ce[u] = q.ce[base + p];  // background CE from the streaming kernel
    }
  }
  __syncthreads();
  // ---- forced list: rank j of each object inside "objects whose best overlap is > 0"; the entries that fall
  // into this slice go into a table indexed by prior ("for j: obj[pr[j]] = j": the last j wins,
  // SSD512.py:552-553 - ranks grow with the object index, so the largest rank wins) ----
  for (int gb = 0; gb < G; gb += kCmThreads) {
    const int g = gb + tid;
    const unsigned long long key = gb == 0 ? key0 : (g < G ? q.gtkey[size_t(n) * q.gmax + g] : 0ull);
    const bool f = key != 0ull;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) S.warp_tot[wid] = __popc(bal);
    __syncthreads();
    int off = S.carry, tot = 0;
#pragma unroll
    for (int w = 0; w < kCmWarps; ++w) {
      const int t = S.warp_tot[w];
      if (w < wid) off += t;
      tot += t;
    }
    if (f) {
      const uint32_t p = 0xffffffffu - uint32_t(key & 0xffffffffull);
      const int j = off + __popc(bal & ((1u << lane) - 1u));
      if (p >= uint32_t(p_lo) && p < uint32_t(p_hi)) atomicMax(&S.patch[p - p_lo], j + 1);
    }
    __syncthreads();
    if (tid == 0) S.carry += tot;
  }
  __syncthreads();
  CM_STAMP_MAX(2);

  // ---- phase A: classes, selection bits, mining candidates ----
  double a_loc = 0.0, a_pos = 0.0, a_neg = 0.0;
  int npos = 0;
  int rec[kCmUnroll], rrank[kCmUnroll];
#pragma unroll
  for (int u = 0; u < kCmUnroll; ++u) {
    const int p = p_lo + u * kCmThreads + tid;
    rec[u] = 0;
    int bin = -1;
    if (p < p_hi) {
      const int forced = S.patch[u * kCmThreads + tid];
      if (forced) {  // index_fill_(0, prior_for_each_object, 1.0) and the object override
        ov[u] = 1.0f;
        ob[u] = forced - 1;
        q.ov[base + p] = 1.0f;
        q.obj[base + p] = forced - 1;
      }
      int cls = 0;
      if (!(ov[u] < q.thr_pos) && G > 0) cls = s_label[ob[u]];
      const bool pos = cls > 0 && !ex[u];
      const bool isneg = ov[u] < q.thr_neg;
      uint8_t selbits = pos ? 1 : 0;
      float v = -1.f;
      if (cls > 0) {
        rec[u] = min(cls, q.C - 1) | (pos ? 0x8000 : 0);  // C <= 16384 elsewhere; classes fit 15 bits
      } else if (focal) {
        if (isneg) {  // target class is 0 there (thr_neg < thr_pos)
          selbits |= 2;
          a_neg += double(focal_bg(q, ce[u]));
        }
      } else if (q.cls_kind == SBOD_CLS_CE_MINE_NONPOS) {
        if (!ex[u]) v = ce[u];
      } else {  // MINE_NEG, MINE_BATCH: only true_neg == -1 rows are candidates
        if (isneg) v = ce[u];
      }
      if (!focal && v >= 0.f) {
        a_neg += 1.0;  // mining modes: this partial counts the candidates (NaN CEs are none)
        bin = sel_bin(v);
        selbits |= 4;  // "mining candidate": the backward compares its CE with the threshold
      }
      q.sel[base + p] = selbits;
      if (!focal) q.cand[base + p] = v;
    }
    if (!focal) {
      // bin 0 holds everything below 2^-10 - most candidates of a trained model: one shared-memory atomic per warp
      const unsigned low = __ballot_sync(0xffffffffu, bin == 0);
      if (bin > 0) atomicAdd(&S.hist[bin], 1u);
      else if (bin == 0 && lane == __ffs(low) - 1) atomicAdd(&S.hist[0], unsigned(__popc(low)));
    }
    const unsigned bal = __ballot_sync(0xffffffffu, rec[u] != 0);
    rrank[u] = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) S.cnt[u][wid] = uint16_t(__popc(bal));
  }
  __syncthreads();
  // list offsets of the (unroll slot, warp) pairs: exclusive scan of the 64 counts by warp 0; the list order
  // is deterministic (unroll slot, warp, ballot rank), so the sums are reproducible
  if (wid == 0) {
    static_assert(kCmUnroll * kCmWarps == 64, "two (slot, warp) pairs per lane");
    uint16_t* flat = &S.cnt[0][0];
    const int c0 = flat[2 * lane], c1 = flat[2 * lane + 1];
    int inc = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    flat[2 * lane] = uint16_t(inc - c0 - c1);
    flat[2 * lane + 1] = uint16_t(inc - c1);
    if (lane == 31) S.fg_n = inc;
  }
  // the slice's histogram joins the image's: one L2 reduction per non-empty bin
  if (!focal) {
#pragma unroll 4
    for (int i = tid; i < kBins; i += kCmThreads) {
      const unsigned int c = S.hist[i];
      if (c) atomicAdd(&gh[i], c);
    }
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < kCmUnroll; ++u) {
    if (rec[u]) {
      const int slot = int(S.cnt[u][wid]) + rrank[u];
      S.fg_idx[slot] = uint16_t(u * kCmThreads + tid);
      S.fg_cls[slot] = uint16_t(rec[u]);
      S.fg_obj[slot] = uint16_t(ob[u]);
    }
  }
  __syncthreads();
  CM_STAMP_MAX(3);
  // ---- phase B: the recorded foreground rows (2-3 % of the priors), one per thread, every load of a row
  // independent of the others ----
  {
    const int total = S.fg_n;
    for (int i = tid; i < total; i += kCmThreads) {
      const int pp = p_lo + int(S.fg_idx[i]);
      const int c = S.fg_cls[i] & 0x7fff;
      const bool pos = (S.fg_cls[i] & 0x8000) != 0;
      const float xc = ld_stream_f32(q.scores + (base + pp) * q.C + c);
      const float lse = q.lse[base + pp];  // (only the foreground rows need it: L2, in flight with the logit)
      float4 pred = make_float4(0.f, 0.f, 0.f, 0.f), pcx = pred, gbox = pred;
      if (pos) {
        pred = reinterpret_cast<const float4*>(q.locs)[base + pp];
        pcx = prior_cxcy_of(q, n, pp);
        gbox = q.gt_boxes[g0 + S.fg_obj[i]];
      }
      const float cet = lse - xc;  // CE against the true class
      q.ce[base + pp] = cet;
      if (pos) {
        ++npos;
        a_pos += focal ? double(focal_fg(q, cet)) : double(cet);
        const LocTerm lt = loc_term<false>(q, pred, pcx, gbox);
        a_loc += double(lt.loss);
      }
    }
  }
  CM_STAMP_MAX(4);
  // ---- partial sums of the slice ----
  double tot[4] = {a_loc, a_pos, a_neg, double(npos)};
  block_sum4_to_thread0(tot, S.red4);
  if (tid == 0) {
    double* bp = q.blockpart + (size_t(n) * slices + blockIdx.x) * 4;
    bp[0] = tot[0]; bp[1] = tot[1]; bp[2] = tot[2]; bp[3] = tot[3];
  }
  CM_STAMP_MAX(5);
}

// ------------------------------------------------------------------------------------------
// mine_kernel: one CTA of 1024 threads per image, programmatic dependent launch behind classify_kernel.
// n_pos and the other sums of the image from the slices' partials (fixed order), sum of its 3 * n_pos largest
// candidate CEs (select_topk_sum), the threshold for the backward; then a ticket of the batch: the last image
// folds the batch in image order (deterministic), mines the batch when the criterion does that
// (SSD300.py:580-588), exchanges the sums with the other ranks if the batch is sharded, and finalises the loss.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMineThreads, 1) mine_kernel(const LossParams q, const int slices) {
  __shared__ MineShared S;
  const int n = blockIdx.x, tid = threadIdx.x;
  const size_t base = size_t(n) * q.P;
  const bool focal = q.cls_kind == SBOD_CLS_FOCAL_SUM || q.cls_kind == SBOD_CLS_FOCAL_NORM;
  const int G = q.gt_offsets[n + 1] - q.gt_offsets[n];
  asm volatile("griddepcontrol.wait;" ::: "memory");
  CM_STAMP_SET(8 + 4 * (n & 63));
  const bool select = !focal && q.cls_kind != SBOD_CLS_CE_MINE_BATCH;
  // this thread's four histogram bins are requested before the partial sums (one round trip less in the chain)
  uint4 hbins = make_uint4(0u, 0u, 0u, 0u);
  if (select) hbins = __ldcg(reinterpret_cast<const uint4*>(q.sel_hist + size_t(n) * kBins + (kBins - 4 * (tid + 1))));
  // leave the per-object keys clean for the next call (every slice has read them)
  for (int g = tid; g < G; g += kMineThreads) q.gtkey[size_t(n) * q.gmax + g] = 0ull;
  if (tid < 128) {  // warp w adds component w of the slices' partials: lanes stride, then a fixed shuffle tree
    const int comp = tid >> 5, l = tid & 31;
    const double* bp = q.blockpart + size_t(n) * slices * 4 + comp;
    double acc = 0.0;
    for (int b = l; b < slices; b += 32) acc += __ldcg(bp + b * 4);
    acc = warp_sum(acc);  // (deterministic: the same tree every run)
    if (l == 0) S.red4[comp] = acc;
  }
  __syncthreads();
  const double img[4] = {S.red4[0], S.red4[1], S.red4[2], S.red4[3]};
  __syncthreads();
  double t_neg = img[2];  // focal: sum over the negatives; MINE_BATCH: number of candidates (batch tail)
  // what the backward needs to know about the image's mined negatives: a candidate row (flag bit 2) is mined
  // iff its CE > thr, or == thr when every tie is taken; +inf = none
  float sel_thr = __int_as_float(0x7f800000), sel_ties = 0.f;
  if (select) {
    const long long n_cand = (long long)(img[2] + 0.5);
    long long k = (long long)(q.ratio) * (long long)(img[3] + 0.5);
    if (k > n_cand) k = n_cand;
    const TopkSum r = select_topk_sum(S, q.cand + base, q.P, q.sel_hist + size_t(n) * kBins, k, q.sel + base, &hbins);
    t_neg = r.sum;
    sel_thr = r.thr;
    sel_ties = r.ties;
  }
  const bool batch = q.cls_kind == SBOD_CLS_CE_MINE_BATCH;
  if (batch) {
    // SSD300 mines over the whole batch (SSD300.py:580-588). Every image's CTA can tell the batch's threshold bin
    // on its own: the slices' partials and the (single) histogram are complete. It then does its image's share
    // of the one pass - the sum above the bin, the bin's values - so the last image only merges 32 short lists
    // instead of scanning N * P candidates alone.
    double c2 = 0.0, c3 = 0.0;  // candidates / positives of the batch: integers, exact in any order
    for (int e = tid; e < q.N * slices; e += kMineThreads) {
      c2 += __ldcg(q.blockpart + size_t(e) * 4 + 2);
      c3 += __ldcg(q.blockpart + size_t(e) * 4 + 3);
    }
    c2 = block_sum(c2, S.red);
    c3 = block_sum(c3, S.red);
    const long long n_cand_b = (long long)(c2 + 0.5);
    long long k_b = (long long)(q.ratio) * (long long)(c3 + 0.5);
    if (k_b > n_cand_b) k_b = n_cand_b;
    int segn = 0;
    double above_n = 0.0;
    if (k_b > 0 && q.N <= kMineThreads) {
      find_threshold_bin(S, q.sel_hist, k_b, false, nullptr);
      const int d0 = S.misc[0], n0 = S.misc[2];
      __syncthreads();
      if (n0 <= kSelCap) {
        double above = 0.0;
        segn = scan_collect(S, q.cand + base, q.P, d0, true, above);
        above_n = block_sum(above, S.red);
        for (int i = tid; i < segn && i < kSelCap; i += kMineThreads) q.sel_seg[size_t(n) * kSelCap + i] = S.vals[i];
      } else {
        segn = -1;
      }
    } else if (k_b > 0) {
      segn = -1;
    }
    if (tid == 0) {
      q.sel_segn[n] = segn;
      q.sel_above[n] = above_n;
    }
    __threadfence();  // the segment before this image's ticket
    __syncthreads();
  }
  if (tid == 0) {
    q.sel_thr[2 * n] = sel_thr;
    q.sel_thr[2 * n + 1] = sel_ties;
    q.partials[n * 4 + 0] = img[0];
    q.partials[n * 4 + 1] = img[1];
    q.partials[n * 4 + 2] = t_neg;
    q.partials[n * 4 + 3] = img[3];
    __threadfence();
    const unsigned int ticket = atomicAdd(&q.counters[0], 1u);
    S.last = (ticket == unsigned(q.N) - 1u) ? 1 : 0;
  }
  __syncthreads();
  CM_STAMP_SET(8 + 4 * (n & 63) + 3);
  if (!S.last) return;
  // ---- the last image: fold the batch in image order (deterministic) ----
  __threadfence();
  if (tid < 4) {
    double acc = 0.0;
#pragma unroll 8
    for (int i = 0; i < q.N; ++i) acc += __ldcg(q.partials + i * 4 + tid);
    S.red4[tid] = acc;
  }
  __syncthreads();
  double tot[4] = {S.red4[0], S.red4[1], S.red4[2], S.red4[3]};
  __syncthreads();
  if (q.cls_kind == SBOD_CLS_CE_MINE_BATCH) {
    // SSD300: the 3 * n_pos largest CEs over every candidate of the BATCH; one threshold for all images
    const long long n_cand = (long long)(tot[2] + 0.5);
    long long k = (long long)(q.ratio) * (long long)(tot[3] + 0.5);
    if (k > n_cand) k = n_cand;
    TopkSum r;
    r.sum = 0.0;
    r.thr = __int_as_float(0x7f800000);
    r.ties = 0.f;
    bool merged = false;
    if (k > 0 && q.N <= kMineThreads) {
      find_threshold_bin(S, q.sel_hist, k, false, nullptr);  // (the same bin every image found)
      const int d0 = S.misc[0], remaining = S.misc[1], n0 = S.misc[2];
      __syncthreads();
      if (n0 <= kSelCap) {
        // merge the images' lists in image order (deterministic), then resolve the k-th value inside the bin
        merged = true;
        int tot_n;
        const int c = tid < q.N ? __ldcg(q.sel_segn + tid) : 0;
        const int off = mine_excl_scan(c, S.warp_tot, &tot_n);
        if (tid < q.N) S.segoff[tid] = off;
        if (tid == 0) S.segoff[q.N] = tot_n;
        if (tid == 0) {
          double a = 0.0;
          for (int i = 0; i < q.N; ++i) a += __ldcg(q.sel_above + i);
          S.red4[0] = a;
        }
        __syncthreads();
        for (int i = 0; i < q.N; ++i) {
          const int o = S.segoff[i], cn = S.segoff[i + 1] - o;
          for (int j = tid; j < cn; j += kMineThreads) S.vals[o + j] = __ldcg(q.sel_seg + size_t(i) * kSelCap + j);
        }
        __syncthreads();
        const double above = S.red4[0];
        const BinResolve br = resolve_in_bin(S, nullptr, 0, d0, n0, true, remaining);
        r.sum = above + block_sum(br.inbin, S.red) + double(br.take_ties) * double(br.thr);
        r.thr = br.thr;
        r.ties = br.all_ties ? 1.f : 0.f;
        if (!br.all_ties) mark_partial_ties(S, q.cand, (long long)q.N * q.P, br.thr, br.take_ties, q.sel);
        for (int b = tid; b < kBins; b += kMineThreads) q.sel_hist[b] = 0u;  // clean for the next call
      }
    }
    if (!merged) r = select_topk_sum(S, q.cand, (long long)q.N * q.P, q.sel_hist, k, q.sel);  // the whole batch here
    tot[2] = r.sum;
    for (int i = tid; i < q.N; i += kMineThreads) {
      q.sel_thr[2 * i] = r.thr;
      q.sel_thr[2 * i + 1] = r.ties;
    }
  }
  if (tid < 32 && q.comm) {
    // sharded batch: the four sums of this rank meet those of the other ranks through the NVLink mailboxes
    // (every rank adds them in rank order); the loss is then formed from the global sums, and the backward
    // scales by the global 1 / n_pos
    comm_allreduce_sum(q.comm, tot, 4);
  }
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) q.sums[i] = tot[i];
    q.counters[0] = 0u;
    finalize_loss(q, tot, q.loss);
  }
  CM_STAMP_MAX(7);
}

__global__ void finalize_kernel(const LossParams q) {
  if (threadIdx.x == 0 && blockIdx.x == 0) finalize_loss(q, q.sums, q.loss);
}

// ------------------------------------------------------------------------------------------
// loss_bwd_kernel: one CTA per tile of rows_per_tile priors.
//   sparse mode (CE + mining): rows are zero except positives / mined negatives
//   dense mode (focal): most rows carry gradient -> tile loaded with TMA, transformed in place
// ------------------------------------------------------------------------------------------
struct BwdParams {
  LossParams q;
  const float* grad_loss;
  float* grad_locs;
  float* grad_scores;
  int dense;
};

template <int kC>  // kC > 0 (odd): compile-time class count - the row transform is unrolled with immediate offsets
__global__ void __launch_bounds__(kRows) loss_bwd_kernel(const BwdParams bp) {
  const LossParams& q = bp.q;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tilebuf = reinterpret_cast<float*>(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + size_t(q.stage_floats) * 4);
  __shared__ int s_cls[kRows];
  __shared__ uint8_t s_sel[kRows];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const double npos_tot = q.sums[3];
  const float gout = bp.grad_loss ? *bp.grad_loss : 1.f;
  const bool focal = q.cls_kind == SBOD_CLS_FOCAL_SUM || q.cls_kind == SBOD_CLS_FOCAL_NORM;
  const float conf_scale =
      (q.cls_kind == SBOD_CLS_FOCAL_SUM) ? gout : float(double(gout) / npos_tot);
  const float loc_scale = float(double(gout) * double(q.reg_weight) /
                                (q.reg_kind == SBOD_REG_L1_ELEM_MEAN ? 4.0 * npos_tot : npos_tot));
  if (bp.dense && tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  int phase = 0;
  for (int tile = blockIdx.x; tile < q.n_tiles; tile += gridDim.x) {
    const TileCoord tc = tile_coord(q, tile);
    const int n = tc.n;
    const int g0 = q.gt_offsets[n];
    const int G = q.gt_offsets[n + 1] - g0;
    const bool valid = tid < tc.rows;
    const int p = tc.p0 + tid;
    const size_t np = size_t(n) * q.P + (valid ? p : tc.p0);
    const size_t first = (size_t(n) * q.P + tc.p0) * size_t(q.C);
    const size_t n_elem = size_t(tc.rows) * q.C;
    const size_t total = size_t(q.N) * q.P * size_t(q.C);
    const uint32_t head = uint32_t(first & 3);

    // the previous tile's bulk store must have finished reading smem before we overwrite it
    if (tid == 0) tma_store_wait_read<0>();
    __syncthreads();
    if (bp.dense && tid == 0) {
      const TileSpan s = make_tile_span(q.scores, first, n_elem, total);
      for (uint32_t i = 0; i < s.tail_floats; ++i)
        tilebuf[s.bulk_bytes / 4 + i] = q.scores[(s.src16 - q.scores) + s.bulk_bytes / 4 + i];
      if (s.bulk_bytes) {
        mbar_arrive_expect_tx(bar, s.bulk_bytes);
        tma_load_1d(tilebuf, s.src16, s.bulk_bytes, bar);  // (no L2 hint here: evict-first measured 350 vs 285 us)
      } else {
        mbar_arrive(bar);
      }
    }

    // per-row bookkeeping + grad wrt locs
    uint8_t selbits = 0;
    int cls = 0;
    if (valid) {
      selbits = q.sel[np];
      const float ov = q.ov[np];
      const int obj = q.obj[np];
      int64_t lab = 0;
      if (G > 0) lab = map_label(q, q.gt_labels[g0 + obj]);
      cls = (ov < q.thr_pos) ? 0 : int(lab);
      cls = min(max(cls, 0), q.C - 1);
      float4 gl = make_float4(0.f, 0.f, 0.f, 0.f);
      if (selbits & 1) {
        const float4 pred = reinterpret_cast<const float4*>(q.locs)[np];
        const LocTerm lt = loc_term<true>(q, pred, prior_cxcy_of(q, n, p), q.gt_boxes[g0 + obj]);
        gl = make_float4(lt.grad.x * loc_scale, lt.grad.y * loc_scale, lt.grad.z * loc_scale,
                         lt.grad.w * loc_scale);
      }
      if (bp.grad_locs) reinterpret_cast<float4*>(bp.grad_locs)[np] = gl;
    }
    s_sel[tid] = valid ? selbits : 0;
    s_cls[tid] = cls;

    if (!bp.grad_scores) continue;

    if (!bp.dense) {
      // zero the tile, then patch the selected rows (a warp per row, coalesced over classes)
      float4* t4 = reinterpret_cast<float4*>(tilebuf);
      const int n4 = int((head + n_elem + 3) / 4);
      for (int i = tid; i < n4; i += kRows) t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      __syncthreads();
      for (int r = wid; r < tc.rows; r += kRows / 32) {
        if (!s_sel[r]) continue;
        const size_t rnp = size_t(n) * q.P + tc.p0 + r;
        const float lse = q.lse[rnp];
        const int rc = s_cls[r];
        const float* x = q.scores + rnp * q.C;
        float* o = tilebuf + head + size_t(r) * q.C;
        for (int k = lane; k < q.C; k += 32)
          o[k] = conf_scale * (__expf(x[k] - lse) - (k == rc ? 1.f : 0.f));
      }
    } else {
      mbar_wait(bar, phase);
      phase ^= 1;
      if (valid) {
        float* row = tilebuf + head + size_t(tid) * q.C;
        const float lse = q.lse[np];
        int gcd = 1;
        while (gcd < 32 && (q.C % (gcd * 2)) == 0) gcd *= 2;
        const int rot = (lane * gcd) >> 5;
        float coef_soft = 0.f, coef_hot = 0.f;  // grad = coef_soft*softmax - coef_hot*onehot
        // The row's cross entropy against its target class, as the forward left it: for the background rows it is
        // log(1 + sum of the other terms) at full relative precision, so the target element softmax_t - 1 is formed
        // as expm1(-ce) instead of exp(x_t - lse) - 1 (lse is an fp32 number of the magnitude of the logits: the
        // difference of two nearly equal probabilities would keep 3 digits at |x| ~ 16).
        const float ce_t = selbits ? q.ce[np] : 0.f;
        if (selbits) {
          if (focal) {
            // L = A * w(pt)^g * ce ; ce = -log pt ; d ce/dx_k = softmax_k - onehot_k ; d pt/dx_k = -pt*(..)
            const float pt = __expf(-ce_t);
            const float ce = ce_t;
            float dL_dce;
            if (selbits & 1) {  // foreground: A=alpha, w = 1-pt  -> dw/dce = pt
              const float w = 1.f - pt;
              dL_dce = q.falpha * (pow_gamma(w, q.fgamma) +
                                   ce * q.fgamma * pow_gamma(w, q.fgamma - 1.f) * pt);
            } else {  // background: A=1-alpha, w = pt -> dw/dce = -pt
              dL_dce = (1.f - q.falpha) * (pow_gamma(pt, q.fgamma) -
                                           ce * q.fgamma * pow_gamma(pt, q.fgamma - 1.f) * pt);
            }
            coef_soft = coef_hot = conf_scale * dL_dce;
          } else {
            coef_soft = coef_hot = conf_scale;
          }
        }
        if (kC) {
          // odd class count: a warp's 32 rows start in 32 different banks, no rotation needed; five
          // instructions per logit (LDS, FFMA, EX2, FMUL, STS), the one-hot term patched in afterwards
          if (coef_soft == 0.f) {
#pragma unroll
            for (int k = 0; k < kC; ++k) row[k] = 0.f;
          } else {
            const float nl2 = -lse * kLog2e;
#pragma unroll
            for (int k = 0; k < kC; ++k) row[k] = coef_soft * ex2_approx(fmaf(row[k], kLog2e, nl2));
            row[cls] = coef_hot * expm1_neg(ce_t);
          }
        } else {
          for (int kk = 0; kk < q.C; ++kk) {
            int k = kk + rot;
            if (k >= q.C) k -= q.C;
            const float sm = __expf(row[k] - lse);
            row[k] = k == cls ? coef_hot * expm1_neg(ce_t) : coef_soft * sm;
          }
        }
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      // aligned interior by bulk store; ragged head/tail by plain stores
      const size_t e0 = first, e1 = first + n_elem;
      const size_t a0 = (e0 + 3) & ~size_t(3), a1 = e1 & ~size_t(3);
      const size_t s0 = e0 & ~size_t(3);  // smem index 0 <-> global element s0
      if (a1 > a0) {
        tma_store_1d(bp.grad_scores + a0, tilebuf + (a0 - s0), uint32_t((a1 - a0) * 4));
        tma_store_commit();
      }
      for (size_t e = e0; e < (a0 < e1 ? a0 : e1); ++e) bp.grad_scores[e] = tilebuf[e - s0];
      for (size_t e = (a1 > a0 ? a1 : (a0 < e1 ? a0 : e1)); e < e1; ++e)
        bp.grad_scores[e] = tilebuf[e - s0];
    }
  }
  if (tid == 0) tma_store_wait_all<0>();
}

// ------------------------------------------------------------------------------------------
// Sparse backward (CE + hard-negative mining): only positives and mined negatives carry gradient
// (about 4*n_pos rows of N*P), so grad_scores is written as
//   zero_fill_kernel  : the whole tensor from one zeroed shared-memory tile with back-to-back bulk
//                       TMA stores (the source never changes, so nothing waits until the end), then
//   bwd_patch_kernel  : the selected rows (a warp per row, coalesced over classes) and grad_locs.
// ------------------------------------------------------------------------------------------
constexpr int kZeroTileBytes = 32 * 1024;

__global__ void __launch_bounds__(128) zero_fill_kernel(float* __restrict__ dst, size_t n_floats) {
  __shared__ __align__(128) unsigned char s_zero[kZeroTileBytes];
  const int tid = threadIdx.x;
  float4* z4 = reinterpret_cast<float4*>(s_zero);
  for (int i = tid; i < kZeroTileBytes / 16; i += blockDim.x) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async();
  __syncthreads();
  // dst is 16-byte aligned; the bulk part covers floor16(bytes), the tail (< 16 B) is plain stores
  const size_t total_bytes = n_floats * 4;
  const size_t bulk_bytes = total_bytes & ~size_t(15);
  const size_t n_chunks = (bulk_bytes + kZeroTileBytes - 1) / kZeroTileBytes;
  if (tid == 0) {
    unsigned char* base = reinterpret_cast<unsigned char*>(dst);
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
      const size_t off = c * kZeroTileBytes;
      const size_t len = (bulk_bytes - off) < size_t(kZeroTileBytes) ? (bulk_bytes - off) : size_t(kZeroTileBytes);
      tma_store_1d_hint(base + off, s_zero, uint32_t(len), l2_policy_evict_first());
      tma_store_commit();
    }
    tma_store_wait_all<0>();
  }
  if (blockIdx.x == 0 && tid < int((total_bytes - bulk_bytes) / 4)) dst[bulk_bytes / 4 + tid] = 0.f;
}

struct BwdParams;
template <int kC>  // kC > 0: compile-time class count (row offsets and chunk tests become immediates)
__global__ void __launch_bounds__(256, 4) bwd_patch_kernel(const LossParams q, const float* __restrict__ grad_loss,
                                                        float* __restrict__ grad_locs,
                                                        float* __restrict__ grad_scores) {
  const int lane = threadIdx.x & 31;
  const int C = kC ? kC : q.C;
  const size_t total = size_t(q.N) * q.P;
  asm volatile("griddepcontrol.wait;" ::: "memory");  // (programmatic dependent launch behind the forward)
  const double npos_tot = q.sums[3];
  const float gout = grad_loss ? *grad_loss : 1.f;
  const float conf_scale = float(double(gout) / npos_tot);
  const float loc_scale = float(double(gout) * double(q.reg_weight) /
                                (q.reg_kind == SBOD_REG_L1_ELEM_MEAN ? 4.0 * npos_tot : npos_tot));
  const size_t warp0 = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = (size_t(gridDim.x) * blockDim.x) >> 5;
  // the selection flags, log-sum-exps and background CEs of the next group are requested one iteration ahead
  uint8_t sel_next = 0;
  float lse_next = 0.f, ce_next = 0.f;
  if (warp0 * 32 + lane < total) {
    sel_next = q.sel[warp0 * 32 + lane];
    lse_next = q.lse[warp0 * 32 + lane];
    ce_next = q.ce[warp0 * 32 + lane];
  }
  for (size_t base = warp0 * 32; base < total; base += n_warps * 32) {
    const size_t i = base + lane;
    const bool in = i < total;
    uint8_t selbits = sel_next;
    const float lse = lse_next;
    const float ce_bg = ce_next;
    {
      const size_t inext = i + n_warps * 32;
      sel_next = 0;
      if (inext < total) {
        sel_next = q.sel[inext];
        lse_next = q.lse[inext];
        ce_next = q.ce[inext];
      }
    }
    // flag bit 2: mining candidate - mined iff its (background) CE beats the image's threshold
    // (classify_kernel / mine_kernel leaves the threshold instead of marking ~3 n_pos rows per image)
    if (in && (selbits & 4)) {
      // (the image of the group's first row with one division per warp; a group rarely straddles two images)
      const int n0 = int(base / q.P);
      const int n = q.P < 32 ? int(i / q.P) : (i < size_t(n0 + 1) * q.P ? n0 : n0 + 1);
      const float thr = q.sel_thr[2 * n], ties = q.sel_thr[2 * n + 1];
      const bool mined = (selbits & 2) || ce_bg > thr || (ce_bg == thr && ties != 0.f);
      selbits = mined ? 2 : 0;
    }
    // a mined negative has target class 0; only positives need their object's label and box
    int cls = 0;
    if (in && (selbits & 1)) {
      const int n = int(i / q.P);
      const int p = int(i - size_t(n) * q.P);
      const int g0 = q.gt_offsets[n];
      const int obj = q.obj[i];
      cls = int(map_label(q, q.gt_labels[g0 + obj]));
      cls = min(max(cls, 0), C - 1);
      if (grad_locs) {
        const float4 pred = reinterpret_cast<const float4*>(q.locs)[i];
        const LocTerm lt = loc_term<true>(q, pred, prior_cxcy_of(q, n, p), q.gt_boxes[g0 + obj]);
        reinterpret_cast<float4*>(grad_locs)[i] = make_float4(lt.grad.x * loc_scale, lt.grad.y * loc_scale,
                                                             lt.grad.z * loc_scale, lt.grad.w * loc_scale);
      }
    } else if (in && grad_locs) {
      reinterpret_cast<float4*>(grad_locs)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (!grad_scores) continue;
    // The target element of a selected row, scale * (softmax_t - 1), from the row's cross entropy against its target
    // (background CE of a mined negative, true-class CE of a positive: both in `ce`) as expm1(-ce): exp(x_t - lse) - 1
    // would difference two nearly equal numbers when the row is confident.
    const float g_target = selbits ? conf_scale * expm1_neg(ce_bg) : 0.f;
    unsigned m = __ballot_sync(0xffffffffu, selbits != 0);
    // four selected rows per round, every 32-class chunk of them loaded before anything is stored
    // (loads and stores may alias as far as the compiler knows): up to 16 loads in flight per lane
    while (m) {
      int src[4];
      int rc[4];
      float rl[4], rt[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        src[u] = -1;
        if (m) {
          src[u] = __ffs(m) - 1;
          m &= m - 1;
        }
        const int sl = src[u] < 0 ? 0 : src[u];
        rc[u] = __shfl_sync(0xffffffffu, cls, sl);
        rl[u] = __shfl_sync(0xffffffffu, lse, sl);
        rt[u] = __shfl_sync(0xffffffffu, g_target, sl);
      }
      // the lane's slice of each row: one 64-bit base per row, then immediate offsets (32 c floats per chunk)
      const float* rp[4];
      float* gp[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t off = (base + size_t(src[u] < 0 ? 0 : src[u])) * size_t(C) + lane;
        rp[u] = q.scores + off;
        gp[u] = grad_scores + off;
      }
      if (C <= 4 * 32) {
        constexpr int kChunks = kC ? (kC + 31) / 32 : 4;
        float x[kChunks][4];
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          // chunk c exists (warp-uniform) and the lane's class 32 c + lane is inside the row
          const bool has = kC ? (32 * c + 31 < kC || lane < kC - 32 * c) : (32 * c + lane < C);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            x[c][u] = 0.f;
            if (has && src[u] >= 0) x[c][u] = ld_stream_f32(rp[u] + 32 * c);
          }
        }
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          const bool has = kC ? (32 * c + 31 < kC || lane < kC - 32 * c) : (32 * c + lane < C);
          const int k = 32 * c + lane;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (has && src[u] >= 0) gp[u][32 * c] = k == rc[u] ? rt[u] : conf_scale * __expf(x[c][u] - rl[u]);
        }
      } else {
        for (int kb = 0; kb < C; kb += 32) {
          const int k = kb + lane;
          float x[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) x[u] = (src[u] >= 0 && k < C) ? rp[u][kb] : 0.f;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (src[u] >= 0 && k < C) gp[u][kb] = k == rc[u] ? rt[u] : conf_scale * __expf(x[u] - rl[u]);
        }
      }
    }
  }
}

// expand the per-prior state into the reference's int64 tensors
__global__ void targets_kernel(const LossParams q, int64_t* cls_out, int64_t* neg_out) {
  const size_t total = size_t(q.N) * q.P;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total;
       i += size_t(gridDim.x) * blockDim.x) {
    const int n = int(i / q.P);
    const int g0 = q.gt_offsets[n];
    const int G = q.gt_offsets[n + 1] - g0;
    const float ov = q.ov[i];
    int64_t lab_raw = 0;
    if (G > 0) lab_raw = q.gt_labels[g0 + q.obj[i]];
    int64_t lab = (ov < q.thr_pos) ? 0 : lab_raw;
    if (q.binarize) lab = lab > 0 ? 1 : 0;  // RefineDet512.py:778-781: threshold first, then binarise
    if (cls_out) cls_out[i] = lab;
    if (neg_out) neg_out[i] = (ov < q.thr_neg) ? -1 : lab_raw;
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Tiling {
  int rows, stages;
  uint32_t stage_floats;
  size_t smem;
};

static Tiling choose_tiling(int C, int max_stages) {
  Tiling t;
  const size_t budget = 100 * 1024;  // two CTAs per SM
  t.rows = kRows;
  while (t.rows > 4 && (size_t(t.rows) * C + 8) * 4 > budget) t.rows /= 2;
  t.stage_floats = uint32_t(align_up(size_t(t.rows) * C + 8, 32));
  t.stages = int(budget / (size_t(t.stage_floats) * 4));
  if (t.stages > max_stages) t.stages = max_stages;
  if (t.stages < 1) t.stages = 1;
  t.smem = size_t(t.stages) * t.stage_floats * 4 + kMaxStages * 8;
  return t;
}

// Process-wide switches (A/B measurements; defaults are the fast settings).
static int g_opt_pdl = 1;           // programmatic dependent launch between the kernels of one call
static int g_opt_peer_exchange = 1; // one-shot NVLink exchange of the loss sums when a communicator is attached

static size_t cm_smem_bytes(const LossParams& q) {
  return ((sizeof(CmShared) + 127) & ~size_t(127)) + align_up(size_t(q.gmax) * 4, 128) + 128;
}
static int cm_slices(int P) { return (P + kCmSlice - 1) / kCmSlice; }

static int fill_params(const sbod_loss_desc* d, LossParams& q, bool need_scores) {
  if (!d) return SBOD_ERR_INVALID;
  if (d->N <= 0 || d->P <= 0 || d->gmax < 0) return SBOD_ERR_INVALID;
  if (need_scores && (d->C <= 0 || !d->scores || !d->locs)) return SBOD_ERR_INVALID;
  if (!d->gt_boxes || !d->gt_labels || !d->gt_offsets || !d->priors_xy || !d->priors_cxcy)
    return SBOD_ERR_INVALID;
  if (!d->ov || !d->obj) return SBOD_ERR_INVALID;
  if (need_scores && (!d->lse || !d->ce || !d->sel || !d->sel_thr || !d->partials || !d->sums || !d->loss))
    return SBOD_ERR_INVALID;
  if (need_scores && (reinterpret_cast<uintptr_t>(d->scores) & 15)) return SBOD_ERR_ALIGNMENT;
  if (need_scores && (reinterpret_cast<uintptr_t>(d->locs) & 15)) return SBOD_ERR_ALIGNMENT;
  if ((long long)d->N * d->P >= (1ll << 31)) return SBOD_ERR_UNSUPPORTED;
  if (d->gmax > 8192) return SBOD_ERR_UNSUPPORTED;
  if (need_scores && d->C > 16384) return SBOD_ERR_UNSUPPORTED;
  if (d->reg_kind < 0 || d->reg_kind > SBOD_REG_CIOU) return SBOD_ERR_INVALID;
  if (d->cls_kind < 0 || d->cls_kind > SBOD_CLS_FOCAL_NORM) return SBOD_ERR_INVALID;
  q.locs = d->locs;
  q.scores = d->scores;
  q.priors_cxcy = reinterpret_cast<const float4*>(d->priors_cxcy);
  q.priors_xy = reinterpret_cast<const float4*>(d->priors_xy);
  q.anchors_xy = reinterpret_cast<const float4*>(d->anchors_xy);
  q.gt_boxes = reinterpret_cast<const float4*>(d->gt_boxes);
  q.gt_labels = d->gt_labels;
  q.gt_offsets = d->gt_offsets;
  q.exclude = d->exclude;
  q.N = d->N; q.P = d->P; q.C = d->C; q.gmax = d->gmax > 0 ? d->gmax : 1;
  q.thr_pos = d->thr_pos; q.thr_neg = d->thr_neg;
  q.reg_kind = d->reg_kind; q.cls_kind = d->cls_kind;
  q.binarize = d->binarize_labels; q.ratio = d->neg_pos_ratio;
  q.reg_weight = d->reg_weight; q.beta = d->smooth_l1_beta;
  q.falpha = d->focal_alpha; q.fgamma = d->focal_gamma;
  q.ov = d->ov; q.obj = d->obj; q.lse = d->lse; q.ce = d->ce; q.sel = d->sel;
  q.partials = d->partials; q.sums = d->sums; q.loss = d->loss;
  q.sel_thr = d->sel_thr;
  q.prefill = d->grad_scores_prefill;
  q.comm = (need_scores && g_opt_peer_exchange) ? static_cast<const CommDev*>(d->comm) : nullptr;
  if (q.comm && d->cls_kind == SBOD_CLS_CE_MINE_BATCH) return SBOD_ERR_INVALID;
  if (q.prefill && (reinterpret_cast<uintptr_t>(q.prefill) & 15)) return SBOD_ERR_ALIGNMENT;
  // workspace carve-up
  const size_t need = sbod_loss_workspace_bytes(d);
  if (!d->workspace || d->workspace_bytes < need) return SBOD_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(d->workspace) & 255) return SBOD_ERR_WORKSPACE;
  unsigned char* w = static_cast<unsigned char*>(d->workspace);
  q.counters = reinterpret_cast<unsigned int*>(w);
  w += 256;
  q.gtkey = reinterpret_cast<unsigned long long*>(w);
  w += align_up(size_t(q.N) * q.gmax * 8, 256);
  q.match_q = reinterpret_cast<unsigned int*>(w);
  w += align_up(size_t(q.N) * 4, 256);
  q.sel_hist = reinterpret_cast<unsigned int*>(w);
  w += align_up(size_t(q.N) * kBins * 4, 256);
  q.cand = reinterpret_cast<float*>(w);
  w += align_up(size_t(q.N) * q.P * 4, 256);
  q.blockpart = reinterpret_cast<double*>(w);
  w += align_up(size_t(q.N) * size_t(cm_slices(q.P)) * 32, 256);
  q.sel_seg = reinterpret_cast<float*>(w);
  w += align_up(d->cls_kind == SBOD_CLS_CE_MINE_BATCH ? size_t(q.N) * kSelCap * 4 : 0, 256);
  q.sel_segn = reinterpret_cast<int*>(w);
  w += align_up(size_t(q.N) * 4, 256);
  q.sel_above = reinterpret_cast<double*>(w);
  Tiling t = choose_tiling(need_scores ? d->C : 1, kMaxStages);
  // the two-threads-per-row layout is bank-conflict free only for odd C; even C is merely slower in smem
  q.fast = (need_scores && d->C >= 2 && d->C <= 128) ? 1 : 0;
  q.ctas_per_sm = 2;
  if (q.fast) {  // 128-row tiles; two CTAs per SM whenever two stages fit in ~90 KB
    t.rows = kTileRows;
    t.stage_floats = uint32_t(align_up(size_t(kTileRows) * d->C + 8, 32));
    const size_t sb = size_t(t.stage_floats) * 4;
    t.stages = int((90 * 1024) / sb);
    if (t.stages < 2) {
      t.stages = int((190 * 1024) / sb);
      q.ctas_per_sm = 1;
    }
    if (t.stages > kMaxStages) t.stages = kMaxStages;
    t.smem = size_t(t.stages) * sb + kMaxStages * 8;
  }
  q.rows_per_tile = t.rows;
  q.tiles_per_image = (q.P + t.rows - 1) / t.rows;
  q.n_tiles = q.tiles_per_image * q.N;
  q.n_stages = t.stages;
  q.stage_floats = t.stage_floats;
  q.with_scores = need_scores ? 1 : 0;
  q.debug_skip = 0;
#ifdef SBOD_DEBUG_HOOKS  // profiling builds only (tools/): release builds never read the environment
  {
    const char* e = getenv("SBOD_DEBUG_SKIP");
    q.debug_skip = e ? atoi(e) : 0;
  }
#endif
  return SBOD_OK;
}

}  // namespace sbod

using namespace sbod;

extern "C" size_t sbod_loss_workspace_bytes(const sbod_loss_desc* d) {
  if (!d) return 0;
  const int gmax = d->gmax > 0 ? d->gmax : 1;
  return sbod_loss_workspace_zero_bytes(d) + align_up(size_t(d->N) * d->P * 4, 256) +
         align_up(size_t(d->N) * size_t(cm_slices(d->P)) * 32, 256) +
         align_up(d->cls_kind == SBOD_CLS_CE_MINE_BATCH ? size_t(d->N) * kSelCap * 4 : 0, 256) +
         align_up(size_t(d->N) * 4, 256) + align_up(size_t(d->N) * 8, 256);
}

// leading bytes of the loss workspace that carry the zero contract: counters, per-object keys, ticket queues
extern "C" size_t sbod_loss_workspace_zero_bytes(const sbod_loss_desc* d) {
  if (!d) return 0;
  const int gmax = d->gmax > 0 ? d->gmax : 1;
  return 256 + align_up(size_t(d->N) * gmax * 8, 256) + align_up(size_t(d->N) * 4, 256) +
         align_up(size_t(d->N) * kBins * 4, 256);
}

extern "C" int sbod_set_option(int key, int value) {
  if (key == SBOD_OPT_PDL) g_opt_pdl = value ? 1 : 0;
  else if (key == SBOD_OPT_PEER_EXCHANGE) g_opt_peer_exchange = value ? 1 : 0;
  else return SBOD_ERR_INVALID;
  return SBOD_OK;
}

// Workspace contract: the first 256 bytes (counters) and the gtkey block must be zero on the
// first call; every call leaves them zero again. sbod_workspace_init does the first zeroing.
extern "C" int sbod_workspace_init(void* workspace, size_t bytes, sbod_stream_t stream) {
  if (!workspace) return SBOD_ERR_INVALID;
  SBOD_CUDA_TRY(cudaMemsetAsync(workspace, 0, bytes, static_cast<cudaStream_t>(stream)));
  return SBOD_OK;
}

static int set_kernel_attrs() {
  static DeviceOnce attr_once;
  if (!attr_once.pending()) return SBOD_OK;
  SBOD_CUDA_TRY(cudaFuncSetAttribute(match_lse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  SBOD_CUDA_TRY(cudaFuncSetAttribute(match_lse_fast_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  SBOD_CUDA_TRY(cudaFuncSetAttribute(match_lse_fast_kernel<81>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  SBOD_CUDA_TRY(cudaFuncSetAttribute(match_lse_fast_kernel<21>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  SBOD_CUDA_TRY(cudaFuncSetAttribute(loss_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  SBOD_CUDA_TRY(cudaFuncSetAttribute(loss_bwd_kernel<81>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  SBOD_CUDA_TRY(cudaFuncSetAttribute(loss_bwd_kernel<21>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  SBOD_CUDA_TRY(cudaFuncSetAttribute(forced_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  SBOD_CUDA_TRY(cudaFuncSetAttribute(classify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  // without a preference the driver picks the smallest shared-memory carve-out that fits ONE CTA per SM
  SBOD_CUDA_TRY(cudaFuncSetAttribute(classify_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  SBOD_CUDA_TRY(cudaFuncSetAttribute(match_lse_fast_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  SBOD_CUDA_TRY(cudaFuncSetAttribute(match_lse_fast_kernel<81>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  SBOD_CUDA_TRY(cudaFuncSetAttribute(match_lse_fast_kernel<21>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  attr_once.mark();
  return SBOD_OK;
}

static int launch_match(const LossParams& q, cudaStream_t st) {
  int rc = set_kernel_attrs();
  if (rc) return rc;
  if (q.fast) {
    const size_t smem = size_t(q.n_stages) * q.stage_floats * 4 + kMaxStages * 8;
    int grid = sm_count() * q.ctas_per_sm;
    if (grid > q.n_tiles) grid = q.n_tiles;
    if (q.C == 81) match_lse_fast_kernel<81><<<grid, kFastThreads, smem, st>>>(q);       // COCO
    else if (q.C == 21) match_lse_fast_kernel<21><<<grid, kFastThreads, smem, st>>>(q);  // VOC
    else match_lse_fast_kernel<0><<<grid, kFastThreads, smem, st>>>(q);
    SBOD_LAUNCH_CHECK();
    return SBOD_OK;
  }
  const Tiling t = choose_tiling(q.with_scores ? q.C : 1, kMaxStages);
  const size_t smem = q.with_scores ? t.smem : 0;
  int ctas_per_sm = 8;
  if (smem > 0) {
    ctas_per_sm = int((220 * 1024) / (smem + 9 * 1024));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 8) ctas_per_sm = 8;
  }
  int grid = sm_count() * ctas_per_sm;
  if (grid > q.n_tiles) grid = q.n_tiles;
  match_lse_kernel<<<grid, kRows, smem, st>>>(q);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

static int launch_mine(const LossParams& q, cudaStream_t st) {
  int rc = set_kernel_attrs();
  if (rc) return rc;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(cm_slices(q.P)), unsigned(q.N), 1);
    cfg.blockDim = dim3(kCmThreads, 1, 1);
    cfg.dynamicSmemBytes = cm_smem_bytes(q);
    cfg.stream = st;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = g_opt_pdl ? 1 : 0;
    SBOD_CUDA_TRY(cudaLaunchKernelEx(&cfg, classify_kernel, q));
    cfg.gridDim = dim3(unsigned(q.N), 1, 1);
    cfg.blockDim = dim3(kMineThreads, 1, 1);
    cfg.dynamicSmemBytes = 0;
    SBOD_CUDA_TRY(cudaLaunchKernelEx(&cfg, mine_kernel, q, cm_slices(q.P)));
  }
  return SBOD_OK;
}

extern "C" int sbod_loss_forward(const sbod_loss_desc* d, sbod_stream_t stream) {
  LossParams q;
  int rc = fill_params(d, q, true);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rc = launch_match(q, st);
  if (rc) return rc;
  return launch_mine(q, st);
}

// Profiling / bench hook: launch ONE stage of sbod_loss_forward: 0 = the match + log-sum-exp kernel,
// 1 = classify_kernel / mine_kernel (forced-match override + classification + mining + reduction). Stage 0 may be
// repeated; stage 1 must follow before the workspace is used by a full forward again.
extern "C" int sbod_loss_forward_stage(const sbod_loss_desc* d, int stage, sbod_stream_t stream) {
  LossParams q;
  int rc = fill_params(d, q, true);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (stage == 0) return launch_match(q, st);
  if (stage == 1) return launch_mine(q, st);
  return SBOD_ERR_INVALID;
}

#ifdef SBOD_DEBUG_HOOKS
extern "C" __attribute__((visibility("default"))) int sbod_debug_cm_times(unsigned long long* out, int reset) {
  if (out) cudaMemcpyFromSymbol(out, g_cm_times, sizeof(g_cm_times));
  if (reset) {
    unsigned long long init[8 + 4 * 64];
    for (int i = 0; i < 8 + 4 * 64; ++i) init[i] = (i < 2) ? ~0ull : 0ull;
    cudaMemcpyToSymbol(g_cm_times, init, sizeof(init));
  }
  return 0;
}
#endif

extern "C" int sbod_loss_finalize(const sbod_loss_desc* d, sbod_stream_t stream) {
  LossParams q;
  int rc = fill_params(d, q, true);
  if (rc) return rc;
  finalize_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(q);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" int sbod_loss_backward(const sbod_loss_desc* d, const float* grad_loss,
                                  float* grad_locs, float* grad_scores, sbod_stream_t stream) {
  LossParams q;
  int rc = fill_params(d, q, true);
  if (rc) return rc;
  if (grad_scores && (reinterpret_cast<uintptr_t>(grad_scores) & 15)) return SBOD_ERR_ALIGNMENT;
  if (grad_locs && (reinterpret_cast<uintptr_t>(grad_locs) & 15)) return SBOD_ERR_ALIGNMENT;
  BwdParams bp;
  const Tiling t = choose_tiling(q.C, 1);
  q.n_stages = 1;
  q.stage_floats = t.stage_floats;
  q.rows_per_tile = t.rows;
  q.tiles_per_image = (q.P + t.rows - 1) / t.rows;
  q.n_tiles = q.tiles_per_image * q.N;
  bp.q = q;
  bp.grad_loss = grad_loss;
  bp.grad_locs = grad_locs;
  bp.grad_scores = grad_scores;
  bp.dense = (q.cls_kind == SBOD_CLS_FOCAL_SUM || q.cls_kind == SBOD_CLS_FOCAL_NORM) ? 1 : 0;
  const size_t smem = size_t(t.stage_floats) * 4 + 16;
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    SBOD_CUDA_TRY(cudaFuncSetAttribute(loss_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SBOD_CUDA_TRY(cudaFuncSetAttribute(loss_bwd_kernel<81>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SBOD_CUDA_TRY(cudaFuncSetAttribute(loss_bwd_kernel<21>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_once.mark();
  }
  int ctas_per_sm = int((220 * 1024) / (smem + 4 * 1024));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm > 8) ctas_per_sm = 8;
  int grid = sm_count() * ctas_per_sm;
  if (grid > q.n_tiles) grid = q.n_tiles;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!bp.dense) {
    const bool prefilled = q.fast && grad_scores && grad_scores == q.prefill;  // zeroed by the forward
    if (grad_scores && !prefilled) {
      zero_fill_kernel<<<sm_count() * 4, 128, 0, st>>>(grad_scores, size_t(q.N) * q.P * size_t(q.C));
      SBOD_LAUNCH_CHECK();
    }
    const size_t rows = size_t(q.N) * q.P;
    int pgrid = int((rows + 255) / 256);
    if (pgrid > sm_count() * 8) pgrid = sm_count() * 8;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(pgrid), 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = g_opt_pdl ? 1 : 0;
    if (q.C == 81) SBOD_CUDA_TRY(cudaLaunchKernelEx(&cfg, bwd_patch_kernel<81>, q, grad_loss, grad_locs, grad_scores));
    else if (q.C == 21) SBOD_CUDA_TRY(cudaLaunchKernelEx(&cfg, bwd_patch_kernel<21>, q, grad_loss, grad_locs, grad_scores));
    else SBOD_CUDA_TRY(cudaLaunchKernelEx(&cfg, bwd_patch_kernel<0>, q, grad_loss, grad_locs, grad_scores));
    return SBOD_OK;
  }
  if (q.C == 81) loss_bwd_kernel<81><<<grid, kRows, smem, st>>>(bp);
  else if (q.C == 21) loss_bwd_kernel<21><<<grid, kRows, smem, st>>>(bp);
  else loss_bwd_kernel<0><<<grid, kRows, smem, st>>>(bp);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

extern "C" int sbod_loss_targets(const sbod_loss_desc* d, int64_t* cls_out, int64_t* neg_out,
                                 sbod_stream_t stream) {
  LossParams q;
  int rc = fill_params(d, q, false);
  if (rc) return rc;
  const size_t total = size_t(q.N) * q.P;
  int grid = int((total + 255) / 256);
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  targets_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(q, cls_out, neg_out);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}

// ---- stand-alone batched assignment (no scores) -------------------------------------------
extern "C" size_t sbod_assign_workspace_bytes(int N, int gmax) {
  if (gmax < 1) gmax = 1;
  return 256 + align_up(size_t(N) * gmax * 8, 256);
}

extern "C" int sbod_assign(const float* gt_boxes, const int64_t* gt_labels,
                           const int32_t* gt_offsets, int N, int gmax, const float* anchors_xy,
                           int per_image_anchors, int P, float thr_pos, float thr_neg,
                           float* ov_out, int32_t* obj_out, int64_t* cls_out, int64_t* neg_out,
                           void* workspace, size_t workspace_bytes, sbod_stream_t stream) {
  if (!gt_boxes || !gt_labels || !gt_offsets || !anchors_xy || !ov_out || !obj_out)
    return SBOD_ERR_INVALID;
  if (N <= 0 || P <= 0 || gmax < 0) return SBOD_ERR_INVALID;
  if (gmax > 8192) return SBOD_ERR_UNSUPPORTED;
  if ((long long)N * P >= (1ll << 31)) return SBOD_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < sbod_assign_workspace_bytes(N, gmax))
    return SBOD_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return SBOD_ERR_WORKSPACE;
  LossParams q;
  memset(&q, 0, sizeof(q));
  q.gt_boxes = reinterpret_cast<const float4*>(gt_boxes);
  q.gt_labels = gt_labels;
  q.gt_offsets = gt_offsets;
  q.priors_xy = reinterpret_cast<const float4*>(anchors_xy);
  q.anchors_xy = per_image_anchors ? reinterpret_cast<const float4*>(anchors_xy) : nullptr;
  q.N = N; q.P = P; q.C = 1; q.gmax = gmax > 0 ? gmax : 1;
  q.thr_pos = thr_pos; q.thr_neg = thr_neg;
  q.ov = ov_out; q.obj = obj_out;
  unsigned char* w = static_cast<unsigned char*>(workspace);
  q.counters = reinterpret_cast<unsigned int*>(w);
  q.gtkey = reinterpret_cast<unsigned long long*>(w + 256);
  q.rows_per_tile = kRows;
  q.tiles_per_image = (P + kRows - 1) / kRows;
  q.n_tiles = q.tiles_per_image * N;
  q.n_stages = 1;
  q.stage_floats = 0;
  q.with_scores = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  q.fast = 0;
  q.ctas_per_sm = 8;
  int rc = launch_match(q, st);
  if (rc) return rc;
  forced_match_kernel<<<q.N, 256, size_t(q.gmax) * 8, st>>>(q);
  SBOD_LAUNCH_CHECK();
  if (cls_out || neg_out) {
    const size_t total = size_t(N) * P;
    int grid = int((total + 255) / 256);
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    targets_kernel<<<grid, 256, 0, st>>>(q, cls_out, neg_out);
    SBOD_LAUNCH_CHECK();
  }
  return SBOD_OK;
}

// ---- end-to-end helper with HOST buffers (bench.py e2e leg) ----------------------------------
namespace {
struct HostArena {
  size_t locs, scores, pcx, pxy, gtb, gtl, gto, ov, obj, lse, ce, sel, sel_thr, partials, sums, loss, ws, total;
};
HostArena host_arena_layout(const sbod_loss_desc* d, int T) {
  HostArena a;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += align_up(bytes, 256); return at; };
  const size_t NP = size_t(d->N) * d->P;
  a.locs = take(NP * 16);
  a.scores = take(NP * size_t(d->C) * 4);
  a.pcx = take(size_t(d->P) * 16);
  a.pxy = take(size_t(d->P) * 16);
  a.gtb = take(size_t(T > 0 ? T : 1) * 16);
  a.gtl = take(size_t(T > 0 ? T : 1) * 8);
  a.gto = take(size_t(d->N + 1) * 4);
  a.ov = take(NP * 4);
  a.obj = take(NP * 4);
  a.lse = take(NP * 4);
  a.ce = take(NP * 4);
  a.sel = take(NP);
  a.sel_thr = take(size_t(d->N) * 8);
  a.partials = take(size_t(d->N) * 32);
  a.sums = take(32);
  a.loss = take(16);
  a.ws = take(sbod_loss_workspace_bytes(d));
  a.total = o;
  return a;
}
}  // namespace

extern "C" size_t sbod_loss_forward_host_arena_bytes(const sbod_loss_desc* d, int T) {
  if (!d) return 0;
  return host_arena_layout(d, T).total;
}

extern "C" int sbod_loss_forward_host(const sbod_loss_desc* h, int T, float* loss_host,
                                      void* dev_arena, size_t arena_bytes, sbod_stream_t stream) {
  if (!h || !loss_host || !dev_arena || T < 0) return SBOD_ERR_INVALID;
  if (h->anchors_xy || h->exclude) return SBOD_ERR_UNSUPPORTED;
  const HostArena a = host_arena_layout(h, T);
  if (arena_bytes < a.total) return SBOD_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(dev_arena) & 255) return SBOD_ERR_WORKSPACE;
  unsigned char* base = static_cast<unsigned char*>(dev_arena);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t NP = size_t(h->N) * h->P;
  SBOD_CUDA_TRY(cudaMemcpyAsync(base + a.locs, h->locs, NP * 16, cudaMemcpyHostToDevice, st));
  SBOD_CUDA_TRY(cudaMemcpyAsync(base + a.scores, h->scores, NP * size_t(h->C) * 4, cudaMemcpyHostToDevice, st));
  SBOD_CUDA_TRY(cudaMemcpyAsync(base + a.pcx, h->priors_cxcy, size_t(h->P) * 16, cudaMemcpyHostToDevice, st));
  SBOD_CUDA_TRY(cudaMemcpyAsync(base + a.pxy, h->priors_xy, size_t(h->P) * 16, cudaMemcpyHostToDevice, st));
  if (T > 0) {
    SBOD_CUDA_TRY(cudaMemcpyAsync(base + a.gtb, h->gt_boxes, size_t(T) * 16, cudaMemcpyHostToDevice, st));
    SBOD_CUDA_TRY(cudaMemcpyAsync(base + a.gtl, h->gt_labels, size_t(T) * 8, cudaMemcpyHostToDevice, st));
  }
  SBOD_CUDA_TRY(cudaMemcpyAsync(base + a.gto, h->gt_offsets, size_t(h->N + 1) * 4, cudaMemcpyHostToDevice, st));
  SBOD_CUDA_TRY(cudaMemsetAsync(base + a.ws, 0, sbod_loss_workspace_zero_bytes(h), st));
  sbod_loss_desc d = *h;
  d.grad_scores_prefill = nullptr;
  d.locs = reinterpret_cast<const float*>(base + a.locs);
  d.scores = reinterpret_cast<const float*>(base + a.scores);
  d.priors_cxcy = reinterpret_cast<const float*>(base + a.pcx);
  d.priors_xy = reinterpret_cast<const float*>(base + a.pxy);
  d.gt_boxes = reinterpret_cast<const float*>(base + a.gtb);
  d.gt_labels = reinterpret_cast<const int64_t*>(base + a.gtl);
  d.gt_offsets = reinterpret_cast<const int32_t*>(base + a.gto);
  d.ov = reinterpret_cast<float*>(base + a.ov);
  d.obj = reinterpret_cast<int32_t*>(base + a.obj);
  d.lse = reinterpret_cast<float*>(base + a.lse);
  d.ce = reinterpret_cast<float*>(base + a.ce);
  d.sel = reinterpret_cast<uint8_t*>(base + a.sel);
  d.sel_thr = reinterpret_cast<float*>(base + a.sel_thr);
  d.partials = reinterpret_cast<double*>(base + a.partials);
  d.sums = reinterpret_cast<double*>(base + a.sums);
  d.loss = reinterpret_cast<float*>(base + a.loss);
  d.workspace = base + a.ws;
  d.workspace_bytes = sbod_loss_workspace_bytes(h);
  int rc = sbod_loss_forward(&d, stream);
  if (rc) return rc;
  SBOD_CUDA_TRY(cudaMemcpyAsync(loss_host, d.loss, 16, cudaMemcpyDeviceToHost, st));
  return SBOD_OK;
}
