// row_stream.cuh — the fast streaming layout shared by match_lse_fast_kernel and
// detect_score_fast_kernel: a 128-row tile of [*, C] logits lands in shared memory by bulk TMA and
// is consumed by 256 threads, TWO threads per row (even / odd class indices), which doubles the
// warps per byte of staged tile compared with one thread per row.
//
// Bank-conflict-free mapping (needs C odd): a warp pair owns 32 consecutive rows, warp parity picks
// the row parity, lane parity picks the element parity. For a fixed step j the 32 lanes of a warp
// touch word  head + C*(r0 + 2i) + 2j + h  (i = lane/2, h = lane&1); C odd => C*2i mod 32 runs over
// the 16 even residues, so 2i' + h covers all 32 banks exactly once.
#pragma once

#include "common.cuh"

namespace sbod {

constexpr int kTileRows = 128;
constexpr int kStreamThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

SBOD_DEVINL float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

SBOD_DEVINL void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// thread t (0..255) of the stream group -> (row in tile, element parity)
SBOD_DEVINL void stream_map(int t, int& row, int& h) {
  const int w = t >> 5, l = t & 31;
  row = 32 * (w >> 1) + 2 * (l >> 1) + (w & 1);
  h = l & 1;
}

// max over this thread's elements rp[0], rp[2], ..., rp[2*(nh-1)], combined with the partner lane
SBOD_DEVINL float half_row_max(const float* rp, int nh) {
  const float NEG = -__int_as_float(0x7f800000);
  float m0 = NEG, m1 = NEG, m2 = NEG, m3 = NEG;
  int j = 0;
  for (; j + 8 <= nh; j += 8) {
    const float v0 = rp[2 * j], v1 = rp[2 * j + 2], v2 = rp[2 * j + 4], v3 = rp[2 * j + 6];
    const float v4 = rp[2 * j + 8], v5 = rp[2 * j + 10], v6 = rp[2 * j + 12], v7 = rp[2 * j + 14];
    m0 = fmaxf(m0, fmaxf(v0, v1));
    m1 = fmaxf(m1, fmaxf(v2, v3));
    m2 = fmaxf(m2, fmaxf(v4, v5));
    m3 = fmaxf(m3, fmaxf(v6, v7));
  }
  for (; j < nh; ++j) m0 = fmaxf(m0, rp[2 * j]);
  float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
  return fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
}

// sum of exp(x - mx) over this thread's elements, combined with the partner lane.
// nmx2 = -mx * log2(e)
SBOD_DEVINL float half_row_sumexp(const float* rp, int nh, float nmx2) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int j = 0;
  for (; j + 8 <= nh; j += 8) {
    const float v0 = rp[2 * j], v1 = rp[2 * j + 2], v2 = rp[2 * j + 4], v3 = rp[2 * j + 6];
    const float v4 = rp[2 * j + 8], v5 = rp[2 * j + 10], v6 = rp[2 * j + 12], v7 = rp[2 * j + 14];
    s0 += ex2_approx(fmaf(v0, kLog2e, nmx2)) + ex2_approx(fmaf(v1, kLog2e, nmx2));
    s1 += ex2_approx(fmaf(v2, kLog2e, nmx2)) + ex2_approx(fmaf(v3, kLog2e, nmx2));
    s2 += ex2_approx(fmaf(v4, kLog2e, nmx2)) + ex2_approx(fmaf(v5, kLog2e, nmx2));
    s3 += ex2_approx(fmaf(v6, kLog2e, nmx2)) + ex2_approx(fmaf(v7, kLog2e, nmx2));
  }
  for (; j < nh; ++j) s0 += ex2_approx(fmaf(rp[2 * j], kLog2e, nmx2));
  const float s = (s0 + s1) + (s2 + s3);
  return s + __shfl_xor_sync(0xffffffffu, s, 1);
}

// Compile-time C: sum of exp(x - shift) over one row shared by the two threads of a pair, fully
// unrolled (no loop control, immediate offsets). rbase = first logit of the row, h = element parity
// of this thread, nshift2 = -shift * log2(e). The caller checks the result for overflow.
template <int kC>
SBOD_DEVINL float pair_row_sumexp_fixed(const float* rbase, int h, float nshift2) {
  constexpr int kBoth = kC / 2;  // elements both parities own; an odd C gives parity 0 one more
  const float* rp = rbase + h;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int j = 0; j < kBoth; j += 4) {
    s0 += ex2_approx(fmaf(rp[2 * j], kLog2e, nshift2));
    if (j + 1 < kBoth) s1 += ex2_approx(fmaf(rp[2 * j + 2], kLog2e, nshift2));
    if (j + 2 < kBoth) s2 += ex2_approx(fmaf(rp[2 * j + 4], kLog2e, nshift2));
    if (j + 3 < kBoth) s3 += ex2_approx(fmaf(rp[2 * j + 6], kLog2e, nshift2));
  }
  if ((kC & 1) && h == 0) s1 += ex2_approx(fmaf(rp[2 * kBoth], kLog2e, nshift2));
  const float s = (s0 + s1) + (s2 + s3);
  return s + __shfl_xor_sync(0xffffffffu, s, 1);
}

// Compile-time C variants of half_row_max / half_row_sumexp_mask for the eval path: fully unrolled over
// the kC / 2 elements both parities own (+ one more for parity 0 when kC is odd), immediate offsets and
// immediate mask bits. rbase = first logit of the row, h = element parity of this thread.
template <int kC>
SBOD_DEVINL float pair_row_max_fixed(const float* rbase, int h) {
  constexpr int kBoth = kC / 2;
  const float NEG = -__int_as_float(0x7f800000);
  const float* rp = rbase + h;
  float a0 = NEG, a1 = NEG, a2 = NEG, a3 = NEG;
#pragma unroll
  for (int j = 0; j < kBoth; j += 4) {
    a0 = fmaxf(a0, rp[2 * j]);
    if (j + 1 < kBoth) a1 = fmaxf(a1, rp[2 * j + 2]);
    if (j + 2 < kBoth) a2 = fmaxf(a2, rp[2 * j + 4]);
    if (j + 3 < kBoth) a3 = fmaxf(a3, rp[2 * j + 6]);
  }
  if ((kC & 1) && h == 0) a1 = fmaxf(a1, rp[2 * kBoth]);
  const float m = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
  return fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
}

template <int kC>
SBOD_DEVINL float pair_row_sumexp_mask_fixed(const float* rbase, int h, float nmx2, float floor_e, uint32_t& m0,
                                             uint32_t& m1) {
  constexpr int kBoth = kC / 2;
  const float* rp = rbase + h;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  m0 = 0u;
  m1 = 0u;
#pragma unroll
  for (int j = 0; j < kBoth; ++j) {
    const float e = ex2_approx(fmaf(rp[2 * j], kLog2e, nmx2));
    if ((j & 3) == 0) s0 += e;
    else if ((j & 3) == 1) s1 += e;
    else if ((j & 3) == 2) s2 += e;
    else s3 += e;
    if (e > floor_e) {
      if (j < 32) m0 |= 1u << j;
      else m1 |= 1u << (j - 32);
    }
  }
  if ((kC & 1) && h == 0) {
    const float e = ex2_approx(fmaf(rp[2 * kBoth], kLog2e, nmx2));
    s1 += e;
    if (e > floor_e) {
      if (kBoth < 32) m0 |= 1u << kBoth;
      else m1 |= 1u << (kBoth - 32);
    }
  }
  const float s = (s0 + s1) + (s2 + s3);
  return s + __shfl_xor_sync(0xffffffffu, s, 1);
}

// As above, and additionally records in (m0, m1) which of this thread's elements (bit j, j < 64)
// have exp(x - mx) > floor. Since the row sum is >= 1, prob = e/sum <= e, so the mask is a
// superset of the classes whose probability exceeds `floor`.
SBOD_DEVINL float half_row_sumexp_mask(const float* rp, int nh, float nmx2, float floor_e,
                                       uint32_t& m0, uint32_t& m1) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  m0 = 0u;
  m1 = 0u;
  int j = 0;
  if (nh >= 32) {
#pragma unroll
    for (int u = 0; u < 32; u += 4) {
      const float e0 = ex2_approx(fmaf(rp[2 * u], kLog2e, nmx2));
      const float e1 = ex2_approx(fmaf(rp[2 * u + 2], kLog2e, nmx2));
      const float e2 = ex2_approx(fmaf(rp[2 * u + 4], kLog2e, nmx2));
      const float e3 = ex2_approx(fmaf(rp[2 * u + 6], kLog2e, nmx2));
      s0 += e0; s1 += e1; s2 += e2; s3 += e3;
      if (e0 > floor_e) m0 |= 1u << u;
      if (e1 > floor_e) m0 |= 1u << (u + 1);
      if (e2 > floor_e) m0 |= 1u << (u + 2);
      if (e3 > floor_e) m0 |= 1u << (u + 3);
    }
    j = 32;
    if (nh >= 64) {
#pragma unroll
      for (int u = 0; u < 32; u += 4) {
        const float e0 = ex2_approx(fmaf(rp[2 * (32 + u)], kLog2e, nmx2));
        const float e1 = ex2_approx(fmaf(rp[2 * (32 + u) + 2], kLog2e, nmx2));
        const float e2 = ex2_approx(fmaf(rp[2 * (32 + u) + 4], kLog2e, nmx2));
        const float e3 = ex2_approx(fmaf(rp[2 * (32 + u) + 6], kLog2e, nmx2));
        s0 += e0; s1 += e1; s2 += e2; s3 += e3;
        if (e0 > floor_e) m1 |= 1u << u;
        if (e1 > floor_e) m1 |= 1u << (u + 1);
        if (e2 > floor_e) m1 |= 1u << (u + 2);
        if (e3 > floor_e) m1 |= 1u << (u + 3);
      }
      j = 64;
    }
  }
  for (; j < nh; ++j) {
    const float e = ex2_approx(fmaf(rp[2 * j], kLog2e, nmx2));
    s0 += e;
    if (e > floor_e) {
      if (j < 32) m0 |= 1u << j;
      else m1 |= 1u << (j - 32);
    }
  }
  const float s = (s0 + s1) + (s2 + s3);
  return s + __shfl_xor_sync(0xffffffffu, s, 1);
}

// Eval path, softmax rows: sum of exp(x - shift) over ALL elements of a row and the maximum of exp(x - shift)
// over its FOREGROUND elements (k >= 1), shared by the two threads of a pair. kC > 0: compile-time class
// count (fully unrolled, immediate offsets); kC == 0: run-time C. rbase = first logit of the row, h = element
// parity of this thread, nshift2 = -shift * log2(e). The caller checks the sum for overflow.
template <int kC>
SBOD_DEVINL void pair_row_sum_fgmax(const float* rbase, int h, int C, float nshift2, float& sum, float& fgmax) {
  const int Cc = kC ? kC : C;
  const int both = Cc / 2;
  const float* rp = rbase + h;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  float m0 = 0.f, m1 = 0.f;  // exp() > 0: 0 is the neutral element
  if (kC) {
#pragma unroll
    for (int j = 0; j < kC / 2; ++j) {
      const float e = ex2_approx(fmaf(rp[2 * j], kLog2e, nshift2));
      if ((j & 3) == 0) s0 += e;
      else if ((j & 3) == 1) s1 += e;
      else if ((j & 3) == 2) s2 += e;
      else s3 += e;
      if (j == 0) {
        if (h == 1) m0 = e;  // element 0 of parity 0 is the background
      } else if (j & 1) {
        m1 = fmaxf(m1, e);
      } else {
        m0 = fmaxf(m0, e);
      }
    }
  } else {
    for (int j = 0; j < both; ++j) {
      const float e = ex2_approx(fmaf(rp[2 * j], kLog2e, nshift2));
      s0 += e;
      if (j > 0 || h == 1) m0 = fmaxf(m0, e);
    }
  }
  if ((Cc & 1) && h == 0) {  // an odd C gives parity 0 one more element (never the background: Cc >= 3 then)
    const float e = ex2_approx(fmaf(rp[2 * both], kLog2e, nshift2));
    s1 += e;
    if (both > 0) m1 = fmaxf(m1, e);
  }
  const float s = (s0 + s1) + (s2 + s3);
  const float m = fmaxf(m0, m1);
  sum = s + __shfl_xor_sync(0xffffffffu, s, 1);
  fgmax = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
}

// generic tile geometry shared by the streaming kernels
struct StreamTile {
  int n, p0, rows;
};
SBOD_DEVINL StreamTile stream_tile(int tile, int tiles_per_image, int rows_per_tile, int P) {
  StreamTile t;
  t.n = tile / tiles_per_image;
  t.p0 = (tile - t.n * tiles_per_image) * rows_per_tile;
  t.rows = min(rows_per_tile, P - t.p0);
  return t;
}

// arm `bar` and start the bulk copy of rows [p0, p0+rows) of image n of a dense [N,P,C] tensor
SBOD_DEVINL void stream_issue(const float* base, int N, int P, int C, const StreamTile& t,
                              float* stage, uint64_t* bar) {
  const size_t first = (size_t(t.n) * P + t.p0) * size_t(C);
  const size_t total = size_t(N) * P * size_t(C);
  const TileSpan s = make_tile_span(base, first, size_t(t.rows) * C, total);
  for (uint32_t i = 0; i < s.tail_floats; ++i)
    stage[s.bulk_bytes / 4 + i] = base[(s.src16 - base) + s.bulk_bytes / 4 + i];
  if (s.bulk_bytes) {
    mbar_arrive_expect_tx(bar, s.bulk_bytes);
    tma_load_1d_hint(stage, s.src16, s.bulk_bytes, bar, l2_policy_evict_first());  // (streamed once)
  } else {
    mbar_arrive(bar);
  }
}

// contiguous, balanced range of tiles owned by CTA b of g
SBOD_DEVINL void tile_range(int n_tiles, int b, int g, int& t0, int& t1) {
  t0 = int((long long)n_tiles * b / g);
  t1 = int((long long)n_tiles * (b + 1) / g);
}

}  // namespace sbod
