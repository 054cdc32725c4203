// map.cu — VOC07 11-point mean average precision of a set of detections (sm_100a).
//
// Replaces metrics.calculate_mAP (reference metrics.py:8-145), the consumer of detect()'s output in
// every evaluate(): there a Python loop per detection with one find_jaccard_overlap call each.
//
//   map_match_kernel   one CTA per image. A detection's verdict (true positive / false positive /
//                      ignored because it hit a "difficult" object) depends only on the detections of
//                      the same image and class that precede it in descending-score order
//                      (metrics.py:84-118 keeps one "already detected" flag per object), so images are
//                      independent. Each detection gets its rank inside (image, class) by counting
//                      (score desc, index asc); round r then settles every detection of rank r in
//                      parallel — within a round different threads touch objects of different classes.
//   global sort        key = class << 32 | descending-score bits, stable (ties keep the order of the
//                      concatenated detections): cub::DeviceRadixSort, 48 key bits.
//   map_ap_kernel      one CTA per class: segment of the class by binary search, cumulative TP / FP in
//                      sorted order (block scan with carry), precision / recall in fp32 exactly as the
//                      reference's tensor expressions, max precision at recall >= t for the 11
//                      thresholds, their mean.
#include <cub/device/device_radix_sort.cuh>
#include <stdint.h>

#include "../../include/sbod.h"
#include "common.cuh"

namespace sbod {

constexpr int kMapThreads = 256;

struct MapParams {
  const float4* det_boxes;
  const int64_t* det_labels;
  const float* det_scores;
  const int32_t* det_offsets;  // [N+1]
  const float4* gt_boxes;
  const int64_t* gt_labels;
  const uint8_t* gt_difficult;
  const int32_t* gt_offsets;  // [N+1]
  int N, D, T, n_classes;
  double threshold;
  float recall_thr[11];
  uint8_t* verdict;            // [D] 1 = TP, 2 = FP, 0 = ignored
  int32_t* rank;               // [D] rank inside (image, class)
  unsigned long long* keys_in; // [D]
  unsigned long long* keys_out;
  uint32_t* idx_in;            // [D]
  uint32_t* idx_out;
  float* out_ap;               // [n_classes - 1]
};

// metrics.find_jaccard_overlap(detection[1,4], objects[k,4]) for one pair: IoU with eps, a zero-size
// detection gives 0, a zero-size object -1 (applied last)                       metrics.py:208-252
SBOD_DEVINL float map_iou(const float4 d, const float4 o) {
  float v = iou_metrics_rn(d, box_area_rn(d), o, box_area_rn(o));
  if (gt_is_zero(d)) v = 0.f;
  if (anchor_is_zero(o)) v = -1.f;
  return v;
}

__global__ void __launch_bounds__(kMapThreads) map_match_kernel(const MapParams q) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint8_t* s_found = smem_raw;  // [G] object already matched by an earlier detection
  __shared__ int s_max_rank;
  const int n = blockIdx.x, tid = threadIdx.x;
  const int d0 = q.det_offsets[n], d1 = q.det_offsets[n + 1];
  const int g0 = q.gt_offsets[n], G = q.gt_offsets[n + 1] - g0;
  for (int g = tid; g < G; g += kMapThreads) s_found[g] = 0;
  if (tid == 0) s_max_rank = -1;
  __syncthreads();
  // rank inside (image, class): detections that precede this one in (score desc, index asc)
  for (int d = d0 + tid; d < d1; d += kMapThreads) {
    const int64_t c = q.det_labels[d];
    const float s = q.det_scores[d];
    int r = 0;
    for (int e = d0; e < d1; ++e) {
      if (q.det_labels[e] != c) continue;
      const float se = q.det_scores[e];
      if (se > s || (se == s && e < d)) ++r;
    }
    q.rank[d] = r;
    atomicMax(&s_max_rank, r);
    // the global order: class ascending, score descending, then the concatenation order (stable sort)
    q.keys_in[d] = (static_cast<unsigned long long>(uint32_t(c) & 0xffffu) << 32) |
                   static_cast<unsigned long long>(0xffffffffu - float_sortable_u32(s));
    q.idx_in[d] = uint32_t(d);
  }
  __syncthreads();
  const int rounds = s_max_rank + 1;
  for (int r = 0; r < rounds; ++r) {
    for (int d = d0 + tid; d < d1; d += kMapThreads) {
      if (q.rank[d] != r) continue;
      const int64_t c = q.det_labels[d];
      const float4 box = q.det_boxes[d];
      float best = 0.f;
      int arg = -1;  // first object of the class with the largest overlap (torch.max: first index)
      for (int g = 0; g < G; ++g) {
        if (q.gt_labels[g0 + g] != c) continue;
        const float v = map_iou(box, q.gt_boxes[g0 + g]);
        if (arg < 0 || v > best) {
          best = v;
          arg = g;
        }
      }
      uint8_t verdict = 2;  // no object of this class in the image, or not enough overlap
      if (arg >= 0 && double(best) > q.threshold) {
        if (q.gt_difficult[g0 + arg]) verdict = 0;  // matched a difficult object: neither TP nor FP
        else if (!s_found[arg]) {
          verdict = 1;
          s_found[arg] = 1;
        }
      }
      q.verdict[d] = verdict;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kMapThreads) map_ap_kernel(const MapParams q) {
  __shared__ int s_scan[kMapThreads];
  __shared__ int s_scan2[kMapThreads];
  __shared__ float s_best[11][kMapThreads / 32];
  __shared__ int s_carry[2];
  __shared__ int s_easy;
  const int c = blockIdx.x + 1, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // segment of class c in the sorted keys: [lower_bound(c << 32), lower_bound((c + 1) << 32))
  auto lower_bound = [&](unsigned long long key) {
    int lo = 0, hi = q.D;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (q.keys_out[mid] < key) lo = mid + 1;
      else hi = mid;
    }
    return lo;
  };
  const int s0 = lower_bound(static_cast<unsigned long long>(c) << 32);
  const int s1 = lower_bound(static_cast<unsigned long long>(c + 1) << 32);
  // objects of the class that are not "difficult"                             metrics.py:60
  if (tid == 0) {
    s_easy = 0;
    s_carry[0] = s_carry[1] = 0;
  }
  __syncthreads();
  int easy = 0;
  for (int g = tid; g < q.T; g += kMapThreads)
    if (q.gt_labels[g] == c && !q.gt_difficult[g]) ++easy;
  easy = warp_sum(easy);
  if (lane == 0 && easy) atomicAdd(&s_easy, easy);
  __syncthreads();
  const float n_easy = float(s_easy);
  float best[11];
#pragma unroll
  for (int i = 0; i < 11; ++i) best[i] = -1.f;  // -1: no detection reaches this recall
  for (int base = s0; base < s1; base += kMapThreads) {
    const int i = base + tid;
    int tp = 0, fp = 0;
    if (i < s1) {
      const uint8_t v = q.verdict[q.idx_out[i]];
      tp = v == 1;
      fp = v == 2;
    }
    // inclusive block scan of (tp, fp)
    s_scan[tid] = tp;
    s_scan2[tid] = fp;
    __syncthreads();
    for (int o = 1; o < kMapThreads; o <<= 1) {
      const int a = tid >= o ? s_scan[tid - o] : 0;
      const int b = tid >= o ? s_scan2[tid - o] : 0;
      __syncthreads();
      s_scan[tid] += a;
      s_scan2[tid] += b;
      __syncthreads();
    }
    if (i < s1) {
      // fp32 like the reference's tensors: cumsum, tp / (tp + fp + 1e-10), tp / n_easy  (metrics.py:121-125)
      const float ctp = float(s_carry[0] + s_scan[tid]), cfp = float(s_carry[1] + s_scan2[tid]);
      const float precision = __fdiv_rn(ctp, __fadd_rn(__fadd_rn(ctp, cfp), 1e-10f));
      const float recall = __fdiv_rn(ctp, n_easy);  // 0 / 0 = NaN: no threshold is reached
#pragma unroll
      for (int t = 0; t < 11; ++t)
        if (recall >= q.recall_thr[t]) best[t] = fmaxf(best[t], precision);
    }
    __syncthreads();
    if (tid == kMapThreads - 1) {
      s_carry[0] += s_scan[tid];
      s_carry[1] += s_scan2[tid];
    }
    __syncthreads();
  }
#pragma unroll
  for (int t = 0; t < 11; ++t) {
    const float m = warp_max(best[t]);
    if (lane == 0) s_best[t][wid] = m;
  }
  __syncthreads();
  if (tid == 0) {
    float acc = 0.f;  // precisions.mean() over the 11 thresholds
    for (int t = 0; t < 11; ++t) {
      float m = -1.f;
      for (int w = 0; w < kMapThreads / 32; ++w) m = fmaxf(m, s_best[t][w]);
      acc = __fadd_rn(acc, m < 0.f ? 0.f : m);
    }
    // a class without detections keeps AP = 0 (metrics.py:73 `continue`)
    q.out_ap[c - 1] = s1 > s0 ? __fdiv_rn(acc, 11.f) : 0.f;
  }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct MapLayout {
  size_t verdict, rank, keys_in, keys_out, idx_in, idx_out, cub, cub_bytes, total;
};

static MapLayout map_layout(int D) {
  MapLayout l;
  const size_t d = size_t(D > 0 ? D : 1);
  size_t b = 0;
  l.verdict = b; b += align_up(d, 256);
  l.rank = b; b += align_up(d * 4, 256);
  l.keys_in = b; b += align_up(d * 8, 256);
  l.keys_out = b; b += align_up(d * 8, 256);
  l.idx_in = b; b += align_up(d * 4, 256);
  l.idx_out = b; b += align_up(d * 4, 256);
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, static_cast<const unsigned long long*>(nullptr),
                                  static_cast<unsigned long long*>(nullptr), static_cast<const uint32_t*>(nullptr),
                                  static_cast<uint32_t*>(nullptr), int(d), 0, 48);
  l.cub = b;
  l.cub_bytes = cub_bytes;
  b += align_up(cub_bytes, 256);
  l.total = b;
  return l;
}

}  // namespace sbod

using namespace sbod;

extern "C" size_t sbod_map_workspace_bytes(int n_detections) { return map_layout(n_detections).total; }

extern "C" int sbod_map(const float* det_boxes, const int64_t* det_labels, const float* det_scores,
                        const int32_t* det_offsets, int n_detections, const float* true_boxes,
                        const int64_t* true_labels, const uint8_t* true_difficulties, const int32_t* gt_offsets,
                        int n_objects, int n_images, int gmax, int n_classes, double threshold,
                        const float* recall_thresholds11, float* out_ap, void* workspace, size_t workspace_bytes,
                        sbod_stream_t stream) {
  if (n_images < 0 || n_detections < 0 || n_objects < 0 || n_classes < 2 || n_classes > 65535 || !out_ap ||
      !recall_thresholds11 || !det_offsets || !gt_offsets)
    return SBOD_ERR_INVALID;
  if (n_detections > 0 && (!det_boxes || !det_labels || !det_scores)) return SBOD_ERR_INVALID;
  if (n_objects > 0 && (!true_boxes || !true_labels || !true_difficulties)) return SBOD_ERR_INVALID;
  const MapLayout l = map_layout(n_detections);
  if (!workspace || workspace_bytes < l.total || (reinterpret_cast<uintptr_t>(workspace) & 255)) return SBOD_ERR_WORKSPACE;
  if (gmax < 0 || gmax > 200 * 1024) return SBOD_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned char* w = static_cast<unsigned char*>(workspace);
  MapParams q;
  q.det_boxes = reinterpret_cast<const float4*>(det_boxes);
  q.det_labels = det_labels;
  q.det_scores = det_scores;
  q.det_offsets = det_offsets;
  q.gt_boxes = reinterpret_cast<const float4*>(true_boxes);
  q.gt_labels = true_labels;
  q.gt_difficult = true_difficulties;
  q.gt_offsets = gt_offsets;
  q.N = n_images; q.D = n_detections; q.T = n_objects; q.n_classes = n_classes;
  q.threshold = threshold;
  for (int i = 0; i < 11; ++i) q.recall_thr[i] = recall_thresholds11[i];
  q.verdict = w + l.verdict;
  q.rank = reinterpret_cast<int32_t*>(w + l.rank);
  q.keys_in = reinterpret_cast<unsigned long long*>(w + l.keys_in);
  q.keys_out = reinterpret_cast<unsigned long long*>(w + l.keys_out);
  q.idx_in = reinterpret_cast<uint32_t*>(w + l.idx_in);
  q.idx_out = reinterpret_cast<uint32_t*>(w + l.idx_out);
  q.out_ap = out_ap;
  if (n_detections > 0 && n_images > 0) {
    static DeviceOnce attr_once;
    if (attr_once.pending()) {
      SBOD_CUDA_TRY(cudaFuncSetAttribute(map_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_once.mark();
    }
    map_match_kernel<<<n_images, kMapThreads, size_t(gmax > 0 ? gmax : 1), st>>>(q);
    SBOD_LAUNCH_CHECK();
    size_t cub_bytes = l.cub_bytes;
    SBOD_CUDA_TRY(cub::DeviceRadixSort::SortPairs(w + l.cub, cub_bytes, q.keys_in, q.keys_out, q.idx_in, q.idx_out,
                                                  n_detections, 0, 48, st));
  }
  map_ap_kernel<<<n_classes - 1, kMapThreads, 0, st>>>(q);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}
