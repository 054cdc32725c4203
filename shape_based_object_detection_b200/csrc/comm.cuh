// comm.cuh — one-shot exchange of a few doubles between the GPUs of one node over NVLink / NVSwitch
// peer memory, callable from INSIDE a kernel (sm_100a). Used for the only data of the path that
// crosses GPUs: the loss sums of a sharded batch.
//
// Every rank owns a "mailbox" in its own HBM:  slot[parity][source rank] = 7 doubles + an epoch flag.
// A call with epoch e: each rank stores its contribution into slot[e & 1][own rank] of EVERY rank's
// mailbox (peer-mapped pointers, plain stores over NVLink), fences at system scope and stores the flag;
// then it polls its OWN mailbox until all `world` flags of that parity show e, and adds the
// contributions in rank order - the same sequence of additions on every rank, so all ranks hold
// bit-identical sums. The epoch lives in device memory and is advanced by the kernel itself, so a
// captured CUDA graph replays correctly. Two parities suffice: a rank cannot finish call e + 1 (which
// it must before it overwrites parity e & 1 again in call e + 2) without every other rank having
// posted e + 1, i.e. having finished reading call e.
#pragma once

#include "common.cuh"

namespace sbod {

constexpr int kCommMaxWorld = 16;
constexpr int kCommSlotDoubles = 8;  // 7 payload doubles + 1 epoch word

struct CommDev {
  int rank, world;
  unsigned long long* epoch;            // [1] local call counter
  double* mailbox;                      // local: [2][world][kCommSlotDoubles]
  double* peer_mailbox[kCommMaxWorld];  // every rank's mailbox as mapped into this process (own included)
};

SBOD_DEVINL void st_relaxed_sys_f64(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
SBOD_DEVINL void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
SBOD_DEVINL unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
SBOD_DEVINL double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// All-reduce (sum) of vals[0..k), k <= 7, across the ranks of `c`. Call with (at least) the first warp of
// ONE CTA per rank, all 32 lanes of that warp; the result is valid in every lane on return.
SBOD_DEVINL void comm_allreduce_sum(const CommDev* __restrict__ cp, double* vals, int k) {
  const CommDev c = *cp;
  const int lane = threadIdx.x & 31;
  const unsigned long long e = *c.epoch + 1ull;
  const int parity = int(e & 1ull);
  const size_t slot_own = (size_t(parity) * c.world + c.rank) * kCommSlotDoubles;
  if (lane < c.world) {  // lane r posts to rank r
    double* dst = c.peer_mailbox[lane] + slot_own;
    for (int i = 0; i < k; ++i) st_relaxed_sys_f64(dst + i, vals[i]);
    // (the release store orders this lane's payload stores before the flag: no separate system fence)
    st_release_sys_u64(reinterpret_cast<unsigned long long*>(dst + kCommSlotDoubles - 1), e);
  }
  __syncwarp();
  double mine[kCommSlotDoubles - 1];
  if (lane < c.world) {  // lane r waits for rank r's contribution
    const double* src = c.mailbox + (size_t(parity) * c.world + lane) * kCommSlotDoubles;
    while (ld_acquire_sys_u64(reinterpret_cast<const unsigned long long*>(src + kCommSlotDoubles - 1)) != e) {
    }
    for (int i = 0; i < k; ++i) mine[i] = ld_relaxed_sys_f64(src + i);
  }
  for (int i = 0; i < k; ++i) {  // rank order: identical on every rank
    double acc = 0.0;
    for (int r = 0; r < c.world; ++r) acc += __shfl_sync(0xffffffffu, lane < c.world ? mine[i] : 0.0, r);
    vals[i] = acc;
  }
  __syncwarp();
  if (lane == 0) *c.epoch = e;
}

}  // namespace sbod
