// comm.cu — host side of the in-kernel loss-sum exchange (comm.cuh): mailbox allocation, CUDA IPC export /
// import between the processes (one per GPU) of a node. The kernels that use it: classify_mine_kernel's
// last-CTA epilogue (loss.cu) and fcos_finalize_kernel (fcos.cu).
#include <string.h>

#include "comm.cuh"

namespace sbod {

struct Comm {
  CommDev host;        // host copy of the device descriptor
  CommDev* dev;        // device copy handed to the kernels
  void* local;         // own allocation: epoch word (first 256 bytes) + mailbox
  void* peers[kCommMaxWorld];  // bases of the opened peer allocations (nullptr for the own rank)
  int device;
};

static size_t comm_bytes(int world) { return 256 + size_t(2) * world * kCommSlotDoubles * sizeof(double); }

}  // namespace sbod

using namespace sbod;

extern "C" size_t sbod_comm_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

extern "C" int sbod_comm_create(int rank, int world, void** comm_out, void* handle_out) {
  if (!comm_out || !handle_out || world < 1 || world > kCommMaxWorld || rank < 0 || rank >= world)
    return SBOD_ERR_INVALID;
  Comm* c = new Comm();
  memset(c, 0, sizeof(Comm));
  cudaGetDevice(&c->device);
  cudaError_t e = cudaMalloc(&c->local, comm_bytes(world));
  if (e == cudaSuccess) e = cudaMemset(c->local, 0, comm_bytes(world));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&c->dev), sizeof(CommDev));
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->local);
  if (e != cudaSuccess) {
    if (c->local) cudaFree(c->local);
    if (c->dev) cudaFree(c->dev);
    delete c;
    return int(e);
  }
  memcpy(handle_out, &h, sizeof(h));
  c->host.rank = rank;
  c->host.world = world;
  c->host.epoch = reinterpret_cast<unsigned long long*>(c->local);
  c->host.mailbox = reinterpret_cast<double*>(static_cast<unsigned char*>(c->local) + 256);
  *comm_out = c;
  return SBOD_OK;
}

// all_handles: world * sbod_comm_handle_bytes() bytes, rank-major (gathered by the caller, e.g. with
// torch.distributed.all_gather). Every rank must have created its comm before any rank connects.
extern "C" int sbod_comm_connect(void* comm, const void* all_handles) {
  if (!comm || !all_handles) return SBOD_ERR_INVALID;
  Comm* c = static_cast<Comm*>(comm);
  const unsigned char* hs = static_cast<const unsigned char*>(all_handles);
  for (int r = 0; r < c->host.world; ++r) {
    void* base = c->local;
    if (r != c->host.rank) {
      cudaIpcMemHandle_t h;
      memcpy(&h, hs + size_t(r) * sizeof(h), sizeof(h));
      SBOD_CUDA_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
      c->peers[r] = base;
    }
    c->host.peer_mailbox[r] = reinterpret_cast<double*>(static_cast<unsigned char*>(base) + 256);
  }
  SBOD_CUDA_TRY(cudaMemcpy(c->dev, &c->host, sizeof(CommDev), cudaMemcpyHostToDevice));
  return SBOD_OK;
}

extern "C" const void* sbod_comm_device_ptr(void* comm) {
  return comm ? static_cast<Comm*>(comm)->dev : nullptr;
}

extern "C" int sbod_comm_destroy(void* comm) {
  if (!comm) return SBOD_OK;
  Comm* c = static_cast<Comm*>(comm);
  for (int r = 0; r < c->host.world; ++r)
    if (c->peers[r]) cudaIpcCloseMemHandle(c->peers[r]);
  if (c->dev) cudaFree(c->dev);
  if (c->local) cudaFree(c->local);
  delete c;
  return SBOD_OK;
}

// Stand-alone all-reduce of k <= 7 doubles through the mailboxes (tests; the loss kernels call the device
// function from their own epilogues instead of launching this).
__global__ void comm_allreduce_kernel(const CommDev* c, double* vals, int k) {
  double v[kCommSlotDoubles - 1];
  for (int i = 0; i < k; ++i) v[i] = vals[i];
  comm_allreduce_sum(c, v, k);
  if (threadIdx.x == 0)
    for (int i = 0; i < k; ++i) vals[i] = v[i];
}

extern "C" int sbod_comm_allreduce(void* comm, double* vals_dev, int k, sbod_stream_t stream) {
  if (!comm || !vals_dev || k < 1 || k > kCommSlotDoubles - 1) return SBOD_ERR_INVALID;
  comm_allreduce_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<Comm*>(comm)->dev, vals_dev, k);
  SBOD_LAUNCH_CHECK();
  return SBOD_OK;
}
