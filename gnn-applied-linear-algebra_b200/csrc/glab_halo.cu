// glab_halo.cu -- peer-memory halo exchange for the row-block partitioned operators.
//
// One process per GPU.  Each rank allocates its gathered vectors (n_local + n_halo rows) with
// glab_ipc_alloc so that neighbours can map them (CUDA IPC over NVLink / NVSwitch).  After a
// step has produced new values, glab_halo_push_* packs this rank's boundary rows and stores
// them DIRECTLY into the peer's halo tail through the mapped pointer, then publishes a
// monotonically increasing epoch flag with a system-scope release.  The consumer orders its
// next kernel behind glab_halo_wait (a one-thread acquire spin on its own flag word) -- there is
// no NCCL call and no host round trip on the data path.
#include <cstring>
#include "glab_common.cuh"

namespace glab {

struct PushArgs {
  glab_push_desc d[GLAB_MAX_PEERS];
};

template <typename T, int K>
__global__ void k_halo_push(const T* __restrict__ src, PushArgs a, unsigned int* done_counter,
                            uint32_t* pushed_local) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && pushed_local) *pushed_local += 1u;
  const glab_push_desc d = a.d[blockIdx.y];
  push_rows<T, K>(src, d, (int)(blockIdx.x * blockDim.x + threadIdx.x), (int)(gridDim.x * blockDim.x));
  if (d.flag == nullptr) return;
  // the last CTA of this peer's slice publishes the arrival after all stores are visible
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(done_counter + blockIdx.y, 1u);
    if (t == gridDim.x - 1) {
      done_counter[blockIdx.y] = 0u;
      __threadfence_system();
      asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(d.flag) : "memory");
    }
  }
}

struct WaitArgs {
  uint32_t* flag[GLAB_MAX_PEERS];
};

__global__ void k_bump(uint32_t* p) { *p += 1u; }

__global__ void k_halo_wait(WaitArgs a, int n, const uint32_t* pushed_local) {
  const int i = threadIdx.x;
  if (i >= n) return;
  const uint32_t want = *pushed_local;
  uint32_t v;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a.flag[i]) : "memory");
    if ((int32_t)(v - want) >= 0) break;
    __nanosleep(32);
  } while (true);
}

static unsigned int* done_counters() {
  // GLAB_MAX_PEERS 4-byte counters per process/device, zero-initialised; pushes on one stream
  // are ordered, and every launch leaves them at zero.
  static unsigned int* ctrs[kMaxDevices] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  unsigned int*& ctr = ctrs[dev % kMaxDevices];
  if (!ctr) {
    if (cudaMalloc(&ctr, 256) != cudaSuccess) return nullptr;
    cudaMemset(ctr, 0, 256);
  }
  return ctr;
}

template <typename T>
static int halo_push(const T* src, int k, int n_peers, const glab_push_desc* descs,
                     uint32_t* pushed_local, void* stream) {
  if (n_peers < 0 || n_peers > GLAB_MAX_PEERS || (n_peers > 0 && (!src || !descs))) return GLAB_E_ARG;
  if (n_peers == 0) {
    if (pushed_local) k_bump<<<1, 1, 0, as_stream(stream)>>>(pushed_local);
    return (int)cudaGetLastError();
  }
  PushArgs a;
  int64_t mx = 1;
  for (int q = 0; q < n_peers; ++q) {
    a.d[q] = descs[q];
    if (descs[q].count < 0 || (descs[q].count > 0 && ((!descs[q].send_idx && descs[q].first_row < 0) || !descs[q].dst)))
      return GLAB_E_ARG;
    if (descs[q].count > mx) mx = descs[q].count;
  }
  unsigned int* ctr = done_counters();
  if (!ctr) return GLAB_E_NOMEM;
  int64_t b = (mx + 255) / 256;
  if (b > 32) b = 32;
  dim3 grid((unsigned)b, (unsigned)n_peers);
  cudaStream_t st = as_stream(stream);
  switch (k) {
    case 1: k_halo_push<T, 1><<<grid, 256, 0, st>>>(src, a, ctr, pushed_local); break;
    case 2: k_halo_push<T, 2><<<grid, 256, 0, st>>>(src, a, ctr, pushed_local); break;
    case 4: k_halo_push<T, 4><<<grid, 256, 0, st>>>(src, a, ctr, pushed_local); break;
    case 8: k_halo_push<T, 8><<<grid, 256, 0, st>>>(src, a, ctr, pushed_local); break;
    default: return GLAB_E_ARG;
  }
  return (int)cudaGetLastError();
}

}  // namespace glab

using namespace glab;

extern "C" int glab_ipc_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int glab_ipc_alloc(int64_t bytes, void** dev_ptr, void* handle_out) {
  if (bytes <= 0 || !dev_ptr || !handle_out) return GLAB_E_ARG;
  void* p = nullptr;
  GLAB_CUDA(cudaMalloc(&p, (size_t)bytes));
  cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return (int)e;
  }
  memcpy(handle_out, &h, sizeof(h));
  *dev_ptr = p;
  return 0;
}

extern "C" int glab_ipc_open(const void* handle, void** peer_ptr) {
  if (!handle || !peer_ptr) return GLAB_E_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return GLAB_E_PEER;
  }
  return 0;
}

extern "C" int glab_ipc_close(void* peer_ptr) {
  if (!peer_ptr) return 0;
  GLAB_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return 0;
}

extern "C" int glab_ipc_free(void* dev_ptr) {
  if (!dev_ptr) return 0;
  GLAB_CUDA(cudaFree(dev_ptr));
  return 0;
}

extern "C" int glab_halo_push_f32(const float* src, int k, int n, const glab_push_desc* d,
                                  uint32_t* pushed, void* s) {
  return halo_push<float>(src, k, n, d, pushed, s);
}
extern "C" int glab_halo_push_f64(const double* src, int k, int n, const glab_push_desc* d,
                                  uint32_t* pushed, void* s) {
  return halo_push<double>(src, k, n, d, pushed, s);
}
extern "C" int glab_halo_wait(int n, uint32_t* const* flags, const uint32_t* pushed, void* s) {
  if (n < 0 || n > GLAB_MAX_PEERS || (n > 0 && (!flags || !pushed))) return GLAB_E_ARG;
  if (n == 0) return 0;
  WaitArgs a;
  for (int i = 0; i < n; ++i) {
    if (!flags[i]) return GLAB_E_ARG;
    a.flag[i] = flags[i];
  }
  k_halo_wait<<<1, 32, 0, as_stream(s)>>>(a, n, pushed);
  return (int)cudaGetLastError();
}
