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

template <typename T, int K>
__global__ void k_halo_push(const T* __restrict__ src, const int32_t* __restrict__ send_idx,
                            int64_t count, T* __restrict__ dst, int64_t dst_offset,
                            uint32_t* flag, uint32_t flag_value, unsigned int* done_counter) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    T v[K];
    load_vec_rw<T, K>(v, src + (size_t)send_idx[i] * K);
    store_vec<T, K>(dst + (size_t)(dst_offset + i) * K, v);
  }
  if (flag == nullptr) return;
  // last CTA to finish publishes the flag after all peer stores are visible system-wide
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(done_counter, 1u);
    if (t == gridDim.x - 1) {
      *done_counter = 0u;
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(flag_value) : "memory");
    }
  }
}

__global__ void k_halo_wait(uint32_t* flag, uint32_t flag_value) {
  uint32_t v;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if ((int32_t)(v - flag_value) >= 0) break;
    __nanosleep(64);
  } while (true);
}

static unsigned int* done_counter_for_device() {
  // one 4-byte counter per process/device, zero-initialised; pushes on one stream are ordered.
  static unsigned int* ctr = nullptr;
  if (!ctr) {
    if (cudaMalloc(&ctr, 256) != cudaSuccess) return nullptr;
    cudaMemset(ctr, 0, 256);
  }
  return ctr;
}

template <typename T>
static int halo_push(const T* src, const int32_t* send_idx, int64_t count, int k, T* dst,
                     int64_t dst_offset, uint32_t* flag, uint32_t flag_value, void* stream) {
  if (count < 0 || (count > 0 && (!src || !send_idx || !dst))) return GLAB_E_ARG;
  unsigned int* ctr = done_counter_for_device();
  if (!ctr) return GLAB_E_NOMEM;
  int64_t b = (count + 255) / 256;
  if (b < 1) b = 1;
  if (b > 64) b = 64;
  cudaStream_t st = as_stream(stream);
  switch (k) {
    case 1: k_halo_push<T, 1><<<(int)b, 256, 0, st>>>(src, send_idx, count, dst, dst_offset, flag, flag_value, ctr); break;
    case 2: k_halo_push<T, 2><<<(int)b, 256, 0, st>>>(src, send_idx, count, dst, dst_offset, flag, flag_value, ctr); break;
    case 4: k_halo_push<T, 4><<<(int)b, 256, 0, st>>>(src, send_idx, count, dst, dst_offset, flag, flag_value, ctr); break;
    case 8: k_halo_push<T, 8><<<(int)b, 256, 0, st>>>(src, send_idx, count, dst, dst_offset, flag, flag_value, ctr); break;
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

extern "C" int glab_halo_push_f32(const float* src, const int32_t* idx, int64_t count, int k,
                                  float* dst, int64_t off, uint32_t* flag, uint32_t val, void* s) {
  return halo_push<float>(src, idx, count, k, dst, off, flag, val, s);
}
extern "C" int glab_halo_push_f64(const double* src, const int32_t* idx, int64_t count, int k,
                                  double* dst, int64_t off, uint32_t* flag, uint32_t val, void* s) {
  return halo_push<double>(src, idx, count, k, dst, off, flag, val, s);
}
extern "C" int glab_halo_wait(uint32_t* flag, uint32_t val, void* s) {
  if (!flag) return GLAB_E_ARG;
  k_halo_wait<<<1, 1, 0, as_stream(s)>>>(flag, val);
  return (int)cudaGetLastError();
}
