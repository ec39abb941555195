// glab_common.cuh -- shared device/host helpers of libglab_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/glab.h"

#define GLAB_CUDA(expr)                                   \
  do {                                                    \
    cudaError_t _e = (expr);                              \
    if (_e != cudaSuccess) return (int)_e;                \
  } while (0)

// The CSR structure of one operator (or one row block of it) resident in HBM.
//   rowptr [n_rows+1] int32, colidx [nnz] int32 (+16 B pad), perm [nnz] int32 or NULL.
// coldelta [nnz] int16 (+16 B pad) or NULL: 2-byte column indices relative to the row, streamed
// by the pipeline kernels instead of colidx (halves the index traffic) for every 256-row tile
// whose flag in tile16 is set -- all tiles of a banded operator; all but the wrap-around / halo
// tiles of a periodic or row-partitioned one.
// perm[slot] = index of the caller's edge that landed in CSR slot `slot`; NULL when the
// caller's COO was already row-sorted (all reference generators emit it that way), in which
// case CSR slot order == edge order and per-edge arrays are used zero-copy.
struct glab_plan {
  int64_t n_rows, n_cols, nnz;
  int device;
  int sm_count;
  int32_t* rowptr;
  int32_t* colidx;
  int32_t* perm;
  int16_t* coldelta;       // colidx[slot] - row as int16 (valid in the row tiles flagged in tile16), or NULL
  uint8_t* tile16;         // [ceil(n_rows/256)]: 1 = every |col - row| of the tile's rows fits int16
  int64_t tiles16, tiles_total;
  int idx16_halo;          // 1 = the fused halo kernels stream coldelta too (GLAB_IDX16_HALO, default 1)
  int32_t max_row_nnz;
  int32_t band_local;      // max |col - row| over the columns < n_rows (the local block of a row partition)
  uint32_t* ms_state;      // multi-sweep kernels: [tiles_total] tile completion counters, then epoch / ticket words
  void* owned_vals;        // optional plan-owned copy of the values (glab_plan_adopt_vals_*), else NULL
  size_t owned_vals_bytes;
  void* retired[4];        // allocations replaced by glab_plan_adopt_vals_*, kept until the plan dies
};

namespace glab {

constexpr int kThreads = 256;          // threads per CTA of every row-tile kernel
constexpr int kMaxReduceBlocks = 4096; // upper bound on the persistent grid of reducing kernels
constexpr int kMaxDevices = 64;        // per-device caches of one-time kernel attributes

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- streaming 128-bit loads / stores (CSR arrays are read exactly once: keep them out of
// L1 so the gathered vector window stays resident there) -------------------------------------
__device__ __forceinline__ int4 ld_stream16(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

template <typename T> __device__ __forceinline__ T ldg(const T* p) { return __ldg(p); }

// K contiguous elements of T (K*sizeof(T) is 4,8,16,32 or 64 bytes, naturally aligned).
template <typename T, int K> struct Vec { T v[K]; };

template <typename T, int K>
__device__ __forceinline__ void load_vec(T (&dst)[K], const T* __restrict__ p) {
  constexpr int bytes = K * (int)sizeof(T);
  if constexpr (bytes >= 16) {
    constexpr int per = 16 / (int)sizeof(T);
#pragma unroll
    for (int i = 0; i < bytes / 16; ++i) {
      int4 q = __ldg(reinterpret_cast<const int4*>(p) + i);
      const T* t = reinterpret_cast<const T*>(&q);
#pragma unroll
      for (int j = 0; j < per; ++j) dst[i * per + j] = t[j];
    }
  } else if constexpr (bytes == 8) {
    int2 q = __ldg(reinterpret_cast<const int2*>(p));
    const T* t = reinterpret_cast<const T*>(&q);
#pragma unroll
    for (int j = 0; j < K; ++j) dst[j] = t[j];
  } else {
    dst[0] = __ldg(p);
  }
}

// Plain (coherent) variant for vectors that the same kernel also writes (r, x in cheby_next).
template <typename T, int K>
__device__ __forceinline__ void load_vec_rw(T (&dst)[K], const T* p) {
  constexpr int bytes = K * (int)sizeof(T);
  if constexpr (bytes >= 16) {
    constexpr int per = 16 / (int)sizeof(T);
#pragma unroll
    for (int i = 0; i < bytes / 16; ++i) {
      int4 q = *(reinterpret_cast<const int4*>(p) + i);
      const T* t = reinterpret_cast<const T*>(&q);
#pragma unroll
      for (int j = 0; j < per; ++j) dst[i * per + j] = t[j];
    }
  } else if constexpr (bytes == 8) {
    int2 q = *reinterpret_cast<const int2*>(p);
    const T* t = reinterpret_cast<const T*>(&q);
#pragma unroll
    for (int j = 0; j < K; ++j) dst[j] = t[j];
  } else {
    dst[0] = *p;
  }
}

template <typename T, int K>
__device__ __forceinline__ void store_vec(T* p, const T (&src)[K]) {
  constexpr int bytes = K * (int)sizeof(T);
  if constexpr (bytes >= 16) {
    constexpr int per = 16 / (int)sizeof(T);
#pragma unroll
    for (int i = 0; i < bytes / 16; ++i) {
      int4 q;
      T* t = reinterpret_cast<T*>(&q);
#pragma unroll
      for (int j = 0; j < per; ++j) t[j] = src[i * per + j];
      *(reinterpret_cast<int4*>(p) + i) = q;
    }
  } else if constexpr (bytes == 8) {
    int2 q;
    T* t = reinterpret_cast<T*>(&q);
#pragma unroll
    for (int j = 0; j < K; ++j) t[j] = src[j];
    *reinterpret_cast<int2*>(p) = q;
  } else {
    *p = src[0];
  }
}

// Cooperative coalesced copy of global elements [e_begin, e_end) of `g` into shared memory
// with 16-byte loads.  The global start address is rounded DOWN to 16 bytes, so element e lands
// at s[(e - e_begin) + off] with the returned `off` (< 16/sizeof(E)).  `s` must be 16-byte
// aligned and hold (e_end - e_begin) + 32/sizeof(E) elements.  Reads at most 15 bytes before
// g+e_begin and after g+e_end, always inside the same 16-byte granule, hence inside the
// allocation (cudaMalloc / torch blocks are >=256-byte granular).
template <typename E>
__device__ __forceinline__ int stage_to_smem(const E* __restrict__ g, int e_begin, int e_end, E* s,
                                             int tid, int nthreads) {
  const uintptr_t p0 = reinterpret_cast<uintptr_t>(g + e_begin);
  const uintptr_t a0 = p0 & ~static_cast<uintptr_t>(15);
  const int lead = static_cast<int>(p0 - a0);
  const int nvec = (lead + (e_end - e_begin) * (int)sizeof(E) + 15) >> 4;
  const int4* gv = reinterpret_cast<const int4*>(a0);
  int4* sv = reinterpret_cast<int4*>(s);
  for (int i = tid; i < nvec; i += nthreads) sv[i] = ld_stream16(gv + i);
  return lead / (int)sizeof(E);
}

// Copy this rank's boundary rows of `src` ([*, K]) into a neighbour's halo tail through the
// peer-mapped pointer of `d`, cooperatively by `nthreads` threads.  Contiguous send blocks go as
// 16-byte vectors with 4 independent loads in flight per thread; arbitrary row lists gather
// through send_idx.  L2-coherent loads: the rows were just written by other SMs of this GPU.
template <typename T, int K>
__device__ __forceinline__ void push_rows(const T* __restrict__ src, const glab_push_desc& d, int tid,
                                          int nthreads) {
  T* __restrict__ dst = reinterpret_cast<T*>(d.dst) + (size_t)d.dst_offset * K;
  if (d.first_row >= 0) {
    const T* __restrict__ s = src + (size_t)d.first_row * K;
    const int64_t total = d.count * K;
    constexpr int per = 16 / (int)sizeof(T);
    if ((reinterpret_cast<uintptr_t>(s) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      const int64_t nvec = total / per;
      const int4* sv = reinterpret_cast<const int4*>(s);
      int4* dv = reinterpret_cast<int4*>(dst);
      int64_t i = tid;
      for (; i + 3 * (int64_t)nthreads < nvec; i += 4 * (int64_t)nthreads) {
        const int4 a0 = __ldcg(sv + i), a1 = __ldcg(sv + i + nthreads), a2 = __ldcg(sv + i + 2 * nthreads),
                   a3 = __ldcg(sv + i + 3 * nthreads);
        dv[i] = a0;
        dv[i + nthreads] = a1;
        dv[i + 2 * nthreads] = a2;
        dv[i + 3 * nthreads] = a3;
      }
      for (; i < nvec; i += nthreads) dv[i] = __ldcg(sv + i);
      for (int64_t j = nvec * per + tid; j < total; j += nthreads) dst[j] = __ldcg(s + j);
    } else {
      for (int64_t j = tid; j < total; j += nthreads) dst[j] = __ldcg(s + j);
    }
    return;
  }
  for (int64_t i = tid; i < d.count; i += nthreads) {
    const size_t r = (size_t)__ldg(d.send_idx + i);
#pragma unroll
    for (int c = 0; c < K; ++c) dst[(size_t)i * K + c] = __ldcg(src + r * K + c);
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace glab
