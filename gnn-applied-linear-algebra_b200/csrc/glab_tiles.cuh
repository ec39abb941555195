// glab_tiles.cuh -- the row-tile machinery shared by every fused layer kernel.
//
// HBM layout: int32 CSR (rowptr, colidx) + contiguous values, dense vectors row-major [n, K].
// One CTA of 256 threads owns a tile of R = 256*RPT consecutive rows.  The tile's CSR slots
// form ONE contiguous range [rowptr[r0], rowptr[r1]) of colidx / vals, which the CTA streams
// into shared memory with 16-byte no-allocate loads (fully coalesced, every byte of the CSR
// arrays crosses the memory system exactly once), in chunks of `cap` slots if the tile is
// larger than the staging buffer.  Then each thread walks ITS row's slots in shared memory --
// sequentially, in edge order, which is the accumulation order of scatter_add_ in the
// reference -- and gathers x[col] through the read-only path; for the stencil operators the
// gather window of a tile is a few KB and lives in L1/L2, so x costs one compulsory HBM read.
// The fused epilogue (Jacobi / Chebyshev / power method ...) runs on the row sum while it is
// still in registers and writes its result vectors coalesced.
#pragma once
#include "glab_common.cuh"


namespace glab {

// Epilogues with a CTA-wide prologue (run by every thread before the roles split) specialise this.
template <class E> struct has_prologue { static constexpr bool value = false; };


template <typename T> struct TileArgs {
  const int32_t* __restrict__ rowptr;
  const int32_t* __restrict__ colidx;
  const T* __restrict__ vals;
  int row_begin, row_end;
  int cap;  // staging capacity in CSR slots (multiple of 32)
  const int16_t* __restrict__ coldelta;  // col - row per slot (pipeline kernels with IDX != 0), else NULL
  const uint8_t* __restrict__ tile16;    // per 256-row tile: 1 = stream coldelta (IDX == 2), else NULL
};

__host__ __device__ inline size_t round16(size_t x) { return (x + 15) & ~(size_t)15; }

template <typename T> __host__ __device__ inline size_t tile_smem_bytes(int cap, int narr) {
  return round16((size_t)(cap + 8) * 4) + (size_t)narr * round16((size_t)(cap + 8) * sizeof(T));
}

// ------------------------------------------------------------------------------------------
// SpMV-bearing layers: acc[K] = sum_j A_ij * x[j, 0..K) per row, then Epi::row(r, acc).
// ------------------------------------------------------------------------------------------
template <typename T, int K, int RPT, class Epi>
__global__ void __launch_bounds__(kThreads)
k_row_tiles(TileArgs<T> a, const T* __restrict__ x, Epi epi, int ntiles) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int32_t* scol = reinterpret_cast<int32_t*>(smem_raw);
  T* sval = reinterpret_cast<T*>(smem_raw + round16((size_t)(a.cap + 8) * 4));
  constexpr int R = kThreads * RPT;
  const int tid = threadIdx.x;
  typename Epi::State st;
  if constexpr (has_prologue<Epi>::value) epi.prologue(st);
  epi.init(st);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r0 = a.row_begin + tile * R;
    const int r1 = min(r0 + R, a.row_end);
    int rs[RPT], re[RPT];
    T acc[RPT][K];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int r = r0 + i * kThreads + tid;
      rs[i] = re[i] = 0;
      if (r < r1) {
        rs[i] = __ldg(a.rowptr + r);
        re[i] = __ldg(a.rowptr + r + 1);
      }
#pragma unroll
      for (int c = 0; c < K; ++c) acc[i][c] = T(0);
    }
    const int e0 = __ldg(a.rowptr + r0);
    const int e1 = __ldg(a.rowptr + r1);
    for (int c0 = e0; c0 < e1; c0 += a.cap) {
      const int c1 = min(c0 + a.cap, e1);
      __syncthreads();  // everyone is done reading the previous chunk / tile
      const int offc = stage_to_smem<int32_t>(a.colidx, c0, c1, scol, tid, kThreads) - c0;
      const int offv = stage_to_smem<T>(a.vals, c0, c1, sval, tid, kThreads) - c0;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int lo = max(rs[i], c0), hi = min(re[i], c1);
#pragma unroll 4
        for (int j = lo; j < hi; ++j) {
          const int col = scol[j + offc];
          const T v = sval[j + offv];
          T xv[K];
          load_vec<T, K>(xv, x + (size_t)col * K);
#pragma unroll
          for (int c = 0; c < K; ++c) acc[i][c] = acc[i][c] + v * xv[c];  // -fmad=false: mul, add
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int r = r0 + i * kThreads + tid;
      if (r < r1) epi.row(st, r, acc[i]);
    }
  }
  epi.finish(st);
}

// ------------------------------------------------------------------------------------------
// Deterministic grid reduction of up to two fp64 partial sums per thread:
// warp shuffle -> CTA -> per-CTA slot in `ws` -> the last CTA to arrive (ticket) adds the
// slots in a fixed order and writes out[0..1].  ws layout: [0] ticket (uint32), doubles from
// byte 64: partial[2*cta + {0,1}].
// ------------------------------------------------------------------------------------------
// Device-side copy of glab_peer_reduce (passed by value inside the epilogue).
struct PeerReduceDev {
  int world, rank;
  double* mail_local;
  uint32_t* flag_local;
  double* mail_peer[GLAB_MAX_PEERS];
  uint32_t* flag_peer[GLAB_MAX_PEERS];
  uint32_t* parity_counter;
  uint32_t* status;
  unsigned long long timeout_ns;
};

__device__ __forceinline__ void grid_reduce2(double s0, double s1, void* ws, double* out,
                                             const PeerReduceDev* pr = nullptr) {
  __shared__ double red[2][16];  // up to 16 warps per CTA
  __shared__ bool is_last;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(ws);
  double* partial = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(ws) + 64);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  if (lane == 0) { red[0][w] = s0; red[1][w] = s1; }
  __syncthreads();
  if (tid == 0) {
    double t0 = 0, t1 = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { t0 += red[0][i]; t1 += red[1][i]; }
    partial[2 * blockIdx.x] = t0;
    partial[2 * blockIdx.x + 1] = t1;
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double t0 = 0, t1 = 0;
    for (int i = tid; i < (int)gridDim.x; i += (int)blockDim.x) {
      t0 += __ldcg(partial + 2 * i);
      t1 += __ldcg(partial + 2 * i + 1);
    }
    t0 = warp_sum(t0);
    t1 = warp_sum(t1);
    __syncthreads();
    if (lane == 0) { red[0][w] = t0; red[1][w] = t1; }
    __syncthreads();
    if (tid == 0) {
      double u0 = 0, u1 = 0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { u0 += red[0][i]; u1 += red[1][i]; }
      out[0] = u0;
      out[1] = u1;
      *ticket = 0u;  // re-arm for the next stream-ordered call
      if (pr != nullptr && pr->world > 0) {
        // publish this rank's partial sums into every rank's mailbox (peer stores), then bump the
        // arrival counters with a system-scope release
        const uint32_t par = *pr->parity_counter & 1u;
        for (int q = 0; q < pr->world; ++q) {
          double* m = pr->mail_peer[q] + ((size_t)par * GLAB_MAX_PEERS + pr->rank) * 2;
          m[0] = u0;
          m[1] = u1;
        }
        __threadfence_system();
        for (int q = 0; q < pr->world; ++q)
          asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(pr->flag_peer[q] + 4 * pr->rank) : "memory");
        *pr->parity_counter += 1u;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Per-edge-output tiles (AMG setup): phase 1 reduces over each row (optional), phase 2
// computes one value per CSR slot INTO the staged value buffer, and the CTA then writes the
// chunk out coalesced, in the caller's edge order (out[perm[slot]]).
//   Op::kReduce        -- whether phase 1 is needed
//   Op::kNarr          -- number of per-slot value arrays staged (1: vals, 2: vals + aux)
//   Op::RowState, op.begin_row(st, r), op.accumulate(st, v, aux, col),
//   op.end_row(st, r), op.edge(st, v, aux, col) -> T
// ------------------------------------------------------------------------------------------
template <typename T, class Op>
__global__ void __launch_bounds__(kThreads)
k_edge_tiles(TileArgs<T> a, const T* __restrict__ aux, const int32_t* __restrict__ perm, Op op,
             T* __restrict__ out, int ntiles) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int32_t* scol = reinterpret_cast<int32_t*>(smem_raw);
  T* sval = reinterpret_cast<T*>(smem_raw + round16((size_t)(a.cap + 8) * 4));
  T* saux = sval + round16((size_t)(a.cap + 8) * sizeof(T)) / sizeof(T);
  const int tid = threadIdx.x;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r0 = a.row_begin + tile * kThreads;
    const int r1 = min(r0 + kThreads, a.row_end);
    const int r = r0 + tid;
    int rs = 0, re = 0;
    if (r < r1) {
      rs = __ldg(a.rowptr + r);
      re = __ldg(a.rowptr + r + 1);
    }
    const int e0 = __ldg(a.rowptr + r0);
    const int e1 = __ldg(a.rowptr + r1);
    const bool single = (e1 - e0) <= a.cap;
    typename Op::RowState st;
    op.begin_row(st, r < r1 ? r : r0);
    int offc = 0, offv = 0, offa = 0;
    if (Op::kReduce) {
      for (int c0 = e0; c0 < e1; c0 += a.cap) {
        const int c1 = min(c0 + a.cap, e1);
        __syncthreads();
        offc = stage_to_smem<int32_t>(a.colidx, c0, c1, scol, tid, kThreads) - c0;
        offv = stage_to_smem<T>(a.vals, c0, c1, sval, tid, kThreads) - c0;
        if (Op::kNarr > 1) offa = stage_to_smem<T>(aux, c0, c1, saux, tid, kThreads) - c0;
        __syncthreads();
        const int lo = max(rs, c0), hi = min(re, c1);
        for (int j = lo; j < hi; ++j)
          op.accumulate(st, sval[j + offv], Op::kNarr > 1 ? saux[j + offa] : T(0), scol[j + offc]);
      }
    }
    op.end_row(st, r < r1 ? r : r0);
    for (int c0 = e0; c0 < e1; c0 += a.cap) {
      const int c1 = min(c0 + a.cap, e1);
      if (!(Op::kReduce && single)) {
        __syncthreads();
        offc = stage_to_smem<int32_t>(a.colidx, c0, c1, scol, tid, kThreads) - c0;
        offv = stage_to_smem<T>(a.vals, c0, c1, sval, tid, kThreads) - c0;
        if (Op::kNarr > 1) offa = stage_to_smem<T>(aux, c0, c1, saux, tid, kThreads) - c0;
        __syncthreads();
      }
      const int lo = max(rs, c0), hi = min(re, c1);
      for (int j = lo; j < hi; ++j)
        sval[j + offv] =
            op.edge(st, sval[j + offv], Op::kNarr > 1 ? saux[j + offa] : T(0), scol[j + offc]);
      __syncthreads();
      if (perm == nullptr) {
        for (int j = c0 + tid; j < c1; j += kThreads) out[j] = sval[j + offv];
      } else {
        for (int j = c0 + tid; j < c1; j += kThreads) out[__ldg(perm + j)] = sval[j + offv];
      }
    }
    __syncthreads();
  }
}

}  // namespace glab
