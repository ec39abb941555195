// glab_amg.cu -- AMG setup kernels: classical / smoothed-aggregation strength of connection,
// direct-interpolation weights, and the per-edge message column.  All are row-local single
// passes over the CSR slots with bit-exact element-wise chains (IEEE div, no FMA contraction,
// the reference's operation order); citations are in include/glab.h.
#include <cstdlib>
#include "glab_pipe.cuh"

namespace glab {

// S_ij = relu(((-1*A_ij)/v_i) - theta),  v_i = max_k(-A_ik)  (SOCClassicGNN.py:69, :125)
template <typename T> struct OpSocClassic {
  static constexpr bool kReduce = true;
  static constexpr int kNarr = 1;
  static constexpr bool kNeedCol = false;
  T theta;
  T* rowmax;  // optional [n_rows]
  struct RowState { T v; bool any; };
  __device__ void begin_row(RowState& s, int) const { s.v = T(0); s.any = false; }
  __device__ void accumulate(RowState& s, T a, T, int) const {
    const T m = T(-1) * a;
    // amax semantics of scatter_reduce_: NaN propagates
    if (!s.any) { s.v = m; s.any = true; }
    else if (m > s.v || m != m) s.v = m;
  }
  __device__ void end_row(RowState& s, int r) const {
    if (!s.any) s.v = T(0);  // torch_scatter: rows without edges -> 0
    if (rowmax) rowmax[r] = s.v;
  }
  __device__ T edge(const RowState& s, T a, T, int) const {
    const T q = (T(-1) * a) / s.v;
    const T d = q - theta;
    return (d <= T(0)) ? T(0) : d;  // relu; NaN stays NaN like torch.relu
  }
};

// S_ij = (A_ij*A_ij)/(A_ii*A_jj)  (SOCSAGNN.py:67)
template <typename T> struct OpSocSA {
  static constexpr bool kReduce = false;
  static constexpr int kNarr = 1;
  static constexpr bool kNeedCol = true;
  const T* diag;
  struct RowState { T dii; };
  __device__ void begin_row(RowState& s, int r) const { s.dii = __ldg(diag + r); }
  __device__ void accumulate(RowState&, T, T, int) const {}
  __device__ void end_row(RowState&, int) const {}
  __device__ T edge(const RowState& s, T a, T, int col) const {
    return (a * a) / (s.dii * __ldg(diag + col));
  }
};

// DirectInterpGNN.py:89-94, :127, :150
template <typename T> struct OpDirectInterp {
  static constexpr bool kReduce = true;
  static constexpr int kNarr = 2;
  static constexpr bool kNeedCol = true;
  const T* diag;
  const T* cflag;
  struct RowState { T num, den, alpha, omc; };
  __device__ void begin_row(RowState& s, int) const { s.num = T(0); s.den = T(0); }
  __device__ void accumulate(RowState& s, T a, T sij, int col) const {
    s.num = s.num + a;
    s.den = s.den + (a * sij) * __ldg(cflag + col);
  }
  __device__ void end_row(RowState& s, int r) const {
    const T gamma = s.num / s.den;
    s.alpha = (T(1) / __ldg(diag + r)) * gamma;
    s.omc = T(1) - __ldg(cflag + r);
  }
  __device__ T edge(const RowState& s, T a, T, int) const { return s.omc * ((-a) * s.alpha); }
};

static inline int round_up128(int x) { return (x + 127) / 128 * 128; }

// TMA pipeline variant; returns -1000 when the operator does not fit (caller falls back).
template <typename T, class Op, bool IDX16>
static int launch_edge_pipe_impl(const glab_plan* p, const T* vals, const T* aux, const Op& op, T* out, void* stream);

// Operators that read the column index can stream 16-bit row-relative indices on every tile whose
// columns lie within +-32767 of their rows (plan: coldelta / tile16).  OFF by default (GLAB_EDGE_IDX16=1
// enables it): measured on D4096 fp32 (B200, profiles/r02_kernel_notes.md) soc_sa 0.3075 ms and
// direct_interp 0.3997 ms with 16-bit indices vs 0.3049 / 0.3848 ms with int32 -- these passes are bound by
// the dependent diag[col] gather and the IEEE divisions (issue slots), not by HBM bytes, so streaming
// 2 bytes per edge less buys nothing and the per-tile branch costs a little.
template <typename T, class Op>
static int launch_edge_pipe(const glab_plan* p, const T* vals, const T* aux, const Op& op, T* out,
                            void* stream) {
  if constexpr (Op::kNeedCol) {
    static const bool allow = [] { const char* e = getenv("GLAB_EDGE_IDX16"); return e && atoi(e) != 0; }();
    if (allow && p->coldelta && p->tile16 && p->tiles16 > 0)
      return launch_edge_pipe_impl<T, Op, true>(p, vals, aux, op, out, stream);
  }
  return launch_edge_pipe_impl<T, Op, false>(p, vals, aux, op, out, stream);
}

template <typename T, class Op, bool IDX16>
static int launch_edge_pipe_impl(const glab_plan* p, const T* vals, const T* aux, const Op& op, T* out,
                                 void* stream) {
  if (getenv("GLAB_PIPE") && atoi(getenv("GLAB_PIPE")) == 0) return -1000;
  if (reinterpret_cast<uintptr_t>(p->rowptr) & 15) return -1000;
  const int64_t slots = (int64_t)kThreads * (p->max_row_nnz > 0 ? p->max_row_nnz : 1);
  if (slots > 16384) return -1000;
  EdgePipeLayout L;
  int off = 0;
  L.off_row = off; off += round_up128((kThreads + 1) * 4 + 32);
  L.off_val = off; off += round_up128((int)slots * (int)sizeof(T) + 32);
  L.off_col = off; if (Op::kNeedCol) off += round_up128((int)slots * 4 + 32);
  L.off_aux = off; if (Op::kNarr > 1) off += round_up128((int)slots * (int)sizeof(T) + 32);
  L.stage_bytes = off;
  auto kern = k_edge_pipe<T, Op, IDX16>;
  static int max_smem_dev[kMaxDevices] = {};
  int& max_smem = max_smem_dev[p->device % kMaxDevices];
  if (!max_smem) {
    cudaFuncAttributes fa;
    GLAB_CUDA(cudaFuncGetAttributes(&fa, kern));
    const int m = 227 * 1024 - (int)fa.sharedSizeBytes;
    GLAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, m));
    max_smem = m;
  }
  if (2 * L.stage_bytes + 128 > max_smem) return -1000;
  int stages = (max_smem / 4 - 128) / L.stage_bytes;
  if (stages > 4) stages = 4;
  if (stages < 2) stages = 2;
  L.stages = stages;
  const size_t smem = (size_t)stages * L.stage_bytes + 128;
  int occ = 0;
  GLAB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kPipeThreads, smem));
  if (occ < 1) return -1000;
  const int ntiles = (int)((p->n_rows + kThreads - 1) / kThreads);
  int grid = p->sm_count * occ;
  if (grid > ntiles) grid = ntiles;
  TileArgs<T> a{p->rowptr, p->colidx, vals, 0, (int)p->n_rows, (int)slots, IDX16 ? p->coldelta : nullptr,
                IDX16 ? p->tile16 : nullptr};
  kern<<<grid, kPipeThreads, smem, as_stream(stream)>>>(a, aux ? aux : vals, p->perm, op, out, ntiles, L);
  return (int)cudaGetLastError();
}

template <typename T, class Op>
static int launch_edge_tiles(const glab_plan* p, const T* vals, const T* aux, const Op& op, T* out,
                             void* stream) {
  if (!p) return GLAB_E_ARG;
  if (p->nnz == 0 || p->n_rows == 0) return 0;  // no edges -> no per-edge outputs
  if (!out || !vals) return GLAB_E_ARG;
  {
    const int rc = launch_edge_pipe<T, Op>(p, vals, aux, op, out, stream);
    if (rc != -1000) return rc;
  }
  const int ntiles = (int)((p->n_rows + kThreads - 1) / kThreads);
  int64_t want = (int64_t)kThreads * (p->max_row_nnz > 0 ? p->max_row_nnz : 1);
  int cap = (int)(want < 4096 ? want : 4096);
  cap = (cap + 31) & ~31;
  const size_t smem = tile_smem_bytes<T>(cap, Op::kNarr);
  auto kern = k_edge_tiles<T, Op>;
  static bool attr_done[kMaxDevices] = {};
  if (!attr_done[p->device % kMaxDevices]) {
    GLAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    attr_done[p->device % kMaxDevices] = true;
  }
  TileArgs<T> a{p->rowptr, p->colidx, vals, 0, (int)p->n_rows, cap};
  kern<<<ntiles, kThreads, smem, as_stream(stream)>>>(a, aux, p->perm, op, out, ntiles);
  return (int)cudaGetLastError();
}

// c[e, column + c] = A_slot * x[col(slot), c]; with write_A also out[e, column - 1] = A_slot, so
// that the reference's returned edge_attr = cat([A_ij, c_ij], 1) is produced by ONE pass with
// one (1+K)-wide store per edge.
template <typename T, int K, bool WriteA>
__global__ void k_edge_messages(const int32_t* __restrict__ colidx, const T* __restrict__ vals,
                                const int32_t* __restrict__ perm, const T* __restrict__ x,
                                int64_t nnz, T* __restrict__ out, int64_t ld, int64_t column) {
  const bool packed = WriteA && K == 1 && ld == 2 && column == 1 &&
                      (reinterpret_cast<uintptr_t>(out) % (2 * sizeof(T)) == 0);
  int64_t done = 0;
  if constexpr (WriteA && K == 1) {
    // slot order == edge order (no permutation): four edges per thread -- 16-byte index / value loads,
    // four gathers in flight, one 8-element store -- so that the pass is limited by HBM, not by the
    // dependent load chain of one edge per thread
    constexpr int E = 4;
    if (packed && perm == nullptr && (reinterpret_cast<uintptr_t>(out) % 16 == 0) &&
        (reinterpret_cast<uintptr_t>(vals) % 16 == 0) && (reinterpret_cast<uintptr_t>(colidx) % 16 == 0)) {
      const int64_t groups = nnz / E;
      for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < groups;
           g += (int64_t)gridDim.x * blockDim.x) {
        int32_t c[E];
        T v[E], o[2 * E];
        load_vec<int32_t, E>(c, colidx + g * E);
        load_vec<T, E>(v, vals + g * E);
#pragma unroll
        for (int e = 0; e < E; ++e) o[2 * e + 1] = __ldg(x + c[e]);
#pragma unroll
        for (int e = 0; e < E; ++e) {
          o[2 * e] = v[e];
          o[2 * e + 1] = v[e] * o[2 * e + 1];
        }
        store_vec<T, 2 * E>(out + g * (2 * E), o);
      }
      done = groups * E;
    }
  }
  for (int64_t i = done + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int col = __ldg(colidx + i);
    const T v = __ldg(vals + i);
    T xv[K];
    load_vec<T, K>(xv, x + (size_t)col * K);
    const int64_t e = perm ? (int64_t)__ldg(perm + i) : i;
    if (packed) {
      T pair[2] = {v, v * xv[0]};
      store_vec<T, 2>(out + e * 2, pair);
    } else {
      if (WriteA) out[e * ld + column - 1] = v;
#pragma unroll
      for (int c = 0; c < K; ++c) out[e * ld + column + c] = v * xv[c];
    }
  }
}

template <typename T, bool WriteA>
static int edge_messages(const glab_plan* p, const T* vals, const T* x, int k, T* out, int64_t ld,
                         int64_t column, void* stream) {
  if (!p || ld < 1 || column < (WriteA ? 1 : 0) || column + k > ld) return GLAB_E_ARG;
  if (p->nnz == 0) return 0;
  if (!x || !out || !vals) return GLAB_E_ARG;
  int64_t b = (p->nnz + 255) / 256;
  const int64_t capb = (int64_t)p->sm_count * 32;
  const int grid = (int)(b < capb ? b : capb);
  cudaStream_t st = as_stream(stream);
  switch (k) {
    case 1: k_edge_messages<T, 1, WriteA><<<grid, 256, 0, st>>>(p->colidx, vals, p->perm, x, p->nnz, out, ld, column); break;
    case 2: k_edge_messages<T, 2, WriteA><<<grid, 256, 0, st>>>(p->colidx, vals, p->perm, x, p->nnz, out, ld, column); break;
    case 4: k_edge_messages<T, 4, WriteA><<<grid, 256, 0, st>>>(p->colidx, vals, p->perm, x, p->nnz, out, ld, column); break;
    case 8: k_edge_messages<T, 8, WriteA><<<grid, 256, 0, st>>>(p->colidx, vals, p->perm, x, p->nnz, out, ld, column); break;
    default: return GLAB_E_ARG;
  }
  return (int)cudaGetLastError();
}

// The bare seam: out[i,:] = reduce over the row's slots of src[slot,:]  (scatter sum / max).
// Mode: 0 = sum, 1 = max, 2 = min, 3 = mean (sum / max(count, 1)); empty rows give 0 in every
// mode, like torch_scatter (the 4-way aggregation of TrainableJacobiGNN.py:65-68).
template <typename T, int K, int Mode>
__global__ void k_segment_reduce(const int32_t* __restrict__ rowptr, const T* __restrict__ src,
                                 int64_t n_rows, T* __restrict__ out) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows;
       r += (int64_t)gridDim.x * blockDim.x) {
    const int rs = __ldg(rowptr + r), re = __ldg(rowptr + r + 1);
    T acc[K];
#pragma unroll
    for (int c = 0; c < K; ++c) acc[c] = T(0);
    for (int j = rs; j < re; ++j) {
      T v[K];
      load_vec<T, K>(v, src + (size_t)j * K);
#pragma unroll
      for (int c = 0; c < K; ++c) {
        if (Mode == 1) acc[c] = (j == rs || v[c] > acc[c] || v[c] != v[c]) ? v[c] : acc[c];
        else if (Mode == 2) acc[c] = (j == rs || v[c] < acc[c] || v[c] != v[c]) ? v[c] : acc[c];
        else acc[c] = acc[c] + v[c];
      }
    }
    if (Mode == 3) {
      const T cnt = (T)(re - rs > 0 ? re - rs : 1);
#pragma unroll
      for (int c = 0; c < K; ++c) acc[c] = acc[c] / cnt;
    }
    store_vec<T, K>(out + (size_t)r * K, acc);
  }
}

template <typename T, int Mode>
static int segment_reduce(const glab_plan* p, const T* src, int k, T* out, void* stream) {
  if (!p || !out || (p->nnz > 0 && !src)) return GLAB_E_ARG;
  if (p->n_rows == 0) return 0;
  int64_t b = (p->n_rows + 255) / 256;
  const int64_t capb = (int64_t)p->sm_count * 32;
  const int grid = (int)(b < capb ? b : capb);
  cudaStream_t st = as_stream(stream);
  switch (k) {
    case 1: k_segment_reduce<T, 1, Mode><<<grid, 256, 0, st>>>(p->rowptr, src, p->n_rows, out); break;
    case 2: k_segment_reduce<T, 2, Mode><<<grid, 256, 0, st>>>(p->rowptr, src, p->n_rows, out); break;
    case 4: k_segment_reduce<T, 4, Mode><<<grid, 256, 0, st>>>(p->rowptr, src, p->n_rows, out); break;
    case 8: k_segment_reduce<T, 8, Mode><<<grid, 256, 0, st>>>(p->rowptr, src, p->n_rows, out); break;
    default: return GLAB_E_ARG;
  }
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Fused 4-way aggregation  out[i, :] = [min | mean | sum | max] over the row's slots of src[slot, 0..F)
// in ONE pass over src (TrainableJacobiDiag/TrainableJacobiGNN.py:53-70 and
// DiffCoeffs/LearnDiffusionCoeffs.py:291-342 call torch_scatter.scatter four times on the same
// tensor; their `batch` vector -- vertex or edge -> graph id of a batch of small graphs -- is just
// another sorted index with LONG segments).  Two paths:
//   thread per (row, feature)   short rows (edge -> vertex aggregation): sequential in edge order, so
//                               sum / mean are bit-identical to scatter_add_ on the CPU;
//   warp per row                long rows (vertex / edge -> graph aggregation of batched graphs): lanes
//                               stride over the row (coalesced), shuffle reduction; min / max stay exact,
//                               sum / mean add in a fixed (launch-independent) but different order.
// amin / amax propagate NaN like torch; empty rows give 0 in every block, like torch_scatter.
// ------------------------------------------------------------------------------------------
template <typename T> struct Agg4 {
  T mn, mx, s;
  bool any;
  __device__ __forceinline__ void init() { mn = mx = s = T(0); any = false; }
  __device__ __forceinline__ void add(T v) {
    if (!any) { mn = mx = v; any = true; }
    else {
      if (v < mn || v != v) mn = (mn != mn) ? mn : v;
      if (v > mx || v != v) mx = (mx != mx) ? mx : v;
    }
    s = s + v;
  }
  __device__ __forceinline__ void merge(const Agg4& o) {   // min / max only (NaN sticks)
    if (!o.any) return;
    if (!any) { mn = o.mn; mx = o.mx; any = true; return; }
    if (o.mn < mn || o.mn != o.mn) mn = (mn != mn) ? mn : o.mn;
    if (o.mx > mx || o.mx != o.mx) mx = (mx != mx) ? mx : o.mx;
  }
};

template <typename T>
__global__ void k_segment_agg4_thread(const int32_t* __restrict__ rowptr, const T* __restrict__ src, int F,
                                      int64_t n_rows, T* __restrict__ out) {
  const int64_t total = n_rows * F;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / F;
    const int f = (int)(i - r * F);
    const int rs = __ldg(rowptr + r), re = __ldg(rowptr + r + 1);
    Agg4<T> a;
    a.init();
    for (int j = rs; j < re; ++j) a.add(__ldg(src + (size_t)j * F + f));
    const T cnt = (T)(re - rs > 0 ? re - rs : 1);
    T* o = out + (size_t)r * 4 * F;
    o[f] = a.mn;
    o[F + f] = a.s / cnt;
    o[2 * F + f] = a.s;
    o[3 * F + f] = a.mx;
  }
}

template <typename T>
__global__ void k_segment_agg4_warp(const int32_t* __restrict__ rowptr, const T* __restrict__ src, int F,
                                    int64_t n_rows, T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < n_rows; r += nwarps) {
    const int rs = __ldg(rowptr + r), re = __ldg(rowptr + r + 1);
    const int64_t e0 = (int64_t)rs * F, e1 = (int64_t)re * F;   // the row is one contiguous run of src
    for (int f = 0; f < F; ++f) {
      Agg4<T> a;
      a.init();
      double s = 0.0;    // lane-partial sums in fp64, combined in a fixed order
      for (int64_t e = e0 + f + (int64_t)lane * F; e < e1; e += 32 * (int64_t)F) {
        const T v = __ldg(src + e);
        a.add(v);
        s += (double)v;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Agg4<T> b;
        b.mn = __shfl_xor_sync(0xffffffffu, a.mn, o);
        b.mx = __shfl_xor_sync(0xffffffffu, a.mx, o);
        b.any = __shfl_xor_sync(0xffffffffu, (int)a.any, o) != 0;
        b.s = T(0);
        a.merge(b);
        s += __shfl_xor_sync(0xffffffffu, s, o);
      }
      if (lane == 0) {
        const T cnt = (T)(re - rs > 0 ? re - rs : 1);
        T* o = out + (size_t)r * 4 * F;
        const T sum = (T)s;
        o[f] = a.mn;
        o[F + f] = sum / cnt;
        o[2 * F + f] = sum;
        o[3 * F + f] = a.mx;
      }
    }
  }
}

template <typename T>
static int segment_agg4(const glab_plan* p, const T* src, int F, T* out, void* stream) {
  if (!p || !out || F < 1 || F > 64 || (p->nnz > 0 && !src)) return GLAB_E_ARG;
  if (p->n_rows == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const int64_t capb = (int64_t)p->sm_count * 32;
  if (p->max_row_nnz > 64) {
    int64_t b = (p->n_rows * 32 + 255) / 256;
    const int grid = (int)(b < capb ? b : capb);
    k_segment_agg4_warp<T><<<grid, 256, 0, st>>>(p->rowptr, src, F, p->n_rows, out);
  } else {
    int64_t b = (p->n_rows * F + 255) / 256;
    const int grid = (int)(b < capb ? b : capb);
    k_segment_agg4_thread<T><<<grid, 256, 0, st>>>(p->rowptr, src, F, p->n_rows, out);
  }
  return (int)cudaGetLastError();
}

}  // namespace glab

using namespace glab;

extern "C" int glab_segment_agg4_f32(const glab_plan* p, const float* src, int F, float* out, void* s) {
  return segment_agg4<float>(p, src, F, out, s);
}
extern "C" int glab_segment_agg4_f64(const glab_plan* p, const double* src, int F, double* out, void* s) {
  return segment_agg4<double>(p, src, F, out, s);
}

#define GLAB_AMG_INST(SUF, T)                                                                      \
  extern "C" int glab_soc_classic_##SUF(const glab_plan* p, const T* v, T theta, T* S, T* rowmax,  \
                                        void* s) {                                                 \
    OpSocClassic<T> op{theta, rowmax};                                                             \
    return launch_edge_tiles<T>(p, v, (const T*)nullptr, op, S, s);                                \
  }                                                                                                \
  extern "C" int glab_soc_sa_##SUF(const glab_plan* p, const T* v, const T* diag, T* S, void* s) { \
    if (p && p->nnz > 0 && !diag) return GLAB_E_ARG;                                               \
    OpSocSA<T> op{diag};                                                                           \
    return launch_edge_tiles<T>(p, v, (const T*)nullptr, op, S, s);                                \
  }                                                                                                \
  extern "C" int glab_direct_interp_##SUF(const glab_plan* p, const T* v, const T* S,              \
                                          const T* diag, const T* cflag, T* w, void* s) {          \
    if (p && p->nnz > 0 && (!diag || !cflag || !S)) return GLAB_E_ARG;                            \
    OpDirectInterp<T> op{diag, cflag};                                                             \
    return launch_edge_tiles<T>(p, v, S, op, w, s);                                                \
  }                                                                                                \
  extern "C" int glab_edge_messages_##SUF(const glab_plan* p, const T* v, const T* x, int k,       \
                                          T* out, int64_t ld, int64_t column, void* s) {           \
    return edge_messages<T, false>(p, v, x, k, out, ld, column, s);                                \
  }                                                                                                \
  extern "C" int glab_edge_attr_##SUF(const glab_plan* p, const T* v, const T* x, int k, T* out,   \
                                      int64_t ld, void* s) {                                       \
    return edge_messages<T, true>(p, v, x, k, out, ld, 1, s);                                      \
  }

#define GLAB_SEG_INST(SUF, T)                                                                      \
  extern "C" int glab_segment_sum_##SUF(const glab_plan* p, const T* src, int k, T* out, void* s) { \
    return segment_reduce<T, 0>(p, src, k, out, s);                                                \
  }                                                                                                \
  extern "C" int glab_segment_max_##SUF(const glab_plan* p, const T* src, T* out, void* s) {       \
    return segment_reduce<T, 1>(p, src, 1, out, s);                                                \
  }                                                                                                \
  extern "C" int glab_segment_min_##SUF(const glab_plan* p, const T* src, T* out, void* s) {       \
    return segment_reduce<T, 2>(p, src, 1, out, s);                                                \
  }                                                                                                \
  extern "C" int glab_segment_mean_##SUF(const glab_plan* p, const T* src, int k, T* out, void* s) { \
    return segment_reduce<T, 3>(p, src, k, out, s);                                                \
  }

GLAB_SEG_INST(f32, float)
GLAB_SEG_INST(f64, double)
GLAB_AMG_INST(f32, float)
GLAB_AMG_INST(f64, double)
