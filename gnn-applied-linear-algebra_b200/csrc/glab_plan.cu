// glab_plan.cu -- int64 COO -> int32 CSR plan (+ stable permutation) on the device.
//
// Replaces the grouping that torch_scatter.scatter(index=edgeij_pair[0]) performs implicitly on
// every call (reference call sites listed in include/glab.h).  The plan is built once per
// operator and reused by every layer step.
//
// Fast path (all reference generators emit row-sorted COO, UtilsGNN.py:53-67,74-78): one pass
// that validates indices and detects sortedness, one pass that writes rowptr from the row
// boundaries (no histogram, no scan), one pass that narrows col to int32.  perm stays NULL.
// Slow path (arbitrary edge order): stable LSD radix sort of (row, edge id) with CUB -- setup
// only, never on a layer step -- then the same boundary pass.
#include <cub/device/device_radix_sort.cuh>
#include <cstdlib>
#include <cstring>
#include <new>
#include "glab_common.cuh"

namespace glab {

// flags[0] bit0 = rows not sorted, bit1 = index out of range
__global__ void k_validate(const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                           int64_t nnz, int64_t n_rows, int64_t n_cols, int* flags) {
  int f = 0;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = row[e], c = col[e];
    if (r < 0 || r >= n_rows || c < 0 || c >= n_cols) f |= 2;
    if (e > 0 && row[e - 1] > r) f |= 1;
  }
  if (f) atomicOr(flags, f);
}

// rows sorted ascending: rowptr[r] = first slot whose row is >= r.
template <typename RowT>
__global__ void k_rowptr_from_sorted(const RowT* __restrict__ row, int64_t nnz, int64_t n_rows,
                                     int32_t* __restrict__ rowptr) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e <= nnz;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t prev = (e == 0) ? -1 : (int64_t)row[e - 1];
    const int64_t cur = (e == nnz) ? n_rows : (int64_t)row[e];
    for (int64_t r = prev + 1; r <= cur; ++r) rowptr[r] = (int32_t)e;
  }
}

__global__ void k_narrow(const int64_t* __restrict__ src, int64_t n, int32_t* __restrict__ dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = (int32_t)src[i];
}

__global__ void k_iota_rows(const int64_t* __restrict__ row, int64_t n, int32_t* __restrict__ keys,
                            int32_t* __restrict__ ids) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    keys[i] = (int32_t)row[i];
    ids[i] = (int32_t)i;
  }
}

__global__ void k_permute_cols(const int64_t* __restrict__ col, const int32_t* __restrict__ perm,
                               int64_t n, int32_t* __restrict__ dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = (int32_t)col[perm[i]];
}

__global__ void k_max_row(const int32_t* __restrict__ rowptr, int64_t n_rows, int* out) {
  int m = 0;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows;
       r += (int64_t)gridDim.x * blockDim.x)
    m = max(m, rowptr[r + 1] - rowptr[r]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

// max |col - row| over the local columns (col < n_rows): the dependency band of the multi-sweep kernels
__global__ void k_band(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int64_t n_rows,
                       int* out) {
  int m = 0;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows;
       r += (int64_t)gridDim.x * blockDim.x) {
    const int e1 = rowptr[r + 1];
    for (int e = rowptr[r]; e < e1; ++e) {
      const int64_t c = colidx[e];
      if (c < n_rows) {
        const int64_t d = c > r ? c - r : r - c;
        m = max(m, (int)d);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

// delta[slot] = col - row as int16; a row with a delta that does not fit clears its tile's flag
__global__ void k_col_delta(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                            int64_t n_rows, int16_t* __restrict__ delta, uint8_t* __restrict__ tile16) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows;
       r += (int64_t)gridDim.x * blockDim.x) {
    bool bad = false;
    const int e1 = rowptr[r + 1];
    for (int e = rowptr[r]; e < e1; ++e) {
      const int64_t d = (int64_t)colidx[e] - r;
      bad |= (d < -32767 || d > 32767);
      delta[e] = (int16_t)d;
    }
    if (bad) tile16[r / kThreads] = 0;
  }
}

__global__ void k_validate_csr(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                               int64_t n_rows, int64_t n_cols, int64_t nnz, int* flags) {
  int f = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz;
       i += (int64_t)gridDim.x * blockDim.x)
    if (col[i] < 0 || col[i] >= n_cols) f |= 2;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows;
       r += (int64_t)gridDim.x * blockDim.x)
    if (rowptr[r + 1] < rowptr[r]) f |= 1;
  if (blockIdx.x == 0 && threadIdx.x == 0 && (rowptr[0] != 0 || rowptr[n_rows] != nnz)) f |= 1;
  if (f) atomicOr(flags, f);
}

template <typename T>
__global__ void k_gather_vals(const T* __restrict__ edge_attr, int64_t ld, int64_t column,
                              const int32_t* __restrict__ perm, int64_t nnz, T* __restrict__ vals) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = perm ? (int64_t)perm[i] : i;
    vals[i] = edge_attr[e * ld + column];
  }
}

template <typename T>
__global__ void k_scatter_edges(const T* __restrict__ in, const int32_t* __restrict__ perm,
                                int64_t nnz, T* __restrict__ out, int64_t ld, int64_t column) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = perm ? (int64_t)perm[i] : i;
    out[e * ld + column] = in[i];
  }
}

// ---- interleaved [n, ld] <-> dense column blocks (layout glue of the drop-in API) ------------
template <typename T> struct PackArgs {
  T* part[GLAB_MAX_PARTS];
  int width[GLAB_MAX_PARTS];
  int offset[GLAB_MAX_PARTS];
  int n_parts;
};

// One CTA moves a tile of 256 rows through shared memory so that both sides are coalesced:
// the interleaved side is one contiguous run of 256*ld elements, each dense part a run of 256*w.
template <typename T, bool Pack>
__global__ void __launch_bounds__(256) k_pack_narrow(PackArgs<T> a, int64_t n, int ld, T* __restrict__ inter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);  // [256][ld]
  const int64_t r0 = (int64_t)blockIdx.x * 256;
  const int rows = (int)min((int64_t)256, n - r0);
  const int total = rows * ld;
  if (Pack) {
    for (int j = 0; j < a.n_parts; ++j) {
      const int w = a.width[j], off = a.offset[j];
      const T* __restrict__ src = a.part[j] + r0 * w;
      for (int i = threadIdx.x; i < rows * w; i += 256) tile[(i / w) * ld + off + (i % w)] = src[i];
    }
    __syncthreads();
    T* dst = inter + r0 * ld;
    for (int i = threadIdx.x; i < total; i += 256) dst[i] = tile[i];
  } else {
    const T* __restrict__ src = inter + r0 * ld;
    for (int i = threadIdx.x; i < total; i += 256) tile[i] = src[i];
    __syncthreads();
    for (int j = 0; j < a.n_parts; ++j) {
      const int w = a.width[j], off = a.offset[j];
      T* dst = a.part[j] + r0 * w;
      for (int i = threadIdx.x; i < rows * w; i += 256) dst[i] = tile[(i / w) * ld + off + (i % w)];
    }
  }
}

template <typename T, bool Pack>
static int pack_launch_narrow(const PackArgs<T>& a, int64_t n, int ld, T* inter, void* stream) {
  const int64_t blocks = (n + 255) / 256;
  if (blocks > INT32_MAX) return GLAB_E_RANGE;
  const size_t smem = (size_t)256 * ld * sizeof(T);
  auto kern = k_pack_narrow<T, Pack>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(a, n, (int)ld, inter);
  return (int)cudaGetLastError();
}

// One CTA moves a tile of kPackRows rows through shared memory so that both sides are coalesced: the
// interleaved side is one contiguous run of rows*ld elements, each dense part a run of rows*w.  Full,
// 16-byte-aligned runs move as 16-byte vectors with every load of a phase issued before the first
// store (a copy kernel lives on bytes in flight: 12-48 KB per CTA here).
constexpr int kPackRows = 1024;

template <typename T> __device__ __forceinline__ bool aligned16(const T* p) {
  return (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}

// global run -> shared run (count elements, both 16-byte aligned when vec)
template <typename T> __device__ __forceinline__ void run_to_smem(T* dst, const T* __restrict__ src, int count, bool vec) {
  constexpr int V = 16 / sizeof(T);
  if (vec) {
    const int nv = count / V;
    const uint4* __restrict__ s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    int i = threadIdx.x;
    for (; i + 3 * 256 < nv; i += 4 * 256) {
      const uint4 a = __ldcs(s4 + i), b = __ldcs(s4 + i + 256), c = __ldcs(s4 + i + 512), d = __ldcs(s4 + i + 768);
      d4[i] = a; d4[i + 256] = b; d4[i + 512] = c; d4[i + 768] = d;
    }
    for (; i < nv; i += 256) d4[i] = __ldcs(s4 + i);
    for (int j = nv * V + threadIdx.x; j < count; j += 256) dst[j] = src[j];
  } else {
    for (int i = threadIdx.x; i < count; i += 256) dst[i] = src[i];
  }
}
template <typename T> __device__ __forceinline__ void smem_to_run(T* __restrict__ dst, const T* src, int count, bool vec) {
  constexpr int V = 16 / sizeof(T);
  if (vec) {
    const int nv = count / V;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (int i = threadIdx.x; i < nv; i += 256) __stcs(d4 + i, s4[i]);
    for (int j = nv * V + threadIdx.x; j < count; j += 256) dst[j] = src[j];
  } else {
    for (int i = threadIdx.x; i < count; i += 256) dst[i] = src[i];
  }
}

// Shared layout: the interleaved tile [rows][ld] first, then one dense staging run per part, so that
// both global sides move as vectors and the transposition happens inside shared memory.
template <typename T, bool Pack>
__global__ void __launch_bounds__(256) k_pack(PackArgs<T> a, int64_t n, int ld, T* __restrict__ inter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);                 // [kPackRows][ld]
  T* stage = tile + (size_t)kPackRows * ld;                 // dense parts back to back: part j at kPackRows*offset[j]
  const int64_t r0 = (int64_t)blockIdx.x * kPackRows;
  const int rows = (int)min((int64_t)kPackRows, n - r0);
  const int total = rows * ld;
  T* gi = inter + r0 * ld;
  const bool vi = aligned16(gi);
  if (Pack) {
    for (int j = 0; j < a.n_parts; ++j) {
      const T* src = a.part[j] + r0 * a.width[j];
      run_to_smem(stage + (size_t)kPackRows * a.offset[j], src, rows * a.width[j], aligned16(src));
    }
    __syncthreads();
    for (int j = 0; j < a.n_parts; ++j) {
      const int w = a.width[j], off = a.offset[j];
      const T* st = stage + (size_t)kPackRows * off;
      if (w == 1) {
        for (int i = threadIdx.x; i < rows; i += 256) tile[i * ld + off] = st[i];
      } else {
        for (int i = threadIdx.x; i < rows * w; i += 256) tile[(i / w) * ld + off + (i % w)] = st[i];
      }
    }
    __syncthreads();
    smem_to_run(gi, tile, total, vi);
  } else {
    run_to_smem(tile, gi, total, vi);
    __syncthreads();
    for (int j = 0; j < a.n_parts; ++j) {
      const int w = a.width[j], off = a.offset[j];
      T* st = stage + (size_t)kPackRows * off;
      if (w == 1) {
        for (int i = threadIdx.x; i < rows; i += 256) st[i] = tile[i * ld + off];
      } else {
        for (int i = threadIdx.x; i < rows * w; i += 256) st[i] = tile[(i / w) * ld + off + (i % w)];
      }
    }
    __syncthreads();
    for (int j = 0; j < a.n_parts; ++j) {
      T* dst = a.part[j] + r0 * a.width[j];
      smem_to_run(dst, stage + (size_t)kPackRows * a.offset[j], rows * a.width[j], aligned16(dst));
    }
  }
}

// Single-column parts (k = 1: vertex_attr = [A_ii, b, x], [b, x], [b, x, r, p]): a register transpose, no
// shared memory, no barriers.  A thread owns R = 16 / sizeof(T) consecutive rows: one 16-byte vector per
// part on the dense side, LD vectors (R * LD contiguous elements) on the interleaved side -- both sides
// perfectly coalesced.  col[c] = the part that holds column c (NULL: column not wanted / not written).
template <typename T, int LD> struct ColArgs { T* col[LD]; };

template <typename T, int LD, bool Pack>
__global__ void __launch_bounds__(256) k_pack_cols(ColArgs<T, LD> a, int64_t n, T* __restrict__ inter) {
  constexpr int R = 16 / (int)sizeof(T);
  const int64_t groups = n / R;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    union { uint4 q[LD]; T e[R * LD]; } w;   // interleaved side: row-major [R][LD]
    if (Pack) {
#pragma unroll
      for (int c = 0; c < LD; ++c) {
        union { uint4 q; T e[R]; } v;
        v.q = __ldcs(reinterpret_cast<const uint4*>(a.col[c] + g * R));
#pragma unroll
        for (int r = 0; r < R; ++r) w.e[r * LD + c] = v.e[r];
      }
      uint4* dst = reinterpret_cast<uint4*>(inter + g * R * LD);
#pragma unroll
      for (int i = 0; i < LD; ++i) __stcs(dst + i, w.q[i]);
    } else {
      const uint4* src = reinterpret_cast<const uint4*>(inter + g * R * LD);
#pragma unroll
      for (int i = 0; i < LD; ++i) w.q[i] = __ldcs(src + i);
#pragma unroll
      for (int c = 0; c < LD; ++c) {
        if (a.col[c] == nullptr) continue;
        union { uint4 q; T e[R]; } v;
#pragma unroll
        for (int r = 0; r < R; ++r) v.e[r] = w.e[r * LD + c];
        __stcs(reinterpret_cast<uint4*>(a.col[c] + g * R), v.q);
      }
    }
  }
  // ragged tail: fewer than R rows
  if (blockIdx.x == 0) {
#pragma unroll
    for (int c = 0; c < LD; ++c)   // (static indices: the pointer array stays in the parameter space)
      if (threadIdx.x == c && a.col[c] != nullptr)
        for (int64_t r = groups * R; r < n; ++r) {
          if (Pack) inter[r * LD + c] = a.col[c][r];
          else a.col[c][r] = inter[r * LD + c];
        }
  }
}

// Returns -1000 when the call is not of the single-column shape (caller takes the tiled kernel).
template <typename T, int LD, bool Pack>
static int pack_cols_launch(int64_t n, int n_parts, T* const* parts, const int32_t* widths, const int32_t* offsets,
                            T* inter, void* stream) {
  ColArgs<T, LD> a;
  for (int c = 0; c < LD; ++c) a.col[c] = nullptr;
  if (reinterpret_cast<uintptr_t>(inter) & 15) return -1000;
  for (int j = 0; j < n_parts; ++j) {
    if (widths[j] != 1 || offsets[j] < 0 || offsets[j] >= LD || a.col[offsets[j]] != nullptr) return -1000;
    if (!parts[j] || (reinterpret_cast<uintptr_t>(parts[j]) & 15)) return -1000;
    a.col[offsets[j]] = parts[j];
  }
  if (Pack)
    for (int c = 0; c < LD; ++c)
      if (a.col[c] == nullptr) return -1000;   // a packed block has no holes
  constexpr int R = 16 / (int)sizeof(T);
  int64_t blocks = (n / R + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > (int64_t)148 * 64) blocks = (int64_t)148 * 64;
  k_pack_cols<T, LD, Pack><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(a, n, inter);
  return (int)cudaGetLastError();
}

template <typename T, bool Pack>
static int pack_launch(int64_t n, int64_t ld, int n_parts, T* const* parts, const int32_t* widths,
                       const int32_t* offsets, T* inter, void* stream) {
  if (n < 0 || ld < 1 || ld > 64 || n_parts < 0 || n_parts > GLAB_MAX_PARTS) return GLAB_E_ARG;
  if (n == 0 || n_parts == 0) return 0;
  if (!parts || !widths || !offsets || !inter) return GLAB_E_ARG;
  PackArgs<T> a;
  a.n_parts = n_parts;
  for (int j = 0; j < n_parts; ++j) {
    if (!parts[j] || widths[j] < 1 || offsets[j] < 0 || offsets[j] + widths[j] > ld) return GLAB_E_ARG;
    a.part[j] = parts[j];
    a.width[j] = widths[j];
    a.offset[j] = offsets[j];
  }
  if (ld >= 2 && ld <= 4) {
    int rc = -1000;
    switch ((int)ld) {
      case 2: rc = pack_cols_launch<T, 2, Pack>(n, n_parts, parts, widths, offsets, inter, stream); break;
      case 3: rc = pack_cols_launch<T, 3, Pack>(n, n_parts, parts, widths, offsets, inter, stream); break;
      default: rc = pack_cols_launch<T, 4, Pack>(n, n_parts, parts, widths, offsets, inter, stream); break;
    }
    if (rc != -1000) return rc;
  }
  // tile height: kPackRows while two copies of the tile fit 64 KB of shared memory, else fewer rows --
  // wide interleaved blocks (ld up to 64) keep the old 256-row tiles
  const size_t smem = (size_t)2 * kPackRows * ld * sizeof(T);
  if (smem > 96 * 1024) return pack_launch_narrow<T, Pack>(a, n, (int)ld, inter, stream);
  const int64_t blocks = (n + kPackRows - 1) / kPackRows;
  if (blocks > INT32_MAX) return GLAB_E_RANGE;
  auto kern = k_pack<T, Pack>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(a, n, (int)ld, inter);
  return (int)cudaGetLastError();
}

static int grid_for(int64_t n, int sm_count) {
  int64_t b = (n + 255) / 256;
  int64_t cap = (int64_t)sm_count * 16;
  if (b < 1) b = 1;
  return (int)(b < cap ? b : cap);
}

static void plan_free(glab_plan* p) {
  if (!p) return;
  if (p->rowptr) cudaFree(p->rowptr);
  if (p->colidx) cudaFree(p->colidx);
  if (p->perm) cudaFree(p->perm);
  if (p->coldelta) cudaFree(p->coldelta);
  if (p->tile16) cudaFree(p->tile16);
  if (p->owned_vals) cudaFree(p->owned_vals);
  if (p->ms_state) cudaFree(p->ms_state);
  for (void* q : p->retired)
    if (q) cudaFree(q);
  delete p;
}

static int plan_alloc(int64_t n_rows, int64_t n_cols, int64_t nnz, glab_plan** out) {
  if (!out || n_rows < 0 || n_cols < 0 || nnz < 0) return GLAB_E_ARG;
  if (nnz >= (int64_t)INT32_MAX - 64 || n_rows >= (int64_t)INT32_MAX - 64 ||
      n_cols >= (int64_t)INT32_MAX - 64)
    return GLAB_E_RANGE;
  glab_plan* p = new (std::nothrow) glab_plan();
  if (!p) return GLAB_E_NOMEM;
  p->n_rows = n_rows;
  p->n_cols = n_cols;
  p->nnz = nnz;
  p->rowptr = nullptr;
  p->colidx = nullptr;
  p->perm = nullptr;
  p->coldelta = nullptr;
  p->tile16 = nullptr;
  p->tiles16 = p->tiles_total = 0;
  p->idx16_halo = 0;
  p->max_row_nnz = 0;
  p->band_local = 0;
  p->ms_state = nullptr;
  p->owned_vals = nullptr;
  p->owned_vals_bytes = 0;
  for (void*& q : p->retired) q = nullptr;
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess)
    e = cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, p->device);
  if (e == cudaSuccess) e = cudaMalloc(&p->rowptr, (size_t)(n_rows + 1 + 8) * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&p->colidx, (size_t)(nnz + 8) * sizeof(int32_t));
  const size_t ms_words = (size_t)((n_rows + kThreads - 1) / kThreads) + 32;
  if (e == cudaSuccess) e = cudaMalloc(&p->ms_state, ms_words * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(p->ms_state, 0, ms_words * sizeof(uint32_t));
  if (e != cudaSuccess) {
    plan_free(p);
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? GLAB_E_NOMEM : (int)e;
  }
  *out = p;
  return 0;
}

// 2-byte row-relative column indices: the pipeline kernels stream 2 B instead of 4 B of index per
// nonzero in every 256-row tile whose deltas all fit int16.  GLAB_IDX16 = 0 disables it, 1 keeps it
// only for plans where EVERY tile qualifies (banded operators), 2 (default) also keeps mixed plans
// (periodic wrap-around rows, halo columns of a row block) when at least half of the tiles qualify.
// GLAB_IDX16_HALO = 0 keeps the fused multi-GPU halo kernels on int32 (default 1: they stream 2-byte
// indices in the qualifying tiles like the single-GPU kernels).
// Optional: on allocation failure the plan simply stays on int32.
static int plan_build_coldelta(glab_plan* p, cudaStream_t st) {
  const char* env = getenv("GLAB_IDX16");
  const int mode = env ? atoi(env) : 2;
  if (mode <= 0 || p->nnz == 0 || p->n_rows == 0) return 0;
  const int64_t ntiles = (p->n_rows + kThreads - 1) / kThreads;
  int16_t* d = nullptr;
  uint8_t* f = nullptr;
  if (cudaMalloc(&d, (size_t)(p->nnz + 16) * sizeof(int16_t)) != cudaSuccess ||
      cudaMalloc(&f, (size_t)ntiles + 16) != cudaSuccess) {
    if (d) cudaFree(d);
    cudaGetLastError();
    return 0;
  }
  uint8_t* h = new (std::nothrow) uint8_t[(size_t)ntiles];
  cudaError_t e = h ? cudaMemsetAsync(f, 1, (size_t)ntiles, st) : cudaErrorMemoryAllocation;
  if (e == cudaSuccess) {
    k_col_delta<<<grid_for(p->n_rows, p->sm_count), 256, 0, st>>>(p->rowptr, p->colidx, p->n_rows, d, f);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(h, f, (size_t)ntiles, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  int64_t good = 0;
  if (e == cudaSuccess)
    for (int64_t t = 0; t < ntiles; ++t) good += h[t] ? 1 : 0;
  delete[] h;
  const bool keep = e == cudaSuccess && (good == ntiles || (mode >= 2 && 2 * good >= ntiles && good > 0));
  if (!keep) {
    cudaFree(d);
    cudaFree(f);
    cudaGetLastError();
    return (e == cudaSuccess || e == cudaErrorMemoryAllocation) ? 0 : (int)e;
  }
  p->coldelta = d;
  p->tile16 = f;
  p->tiles16 = good;
  p->tiles_total = ntiles;
  const char* eh = getenv("GLAB_IDX16_HALO");
  p->idx16_halo = eh ? (atoi(eh) != 0) : 1;
  return 0;
}

}  // namespace glab

using namespace glab;

extern "C" int glab_version(void) { return GLAB_VERSION; }

// Mirrors make_pipe_layout / launch_pipe_halo_impl (glab_layers_impl.cuh) with the widest epilogue
// (six k-wide vertex streams) and a margin for the kernels' static shared memory.
extern "C" int glab_halo_fits(const glab_plan* p, int k, int elem_bytes) {
  if (!p || k < 1 || (elem_bytes != 4 && elem_bytes != 8)) return 0;
  auto up = [](int64_t v) { return (v + 127) / 128 * 128; };
  const int64_t slots = (int64_t)256 * (p->max_row_nnz > 0 ? p->max_row_nnz : 1);
  if (slots > 24576) return 0;
  const int64_t stage = up(257 * 4 + 32) + up(slots * 4 + 32) + up(slots * elem_bytes + 32) +
                        6 * up((int64_t)256 * k * elem_bytes + 32);
  return 2 * stage + 128 <= 227 * 1024 - 2048 ? 1 : 0;
}

extern "C" const char* glab_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case GLAB_E_ARG: return "glab: invalid argument";
    case GLAB_E_RANGE: return "glab: index out of range or nnz >= 2^31";
    case GLAB_E_NOMEM: return "glab: out of device memory";
    case GLAB_E_UNSORTED: return "glab: input not row sorted";
    case GLAB_E_PEER: return "glab: peer memory mapping failed";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "glab: unknown error";
  }
}

extern "C" int glab_plan_create(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* row,
                                const int64_t* col, void* stream_, glab_plan** out) {
  if (nnz > 0 && (!row || !col)) return GLAB_E_ARG;
  glab_plan* p = nullptr;
  int rc = plan_alloc(n_rows, n_cols, nnz, &p);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream_);
  int* d_flags = nullptr;  // [0] validation flags, [1] max row nnz, [2] local band
  int h_flags[3] = {0, 0, 0};
  int32_t* keys_in = nullptr;
  int32_t* keys_out = nullptr;
  int32_t* ids_in = nullptr;
  void* tmp = nullptr;
  cudaError_t e = cudaMalloc(&d_flags, 3 * sizeof(int));
  auto fail = [&](int code) {
    if (d_flags) cudaFree(d_flags);
    if (keys_in) cudaFree(keys_in);
    if (keys_out) cudaFree(keys_out);
    if (ids_in) cudaFree(ids_in);
    if (tmp) cudaFree(tmp);
    plan_free(p);
    cudaGetLastError();
    return code;
  };
  if (e != cudaSuccess) return fail((int)e);
  if ((e = cudaMemsetAsync(d_flags, 0, 3 * sizeof(int), st)) != cudaSuccess) return fail((int)e);
  const int g = grid_for(nnz, p->sm_count);
  if (nnz > 0) k_validate<<<g, 256, 0, st>>>(row, col, nnz, n_rows, n_cols, d_flags);
  if ((e = cudaMemcpyAsync(h_flags, d_flags, sizeof(int), cudaMemcpyDeviceToHost, st)) !=
      cudaSuccess)
    return fail((int)e);
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail((int)e);
  if (h_flags[0] & 2) return fail(GLAB_E_RANGE);

  if (!(h_flags[0] & 1)) {
    k_rowptr_from_sorted<int64_t><<<grid_for(nnz + 1, p->sm_count), 256, 0, st>>>(row, nnz, n_rows,
                                                                                 p->rowptr);
    if (nnz > 0) k_narrow<<<g, 256, 0, st>>>(col, nnz, p->colidx);
  } else {
    // stable sort of edge ids by row (LSD radix sort is stable)
    size_t tmp_bytes = 0;
    int bits = 1;
    while (((int64_t)1 << bits) < n_rows) ++bits;
    if ((e = cudaMalloc(&keys_in, (size_t)nnz * 4)) != cudaSuccess) return fail(GLAB_E_NOMEM);
    if ((e = cudaMalloc(&keys_out, (size_t)nnz * 4)) != cudaSuccess) return fail(GLAB_E_NOMEM);
    if ((e = cudaMalloc(&ids_in, (size_t)nnz * 4)) != cudaSuccess) return fail(GLAB_E_NOMEM);
    if ((e = cudaMalloc(&p->perm, (size_t)(nnz + 8) * 4)) != cudaSuccess) return fail(GLAB_E_NOMEM);
    k_iota_rows<<<g, 256, 0, st>>>(row, nnz, keys_in, ids_in);
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, ids_in, p->perm,
                                    (int)nnz, 0, bits, st);
    if ((e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16)) != cudaSuccess) return fail(GLAB_E_NOMEM);
    e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, ids_in, p->perm,
                                        (int)nnz, 0, bits, st);
    if (e != cudaSuccess) return fail((int)e);
    k_rowptr_from_sorted<int32_t><<<grid_for(nnz + 1, p->sm_count), 256, 0, st>>>(keys_out, nnz,
                                                                                 n_rows, p->rowptr);
    k_permute_cols<<<g, 256, 0, st>>>(col, p->perm, nnz, p->colidx);
  }
  if (n_rows > 0) {
    k_max_row<<<grid_for(n_rows, p->sm_count), 256, 0, st>>>(p->rowptr, n_rows, d_flags + 1);
    k_band<<<grid_for(n_rows, p->sm_count), 256, 0, st>>>(p->rowptr, p->colidx, n_rows, d_flags + 2);
  }
  if ((e = cudaMemcpyAsync(h_flags + 1, d_flags + 1, 2 * sizeof(int), cudaMemcpyDeviceToHost, st)) !=
      cudaSuccess)
    return fail((int)e);
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail((int)e);
  if ((e = cudaGetLastError()) != cudaSuccess) return fail((int)e);
  p->max_row_nnz = h_flags[1];
  p->band_local = h_flags[2];
  if (int rcd = plan_build_coldelta(p, st)) return fail(rcd);
  cudaFree(d_flags);
  if (keys_in) cudaFree(keys_in);
  if (keys_out) cudaFree(keys_out);
  if (ids_in) cudaFree(ids_in);
  if (tmp) cudaFree(tmp);
  *out = p;
  return 0;
}

extern "C" int glab_plan_create_csr(int64_t n_rows, int64_t n_cols, int64_t nnz,
                                    const int32_t* rowptr, const int32_t* colidx, void* stream_,
                                    glab_plan** out) {
  if (!rowptr || (nnz > 0 && !colidx)) return GLAB_E_ARG;
  glab_plan* p = nullptr;
  int rc = plan_alloc(n_rows, n_cols, nnz, &p);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream_);
  int* d_flags = nullptr;
  int h_flags[3] = {0, 0, 0};
  cudaError_t e = cudaMalloc(&d_flags, 3 * sizeof(int));
  auto fail = [&](int code) {
    if (d_flags) cudaFree(d_flags);
    plan_free(p);
    cudaGetLastError();
    return code;
  };
  if (e != cudaSuccess) return fail((int)e);
  cudaMemsetAsync(d_flags, 0, 3 * sizeof(int), st);
  cudaMemcpyAsync(p->rowptr, rowptr, (size_t)(n_rows + 1) * 4, cudaMemcpyDeviceToDevice, st);
  if (nnz > 0) cudaMemcpyAsync(p->colidx, colidx, (size_t)nnz * 4, cudaMemcpyDeviceToDevice, st);
  k_validate_csr<<<grid_for(nnz > n_rows ? nnz : n_rows, p->sm_count), 256, 0, st>>>(
      p->rowptr, p->colidx, n_rows, n_cols, nnz, d_flags);
  if (n_rows > 0) {
    k_max_row<<<grid_for(n_rows, p->sm_count), 256, 0, st>>>(p->rowptr, n_rows, d_flags + 1);
    k_band<<<grid_for(n_rows, p->sm_count), 256, 0, st>>>(p->rowptr, p->colidx, n_rows, d_flags + 2);
  }
  if ((e = cudaMemcpyAsync(h_flags, d_flags, 3 * sizeof(int), cudaMemcpyDeviceToHost, st)) !=
      cudaSuccess)
    return fail((int)e);
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail((int)e);
  if ((e = cudaGetLastError()) != cudaSuccess) return fail((int)e);
  if (h_flags[0]) return fail(GLAB_E_RANGE);
  p->max_row_nnz = h_flags[1];
  p->band_local = h_flags[2];
  if (int rcd = plan_build_coldelta(p, st)) return fail(rcd);
  cudaFree(d_flags);
  *out = p;
  return 0;
}

extern "C" int glab_plan_destroy(glab_plan* plan) {
  plan_free(plan);
  return 0;
}

extern "C" int glab_plan_info(const glab_plan* p, int64_t* n_rows, int64_t* n_cols, int64_t* nnz,
                              int32_t* max_row_nnz, int32_t* identity_perm) {
  if (!p) return GLAB_E_ARG;
  if (n_rows) *n_rows = p->n_rows;
  if (n_cols) *n_cols = p->n_cols;
  if (nnz) *nnz = p->nnz;
  if (max_row_nnz) *max_row_nnz = p->max_row_nnz;
  if (identity_perm) *identity_perm = p->perm ? 0 : 1;
  return 0;
}

extern "C" int glab_plan_index_width(const glab_plan* p, int32_t* bytes) {
  if (!p || !bytes) return GLAB_E_ARG;
  *bytes = (p->coldelta && p->tiles16 == p->tiles_total) ? 2 : 4;
  return 0;
}

extern "C" int glab_plan_index16_tiles(const glab_plan* p, int64_t* tiles16, int64_t* tiles_total) {
  if (!p || !tiles16 || !tiles_total) return GLAB_E_ARG;
  *tiles16 = p->coldelta ? p->tiles16 : 0;
  *tiles_total = (p->n_rows + kThreads - 1) / kThreads;
  return 0;
}

extern "C" int glab_plan_csr(const glab_plan* p, const int32_t** rowptr, const int32_t** colidx,
                             const int32_t** perm) {
  if (!p) return GLAB_E_ARG;
  if (rowptr) *rowptr = p->rowptr;
  if (colidx) *colidx = p->colidx;
  if (perm) *perm = p->perm;
  return 0;
}

template <typename T>
static int gather_vals(const glab_plan* p, const T* edge_attr, int64_t ld, int64_t column, T* vals,
                       void* stream) {
  if (!p || !vals || (p->nnz > 0 && !edge_attr) || ld < 1 || column < 0 || column >= ld)
    return GLAB_E_ARG;
  if (p->nnz == 0) return 0;
  k_gather_vals<T><<<grid_for(p->nnz, p->sm_count), 256, 0, as_stream(stream)>>>(
      edge_attr, ld, column, p->perm, p->nnz, vals);
  return (int)cudaGetLastError();
}

template <typename T>
static int scatter_edges(const glab_plan* p, const T* in, T* out, int64_t ld, int64_t column,
                         void* stream) {
  if (!p || (p->nnz > 0 && (!in || !out)) || ld < 1 || column < 0 || column >= ld)
    return GLAB_E_ARG;
  if (p->nnz == 0) return 0;
  k_scatter_edges<T><<<grid_for(p->nnz, p->sm_count), 256, 0, as_stream(stream)>>>(
      in, p->perm, p->nnz, out, ld, column);
  return (int)cudaGetLastError();
}

extern "C" int glab_gather_vals_f32(const glab_plan* p, const float* ea, int64_t ld, int64_t c,
                                    float* v, void* s) { return gather_vals<float>(p, ea, ld, c, v, s); }
extern "C" int glab_gather_vals_f64(const glab_plan* p, const double* ea, int64_t ld, int64_t c,
                                    double* v, void* s) { return gather_vals<double>(p, ea, ld, c, v, s); }
extern "C" int glab_scatter_edges_f32(const glab_plan* p, const float* in, float* out, int64_t ld,
                                      int64_t c, void* s) { return scatter_edges<float>(p, in, out, ld, c, s); }
extern "C" int glab_scatter_edges_f64(const glab_plan* p, const double* in, double* out, int64_t ld,
                                      int64_t c, void* s) { return scatter_edges<double>(p, in, out, ld, c, s); }

extern "C" int glab_pack_f32(int64_t n, int64_t ld, int np, const float* const* parts, const int32_t* w,
                             const int32_t* o, float* dst, void* s) {
  return pack_launch<float, true>(n, ld, np, const_cast<float* const*>(parts), w, o, dst, s);
}
extern "C" int glab_pack_f64(int64_t n, int64_t ld, int np, const double* const* parts, const int32_t* w,
                             const int32_t* o, double* dst, void* s) {
  return pack_launch<double, true>(n, ld, np, const_cast<double* const*>(parts), w, o, dst, s);
}
extern "C" int glab_unpack_f32(int64_t n, int64_t ld, int np, float* const* parts, const int32_t* w,
                               const int32_t* o, const float* src, void* s) {
  return pack_launch<float, false>(n, ld, np, parts, w, o, const_cast<float*>(src), s);
}
extern "C" int glab_unpack_f64(int64_t n, int64_t ld, int np, double* const* parts, const int32_t* w,
                               const int32_t* o, const double* src, void* s) {
  return pack_launch<double, false>(n, ld, np, parts, w, o, const_cast<double*>(src), s);
}

// ---- L2 residency (optional) ------------------------------------------------------------------
template <typename T>
static int adopt_vals(glab_plan* p, const T* vals, const T** out, void* stream) {
  if (!p || !out || (p->nnz > 0 && !vals)) return GLAB_E_ARG;
  // one allocation [colidx | vals] so that a single access-policy window covers the operator
  const size_t col_bytes = ((size_t)(p->nnz + 8) * 4 + 255) & ~(size_t)255;
  const size_t val_bytes = (size_t)(p->nnz + 8) * sizeof(T);
  void* buf = nullptr;
  cudaError_t e = cudaMalloc(&buf, col_bytes + val_bytes);
  if (e != cudaSuccess) { cudaGetLastError(); return GLAB_E_NOMEM; }
  // The replaced allocation may still be referenced (glab_plan_csr pointers handed out earlier, an
  // earlier adoption's values): it is retired, not freed, until the plan is destroyed.
  int slot = -1;
  for (int i = 0; i < 4; ++i)
    if (!p->retired[i]) { slot = i; break; }
  if (slot < 0) { cudaFree(buf); return GLAB_E_ARG; }   // adopted too often
  cudaStream_t st = as_stream(stream);
  e = cudaMemcpyAsync(buf, p->colidx, (size_t)p->nnz * 4, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync((char*)buf + col_bytes, vals, (size_t)p->nnz * sizeof(T), cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { cudaFree(buf); return (int)e; }
  p->retired[slot] = p->colidx;
  p->colidx = reinterpret_cast<int32_t*>(buf);
  p->owned_vals = nullptr;  // the values live inside the colidx allocation: freed with it
  p->owned_vals_bytes = col_bytes + val_bytes;
  *out = reinterpret_cast<const T*>((char*)buf + col_bytes);
  return 0;
}

extern "C" int glab_plan_adopt_vals_f32(glab_plan* p, const float* v, const float** o, void* s) {
  return adopt_vals<float>(p, v, o, s);
}
extern "C" int glab_plan_adopt_vals_f64(glab_plan* p, const double* v, const double** o, void* s) {
  return adopt_vals<double>(p, v, o, s);
}

extern "C" int glab_plan_l2_persist(const glab_plan* p, int enable, void* stream) {
  if (!p) return GLAB_E_ARG;
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof(attr));
  if (enable) {
    int dev = p->device, max_window = 0, max_persist = 0;
    GLAB_CUDA(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev));
    GLAB_CUDA(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
    if (max_window <= 0 || max_persist <= 0) return 0;  // not supported: silently a no-op
    GLAB_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist));
    size_t bytes = p->owned_vals_bytes ? p->owned_vals_bytes : (size_t)p->nnz * 4;
    if (bytes > (size_t)max_window) bytes = (size_t)max_window;
    attr.accessPolicyWindow.base_ptr = p->colidx;
    attr.accessPolicyWindow.num_bytes = bytes;
    double ratio = (double)max_persist * 0.9 / (double)bytes;
    attr.accessPolicyWindow.hitRatio = (float)(ratio > 1.0 ? 1.0 : ratio);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  } else {
    attr.accessPolicyWindow.num_bytes = 0;
  }
  GLAB_CUDA(cudaStreamSetAttribute(as_stream(stream), cudaStreamAttributeAccessPolicyWindow, &attr));
  return 0;
}
