// glab_setup.cu -- AMG setup on the device: the glue either side of the SOC / direct-interpolation
// layers in the reference's two-grid cycle.
//
//   glab_interp_*    prolongator P = [I + W](:, C) assembled SPARSE, sorted COO, from the per-edge
//                    weights of DirectInterpGNN.  Replaces VCycle.py:126-137 (dense eye(n) + W,
//                    column slice, .to_sparse()).
//   glab_spgemm_*    C = X * Y for two CSR plans (row-local accumulation in shared memory, or
//                    expand - sort - compress for dense-ish rows), used twice for the Galerkin
//                    operator A_c = P^T (A P).  Replaces VCycle.py:209 (torch.sparse @).
//   glab_cf_split_*  PMIS coarse/fine splitting on the strength graph with integer keys
//                    (deterministic, order-independent).  Stands in for pyamg's CLJP call at
//                    VCycle.py:114 / DirectInterpGNN.py:194 (un-pinned third-party, absent).
//
// All of this is SETUP work (once per operator), HBM-bound integer/key traffic; the radix sort and
// the scans are CUB device primitives, everything else is hand-written.  Products that meet in one
// entry are summed SEQUENTIALLY in expansion order (X slot order, then Y slot order) on both SpGEMM
// paths, so the result is deterministic and independent of the path and of the launch geometry.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include "glab_common.cuh"

namespace glab {

static int grid1d(int64_t n, int sm_count, int threads = 256) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count * 32;
  if (b < 1) b = 1;
  return (int)(b < cap ? b : cap);
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static int bits_for(int64_t n) {  // bits needed to hold values in [0, n)
  int b = 1;
  while (((int64_t)1 << b) < n) ++b;
  return b;
}

// ============================================================================ prolongator assembly
// Entry (i, j) of W survives in P when j is a coarse point and the weight is not an exact zero
// (.to_sparse() of the reference drops exact zeros, NaN is kept: VCycle.py:129-137).
// mode 0: the Python reference as shipped -- coarse rows keep their (NaN) W entries.
// mode 1: the MATLAB twin -- coarse rows are identity rows (test_direct_interpolation.m:130-132).
template <typename T>
__device__ __forceinline__ bool keeps(T w, T cj) { return cj > T(0) && !(w == T(0)); }

template <typename T>
__global__ void k_interp_count(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                               const T* __restrict__ w, const T* __restrict__ cflag, int mode,
                               int64_t n, int32_t* __restrict__ cnt, int32_t* __restrict__ cid) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (i == n) {  // sentinel so that the exclusive scans deliver the totals in slot n
      cnt[n] = 0;
      cid[n] = 0;
      continue;
    }
    const bool coarse = cflag[i] > T(0);
    int c = coarse ? 1 : 0;
    if (!(coarse && mode == 1)) {
      const int e1 = rowptr[i + 1];
      for (int e = rowptr[i]; e < e1; ++e) c += keeps(w[e], cflag[colidx[e]]) ? 1 : 0;
    }
    cnt[i] = c;
    cid[i] = coarse ? 1 : 0;
  }
}

// Row i of P in ascending coarse-column order: the kept W entries (ascending when the row of A is,
// as in every coalesced operator) with the identity entry (i, cid[i]) merged in at its place.
template <typename T>
__global__ void k_interp_fill(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                              const T* __restrict__ w, const T* __restrict__ cflag, int mode,
                              int64_t n, const int32_t* __restrict__ prow,
                              const int32_t* __restrict__ cid, int64_t* __restrict__ out_row,
                              int64_t* __restrict__ out_col, T* __restrict__ out_val) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const bool coarse = cflag[i] > T(0);
    int64_t o = prow[i];
    bool placed = !coarse;
    if (!(coarse && mode == 1)) {
      const int e1 = rowptr[i + 1];
      for (int e = rowptr[i]; e < e1; ++e) {
        const int j = colidx[e];
        const T we = w[e];
        if (!keeps(we, cflag[j])) continue;
        if (!placed && j > i) {
          out_row[o] = i; out_col[o] = cid[i]; out_val[o] = T(1); ++o;
          placed = true;
        }
        out_row[o] = i; out_col[o] = cid[j]; out_val[o] = we; ++o;
      }
    }
    if (!placed) { out_row[o] = i; out_col[o] = cid[i]; out_val[o] = T(1); }
  }
}

static int64_t interp_workspace_bytes(int64_t n) {
  if (n < 0) return GLAB_E_ARG;
  size_t tb = 0;
  cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, tb, (int32_t*)nullptr, (int32_t*)nullptr, (int)(n + 1),
                                                (cudaStream_t)0);
  if (e != cudaSuccess) return -(int64_t)e;
  return (int64_t)align256(tb + 16);
}

template <typename T>
static int interp_count(const glab_plan* A, const T* w, const T* cflag, int mode, void* ws, int64_t ws_bytes,
                        int32_t* prow, int32_t* cid, int64_t* nnz_p, int64_t* n_coarse, void* stream_) {
  if (!A || !prow || !cid || !nnz_p || !n_coarse || !ws || (mode != 0 && mode != 1)) return GLAB_E_ARG;
  const int64_t n = A->n_rows;
  if ((n > 0 && !cflag) || (A->nnz > 0 && !w)) return GLAB_E_ARG;
  const int64_t need = interp_workspace_bytes(n);
  if (need < 0) return (int)-need;
  if (ws_bytes < need) return GLAB_E_ARG;
  cudaStream_t st = as_stream(stream_);
  k_interp_count<T><<<grid1d(n + 1, A->sm_count), 256, 0, st>>>(A->rowptr, A->colidx, w, cflag, mode, n,
                                                                prow, cid);
  GLAB_CUDA(cudaGetLastError());
  size_t tb = (size_t)ws_bytes;
  GLAB_CUDA(cub::DeviceScan::ExclusiveSum(ws, tb, prow, prow, (int)(n + 1), st));
  tb = (size_t)ws_bytes;
  GLAB_CUDA(cub::DeviceScan::ExclusiveSum(ws, tb, cid, cid, (int)(n + 1), st));
  int32_t totals[2] = {0, 0};
  GLAB_CUDA(cudaMemcpyAsync(&totals[0], prow + n, 4, cudaMemcpyDeviceToHost, st));
  GLAB_CUDA(cudaMemcpyAsync(&totals[1], cid + n, 4, cudaMemcpyDeviceToHost, st));
  GLAB_CUDA(cudaStreamSynchronize(st));
  *nnz_p = totals[0];
  *n_coarse = totals[1];
  return 0;
}

template <typename T>
static int interp_fill(const glab_plan* A, const T* w, const T* cflag, int mode, const int32_t* prow,
                       const int32_t* cid, int64_t* out_row, int64_t* out_col, T* out_val,
                       void* stream_) {
  if (!A || !prow || !cid || (mode != 0 && mode != 1)) return GLAB_E_ARG;
  const int64_t n = A->n_rows;
  if (n == 0) return 0;
  if (!cflag || (A->nnz > 0 && !w) || !out_row || !out_col || !out_val) return GLAB_E_ARG;
  k_interp_fill<T><<<grid1d(n, A->sm_count), 256, 0, as_stream(stream_)>>>(
      A->rowptr, A->colidx, w, cflag, mode, n, prow, cid, out_row, out_col, out_val);
  return (int)cudaGetLastError();
}

// ============================================================================ SpGEMM
// Z = X * Y.  Two paths that produce bit-identical results (each entry's products are added one
// at a time in expansion order: X slot order, then Y slot order, first product not added to 0):
//
//   row-local  one thread per output row keeps the row's distinct columns AND their sums sorted in a
//              private shared-memory strip (insert / accumulate) -- ONE pass: the finished strip is
//              parked at the row's product offset in the workspace (a row has at most as many
//              distinct columns as products) together with its length; the numeric call only
//              compacts the parked rows into the COO output.  The strip height is a template
//              parameter (16 / 32 / 64 distinct columns): shared memory per CTA, and with it the
//              number of rows in flight per SM, follows the operator instead of the worst case
//              (the kernel is a chain of dependent gathers -- rows in flight are its throughput).
//              Used whenever no row has more than kRowCap distinct columns (all stencil /
//              Galerkin operators).
//   ESC        expand - stable radix sort - compress through a caller workspace of ~24 B per
//              product; the general fallback (dense-ish rows).
//
// Workspace header: hdr[0] = number of distinct (row, col) keys (ESC), hdr[1] = sorted buffer
// selector (ESC), hdr[2] = row-local overflow flag, hdr[3] = path taken (1 = row-local, 2 = ESC).
// The row-local path computes the VALUES during the symbolic call: the numeric call must be given
// the same operands (include/glab.h says so).
constexpr int kRowCap = 64;       // distinct columns per output row on the row-local path
constexpr int kRowThreads = 128;  // rows (threads) per CTA on the row-local path

struct SpgemmWs {
  size_t hdr, rowoff, rowcnt, keys[2], vals[2], cub, total;
  size_t cub_bytes;
  bool esc;  // the workspace is large enough for the ESC path
};

struct HeadOp {  // p starts a run of equal keys
  const uint64_t* k;
  __host__ __device__ __forceinline__ bool operator()(int p) const { return p == 0 || k[p] != k[p - 1]; }
};

// A row with at most kRowCap products cannot have more than kRowCap distinct columns: then the
// row-local path is guaranteed and only the header + the row-offset array are needed.
template <typename T>
static int spgemm_layout(int64_t n_rows_x, int64_t n_products, int64_t max_row_products, SpgemmWs* L) {
  if (n_rows_x < 0 || n_products < 0 || max_row_products < 0) return GLAB_E_ARG;
  if (n_products >= (int64_t)INT32_MAX - 64) return GLAB_E_RANGE;
  L->esc = max_row_products > kRowCap;
  size_t scan_b = 0;
  GLAB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_b, (int64_t*)nullptr, (int64_t*)nullptr,
                                          (int)(n_rows_x + 1), (cudaStream_t)0));
  size_t cb = scan_b;
  const size_t P = (size_t)n_products + 2;          // parked rows (row-local) / expanded products (ESC)
  const size_t P2 = L->esc ? P : 0;                  // second buffers: ESC only
  if (L->esc) {
    size_t sort_b = 0, sel_b = 0;
    cub::DoubleBuffer<uint64_t> dk(nullptr, nullptr);
    cub::DoubleBuffer<T> dv(nullptr, nullptr);
    GLAB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_b, dk, dv, (int)n_products, 0, 64, (cudaStream_t)0));
    HeadOp op{nullptr};
    GLAB_CUDA(cub::DeviceSelect::If(nullptr, sel_b, thrust::counting_iterator<int>(0), (int32_t*)nullptr,
                                    (int64_t*)nullptr, (int)n_products, op, (cudaStream_t)0));
    if (sort_b > cb) cb = sort_b;
    if (sel_b > cb) cb = sel_b;
  }
  size_t o = 0;
  L->hdr = o;      o += 256;
  L->rowoff = o;   o += align256((size_t)(n_rows_x + 2) * 8);
  L->rowcnt = o;   o += align256((size_t)(n_rows_x + 2) * 4);
  L->keys[0] = o;  o += align256(P * (L->esc ? 8 : 4));   // row-local: int32 columns of the parked rows
  L->keys[1] = o;  o += align256(P2 * 8);
  L->vals[0] = o;  o += align256(P * sizeof(T));
  L->vals[1] = o;  o += align256(P2 * sizeof(T));
  L->cub = o;      o += align256(cb + 16);
  L->cub_bytes = cb;
  L->total = o;
  return 0;
}

// products of X row r = sum over its slots (r, j) of nnz(Y row j)
__global__ void k_spgemm_count(const int32_t* __restrict__ xrp, const int32_t* __restrict__ xci,
                               const int32_t* __restrict__ yrp, int64_t n, int64_t* __restrict__ cnt,
                               unsigned long long* __restrict__ total) {
  unsigned long long local = 0, biggest = 0;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r <= n;
       r += (int64_t)gridDim.x * blockDim.x) {
    int64_t c = 0;
    if (r < n) {
      const int e1 = xrp[r + 1];
      for (int e = xrp[r]; e < e1; ++e) {
        const int j = xci[e];
        c += yrp[j + 1] - yrp[j];
      }
    }
    if (cnt) cnt[r] = c;
    local += (unsigned long long)c;
    biggest = (unsigned long long)c > biggest ? (unsigned long long)c : biggest;
  }
  if (total) {  // total[0] = sum, total[1] = max over rows
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      local += __shfl_xor_sync(0xffffffffu, local, o);
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, biggest, o);
      biggest = other > biggest ? other : biggest;
    }
    if ((threadIdx.x & 31) == 0 && local) {
      atomicAdd(total, local);
      atomicMax(total + 1, biggest);
    }
  }
}

// ---- row-local path ---------------------------------------------------------------------------
// Thread `tid` owns strip element s at cols[s * kRowThreads + tid] (conflict-free when the threads
// of a warp touch the same s).  Columns arrive mostly ascending, so the insertion point is searched
// from the top.  The finished strip goes to park_col / park_val at the row's product offset.
template <typename T, int CAP>
__global__ void __launch_bounds__(kRowThreads)
k_spgemm_rows(const int32_t* __restrict__ xrp, const int32_t* __restrict__ xci, const T* __restrict__ xv,
              const int32_t* __restrict__ yrp, const int32_t* __restrict__ yci, const T* __restrict__ yv,
              int64_t n, const int64_t* __restrict__ poff, int32_t* __restrict__ rowcnt,
              int32_t* __restrict__ park_col, T* __restrict__ park_val, int64_t* __restrict__ hdr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int32_t* cols = reinterpret_cast<int32_t*>(smem_raw);                                  // [CAP][kRowThreads]
  T* vals = reinterpret_cast<T*>(smem_raw + (size_t)CAP * kRowThreads * sizeof(int32_t));    // same shape
  const int tid = threadIdx.x;
  for (int64_t r = blockIdx.x * (int64_t)kRowThreads + tid; r <= n; r += (int64_t)gridDim.x * kRowThreads) {
    if (r == n) {
      rowcnt[n] = 0;  // sentinel: the exclusive scan leaves nnz(Z) here
      continue;
    }
    int cnt = 0;
    bool over = false;
    const int e1 = __ldg(xrp + r + 1);
    // The row is a chain of dependent gathers (X slot -> Y row bounds -> Y entries); four X slots are
    // resolved together and the next Y entry is fetched while the current one is inserted, so that the
    // chain costs ~3 memory latencies per batch instead of 3 per slot.
    for (int eb = __ldg(xrp + r); eb < e1 && !over; eb += 4) {
      int y0[4], y1[4];
      T a[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const bool live = eb + q < e1;
        y0[q] = live ? __ldg(xci + eb + q) : -1;
        a[q] = live ? __ldg(xv + eb + q) : T(0);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = y0[q];
        y0[q] = j >= 0 ? __ldg(yrp + j) : 0;
        y1[q] = j >= 0 ? __ldg(yrp + j + 1) : 0;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int t = y0[q];
        const int tend = y1[q];
        int cn = 0;
        T vn = T(0);
        if (t < tend) {
          cn = __ldg(yci + t);
          vn = __ldg(yv + t);
        }
        while (t < tend && !over) {
          const int c = cn;
          const T prod = a[q] * vn;
          ++t;
          if (t < tend) {
            cn = __ldg(yci + t);
            vn = __ldg(yv + t);
          }
          int pos = cnt;
          while (pos > 0 && cols[(pos - 1) * kRowThreads + tid] >= c) --pos;
          if (pos < cnt && cols[pos * kRowThreads + tid] == c) {
            vals[pos * kRowThreads + tid] = vals[pos * kRowThreads + tid] + prod;
            continue;
          }
          if (cnt == CAP) {
            over = true;
            break;
          }
          for (int s = cnt; s > pos; --s) {
            cols[s * kRowThreads + tid] = cols[(s - 1) * kRowThreads + tid];
            vals[s * kRowThreads + tid] = vals[(s - 1) * kRowThreads + tid];
          }
          cols[pos * kRowThreads + tid] = c;
          vals[pos * kRowThreads + tid] = prod;
          ++cnt;
        }
      }
    }
    rowcnt[r] = cnt;
    if (over) {
      atomicOr(reinterpret_cast<unsigned long long*>(hdr + 2), 1ull);
    } else {
      int64_t o = poff[r];
      for (int s = 0; s < cnt; ++s, ++o) {
        park_col[o] = cols[s * kRowThreads + tid];
        park_val[o] = vals[s * kRowThreads + tid];
      }
    }
  }
}

template <typename T, int CAP>
static int spgemm_rows_launch(const glab_plan* X, const T* xv, const glab_plan* Y, const T* yv, const int64_t* poff,
                              int32_t* rowcnt, int32_t* park_col, T* park_val, int64_t* hdr, cudaStream_t st) {
  const size_t smem = (size_t)CAP * kRowThreads * (sizeof(int32_t) + sizeof(T));
  auto kern = k_spgemm_rows<T, CAP>;
  GLAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n = X->n_rows;
  int64_t blocks = (n + 1 + kRowThreads - 1) / kRowThreads;
  const int64_t cap = (int64_t)X->sm_count * 32;
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, kRowThreads, smem, st>>>(X->rowptr, X->colidx, xv, Y->rowptr, Y->colidx, yv, n, poff,
                                                     rowcnt, park_col, park_val, hdr);
  return (int)cudaGetLastError();
}

// ---- row-local path, CTA-cooperative variant ------------------------------------------------------
// The thread-per-row kernel above interleaves its gathers with the insertion: every load of a warp
// touches 32 different sectors and waits for the previous insertion.  Here a CTA of 128 threads takes
// 128 consecutive X rows and separates the two:
//   1. the rows' X slots are ONE contiguous range: loaded coalesced, with the Y row bounds of every slot;
//   2. a block scan of the Y row lengths gives every slot the offset of its products -- slot order, then Y
//      order: exactly the documented expansion order, and equal to poff[r] - poff[r0] for each row;
//   3. all threads expand all slots into a shared product list (column, a * y): independent loads, no
//      insertion in between, as many in flight as the LSU takes;
//   4. thread-per-row insertion sort IN PLACE inside the row's own segment of the list (a row has at most
//      as many distinct columns as products, and product t is read before anything is written at
//      index <= t): shared memory only, same sequential sums as the strip kernel -> same bits;
//   5. the whole list goes to the parking area at poff[r0] with coalesced stores (the entries behind a
//      row's distinct count are dead and never read).
// No strip height, no overflow: taken whenever 128 rows' slots and products fit kCoopSmem.
constexpr int kCoopRows = 128;
constexpr int kCoopSmem = 64 * 1024;

template <typename T> __host__ __device__ inline size_t coop_smem_bytes(int s_max, int p_max) {
  return (size_t)(p_max + s_max) * sizeof(T) + (size_t)(p_max + 2 * s_max + kCoopRows + 8 + 8) * 4;
}

template <typename T>
__global__ void __launch_bounds__(kCoopRows)
k_spgemm_rows_coop(const int32_t* __restrict__ xrp, const int32_t* __restrict__ xci, const T* __restrict__ xv,
                   const int32_t* __restrict__ yrp, const int32_t* __restrict__ yci, const T* __restrict__ yv,
                   int64_t n, const int64_t* __restrict__ poff, int32_t* __restrict__ rowcnt,
                   int32_t* __restrict__ park_col, T* __restrict__ park_val, int s_max, int p_max) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* pval = reinterpret_cast<T*>(smem_raw);          // [p_max] products: value
  T* sa = pval + p_max;                              // [s_max] X value of the slot
  int32_t* pcol = reinterpret_cast<int32_t*>(sa + s_max);   // [p_max] products: column
  int32_t* sy0 = pcol + p_max;                       // [s_max] first Y slot of the X slot's row
  int32_t* soff = sy0 + s_max;                       // [s_max + 1] Y row length, then exclusive product offset
  int32_t* rstart = soff + s_max + 4;                // [kCoopRows + 1] first product of each row
  int32_t* wsum = rstart + kCoopRows + 4;            // [4] warp totals of the scan
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (blockIdx.x == 0 && tid == 0) rowcnt[n] = 0;    // sentinel: the exclusive scan leaves nnz(Z) here
  for (int64_t r0 = (int64_t)blockIdx.x * kCoopRows; r0 < n; r0 += (int64_t)gridDim.x * kCoopRows) {
    const int rows = (int)min((int64_t)kCoopRows, n - r0);
    const int e0 = __ldg(xrp + r0), e1 = __ldg(xrp + r0 + rows);
    const int S = e1 - e0;
    // 1. X slots and the bounds of their Y rows
    for (int i = tid; i < S; i += kCoopRows) {
      const int j = __ldg(xci + e0 + i);
      const int y0 = __ldg(yrp + j);
      sy0[i] = y0;
      soff[i] = __ldg(yrp + j + 1) - y0;
      sa[i] = __ldg(xv + e0 + i);
    }
    __syncthreads();
    // 2. exclusive scan of the lengths (thread t owns a chunk of consecutive slots)
    const int chunk = (S + kCoopRows - 1) / kCoopRows;
    const int c0 = min(S, tid * chunk), c1 = min(S, c0 + chunk);
    int mine = 0;
    for (int i = c0; i < c1; ++i) mine += soff[i];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int base = incl - mine;
    for (int w = 0; w < warp; ++w) base += wsum[w];
    const int P = wsum[0] + wsum[1] + wsum[2] + wsum[3];
    for (int i = c0; i < c1; ++i) {
      const int len = soff[i];
      soff[i] = base;
      base += len;
    }
    if (tid == 0) soff[S] = P;
    __syncthreads();
    if (tid <= rows) rstart[tid] = tid < rows ? soff[__ldg(xrp + r0 + tid) - e0] : P;
    if (tid == 0 && rows == kCoopRows) rstart[kCoopRows] = P;
    // 3. expansion: slot order, then Y order
    for (int i = tid; i < S; i += kCoopRows) {
      const int o = soff[i], len = soff[i + 1] - o, y0 = sy0[i];
      const T a = sa[i];
      for (int t = 0; t < len; ++t) {
        pcol[o + t] = __ldg(yci + y0 + t);
        pval[o + t] = a * __ldg(yv + y0 + t);
      }
    }
    __syncthreads();
    // 4. in-place insertion inside the row's segment
    if (tid < rows) {
      const int b = rstart[tid], len = rstart[tid + 1] - b;
      int cnt = 0;
      for (int t = 0; t < len; ++t) {
        const int c = pcol[b + t];
        const T v = pval[b + t];
        int pos = cnt;
        while (pos > 0 && pcol[b + pos - 1] >= c) --pos;
        if (pos < cnt && pcol[b + pos] == c) {
          pval[b + pos] = pval[b + pos] + v;
          continue;
        }
        for (int q = cnt; q > pos; --q) {
          pcol[b + q] = pcol[b + q - 1];
          pval[b + q] = pval[b + q - 1];
        }
        pcol[b + pos] = c;
        pval[b + pos] = v;
        ++cnt;
      }
      rowcnt[r0 + tid] = cnt;
    }
    __syncthreads();
    // 5. park the list (row r's entries land at poff[r] + s because poff[r] - poff[r0] == rstart[r - r0])
    const int64_t pb = __ldg(poff + r0);
    for (int i = tid; i < P; i += kCoopRows) {
      park_col[pb + i] = pcol[i];
      park_val[pb + i] = pval[i];
    }
    __syncthreads();
  }
}

// Returns -1000 if 128 rows' slots / products do not fit the shared-memory budget.
template <typename T>
static int spgemm_rows_coop_launch(const glab_plan* X, const T* xv, const glab_plan* Y, const T* yv,
                                   const int64_t* poff, int32_t* rowcnt, int32_t* park_col, T* park_val,
                                   int64_t max_row_products, cudaStream_t st) {
  const int64_t s_need = (int64_t)kCoopRows * (X->max_row_nnz > 0 ? X->max_row_nnz : 1);
  const int64_t p_need = (int64_t)kCoopRows * (max_row_products > 0 ? max_row_products : 1);
  if (s_need > (1 << 20) || p_need > (1 << 20)) return -1000;
  const int s_max = (int)((s_need + 3) & ~3), p_max = (int)((p_need + 3) & ~3);
  const size_t smem = coop_smem_bytes<T>(s_max, p_max);
  if (smem > (size_t)kCoopSmem) return -1000;
  auto kern = k_spgemm_rows_coop<T>;
  GLAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kCoopSmem));
  const int64_t n = X->n_rows;
  int64_t blocks = (n + kCoopRows - 1) / kCoopRows;
  const int64_t cap = (int64_t)X->sm_count * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned)blocks, kCoopRows, smem, st>>>(X->rowptr, X->colidx, xv, Y->rowptr, Y->colidx, yv, n, poff, rowcnt,
                                                  park_col, park_val, s_max, p_max);
  return (int)cudaGetLastError();
}

// Parked rows -> COO output (row, col as int64).  One warp per 32 consecutive rows: the lanes hold the
// rows' output and parking offsets, walk the rows' OUTPUT range 32 entries at a time (coalesced stores,
// near-contiguous loads) and find each entry's row by a shuffle binary search over the 32 offsets.
template <typename T>
__global__ void __launch_bounds__(256)
k_spgemm_unpark(const int64_t* __restrict__ poff, const int32_t* __restrict__ zrp,
                const int32_t* __restrict__ park_col, const T* __restrict__ park_val, int64_t n,
                int64_t* __restrict__ out_row, int64_t* __restrict__ out_col, T* __restrict__ out_val) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r0 = warp * 32; r0 < n; r0 += nwarps * 32) {
    const int64_t rl = r0 + lane;
    const int64_t rend = r0 + 32 < n ? r0 + 32 : n;
    const int zend = __ldg(zrp + rend);
    const int zl = rl < n ? __ldg(zrp + rl) : zend;
    const int64_t pl = rl < n ? __ldg(poff + rl) : 0;
    const int zbeg = __shfl_sync(0xffffffffu, zl, 0);
    for (int base = zbeg; base < zend; base += 32) {
      const int z = base + lane;
      int lo = 0;  // largest i with zrp[r0 + i] <= z  (empty rows share a start: the last one owns the entry)
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const int zc = __shfl_sync(0xffffffffu, zl, lo + step);   // lo + step <= 31
        if (zc <= z) lo += step;
      }
      const int zrow = __shfl_sync(0xffffffffu, zl, lo);
      const int64_t prow = __shfl_sync(0xffffffffu, pl, lo);
      if (z < zend) {
        const int64_t p = prow + (z - zrow);
        out_row[z] = r0 + lo;
        out_col[z] = (int64_t)__ldg(park_col + p);
        out_val[z] = __ldg(park_val + p);
      }
    }
  }
}

// ---- ESC path ---------------------------------------------------------------------------------
// 8 lanes per X row: the lanes walk the row's slots together and spread over each Y row, so a
// product's position is rowoff[r] + (products of the earlier slots) + (position in the Y row):
// expansion order == X slot order, then Y slot order.
template <typename T>
__global__ void k_spgemm_expand(const int32_t* __restrict__ xrp, const int32_t* __restrict__ xci,
                                const T* __restrict__ xv, const int32_t* __restrict__ yrp,
                                const int32_t* __restrict__ yci, const T* __restrict__ yv, int64_t n,
                                const int64_t* __restrict__ rowoff, int col_bits,
                                uint64_t* __restrict__ keys, T* __restrict__ vals) {
  const int lane = threadIdx.x & 7;
  const int64_t group = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) >> 3;
  for (int64_t r = group; r < n; r += ngroups) {
    int64_t o = rowoff[r];
    const int e1 = xrp[r + 1];
    for (int e = xrp[r]; e < e1; ++e) {
      const int j = xci[e];
      const T a = xv[e];
      const int y0 = yrp[j], y1 = yrp[j + 1];
      for (int t = y0 + lane; t < y1; t += 8) {
        keys[o + (t - y0)] = ((uint64_t)r << col_bits) | (uint64_t)(uint32_t)yci[t];
        vals[o + (t - y0)] = a * yv[t];
      }
      o += y1 - y0;
    }
  }
}

__global__ void k_spgemm_store_hdr(int64_t* hdr, int slot, int64_t value) { hdr[slot] = value; }

template <typename T>
__global__ void k_spgemm_compress(const int64_t* __restrict__ hdr, const uint64_t* __restrict__ k0,
                                  const uint64_t* __restrict__ k1, const T* __restrict__ v0,
                                  const T* __restrict__ v1, int64_t n_products, int64_t nnz_out,
                                  int col_bits, int64_t* __restrict__ out_row,
                                  int64_t* __restrict__ out_col, T* __restrict__ out_val) {
  const int sel = (int)hdr[1];
  const uint64_t* __restrict__ keys = sel ? k1 : k0;
  const T* __restrict__ vals = sel ? v1 : v0;
  // the run starts live in the key buffer the sort left free
  const int32_t* __restrict__ heads = reinterpret_cast<const int32_t*>(sel ? k0 : k1);
  const uint64_t cmask = (col_bits >= 64) ? ~0ull : ((1ull << col_bits) - 1);
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < nnz_out;
       u += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p0 = heads[u];
    const int64_t p1 = (u + 1 < nnz_out) ? (int64_t)heads[u + 1] : n_products;
    T acc = vals[p0];
    for (int64_t p = p0 + 1; p < p1; ++p) acc = acc + vals[p];
    const uint64_t key = keys[p0];
    out_row[u] = (int64_t)(key >> col_bits);
    out_col[u] = (int64_t)(key & cmask);
    out_val[u] = acc;
  }
}

static int spgemm_check(const glab_plan* X, const glab_plan* Y) {
  if (!X || !Y) return GLAB_E_ARG;
  if (X->n_cols != Y->n_rows) return GLAB_E_ARG;
  if (bits_for(X->n_rows) + bits_for(Y->n_cols) > 63) return GLAB_E_RANGE;
  return 0;
}

template <typename T>
static int spgemm_symbolic(const glab_plan* X, const T* xv, const glab_plan* Y, const T* yv, void* ws,
                           int64_t ws_bytes, int64_t n_products, int64_t max_row_products, int64_t* nnz_out,
                           void* stream_) {
  int rc = spgemm_check(X, Y);
  if (rc) return rc;
  if (!nnz_out || !ws) return GLAB_E_ARG;
  if ((X->nnz > 0 && !xv) || (Y->nnz > 0 && !yv)) return GLAB_E_ARG;
  SpgemmWs L;
  rc = spgemm_layout<T>(X->n_rows, n_products, max_row_products, &L);
  if (rc) return rc;
  if (ws_bytes < (int64_t)L.total) return GLAB_E_ARG;
  cudaStream_t st = as_stream(stream_);
  char* base = reinterpret_cast<char*>(ws);
  int64_t* hdr = reinterpret_cast<int64_t*>(base + L.hdr);
  int64_t* rowoff = reinterpret_cast<int64_t*>(base + L.rowoff);
  int32_t* rowcnt = reinterpret_cast<int32_t*>(base + L.rowcnt);
  int32_t* park_col = reinterpret_cast<int32_t*>(base + L.keys[0]);
  T* park_val = reinterpret_cast<T*>(base + L.vals[0]);
  void* ctmp = base + L.cub;
  size_t cb = L.cub_bytes;
  GLAB_CUDA(cudaMemsetAsync(hdr, 0, 256, st));
  if (n_products == 0) {
    *nnz_out = 0;
    return 0;
  }
  const int64_t n = X->n_rows;
  // product offset of every row (the parking slot of its strip; also the expansion offsets of ESC)
  k_spgemm_count<<<grid1d(n + 1, X->sm_count), 256, 0, st>>>(X->rowptr, X->colidx, Y->rowptr, n, rowoff, nullptr);
  GLAB_CUDA(cudaGetLastError());
  GLAB_CUDA(cub::DeviceScan::ExclusiveSum(ctmp, cb, rowoff, rowoff, (int)(n + 1), st));
  // ---- row-local, CTA-cooperative (no strip, no overflow) when 128 rows fit the shared-memory budget
  {
    static const bool coop_ok = [] { const char* e = getenv("GLAB_SPGEMM_COOP"); return !(e && atoi(e) == 0); }();
    rc = coop_ok ? spgemm_rows_coop_launch<T>(X, xv, Y, yv, rowoff, rowcnt, park_col, park_val, max_row_products, st)
                 : -1000;
    if (rc != -1000) {
      if (rc) return rc;
      cb = L.cub_bytes;
      GLAB_CUDA(cub::DeviceScan::ExclusiveSum(ctmp, cb, rowcnt, rowcnt, (int)(n + 1), st));
      int32_t total = 0;
      GLAB_CUDA(cudaMemcpyAsync(&total, rowcnt + n, 4, cudaMemcpyDeviceToHost, st));
      k_spgemm_store_hdr<<<1, 1, 0, st>>>(hdr, 3, 1);
      GLAB_CUDA(cudaStreamSynchronize(st));
      GLAB_CUDA(cudaGetLastError());
      *nnz_out = total;
      return 0;
    }
  }
  // ---- row-local attempts: the smallest strip that can hold the rows, one size up on overflow
  const int ladder[3] = {16, 32, kRowCap};
  int first = max_row_products <= 16 ? 0 : 1;      // more than 16 products: most stencil products fit 32 columns
  if (const char* e = getenv("GLAB_SPGEMM_CAP")) {  // tuning knob: 16 / 32 / 64
    const int v = atoi(e);
    first = v >= 64 ? 2 : v >= 32 ? 1 : 0;
    if (max_row_products > ladder[first] && first < 1) first = 1;
  }
  for (int step = first; step < 3; ++step) {
    GLAB_CUDA(cudaMemsetAsync(hdr + 2, 0, 8, st));
    switch (ladder[step]) {
      case 16: rc = spgemm_rows_launch<T, 16>(X, xv, Y, yv, rowoff, rowcnt, park_col, park_val, hdr, st); break;
      case 32: rc = spgemm_rows_launch<T, 32>(X, xv, Y, yv, rowoff, rowcnt, park_col, park_val, hdr, st); break;
      default: rc = spgemm_rows_launch<T, kRowCap>(X, xv, Y, yv, rowoff, rowcnt, park_col, park_val, hdr, st); break;
    }
    if (rc) return rc;
    cb = L.cub_bytes;
    GLAB_CUDA(cub::DeviceScan::ExclusiveSum(ctmp, cb, rowcnt, rowcnt, (int)(n + 1), st));
    int32_t total = 0;
    int64_t overflow = 0;
    GLAB_CUDA(cudaMemcpyAsync(&total, rowcnt + n, 4, cudaMemcpyDeviceToHost, st));
    GLAB_CUDA(cudaMemcpyAsync(&overflow, hdr + 2, 8, cudaMemcpyDeviceToHost, st));
    GLAB_CUDA(cudaStreamSynchronize(st));
    if (!overflow) {
      k_spgemm_store_hdr<<<1, 1, 0, st>>>(hdr, 3, 1);
      GLAB_CUDA(cudaGetLastError());
      *nnz_out = total;
      return 0;
    }
    if (max_row_products <= ladder[step]) return GLAB_E_ARG;  // cannot happen: <= CAP products cannot overflow
  }
  if (!L.esc) return GLAB_E_ARG;  // cannot happen: <= kRowCap products per row cannot overflow
  // ---- ESC fallback
  uint64_t* k[2] = {reinterpret_cast<uint64_t*>(base + L.keys[0]), reinterpret_cast<uint64_t*>(base + L.keys[1])};
  T* v[2] = {reinterpret_cast<T*>(base + L.vals[0]), reinterpret_cast<T*>(base + L.vals[1])};
  const int col_bits = bits_for(Y->n_cols);   // (rowoff already holds the expansion offsets)
  k_spgemm_expand<T><<<grid1d(n * 8, X->sm_count), 256, 0, st>>>(X->rowptr, X->colidx, xv, Y->rowptr, Y->colidx,
                                                              yv, n, rowoff, col_bits, k[0], v[0]);
  GLAB_CUDA(cudaGetLastError());
  cub::DoubleBuffer<uint64_t> dk(k[0], k[1]);
  cub::DoubleBuffer<T> dv(v[0], v[1]);
  cb = L.cub_bytes;
  GLAB_CUDA(cub::DeviceRadixSort::SortPairs(ctmp, cb, dk, dv, (int)n_products, 0,
                                            col_bits + bits_for(X->n_rows), st));
  const int sel = dk.selector;
  if (dv.selector != sel) return GLAB_E_ARG;  // cannot happen: CUB flips both buffers together
  k_spgemm_store_hdr<<<1, 1, 0, st>>>(hdr, 1, sel);
  k_spgemm_store_hdr<<<1, 1, 0, st>>>(hdr, 3, 2);
  HeadOp op{k[sel]};
  cb = L.cub_bytes;
  GLAB_CUDA(cub::DeviceSelect::If(ctmp, cb, thrust::counting_iterator<int>(0),
                                  reinterpret_cast<int32_t*>(k[sel ^ 1]), hdr, (int)n_products, op, st));
  int64_t h = 0;
  GLAB_CUDA(cudaMemcpyAsync(&h, hdr, 8, cudaMemcpyDeviceToHost, st));
  GLAB_CUDA(cudaStreamSynchronize(st));
  GLAB_CUDA(cudaGetLastError());
  *nnz_out = h;
  return 0;
}

template <typename T>
static int spgemm_numeric(const glab_plan* X, const T* xv, const glab_plan* Y, const T* yv, void* ws,
                          int64_t ws_bytes, int64_t n_products, int64_t max_row_products, int64_t nnz_out,
                          int64_t* out_row, int64_t* out_col, T* out_val, void* stream_) {
  int rc = spgemm_check(X, Y);
  if (rc) return rc;
  if (!ws || nnz_out < 0 || nnz_out > n_products) return GLAB_E_ARG;
  if (nnz_out == 0) return 0;
  if (!out_row || !out_col || !out_val || !xv || !yv) return GLAB_E_ARG;
  SpgemmWs L;
  rc = spgemm_layout<T>(X->n_rows, n_products, max_row_products, &L);
  if (rc) return rc;
  if (ws_bytes < (int64_t)L.total) return GLAB_E_ARG;
  cudaStream_t st = as_stream(stream_);
  char* base = reinterpret_cast<char*>(ws);
  int64_t* hdr = reinterpret_cast<int64_t*>(base + L.hdr);
  int64_t path = 0;  // which path the symbolic phase took (it left the row offsets / sorted runs)
  GLAB_CUDA(cudaMemcpyAsync(&path, hdr + 3, 8, cudaMemcpyDeviceToHost, st));
  GLAB_CUDA(cudaStreamSynchronize(st));
  if (path == 1) {
    k_spgemm_unpark<T><<<grid1d(X->n_rows, X->sm_count), 256, 0, st>>>(
        reinterpret_cast<const int64_t*>(base + L.rowoff), reinterpret_cast<const int32_t*>(base + L.rowcnt),
        reinterpret_cast<const int32_t*>(base + L.keys[0]), reinterpret_cast<const T*>(base + L.vals[0]), X->n_rows,
        out_row, out_col, out_val);
    return (int)cudaGetLastError();
  }
  if (path != 2 || !L.esc) return GLAB_E_ARG;  // symbolic phase did not run on this workspace
  k_spgemm_compress<T><<<grid1d(nnz_out, X->sm_count), 256, 0, st>>>(
      hdr, reinterpret_cast<const uint64_t*>(base + L.keys[0]), reinterpret_cast<const uint64_t*>(base + L.keys[1]),
      reinterpret_cast<const T*>(base + L.vals[0]), reinterpret_cast<const T*>(base + L.vals[1]), n_products, nnz_out,
      bits_for(Y->n_cols), out_row, out_col, out_val);
  return (int)cudaGetLastError();
}

// ============================================================================ PMIS C/F splitting
// state: 0 undecided, 1 coarse, 2 fine.  key_i = (lambda_i + 1) << 32 | mix32(i + seed) with
// lambda_i = number of rows that strongly depend on i; mix32 is a bijection of uint32, so all keys
// are distinct and positive: every comparison is an integer comparison, no ties, no rounding.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {  // murmur3 finaliser (bijective)
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}

struct PmisWs {
  size_t key, maxnbr, lambda, state, counter, total;
};

static int pmis_layout(int64_t n, PmisWs* L) {
  if (n < 0) return GLAB_E_ARG;
  size_t o = 0;
  L->counter = o; o += 256;
  L->key = o;     o += align256((size_t)(n + 1) * 8);
  L->maxnbr = o;  o += align256((size_t)(n + 1) * 8);
  L->lambda = o;  o += align256((size_t)(n + 1) * 4);
  L->state = o;   o += align256((size_t)(n + 1) * 4);
  L->total = o;
  return 0;
}

template <typename T>
__global__ void k_pmis_lambda(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                              const T* __restrict__ S, int64_t n, uint32_t* __restrict__ lambda) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int e1 = rowptr[i + 1];
    for (int e = rowptr[i]; e < e1; ++e)
      if (S[e] > T(0) && colidx[e] != i) atomicAdd(lambda + colidx[e], 1u);
  }
}

__global__ void k_pmis_keys(const uint32_t* __restrict__ lambda, uint32_t seed, int64_t n,
                            unsigned long long* __restrict__ key) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    key[i] = ((unsigned long long)(lambda[i] + 1u) << 32) | (unsigned long long)mix32((uint32_t)i + seed);
}

// largest key among the undecided strong neighbours (either direction) of every undecided vertex
template <typename T>
__global__ void k_pmis_neighbour_max(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                     const T* __restrict__ S, int64_t n, const int32_t* __restrict__ state,
                                     const unsigned long long* __restrict__ key,
                                     unsigned long long* __restrict__ maxnbr) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (state[i] != 0) continue;
    const unsigned long long ki = key[i];
    unsigned long long m = 0;
    const int e1 = rowptr[i + 1];
    for (int e = rowptr[i]; e < e1; ++e) {
      if (!(S[e] > T(0))) continue;
      const int j = colidx[e];
      if (j == i || state[j] != 0) continue;
      const unsigned long long kj = key[j];
      m = kj > m ? kj : m;
      atomicMax(maxnbr + j, ki);
    }
    if (m) atomicMax(maxnbr + i, m);
  }
}

__global__ void k_pmis_select(int64_t n, int32_t* __restrict__ state,
                              const unsigned long long* __restrict__ key,
                              unsigned long long* __restrict__ maxnbr) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (state[i] == 0 && key[i] > maxnbr[i]) state[i] = 1;
    maxnbr[i] = 0;
  }
}

// undecided rows that strongly depend on a coarse point become fine; counts what is left.
// Only 0 -> 2 transitions happen here and only "== 1" is tested, so the kernel is order-independent.
template <typename T>
__global__ void k_pmis_fine(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                            const T* __restrict__ S, int64_t n, int32_t* __restrict__ state,
                            unsigned long long* __restrict__ undecided) {
  unsigned long long left = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (state[i] != 0) continue;
    bool fine = false;
    const int e1 = rowptr[i + 1];
    for (int e = rowptr[i]; e < e1 && !fine; ++e)
      fine = (S[e] > T(0)) && colidx[e] != i && state[colidx[e]] == 1;
    if (fine) state[i] = 2; else ++left;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) left += __shfl_xor_sync(0xffffffffu, left, o);
  if ((threadIdx.x & 31) == 0 && left) atomicAdd(undecided, left);
}

template <typename T>
__global__ void k_pmis_flags(const int32_t* __restrict__ state, int64_t n, T* __restrict__ cflag) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    cflag[i] = state[i] == 1 ? T(1) : T(0);
}

template <typename T>
static int cf_split_pmis(const glab_plan* A, const T* S, uint32_t seed, void* ws, int64_t ws_bytes,
                         T* cflag, int32_t* rounds_out, void* stream_) {
  if (!A || !ws) return GLAB_E_ARG;
  const int64_t n = A->n_rows;
  if (A->n_cols != n) return GLAB_E_ARG;
  if ((n > 0 && !cflag) || (A->nnz > 0 && !S)) return GLAB_E_ARG;
  PmisWs L;
  int rc = pmis_layout(n, &L);
  if (rc) return rc;
  if (ws_bytes < (int64_t)L.total) return GLAB_E_ARG;
  if (rounds_out) *rounds_out = 0;
  if (n == 0) return 0;
  cudaStream_t st = as_stream(stream_);
  char* base = reinterpret_cast<char*>(ws);
  unsigned long long* counter = reinterpret_cast<unsigned long long*>(base + L.counter);
  unsigned long long* key = reinterpret_cast<unsigned long long*>(base + L.key);
  unsigned long long* maxnbr = reinterpret_cast<unsigned long long*>(base + L.maxnbr);
  uint32_t* lambda = reinterpret_cast<uint32_t*>(base + L.lambda);
  int32_t* state = reinterpret_cast<int32_t*>(base + L.state);
  GLAB_CUDA(cudaMemsetAsync(ws, 0, L.total, st));
  const int gv = grid1d(n, A->sm_count);
  if (A->nnz > 0) k_pmis_lambda<T><<<gv, 256, 0, st>>>(A->rowptr, A->colidx, S, n, lambda);
  k_pmis_keys<<<gv, 256, 0, st>>>(lambda, seed, n, key);
  int rounds = 0;
  unsigned long long left = (unsigned long long)n;
  while (left) {
    if (++rounds > 100000) return GLAB_E_ARG;  // cannot happen: the largest undecided key wins each round
    k_pmis_neighbour_max<T><<<gv, 256, 0, st>>>(A->rowptr, A->colidx, S, n, state, key, maxnbr);
    k_pmis_select<<<gv, 256, 0, st>>>(n, state, key, maxnbr);
    GLAB_CUDA(cudaMemsetAsync(counter, 0, 8, st));
    k_pmis_fine<T><<<gv, 256, 0, st>>>(A->rowptr, A->colidx, S, n, state, counter);
    GLAB_CUDA(cudaMemcpyAsync(&left, counter, 8, cudaMemcpyDeviceToHost, st));
    GLAB_CUDA(cudaStreamSynchronize(st));
  }
  k_pmis_flags<T><<<gv, 256, 0, st>>>(state, n, cflag);
  GLAB_CUDA(cudaGetLastError());
  if (rounds_out) *rounds_out = rounds;
  return 0;
}

}  // namespace glab

using namespace glab;

extern "C" int64_t glab_interp_workspace_bytes(int64_t n) { return interp_workspace_bytes(n); }
extern "C" int glab_interp_count_f32(const glab_plan* A, const float* w, const float* c, int mode, void* ws,
                                     int64_t ws_bytes, int32_t* prow, int32_t* cid, int64_t* nnz_p,
                                     int64_t* n_coarse, void* s) {
  return interp_count<float>(A, w, c, mode, ws, ws_bytes, prow, cid, nnz_p, n_coarse, s);
}
extern "C" int glab_interp_count_f64(const glab_plan* A, const double* w, const double* c, int mode, void* ws,
                                     int64_t ws_bytes, int32_t* prow, int32_t* cid, int64_t* nnz_p,
                                     int64_t* n_coarse, void* s) {
  return interp_count<double>(A, w, c, mode, ws, ws_bytes, prow, cid, nnz_p, n_coarse, s);
}
extern "C" int glab_interp_fill_f32(const glab_plan* A, const float* w, const float* c, int mode,
                                    const int32_t* prow, const int32_t* cid, int64_t* orow, int64_t* ocol,
                                    float* oval, void* s) {
  return interp_fill<float>(A, w, c, mode, prow, cid, orow, ocol, oval, s);
}
extern "C" int glab_interp_fill_f64(const glab_plan* A, const double* w, const double* c, int mode,
                                    const int32_t* prow, const int32_t* cid, int64_t* orow, int64_t* ocol,
                                    double* oval, void* s) {
  return interp_fill<double>(A, w, c, mode, prow, cid, orow, ocol, oval, s);
}

extern "C" int glab_spgemm_products(const glab_plan* X, const glab_plan* Y, void* scratch16, int64_t* n_products,
                                    int64_t* max_row_products, void* stream_) {
  int rc = spgemm_check(X, Y);
  if (rc) return rc;
  if (!n_products || !max_row_products || !scratch16) return GLAB_E_ARG;
  cudaStream_t st = as_stream(stream_);
  unsigned long long* total = reinterpret_cast<unsigned long long*>(scratch16);
  GLAB_CUDA(cudaMemsetAsync(total, 0, 16, st));
  if (X->n_rows > 0 && X->nnz > 0)
    k_spgemm_count<<<grid1d(X->n_rows + 1, X->sm_count), 256, 0, st>>>(X->rowptr, X->colidx, Y->rowptr,
                                                                      X->n_rows, nullptr, total);
  unsigned long long h[2] = {0, 0};
  GLAB_CUDA(cudaMemcpyAsync(h, total, 16, cudaMemcpyDeviceToHost, st));
  GLAB_CUDA(cudaStreamSynchronize(st));
  GLAB_CUDA(cudaGetLastError());
  *n_products = (int64_t)h[0];
  *max_row_products = (int64_t)h[1];
  return 0;
}

extern "C" int64_t glab_spgemm_workspace_bytes(int64_t n_rows_x, int64_t n_products, int64_t max_row_products,
                                               int elem_size) {
  SpgemmWs L;
  int rc = (elem_size == 8)   ? spgemm_layout<double>(n_rows_x, n_products, max_row_products, &L)
           : (elem_size == 4) ? spgemm_layout<float>(n_rows_x, n_products, max_row_products, &L)
                              : GLAB_E_ARG;
  return rc ? (int64_t)(rc < 0 ? rc : -rc) : (int64_t)L.total;
}

extern "C" int glab_spgemm_symbolic_f32(const glab_plan* X, const float* xv, const glab_plan* Y, const float* yv,
                                        void* ws, int64_t ws_bytes, int64_t n_products, int64_t max_row_products,
                                        int64_t* nnz_out, void* s) {
  return spgemm_symbolic<float>(X, xv, Y, yv, ws, ws_bytes, n_products, max_row_products, nnz_out, s);
}
extern "C" int glab_spgemm_symbolic_f64(const glab_plan* X, const double* xv, const glab_plan* Y, const double* yv,
                                        void* ws, int64_t ws_bytes, int64_t n_products, int64_t max_row_products,
                                        int64_t* nnz_out, void* s) {
  return spgemm_symbolic<double>(X, xv, Y, yv, ws, ws_bytes, n_products, max_row_products, nnz_out, s);
}
extern "C" int glab_spgemm_numeric_f32(const glab_plan* X, const float* xv, const glab_plan* Y, const float* yv,
                                       void* ws, int64_t ws_bytes, int64_t n_products, int64_t max_row_products,
                                       int64_t nnz_out, int64_t* orow, int64_t* ocol, float* oval, void* s) {
  return spgemm_numeric<float>(X, xv, Y, yv, ws, ws_bytes, n_products, max_row_products, nnz_out, orow, ocol, oval, s);
}
extern "C" int glab_spgemm_numeric_f64(const glab_plan* X, const double* xv, const glab_plan* Y, const double* yv,
                                       void* ws, int64_t ws_bytes, int64_t n_products, int64_t max_row_products,
                                       int64_t nnz_out, int64_t* orow, int64_t* ocol, double* oval, void* s) {
  return spgemm_numeric<double>(X, xv, Y, yv, ws, ws_bytes, n_products, max_row_products, nnz_out, orow, ocol, oval,
                                s);
}

extern "C" int64_t glab_cf_split_workspace_bytes(int64_t n) {
  PmisWs L;
  int rc = pmis_layout(n, &L);
  return rc ? (int64_t)rc : (int64_t)L.total;
}
extern "C" int glab_cf_split_pmis_f32(const glab_plan* A, const float* S, uint32_t seed, void* ws, int64_t ws_bytes,
                                      float* cflag, int32_t* rounds, void* s) {
  return cf_split_pmis<float>(A, S, seed, ws, ws_bytes, cflag, rounds, s);
}
extern "C" int glab_cf_split_pmis_f64(const glab_plan* A, const double* S, uint32_t seed, void* ws, int64_t ws_bytes,
                                      double* cflag, int32_t* rounds, void* s) {
  return cf_split_pmis<double>(A, S, seed, ws, ws_bytes, cflag, rounds, s);
}
