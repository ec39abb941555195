// glab_pipe.cuh -- persistent, warp-specialised, TMA-fed row-tile pipeline (the fast path of
// every fused SpMV-bearing layer on sm_100a).
//
// Why: the plain one-tile-per-CTA kernel (glab_tiles.cuh) is latency bound -- every CTA walks
// the dependent chain rowptr -> CSR stream -> barrier -> gather -> store, and ncu shows ~18
// warps per issue slot parked on the long scoreboard at 49 % of the HBM roofline.  Here the
// streaming part of the traffic (colidx, vals, the tile's rowptr slice and the row-aligned
// slices of the epilogue's input vectors: diag / b / x_i / r / p) never touches a register on
// its way in: ONE producer lane per CTA issues 1-D bulk asynchronous copies
// (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes, i.e. the TMA engine) into a
// ring of `S` shared-memory stages, running S-1 tiles ahead of the 8 consumer warps, with
// full/empty mbarriers per stage.  The consumers only do shared-memory reads, the L1/L2-served
// gathers of x[col], the fused epilogue and coalesced stores.  Grid = (#SMs x CTAs/SM), tiles
// are dealt round-robin so that concurrently running CTAs work on neighbouring row blocks and
// share their gather windows in L2.
#pragma once
#include <type_traits>
#include "glab_tiles.cuh"

namespace glab {

constexpr int kPipeThreads = kThreads + 32;  // 8 consumer warps + 1 producer warp
constexpr int kMaxStreams = 3;               // row-aligned epilogue input vectors staged by TMA
// Byte offset, inside a stage's rowptr region, of the word in which the producer tells the consumers
// whether the tile streams 16-bit indices (IDX == 2).  The rowptr slice occupies at most
// (kThreads + 1) * 4 bytes rounded up to 16 = 1040 of the region's 1152 bytes.
constexpr int kTileFlagOff = ((kThreads + 1) * 4 + 15) & ~15;

struct PipeLayout {
  int stages;
  int stage_bytes;             // multiple of 128
  int off_col, off_val, off_row;
  int off_stream[kMaxStreams];
};

// ---- mbarrier / bulk-copy PTX wrappers ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy through the TMA engine; src/dst 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-
// serialization attribute may start while its predecessor in the stream is still draining.
// Everything it does before pdl_wait() must touch only data that no kernel of the sequence
// writes (rowptr / barrier set-up); pdl_wait() returns once the predecessor grid has completed
// and its writes are visible.  pdl_launch_dependents() lets the successor start early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// [p, p + n*esz) widened to 16-byte granules: returns aligned start, bytes (multiple of 16).
__device__ __forceinline__ void align16(const void* p, int nbytes, const void*& start, uint32_t& bytes) {
  const uintptr_t p0 = reinterpret_cast<uintptr_t>(p);
  const uintptr_t a0 = p0 & ~static_cast<uintptr_t>(15);
  start = reinterpret_cast<const void*>(a0);
  bytes = static_cast<uint32_t>(((p0 - a0) + (uintptr_t)nbytes + 15) & ~static_cast<uintptr_t>(15));
}
__device__ __forceinline__ int lead_elems(const void* p, int esz) {
  return static_cast<int>(reinterpret_cast<uintptr_t>(p) & 15) / esz;
}

// Multi-GPU control block of a fused step (HALO = true).  One launch per sweep:
//   * logical tile order puts the rows that read the halo tail FIRST; the producer lane acquires
//     the neighbours' arrival counters before it feeds the first such tile (the neighbours pushed
//     early in THEIR previous kernel, so this never stalls in steady state);
//   * every finished boundary tile bumps a device counter; CTA 0 is a communication CTA that does
//     no row work: it waits for that counter, then stores this rank's boundary values of the
//     produced vector straight into the neighbours' halo tails (peer stores over NVLink) and
//     release-increments their arrival counters -- all while the other CTAs are still busy with
//     the interior tiles, so neither the exchange latency nor the flag round trip is exposed.
struct HaloCtl {
  int int_tile0, int_tiles;    // interior tiles [int_tile0, int_tile0 + int_tiles)
  int lead_tiles, trail_tile0; // boundary tiles [0, lead_tiles) and [trail_tile0, ntiles)
  int n_wait, n_push, push_k;
  uint32_t* wait_flag[GLAB_MAX_PEERS];
  const uint32_t* wait_target;
  glab_push_desc push[GLAB_MAX_PEERS];
  uint32_t* pushed_counter;
  const void* push_src;
  unsigned int* done_counter;
  uint32_t* status;                 // device word, OR-ed with GLAB_STATUS_* when a wait times out (may be NULL)
  unsigned long long timeout_ns;    // 0 = wait forever
};
struct NoHalo {};

// ---- bounded device-side waits --------------------------------------------------------------
// Every in-kernel wait on a flag that ANOTHER CTA or ANOTHER GPU writes is bounded: after
// `timeout_ns` of wall-clock (%globaltimer) the waiter records GLAB_STATUS_* in the caller's status
// word and carries on, so that a protocol bug or a lost peer ends in an error code on the host
// (glab_halo_status / DistOperator.check) instead of a hung GPU.  The results of that launch are
// then undefined.
__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
struct SpinGuard {
  unsigned long long t0 = 0, limit;
  unsigned int it = 0;
  __device__ __forceinline__ explicit SpinGuard(unsigned long long timeout_ns) : limit(timeout_ns) {}
  // call once per failed poll; true = give up
  __device__ __forceinline__ bool expired() {
    if (limit == 0 || (++it & 1023u) != 0) return false;
    const unsigned long long t = gtime_ns();
    if (t0 == 0) { t0 = t; return false; }
    return t - t0 > limit;
  }
};
__device__ __forceinline__ void flag_timeout(uint32_t* status, uint32_t code) {
  if (status) atomicOr(status, code);
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <typename T, int K>
__device__ __forceinline__ void load_vec_cg(T (&dst)[K], const T* p) {
  constexpr int bytes = K * (int)sizeof(T);
  if constexpr (bytes >= 16) {
    constexpr int per = 16 / (int)sizeof(T);
#pragma unroll
    for (int i = 0; i < bytes / 16; ++i) {
      int4 q = __ldcg(reinterpret_cast<const int4*>(p) + i);
      const T* t = reinterpret_cast<const T*>(&q);
#pragma unroll
      for (int j = 0; j < per; ++j) dst[i * per + j] = t[j];
    }
  } else if constexpr (bytes == 8) {
    int2 q = __ldcg(reinterpret_cast<const int2*>(p));
    const T* t = reinterpret_cast<const T*>(&q);
#pragma unroll
    for (int j = 0; j < K; ++j) dst[j] = t[j];
  } else {
    dst[0] = __ldcg(p);
  }
}

// Row sum of one CSR row of a staged tile: acc[c] = sum_j v_j * x[col_j, c], added one product at a
// time in slot order (the accumulation order of scatter_add_).  W16: the tile's columns are 2-byte
// offsets from the row (else 4-byte absolute indices).  CG: 0 = gather through the read-only path
// (vectors no kernel writes while this one runs), 1 = through L2 only (ld.global.cg), 2 = ordinary
// coherent loads (L1-cached, honour fences: vectors other CTAs write during the launch).  Rows of exactly U entries
// take an unpredicated straight-line path with U gathers in flight.
template <typename T, int K, int U, bool W16, int CG>
__device__ __forceinline__ void row_sum(T (&acc)[K], const unsigned char* __restrict__ cbuf, int cofs,
                                        const T* __restrict__ sval, int rs, int re, int r,
                                        const T* __restrict__ x) {
  auto col_at = [&](int j) -> int {
    if constexpr (W16) return r + (int)reinterpret_cast<const int16_t*>(cbuf)[j + cofs];
    else return reinterpret_cast<const int32_t*>(cbuf)[j + cofs];
  };
  auto gather = [&](T (&d)[K], int col) {
    if constexpr (CG == 1) load_vec_cg<T, K>(d, x + (size_t)col * K);
    else if constexpr (CG == 2) load_vec_rw<T, K>(d, x + (size_t)col * K);
    else load_vec<T, K>(d, x + (size_t)col * K);
  };
  if (re - rs == U) {
    T vv[U];
    T xv[U][K];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int col = col_at(rs + u);
      vv[u] = sval[rs + u];
      gather(xv[u], col);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int c = 0; c < K; ++c) acc[c] = acc[c] + vv[u] * xv[u][c];
    }
  } else {
    for (int base = rs; base < re; base += U) {
      T vv[U];
      T xv[U][K];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (base + u < re) {
          const int col = col_at(base + u);
          vv[u] = sval[base + u];
          gather(xv[u], col);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (base + u < re) {
#pragma unroll
          for (int c = 0; c < K; ++c) acc[c] = acc[c] + vv[u] * xv[u][c];
        }
      }
    }
  }
}

// The few tiles of a row block that read the halo tail: one L2-coherent gather at a time (the tail is
// written by the neighbours while earlier kernels of this rank may still run; these tiles are a
// vanishing share of the work, so they get the smallest body, not the fastest).
template <typename T, int K, bool W16>
__device__ __forceinline__ void row_sum_coherent(T (&acc)[K], const unsigned char* __restrict__ cbuf, int cofs,
                                                 const T* __restrict__ sval, int rs, int re, int r, const T* x) {
  for (int j = rs; j < re; ++j) {
    int col;
    if constexpr (W16) col = r + (int)reinterpret_cast<const int16_t*>(cbuf)[j + cofs];
    else col = reinterpret_cast<const int32_t*>(cbuf)[j + cofs];
    T xv[K];
    load_vec_cg<T, K>(xv, x + (size_t)col * K);
    const T v = sval[j];
#pragma unroll
    for (int c = 0; c < K; ++c) acc[c] = acc[c] + v * xv[c];
  }
}

// Epilogue protocol for the pipeline (all in glab_layers.cu):
//   static constexpr int kStreams;  const T* stream_ptr(i);  int stream_width(i)  (elements/row)
//   State, init(State&), finish(State&)
//   row_staged(State&, int r, const T (&acc)[K], const T* s0, const T* s1, const T* s2)
//       s_i -> this row's elements of stream i in shared memory.
// Host-checked preconditions: rowptr and every stream pointer are 16-byte aligned and
// row_begin * width * sizeof(T) is a multiple of 16, so those slices start on a 16-byte granule
// (only the colidx / vals slices, which start at an arbitrary CSR slot, carry a lead offset).
// U = gathers kept in flight per thread per pass; rows of exactly U entries (every interior row
// of a U-point stencil) take an unpredicated straight-line path.
// IDX: how the column indices are streamed.  0 = 4-byte absolute indices (a.colidx).  1 = 2-byte
// offsets from the row (a.coldelta) in every tile: banded operators, single-GPU plans -- 2 B less
// HBM traffic per nonzero, the only per-nonzero bytes besides the value itself.  2 = per 256-row
// tile, a.tile16 says which of the two arrays the tile uses (periodic wrap-around rows, the halo
// columns of a row block: all other tiles still stream 2-byte indices).
template <typename T, int K, int U, class Epi, bool HALO, int IDX = 0>
__global__ void __launch_bounds__(kPipeThreads, (K * sizeof(T) <= 8) ? 4 : 2)
k_row_pipe(TileArgs<T> a, const T* __restrict__ x, Epi epi, int ntiles, PipeLayout L,
           typename std::conditional<HALO, HaloCtl, NoHalo>::type h) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + L.stages;
  unsigned char* stage0 = smem_raw + 128;  // barriers live in the first 128 bytes (<= 8 stages)
  const int tid = threadIdx.x;
  const int S = L.stages;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, kThreads / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_launch_dependents();

  typename Epi::State st;
  if constexpr (has_prologue<Epi>::value) epi.prologue(st);   // CTA-wide (e.g. the norm summed over the ranks)

  // logical tile -> (physical tile, reads-halo flag); boundary tiles come first
  auto phys = [&](int t, bool& boundary) -> int {
    if constexpr (HALO) {
      const int nb = ntiles - h.int_tiles;
      if (t >= nb) { boundary = false; return h.int_tile0 + (t - nb); }
      boundary = true;
      return t < h.lead_tiles ? t : h.trail_tile0 + (t - h.lead_tiles);
    } else {
      boundary = false;
      return t;
    }
  };
  // CTA roles: with a push to do, CTA 0 only communicates and the rest share the tiles
  int cta = blockIdx.x, ncta = gridDim.x;
  bool comm_cta = false;
  if constexpr (HALO) {
    if (h.n_push > 0 && gridDim.x > 1) {
      comm_cta = (blockIdx.x == 0);
      cta = (int)blockIdx.x - 1;
      ncta = (int)gridDim.x - 1;
    }
  }

  if constexpr (HALO) {
    // A rank with nothing to send still counts this step as one "push" of the produced vector: the
    // counter is the wait target of the next step that gathers it.
    if (!comm_cta && h.n_push == 0 && blockIdx.x == 0 && tid == 0 && h.pushed_counter) {
      pdl_wait();
      *h.pushed_counter += 1u;
    }
  }
  if (comm_cta) {
    pdl_wait();
    epi.init(st);
    if constexpr (HALO) {
      // ---------------------------------------------------------------- communication CTA
      const unsigned int nb = (unsigned int)(ntiles - h.int_tiles);
      if (tid == 0) {
        SpinGuard guard(h.timeout_ns);
        while (ld_acquire_gpu(h.done_counter) < nb) {
          __nanosleep(200);
          if (guard.expired()) { flag_timeout(h.status, GLAB_STATUS_TIMEOUT_TILES); break; }
        }
      }
      __syncthreads();
      __threadfence();
      const T* src = reinterpret_cast<const T*>(h.push_src);
      for (int q = 0; q < h.n_push; ++q) push_rows<T, K>(src, h.push[q], tid, kPipeThreads);
      __syncthreads();  // the release below is cumulative over every thread's peer stores
      if (tid == 0) {
        for (int q = 0; q < h.n_push; ++q)
          if (h.push[q].flag)
            asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(h.push[q].flag) : "memory");
        if (h.pushed_counter) *h.pushed_counter += 1u;
        *h.done_counter = 0u;
      }
    }
  } else if (tid >= kThreads) {
   if constexpr (!HALO) {
    // Single-GPU kernels: one producer lane with a one-tile look-ahead on the CSR extents.  (The
    // batch prefetch below was measured SLOWER here -- L4096 Jacobi 0.135 -> 0.153 ms -- although it
    // is faster for the row blocks of the halo kernels: profiles/r02_halo_gap.md.)
    // ------------------------------------------------------------------ producer warp
    if (tid != kThreads) {
      pdl_wait();
      epi.init(st);
    } else {
      int lt = cta;
      int e0n = 0, e1n = 0;
      bool t16n = (IDX == 1);
      bool bnd = false, waited = false;
      if (lt < ntiles) {
        const int r0 = a.row_begin + phys(lt, bnd) * kThreads;
        e0n = __ldg(a.rowptr + r0);
        e1n = __ldg(a.rowptr + min(r0 + kThreads, a.row_end));
        if constexpr (IDX == 2) t16n = __ldg(a.tile16 + r0 / kThreads) != 0;
      }
      pdl_wait();  // rowptr is constant; everything copied below may come from the previous kernel
      epi.init(st);
      int s = 0;
      uint32_t phase = 0;
      for (; lt < ntiles; lt += ncta) {
        const int r0 = a.row_begin + phys(lt, bnd) * kThreads;
        const int r1 = min(r0 + kThreads, a.row_end);
        const int e0 = e0n, e1 = e1n;
        const bool t16 = t16n;
        const int nt = lt + ncta;
        if (nt < ntiles) {  // prefetch the next tile's extents while this stage drains
          bool b2;
          const int q0 = a.row_begin + phys(nt, b2) * kThreads;
          e0n = __ldg(a.rowptr + q0);
          e1n = __ldg(a.rowptr + min(q0 + kThreads, a.row_end));
          if constexpr (IDX == 2) t16n = __ldg(a.tile16 + q0 / kThreads) != 0;
        }
        if constexpr (HALO) {
          if (bnd && !waited) {  // neighbours' halo rows must have landed before consumers gather them
            const uint32_t want = *reinterpret_cast<const volatile uint32_t*>(h.wait_target);
            for (int i = 0; i < h.n_wait; ++i) {
              SpinGuard guard(h.timeout_ns);
              while ((int32_t)(ld_acquire_sys(h.wait_flag[i]) - want) < 0) {
                __nanosleep(32);
                if (guard.expired()) { flag_timeout(h.status, GLAB_STATUS_TIMEOUT_PEER); break; }
              }
            }
            waited = true;
          }
        }
        mbar_wait(empty + s, phase ^ 1u);
        unsigned char* sb = stage0 + (size_t)s * L.stage_bytes;
        const void *src_c = nullptr, *src_v = nullptr;
        uint32_t nb_c = 0, nb_v = 0, nb_s[kMaxStreams];
        const uint32_t nb_r = (uint32_t)(((r1 - r0 + 1) * 4 + 15) & ~15);
        uint32_t total = nb_r;
        if (e1 > e0) {
          if (IDX != 0 && t16) align16(a.coldelta + e0, (e1 - e0) * 2, src_c, nb_c);
          else align16(a.colidx + e0, (e1 - e0) * 4, src_c, nb_c);
          align16(a.vals + e0, (e1 - e0) * (int)sizeof(T), src_v, nb_v);
          total += nb_c + nb_v;
        }
#pragma unroll
        for (int i = 0; i < kMaxStreams; ++i) {
          nb_s[i] = 0;
          if (i < Epi::kStreams) {
            nb_s[i] = (uint32_t)(((r1 - r0) * epi.stream_width(i) * (int)sizeof(T) + 15) & ~15);
            total += nb_s[i];
          }
        }
        // ordinary shared store; the arrive below releases it to the consumers' acquire on `full`
        if constexpr (IDX == 2) *reinterpret_cast<volatile int*>(sb + L.off_row + kTileFlagOff) = t16 ? 1 : 0;
        mbar_expect_tx(full + s, total);
        bulk_g2s(sb + L.off_row, a.rowptr + r0, nb_r, full + s);
        if (nb_c) {
          bulk_g2s(sb + L.off_col, src_c, nb_c, full + s);
          bulk_g2s(sb + L.off_val, src_v, nb_v, full + s);
        }
#pragma unroll
        for (int i = 0; i < kMaxStreams; ++i)
          if (i < Epi::kStreams)
            bulk_g2s(sb + L.off_stream[i], epi.stream_ptr(i) + (size_t)r0 * epi.stream_width(i), nb_s[i],
                     full + s);
        if (++s == S) { s = 0; phase ^= 1u; }
      }
    }
   } else {
    // ------------------------------------------------------------------ producer warp
    // Lane 0 feeds the ring.  The CSR extents of a tile (two dependent global loads, ~1-3 us under
    // load) are prefetched in batches of 32 tiles, one tile per lane, a whole batch ahead: `cur`
    // holds the batch being issued, `nxt` the following one, whose loads were issued 32 tiles
    // earlier (profiles/r02_halo_gap.md: with a one-tile look-ahead the producer sat on that load for
    // ~40 % of its time and the consumers waited on `full` for 17-30 % of theirs).  Two register
    // sets, because the scoreboard tracks a register for the whole warp: a load into one lane of the
    // set being broadcast would stall every shuffle behind it.
    const int lane = tid - kThreads;
    int cur_e0 = 0, cur_e1 = 0, cur_t16 = (IDX == 1) ? 1 : 0;
    int nxt_e0 = 0, nxt_e1 = 0, nxt_t16 = (IDX == 1) ? 1 : 0;
    auto fetch = [&](int seq, int& f0, int& f1, int& f16) {
      const int lt2 = cta + seq * ncta;
      if (lt2 < ntiles) {
        bool b2;
        const int q0 = a.row_begin + phys(lt2, b2) * kThreads;
        f0 = __ldg(a.rowptr + q0);
        f1 = __ldg(a.rowptr + min(q0 + kThreads, a.row_end));
        if constexpr (IDX == 2) f16 = __ldg(a.tile16 + q0 / kThreads) != 0;
      }
    };
    fetch(lane, nxt_e0, nxt_e1, nxt_t16);
    pdl_wait();  // rowptr is constant; everything copied below may come from the previous kernel
    epi.init(st);
    {
      bool bnd = false, waited = false;
      int s = 0, seq = 0;
      uint32_t phase = 0;
      for (int lt = cta; lt >= 0 && lt < ntiles; lt += ncta, ++seq) {
        const int src = seq & 31;
        if (src == 0) {   // next batch becomes current (its loads are 32 tiles old), the one after is requested
          cur_e0 = nxt_e0; cur_e1 = nxt_e1; cur_t16 = nxt_t16;
          fetch(seq + 32 + lane, nxt_e0, nxt_e1, nxt_t16);
        }
        const int e0 = __shfl_sync(0xffffffffu, cur_e0, src);
        const int e1 = __shfl_sync(0xffffffffu, cur_e1, src);
        const bool t16 = __shfl_sync(0xffffffffu, cur_t16, src) != 0;
        if (lane == 0) {
          const int r0 = a.row_begin + phys(lt, bnd) * kThreads;
          const int r1 = min(r0 + kThreads, a.row_end);
          if constexpr (HALO) {
            if (bnd && !waited) {  // neighbours' halo rows must have landed before consumers gather them
              const uint32_t want = *reinterpret_cast<const volatile uint32_t*>(h.wait_target);
              for (int i = 0; i < h.n_wait; ++i) {
                SpinGuard guard(h.timeout_ns);
                while ((int32_t)(ld_acquire_sys(h.wait_flag[i]) - want) < 0) {
                  __nanosleep(32);
                  if (guard.expired()) { flag_timeout(h.status, GLAB_STATUS_TIMEOUT_PEER); break; }
                }
              }
              waited = true;
            }
          }
          mbar_wait(empty + s, phase ^ 1u);
          unsigned char* sb = stage0 + (size_t)s * L.stage_bytes;
          const void *src_c = nullptr, *src_v = nullptr;
          uint32_t nb_c = 0, nb_v = 0, nb_s[kMaxStreams];
          const uint32_t nb_r = (uint32_t)(((r1 - r0 + 1) * 4 + 15) & ~15);
          uint32_t total = nb_r;
          if (e1 > e0) {
            if (IDX != 0 && t16) align16(a.coldelta + e0, (e1 - e0) * 2, src_c, nb_c);
            else align16(a.colidx + e0, (e1 - e0) * 4, src_c, nb_c);
            align16(a.vals + e0, (e1 - e0) * (int)sizeof(T), src_v, nb_v);
            total += nb_c + nb_v;
          }
#pragma unroll
          for (int i = 0; i < kMaxStreams; ++i) {
            nb_s[i] = 0;
            if (i < Epi::kStreams) {
              nb_s[i] = (uint32_t)(((r1 - r0) * epi.stream_width(i) * (int)sizeof(T) + 15) & ~15);
              total += nb_s[i];
            }
          }
          // ordinary shared store; the arrive below releases it to the consumers' acquire on `full`
          if constexpr (IDX == 2) *reinterpret_cast<volatile int*>(sb + L.off_row + kTileFlagOff) = t16 ? 1 : 0;
          mbar_expect_tx(full + s, total);
          bulk_g2s(sb + L.off_row, a.rowptr + r0, nb_r, full + s);
          if (nb_c) {
            bulk_g2s(sb + L.off_col, src_c, nb_c, full + s);
            bulk_g2s(sb + L.off_val, src_v, nb_v, full + s);
          }
#pragma unroll
          for (int i = 0; i < kMaxStreams; ++i)
            if (i < Epi::kStreams)
              bulk_g2s(sb + L.off_stream[i], epi.stream_ptr(i) + (size_t)r0 * epi.stream_width(i), nb_s[i],
                       full + s);
        }
        __syncwarp();
        if (++s == S) { s = 0; phase ^= 1u; }
      }
    }
   }
  } else {
    // ------------------------------------------------------------------ consumer warps
    pdl_wait();
    epi.init(st);
    int s = 0;
    uint32_t phase = 0;
    for (int lt = cta; lt < ntiles; lt += ncta) {
      bool bnd;
      const int r0 = a.row_begin + phys(lt, bnd) * kThreads;
      const int r1 = min(r0 + kThreads, a.row_end);
      const int r = r0 + tid;
      unsigned char* sb = stage0 + (size_t)s * L.stage_bytes;
      const int32_t* srow = reinterpret_cast<const int32_t*>(sb + L.off_row);
      bool t16 = (IDX == 1);
      mbar_wait(full + s, phase);
      if constexpr (IDX == 2)  // block-uniform; written by the producer, no global latency here
        t16 = *reinterpret_cast<const volatile int*>(sb + L.off_row + kTileFlagOff) != 0;
      if (r < r1) {
        const int e0 = srow[0];
        const int rs = srow[tid], re = srow[tid + 1];
        // the tile's columns are 2-byte offsets from the row or 4-byte absolute indices: ONE block-uniform
        // branch per tile around straight-line row bodies (no per-element select)
        const unsigned char* cbuf = sb + L.off_col;
        const T* sval = reinterpret_cast<const T*>(sb + L.off_val) + lead_elems(a.vals + e0, sizeof(T)) - e0;
        T acc[K];
#pragma unroll
        for (int c = 0; c < K; ++c) acc[c] = T(0);
        if (IDX != 0 && t16) {
          if constexpr (IDX != 0) {
            const int cofs = lead_elems(a.coldelta + e0, 2) - e0;
            // rows that read the halo tail gather through L2 (the tail was written by peers)
            if (HALO && bnd) row_sum_coherent<T, K, true>(acc, cbuf, cofs, sval, rs, re, r, x);
            else row_sum<T, K, U, true, 0>(acc, cbuf, cofs, sval, rs, re, r, x);
          }
        } else {
          if constexpr (IDX != 1) {
            const int cofs = lead_elems(a.colidx + e0, 4) - e0;
            if (HALO && bnd) row_sum_coherent<T, K, false>(acc, cbuf, cofs, sval, rs, re, r, x);
            else row_sum<T, K, U, false, 0>(acc, cbuf, cofs, sval, rs, re, r, x);
          }
        }
        const T* sp[kMaxStreams] = {nullptr, nullptr, nullptr};
#pragma unroll
        for (int i = 0; i < kMaxStreams; ++i)
          if (i < Epi::kStreams)
            sp[i] = reinterpret_cast<const T*>(sb + L.off_stream[i]) + (size_t)tid * epi.stream_width(i);
        epi.row_staged(st, r, acc, sp[0], sp[1], sp[2]);
      }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(empty + s);
      if (++s == S) { s = 0; phase ^= 1u; }
      if constexpr (HALO) {
        if (bnd && h.n_push > 0) {  // tell the communication CTA that this boundary tile is stored
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (tid == 0) {
            __threadfence();
            atomicAdd(h.done_counter, 1u);
          }
        }
      }
    }
  }
  epi.finish(st);

}

// ------------------------------------------------------------------------------------------
// Per-edge-output pipeline (AMG setup): same producer / ring as k_row_pipe, but the consumers
// (1) reduce over their row in shared memory, (2) overwrite the staged values with the per-edge
// result in place, (3) meet at a named barrier and (4) write the tile's contiguous slot range
// out coalesced (out[perm[slot]] when the caller's edge order is not the CSR order).
//   Op::kReduce / kNarr (1: vals, 2: vals + aux) / kNeedCol;  RowState, begin_row, accumulate,
//   end_row, edge -- see glab_amg.cu.
// Rows are read with 16-byte shared-memory loads whenever the row start is 16-byte aligned in
// every staged array (always true for the fixed-degree stencil operators), which avoids the
// 8-way bank conflicts a stride-8 scalar walk would have.
// ------------------------------------------------------------------------------------------
struct EdgePipeLayout {
  int stages, stage_bytes, off_row, off_col, off_val, off_aux;
};

template <typename E, int V> __device__ __forceinline__ void lds_vec(E (&d)[V], const E* p) {
  if constexpr (V * sizeof(E) == 16) {
    const int4 q = *reinterpret_cast<const int4*>(p);
    const E* t = reinterpret_cast<const E*>(&q);
#pragma unroll
    for (int i = 0; i < V; ++i) d[i] = t[i];
  } else if constexpr (V * sizeof(E) == 8) {
    const int2 q = *reinterpret_cast<const int2*>(p);
    const E* t = reinterpret_cast<const E*>(&q);
#pragma unroll
    for (int i = 0; i < V; ++i) d[i] = t[i];
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) d[i] = p[i];
  }
}
template <typename E, int V> __device__ __forceinline__ void sts_vec(E* p, const E (&d)[V]) {
  if constexpr (V * sizeof(E) == 16) {
    int4 q;
    E* t = reinterpret_cast<E*>(&q);
#pragma unroll
    for (int i = 0; i < V; ++i) t[i] = d[i];
    *reinterpret_cast<int4*>(p) = q;
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) p[i] = d[i];
  }
}

// IDX16: tiles flagged in a.tile16 stream 16-bit row-relative column indices (a.coldelta: col - row) in
// place of the int32 ones -- 2 bytes per edge less HBM traffic for the operators that read the column
// (soc_sa, direct_interp); the other tiles (wrap-around / far columns) keep int32.  The flag is per tile,
// so the consumer branches once per tile around two straight-line row bodies.
template <typename T, class Op, bool IDX16 = false>
__global__ void __launch_bounds__(kPipeThreads, 4)
k_edge_pipe(TileArgs<T> a, const T* __restrict__ aux, const int32_t* __restrict__ perm, Op op,
            T* __restrict__ out, int ntiles, EdgePipeLayout L) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + L.stages;
  unsigned char* stage0 = smem_raw + 128;
  const int tid = threadIdx.x;
  const int S = L.stages;
  constexpr int V = 16 / (int)sizeof(T);  // elements per 16-byte value vector

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, kThreads / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (tid >= kThreads) {
    if (tid == kThreads) {  // ---------------------------------------------------- producer
      int tile = blockIdx.x;
      int e0n = 0, e1n = 0;
      if (tile < ntiles) {
        const int r0 = tile * kThreads;
        e0n = __ldg(a.rowptr + r0);
        e1n = __ldg(a.rowptr + min(r0 + kThreads, a.row_end));
      }
      int s = 0;
      uint32_t phase = 0;
      for (; tile < ntiles; tile += gridDim.x) {
        const int r0 = tile * kThreads;
        const int r1 = min(r0 + kThreads, a.row_end);
        const int e0 = e0n, e1 = e1n;
        const int nt = tile + gridDim.x;
        if (nt < ntiles) {
          const int q0 = nt * kThreads;
          e0n = __ldg(a.rowptr + q0);
          e1n = __ldg(a.rowptr + min(q0 + kThreads, a.row_end));
        }
        mbar_wait(empty + s, phase ^ 1u);
        unsigned char* sb = stage0 + (size_t)s * L.stage_bytes;
        const void *src_c = nullptr, *src_v = nullptr, *src_a = nullptr;
        uint32_t nb_c = 0, nb_v = 0, nb_a = 0;
        const uint32_t nb_r = (uint32_t)(((r1 - r0 + 1) * 4 + 15) & ~15);
        uint32_t total = nb_r;
        if (e1 > e0) {
          align16(a.vals + e0, (e1 - e0) * (int)sizeof(T), src_v, nb_v);
          total += nb_v;
          if (Op::kNeedCol) {
            if (IDX16 && __ldg(a.tile16 + tile)) align16(a.coldelta + e0, (e1 - e0) * 2, src_c, nb_c);
            else align16(a.colidx + e0, (e1 - e0) * 4, src_c, nb_c);
            total += nb_c;
          }
          if (Op::kNarr > 1) {
            align16(aux + e0, (e1 - e0) * (int)sizeof(T), src_a, nb_a);
            total += nb_a;
          }
        }
        mbar_expect_tx(full + s, total);
        bulk_g2s(sb + L.off_row, a.rowptr + r0, nb_r, full + s);
        if (nb_v) bulk_g2s(sb + L.off_val, src_v, nb_v, full + s);
        if (nb_c) bulk_g2s(sb + L.off_col, src_c, nb_c, full + s);
        if (nb_a) bulk_g2s(sb + L.off_aux, src_a, nb_a, full + s);
        if (++s == S) { s = 0; phase ^= 1u; }
      }
    }
  } else {
    int s = 0;  // ------------------------------------------------------------------ consumers
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int r0 = tile * kThreads;
      const int r1 = min(r0 + kThreads, a.row_end);
      const int r = r0 + tid;
      unsigned char* sb = stage0 + (size_t)s * L.stage_bytes;
      const int32_t* srow = reinterpret_cast<const int32_t*>(sb + L.off_row);
      mbar_wait(full + s, phase);
      const int e0 = srow[0];
      const int e1 = srow[r1 - r0];
      T* sval = reinterpret_cast<T*>(sb + L.off_val) + lead_elems(a.vals + e0, sizeof(T)) - e0;
      const int32_t* scol = reinterpret_cast<const int32_t*>(sb + L.off_col) + lead_elems(a.colidx + e0, 4) - e0;
      const T* saux = reinterpret_cast<const T*>(sb + L.off_aux) + lead_elems(aux + e0, sizeof(T)) - e0;
      bool t16 = false;
      if constexpr (IDX16 && Op::kNeedCol) t16 = __ldg(a.tile16 + tile) != 0;
      if (t16) {
        if constexpr (IDX16 && Op::kNeedCol) {
          const int16_t* sdel = reinterpret_cast<const int16_t*>(sb + L.off_col) + lead_elems(a.coldelta + e0, 2) - e0;
          if (r < r1) {
            const int rs = srow[tid], re = srow[tid + 1];
            typename Op::RowState st;
            op.begin_row(st, r);
            const bool vec = ((re - rs) % V == 0) && ((reinterpret_cast<uintptr_t>(sval + rs) & 15) == 0) &&
                             ((reinterpret_cast<uintptr_t>(sdel + rs) & (V * 2 - 1)) == 0) &&
                             (Op::kNarr < 2 || (reinterpret_cast<uintptr_t>(saux + rs) & 15) == 0);
            if (vec) {
              if (Op::kReduce) {
                for (int j = rs; j < re; j += V) {
                  T v[V], x2[V];
                  int16_t d[V];
                  lds_vec<T, V>(v, sval + j);
                  lds_vec<int16_t, V>(d, sdel + j);
                  if (Op::kNarr > 1) lds_vec<T, V>(x2, saux + j);
#pragma unroll
                  for (int u = 0; u < V; ++u) op.accumulate(st, v[u], Op::kNarr > 1 ? x2[u] : T(0), r + (int)d[u]);
                }
              }
              op.end_row(st, r);
              for (int j = rs; j < re; j += V) {
                T v[V], x2[V], o[V];
                int16_t d[V];
                lds_vec<T, V>(v, sval + j);
                lds_vec<int16_t, V>(d, sdel + j);
                if (Op::kNarr > 1) lds_vec<T, V>(x2, saux + j);
#pragma unroll
                for (int u = 0; u < V; ++u) o[u] = op.edge(st, v[u], Op::kNarr > 1 ? x2[u] : T(0), r + (int)d[u]);
                sts_vec<T, V>(sval + j, o);
              }
            } else {
              if (Op::kReduce)
                for (int j = rs; j < re; ++j)
                  op.accumulate(st, sval[j], Op::kNarr > 1 ? saux[j] : T(0), r + (int)sdel[j]);
              op.end_row(st, r);
              for (int j = rs; j < re; ++j)
                sval[j] = op.edge(st, sval[j], Op::kNarr > 1 ? saux[j] : T(0), r + (int)sdel[j]);
            }
          }
        }
      } else
      if (r < r1) {
        const int rs = srow[tid], re = srow[tid + 1];
        typename Op::RowState st;
        op.begin_row(st, r);
        // 16-byte alignment of the row start in every staged array -> vector walk
        const bool vec = ((re - rs) % V == 0) && ((reinterpret_cast<uintptr_t>(sval + rs) & 15) == 0) &&
                         (!Op::kNeedCol || (reinterpret_cast<uintptr_t>(scol + rs) & (V * 4 - 1)) == 0) &&
                         (Op::kNarr < 2 || (reinterpret_cast<uintptr_t>(saux + rs) & 15) == 0);
        if (vec) {
          if (Op::kReduce) {
            for (int j = rs; j < re; j += V) {
              T v[V], x2[V];
              int32_t c[V];
              lds_vec<T, V>(v, sval + j);
              if (Op::kNeedCol) lds_vec<int32_t, V>(c, scol + j);
              if (Op::kNarr > 1) lds_vec<T, V>(x2, saux + j);
#pragma unroll
              for (int u = 0; u < V; ++u)
                op.accumulate(st, v[u], Op::kNarr > 1 ? x2[u] : T(0), Op::kNeedCol ? c[u] : 0);
            }
          }
          op.end_row(st, r);
          for (int j = rs; j < re; j += V) {
            T v[V], x2[V], o[V];
            int32_t c[V];
            lds_vec<T, V>(v, sval + j);
            if (Op::kNeedCol) lds_vec<int32_t, V>(c, scol + j);
            if (Op::kNarr > 1) lds_vec<T, V>(x2, saux + j);
#pragma unroll
            for (int u = 0; u < V; ++u)
              o[u] = op.edge(st, v[u], Op::kNarr > 1 ? x2[u] : T(0), Op::kNeedCol ? c[u] : 0);
            sts_vec<T, V>(sval + j, o);
          }
        } else {
          if (Op::kReduce)
            for (int j = rs; j < re; ++j)
              op.accumulate(st, sval[j], Op::kNarr > 1 ? saux[j] : T(0), Op::kNeedCol ? scol[j] : 0);
          op.end_row(st, r);
          for (int j = rs; j < re; ++j)
            sval[j] = op.edge(st, sval[j], Op::kNarr > 1 ? saux[j] : T(0), Op::kNeedCol ? scol[j] : 0);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // whole tile's results are in shared memory
      if (perm == nullptr) {
        for (int j = e0 + tid; j < e1; j += kThreads) out[j] = sval[j];
      } else {
        for (int j = e0 + tid; j < e1; j += kThreads) out[__ldg(perm + j)] = sval[j];
      }
      // our in-place (generic-proxy) writes to the stage must be ordered before the TMA
      // (async-proxy) refill that the producer issues once the stage is released
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(empty + s);
      if (++s == S) { s = 0; phase ^= 1u; }
    }
  }
}

}  // namespace glab
