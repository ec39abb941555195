// glab_multisweep.cuh -- several weighted-Jacobi sweeps in ONE persistent kernel launch.
//
// Why: a sweep on an L2-sized row block (one rank's share of a partitioned operator, a coarse level of
// a multigrid hierarchy) takes 15-25 us, of which ~6 us are the pipeline ramp and tail of the launch
// (measured: t = 5.9 us + bytes / 6.5 TB/s for the single-sweep kernel, profiles/r02_*).  Programmatic
// dependent launch cannot hide them, because the next kernel's CTAs only become resident when the
// previous kernel's CTAs retire.  Here the TMA producer simply keeps running ahead across the sweep
// boundary: the ring of stages never drains between sweeps.
//
// Dataflow instead of a grid barrier.  x ping-pongs between two buffers; tile t of sweep s
//   * gathers rows of x_in that tiles t-dep .. t+dep wrote in sweep s-1      (read after write)
//   * overwrites rows of x_out that the same tiles gathered in sweep s-1     (write after read)
// where dep = tile half-bandwidth of the operator (a plan property).  Every consumer warp
// release-increments its tile's completion counter after its stores; before the producer warp feeds
// tile t of sweep s it acquires the counters of t-dep .. t+dep (one coalesced load per lane) and waits
// until they show sweep s-1 complete.  Tiles are dealt round-robin to the CTAs of a CO-RESIDENT grid
// (cooperative launch), all CTAs advance in lock step, and the dependency is a whole sweep old when
// it is checked, so the wait is normally satisfied on the first poll.  Operators without band
// structure get dep >= ntiles, which degenerates into one full completion check per sweep.
//
// Coherence: vectors that change during the launch are never read through the non-coherent path or
// the async proxy.  Only launch-constant data (CSR arrays, diag, b) is TMA-staged; x is gathered with
// ordinary (L1-cached) global loads and the row's own x_i is a coalesced load issued with the gathers.
//
// The hand-off between CTAs (profiles/r02_multisweep.md has the measurements behind each choice):
//   * A fenced release per consumer warp and tile (red.release.gpu = MEMBAR.ALL.GPU + RED) stalls the
//     warp ~2 us in this kernel -- more than the tile's own 1.3 us -- and halves the throughput; moving
//     it to a helper lane of the producer warp stalls the producer instead.  So the default hand-off
//     relies on the L2 being the GPU's point of coherence instead of on MEMBAR:
//       writer  every lane stores its row of x_out; two tiles later (the store is long past the LSU by
//               then, so the load is not held behind it) it reads its own first element back from L2
//               with ld.global.cg, and one tile after that the warp votes that every lane read the
//               bits it had stored.  That is a VERIFIED statement about the L2 -- which every reader
//               reads -- not an ordering assumption; a mismatch simply repeats the read.  Then lane 0
//               issues a RELAXED reduction on the tile's counter.
//       reader  the producer warp polls the counters with relaxed loads (L2), then performs one
//               acquiring load (LDG.STRONG + CCTL.IVALL: invalidates this SM's L1, so no line cached two
//               sweeps ago, when the same buffer held older values, survives) before it releases the
//               stage to the consumers through the stage's mbarrier.
//     A line of x_in is immutable from the moment the check passes until every tile that reads it has
//     finished the sweep (the write-after-read half of the same dependency).
//   * GLAB_MS_STRICT=1 selects the formally fenced hand-off (red.release.gpu / fence.acq_rel.gpu) at
//     run time; tests run both and require identical bits.  GLAB_MS_GATHER_CG=1 at compile time
//     switches the gathers to ld.global.cg.
//
// Multi-GPU (HALO): per sweep the tiles that read the halo tail or are sent to a neighbour come
// first; their producer acquires the neighbours' arrival counters of that sweep's input buffer; the
// communication CTA pushes this rank's boundary rows of the sweep's output into the neighbours' halo
// tails as soon as the boundary tiles of the sweep are stored, while the interior tiles are still in
// flight.  Same counters and descriptors as the single-sweep kernels (glab_halo_step).
#pragma once
#include "glab_pipe.cuh"

namespace glab {

struct MsCtl {
  uint32_t* tile_done;   // [ntiles] consumer-warp completions, monotonically increasing across launches
  uint32_t* epoch;       // tile_done[*] value at the start of this launch (device word)
  unsigned int* ticket;  // exit ticket (the last CTA advances the epoch)
  int dep;               // dependency half-width in tiles (>= ntiles: every tile; < 0: unchecked, timing experiments)
  int chunk;             // dependencies are verified for this many consecutive tiles of a CTA at once
  int strict;            // 1 = fenced release / acquire on every hand-off (MEMBAR.GPU per warp and tile: ~2x slower)
  uint32_t* status;
  unsigned long long timeout_ns;
};

// One direction of the ping-pong on a row-partitioned operator: the sweep gathers `in`, produces `out`.
struct MsHaloDir {
  int n_wait, n_push;
  uint32_t* wait_flag[GLAB_MAX_PEERS];   // arrival counters of the gathered buffer
  const uint32_t* wait_target;           // how often this rank had pushed that buffer before the launch
  glab_push_desc push[GLAB_MAX_PEERS];   // destinations of the produced buffer
  uint32_t* pushed_counter;              // this rank's push count of the produced buffer
};
struct MsHalo {
  int int_tile0, int_tiles, lead_tiles, trail_tile0;
  MsHaloDir dir[2];                      // [0]: sweeps 0, 2, ... (gather A, produce B); [1]: the odd sweeps
  unsigned int* done_counter;            // [0], [4]: finished boundary tiles of the even / odd sweeps (16 B apart)
};
struct MsNoHalo {};

constexpr int kConsumerWarps = kThreads / 32;

#ifndef GLAB_MS_RING
#define GLAB_MS_RING 1    // 1 = CSR extents prefetched a 32-tile batch ahead, 0 = one tile ahead
#endif
#ifndef GLAB_MS_GATHER_CG
#define GLAB_MS_GATHER_CG 0
#endif
constexpr bool kMsGatherCG = GLAB_MS_GATHER_CG != 0;

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// One element through L2, never elided or hoisted by the compiler (the read-back of a lane's own store).
__device__ __forceinline__ float ld_cg_volatile(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_cg_volatile(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// A comparison whose only purpose is to make an instruction depend on a loaded register (the
// hardware scoreboard then holds the warp until the load has returned).  The pattern is a signalling
// NaN payload no computation here produces; the outcome is not used for control flow.
__device__ __forceinline__ bool is_poison(float a) { return __float_as_uint(a) == 0x7fa5a5a5u; }
__device__ __forceinline__ bool is_poison(double a) { return __double_as_longlong(a) == 0x7ff5a5a5a5a5a5a5ll; }
__device__ __forceinline__ bool bits_equal(float a, float b) { return __float_as_uint(a) == __float_as_uint(b); }
__device__ __forceinline__ bool bits_equal(double a, double b) {
  return __double_as_longlong(a) == __double_as_longlong(b);
}
constexpr int kMsDefer = 3;   // tiles between a warp's stores and the publication of their completion
// own-row / coherent variants of load_vec: plain global loads (L1-cached, honour fences)
template <typename T, int K>
__device__ __forceinline__ void load_vec_ca(T (&dst)[K], const T* p) {
  load_vec_rw<T, K>(dst, p);
}

template <typename T, int K, int U, int IDX, bool HALO>
__global__ void __launch_bounds__(kPipeThreads, (K * sizeof(T) <= 8) ? 4 : 2)
k_jacobi_ms(TileArgs<T> a, T* __restrict__ xa, T* __restrict__ xb, const T* __restrict__ diag,
            const T* __restrict__ b, const T* __restrict__ omega, int nsweeps, int ntiles, PipeLayout L, MsCtl m,
            typename std::conditional<HALO, MsHalo, MsNoHalo>::type h) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + L.stages;
  unsigned char* stage0 = smem_raw + 128;
  const int tid = threadIdx.x;
  const int S = L.stages;   // <= 4: the barriers occupy the first 64 of the 128 reserved bytes
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, kConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();
  const uint32_t base = *reinterpret_cast<const volatile uint32_t*>(m.epoch);

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
  int cta = blockIdx.x, ncta = gridDim.x;
  bool comm_cta = false, has_comm = false;
  if constexpr (HALO) {
    if (gridDim.x > 1 && (h.dir[0].n_push > 0 || h.dir[1].n_push > 0)) {
      has_comm = true;
      comm_cta = (blockIdx.x == 0);
      cta = (int)blockIdx.x - 1;
      ncta = (int)gridDim.x - 1;
    }
  }

  if (comm_cta) {
    if constexpr (HALO) {
      // ------------------------------------------------------------------ communication CTA
      // The boundary tiles of consecutive sweeps can overlap by at most one sweep (a neighbour only
      // pushes sweep s after receiving this rank's push of sweep s - 1), so one counter per sweep
      // parity separates them.
      const unsigned int nb = (unsigned int)(ntiles - h.int_tiles);
      for (int sw = 0; sw < nsweeps; ++sw) {
        const MsHaloDir& d = h.dir[sw & 1];
        if (tid == 0) {
          SpinGuard guard(m.timeout_ns);
          while (ld_acquire_gpu(h.done_counter + 4 * (sw & 1)) < nb * (unsigned int)(sw / 2 + 1)) {
            __nanosleep(100);
            if (guard.expired()) { flag_timeout(m.status, GLAB_STATUS_TIMEOUT_TILES); break; }
          }
        }
        __syncthreads();
        __threadfence();
        const T* src = (sw & 1) ? xa : xb;   // the buffer this sweep produced
        for (int q = 0; q < d.n_push; ++q) push_rows<T, K>(src, d.push[q], tid, kPipeThreads);
        __syncthreads();  // the release below is cumulative over every thread's peer stores
        if (tid == 0) {
          for (int q = 0; q < d.n_push; ++q)
            if (d.push[q].flag)
              asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(d.push[q].flag) : "memory");
        }
      }
      if (tid == 0) {
        h.done_counter[0] = 0u;
        h.done_counter[4] = 0u;
      }
    }
  } else if (tid >= kThreads) {
    // -------------------------------------------------------------------- producer warp
    const int lane = tid - kThreads;
    int sw = 0, lt = cta;
    bool bnd = false;
    // CSR extents prefetched 32 tiles ahead (see k_row_pipe): lane l holds those of this CTA's
    // (32 j + l)-th tile of the launch; the sequence simply continues across sweeps.
    const int per_sweep = (cta < ntiles) ? (ntiles - cta + ncta - 1) / ncta : 0;   // tiles of this CTA per sweep
    int cur_e0 = 0, cur_e1 = 0, cur_t16 = (IDX == 1) ? 1 : 0;
    int nxt_e0 = 0, nxt_e1 = 0, nxt_t16 = (IDX == 1) ? 1 : 0;
    auto fetch = [&](int seq, int& f0, int& f1, int& f16) {
      if (per_sweep > 0 && seq < per_sweep * nsweeps) {
        bool b2;
        const int r0 = phys(cta + (seq % per_sweep) * ncta, b2) * kThreads;
        f0 = __ldg(a.rowptr + r0);
        f1 = __ldg(a.rowptr + min(r0 + kThreads, a.row_end));
        if constexpr (IDX == 2) f16 = __ldg(a.tile16 + r0 / kThreads) != 0;
      }
    };
#if GLAB_MS_RING
    fetch(lane, nxt_e0, nxt_e1, nxt_t16);
#endif
    int seq = 0;
    int s = 0;
    uint32_t phase = 0;
    int full_ok_sweep = 0;       // wide dependency: sweeps <= this value are known to be complete everywhere
    int checked_lt = 0;          // logical tiles of the current sweep below this index have verified dependencies
    int halo_ok_sweep = -1;      // HALO: the neighbours' pushes for sweeps <= this value have arrived
    uint32_t want_in[2] = {0, 0};
    if constexpr (HALO) {
      want_in[0] = *reinterpret_cast<const volatile uint32_t*>(h.dir[0].wait_target);
      want_in[1] = *reinterpret_cast<const volatile uint32_t*>(h.dir[1].wait_target);
    }
    while (lt < ntiles && sw < nsweeps) {
      const int tile = phys(lt, bnd);
      const int r0 = tile * kThreads;
      const int r1 = min(r0 + kThreads, a.row_end);
#if GLAB_MS_RING
      const int srcl = seq & 31;
      if (srcl == 0) {   // next batch becomes current, the one after is requested (two register sets, see k_row_pipe)
        cur_e0 = nxt_e0; cur_e1 = nxt_e1; cur_t16 = nxt_t16;
        fetch(seq + 32 + lane, nxt_e0, nxt_e1, nxt_t16);
      }
      const int e0 = __shfl_sync(0xffffffffu, cur_e0, srcl);
      const int e1 = __shfl_sync(0xffffffffu, cur_e1, srcl);
      const bool t16 = __shfl_sync(0xffffffffu, cur_t16, srcl) != 0;
      ++seq;
#else
      // one-tile look-ahead, every lane loading the same (uniform) addresses
      if (seq == 0) { const int l0 = lane; fetch(0 - l0 + l0, nxt_e0, nxt_e1, nxt_t16); }
      const int e0 = nxt_e0, e1 = nxt_e1;
      const bool t16 = nxt_t16 != 0;
      ++seq;
      fetch(seq, nxt_e0, nxt_e1, nxt_t16);
#endif
      int nlt = lt + ncta, nsw = sw;
      if (nlt >= ntiles) { nlt = cta; nsw = sw + 1; }
      if (sw > 0 && m.dep >= 0 && lt >= checked_lt) {
        // The tiles this one gathers from / overwrites must have finished the previous sweep.  Verified
        // for this tile and the next m.chunk - 1 tiles of this CTA at once: relaxed polls, all lanes in
        // parallel; then the acquire -- a full fence in strict mode, else ONE acquiring load, whose
        // CCTL.IVALL invalidates this SM's L1 before the consumers gather.
        const uint32_t need = base + (uint32_t)(kConsumerWarps * sw);
        const bool wide = (m.dep >= ntiles);
        if (!(wide && full_ok_sweep >= sw)) {
          SpinGuard guard(m.timeout_ns);
          while (true) {
            bool ok = true;
            for (int g = 0; g < m.chunk; ++g) {
              const int lt2 = lt + g * ncta;
              if (lt2 >= ntiles) break;
              bool b2;
              const int t2 = phys(lt2, b2);
              int lo = t2 - m.dep, hi = t2 + m.dep;
              if (wide || lo < 0) lo = 0;
              if (wide || hi > ntiles - 1) hi = ntiles - 1;
              for (int i = lo + lane; i <= hi; i += 32) ok = ok && ((int32_t)(ld_relaxed_gpu(m.tile_done + i) - need) >= 0);
              if (wide) break;
            }
            if (__all_sync(0xffffffffu, ok)) break;
            __nanosleep(64);
            if (guard.expired()) { flag_timeout(m.status, GLAB_STATUS_TIMEOUT_SWEEP); break; }
          }
          if (m.strict) fence_acq_rel_gpu();
          else if (lane == 0) (void)ld_acquire_gpu(m.tile_done + tile);
          __syncwarp();
          if (wide) full_ok_sweep = sw;
        }
        checked_lt = lt + m.chunk * ncta;
      }
      if constexpr (HALO) {
        if (bnd && halo_ok_sweep < sw) {
          // neighbours' halo rows of this sweep's input buffer: pushed before the launch (sweep 0: the
          // producer of the vector; later sweeps: the neighbours' communication CTA of sweep sw - 1).
          // Pushes of the gathered buffer during this launch before sweep sw: buffer A (even sweeps gather
          // it) is produced by the odd sweeps 1, 3, .. < sw -> sw / 2;  buffer B by the even sweeps
          // 0, 2, .. < sw -> (sw + 1) / 2
          const MsHaloDir& d = h.dir[sw & 1];
          const uint32_t target = want_in[sw & 1] + (uint32_t)((sw & 1) ? (sw + 1) / 2 : sw / 2);
          if (lane < d.n_wait) {
            SpinGuard guard(m.timeout_ns);
            while ((int32_t)(ld_acquire_sys(d.wait_flag[lane]) - target) < 0) {
              __nanosleep(32);
              if (guard.expired()) { flag_timeout(m.status, GLAB_STATUS_TIMEOUT_PEER); break; }
            }
          }
          __syncwarp();
          halo_ok_sweep = sw;
        }
      }
      if (lane == 0) {
        mbar_wait(empty + s, phase ^ 1u);
        unsigned char* sb = stage0 + (size_t)s * L.stage_bytes;
        const void *src_c = nullptr, *src_v = nullptr;
        uint32_t nb_c = 0, nb_v = 0;
        const uint32_t nb_r = (uint32_t)(((r1 - r0 + 1) * 4 + 15) & ~15);
        const uint32_t nb_d = (uint32_t)(((r1 - r0) * (int)sizeof(T) + 15) & ~15);
        const uint32_t nb_b = (uint32_t)(((r1 - r0) * K * (int)sizeof(T) + 15) & ~15);
        uint32_t total = nb_r + nb_d + nb_b;
        if (e1 > e0) {
          if (IDX != 0 && t16) align16(a.coldelta + e0, (e1 - e0) * 2, src_c, nb_c);
          else align16(a.colidx + e0, (e1 - e0) * 4, src_c, nb_c);
          align16(a.vals + e0, (e1 - e0) * (int)sizeof(T), src_v, nb_v);
          total += nb_c + nb_v;
        }
        if constexpr (IDX == 2) *reinterpret_cast<volatile int*>(sb + L.off_row + kTileFlagOff) = t16 ? 1 : 0;
        mbar_expect_tx(full + s, total);
        bulk_g2s(sb + L.off_row, a.rowptr + r0, nb_r, full + s);
        if (nb_c) {
          bulk_g2s(sb + L.off_col, src_c, nb_c, full + s);
          bulk_g2s(sb + L.off_val, src_v, nb_v, full + s);
        }
        bulk_g2s(sb + L.off_stream[0], diag + r0, nb_d, full + s);
        bulk_g2s(sb + L.off_stream[1], b + (size_t)r0 * K, nb_b, full + s);
      }
      __syncwarp();
      if (++s == S) { s = 0; phase ^= 1u; }
      lt = nlt;
      if (nsw != sw) checked_lt = 0;
      sw = nsw;
    }
  } else {
    // -------------------------------------------------------------------- consumer warps
    const T w = __ldg(omega);
    int s = 0;
    uint32_t phase = 0;
    // Publication of a finished tile.  A write takes ~2 us to reach L2 under this kernel's load --
    // longer than a tile -- and anything that waits for it on the spot (MEMBAR, or a read-back of the
    // address just stored, which the LSU holds behind the store) stalls the warp's next gathers.  So a
    // finished tile rides a 3-slot shift register: two tiles after its stores the lane reads its own
    // first element back from L2 (ld.global.cg), one tile later the warp votes that every lane saw
    // the bits it had stored -- i.e. the L2, which is what every reader reads, HOLDS the new values
    // -- and only then lane 0 increments the tile's counter (a mismatch just repeats the read).
    int ptile[kMsDefer], ppar[kMsDefer];
    T pexp[kMsDefer];
    T echo = T(0);
    bool echo_valid = false;
#pragma unroll
    for (int i = 0; i < kMsDefer; ++i) { ptile[i] = -1; ppar[i] = 0; pexp[i] = T(0); }
    auto read_back0 = [&]() {   // slot 0: this lane's first stored element, through L2
      const int row = ptile[0] * kThreads + tid;
      echo = pexp[0];
      if (row < a.row_end) echo = ld_cg_volatile((ppar[0] ? xa : xb) + (size_t)row * K);
      echo_valid = true;
    };
    auto publish0 = [&]() {     // slot 0 must hold a tile
      while (true) {
        if (!echo_valid) read_back0();
        if (__all_sync(0xffffffffu, bits_equal(echo, pexp[0])) || m.dep < 0) break;   // (dep < 0: timing experiments)
        echo_valid = false;
      }
      if ((tid & 31) == 0) {
        if (m.strict) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(m.tile_done + ptile[0]) : "memory");
        else asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(m.tile_done + ptile[0]) : "memory");
      }
      ptile[0] = -1;
      echo_valid = false;
    };
    auto shift = [&]() {
#pragma unroll
      for (int i = 0; i + 1 < kMsDefer; ++i) { ptile[i] = ptile[i + 1]; ppar[i] = ppar[i + 1]; pexp[i] = pexp[i + 1]; }
      ptile[kMsDefer - 1] = -1;
    };
    auto push_pending = [&](int tile_id, int parity, T first) {
      if (ptile[0] >= 0) publish0();      // its read-back was issued one tile ago
      shift();
      ptile[kMsDefer - 1] = tile_id;
      ppar[kMsDefer - 1] = parity;
      pexp[kMsDefer - 1] = first;
      if (ptile[0] >= 0) read_back0();    // stored two tiles ago: long past the LSU
    };
    auto publish_all = [&]() {
#pragma unroll
      for (int i = 0; i < kMsDefer; ++i) {
        if (ptile[0] >= 0) publish0();
        shift();
      }
    };
    for (int sw = 0; sw < nsweeps; ++sw) {
      const T* xin = (sw & 1) ? xb : xa;
      T* xout = (sw & 1) ? xa : xb;
      for (int lt = cta; lt < ntiles; lt += ncta) {
        bool bnd;
        const int tile = phys(lt, bnd);
        const int r0 = tile * kThreads;
        const int r1 = min(r0 + kThreads, a.row_end);
        const int r = r0 + tid;
        unsigned char* sb = stage0 + (size_t)s * L.stage_bytes;
        const int32_t* srow = reinterpret_cast<const int32_t*>(sb + L.off_row);
        bool t16 = (IDX == 1);
        // never block while holding an unpublished completion (another CTA may be waiting for it)
        if (!mbar_test(full + s, phase)) publish_all();
        mbar_wait(full + s, phase);
        if constexpr (IDX == 2) t16 = *reinterpret_cast<const volatile int*>(sb + L.off_row + kTileFlagOff) != 0;
        T o[K];
#pragma unroll
        for (int c = 0; c < K; ++c) o[c] = T(0);
        if (r < r1) {
          T xx[K];
          if constexpr (kMsGatherCG) load_vec_cg<T, K>(xx, xin + (size_t)r * K);   // in flight together with the gathers
          else load_vec_ca<T, K>(xx, xin + (size_t)r * K);
          const int e0 = srow[0];
          const int rs = srow[tid], re = srow[tid + 1];
          const unsigned char* cbuf = sb + L.off_col;
          const T* sval = reinterpret_cast<const T*>(sb + L.off_val) + lead_elems(a.vals + e0, sizeof(T)) - e0;
          T acc[K];
#pragma unroll
          for (int c = 0; c < K; ++c) acc[c] = T(0);
          // rows that read the halo tail (written by other GPUs): always through L2
          if (IDX != 0 && t16) {
            if constexpr (IDX != 0) {
              const int cofs = lead_elems(a.coldelta + e0, 2) - e0;
              if (HALO && bnd) row_sum_coherent<T, K, true>(acc, cbuf, cofs, sval, rs, re, r, xin);
              else row_sum<T, K, U, true, kMsGatherCG ? 1 : 2>(acc, cbuf, cofs, sval, rs, re, r, xin);
            }
          } else {
            if constexpr (IDX != 1) {
              const int cofs = lead_elems(a.colidx + e0, 4) - e0;
              if (HALO && bnd) row_sum_coherent<T, K, false>(acc, cbuf, cofs, sval, rs, re, r, xin);
              else row_sum<T, K, U, false, kMsGatherCG ? 1 : 2>(acc, cbuf, cofs, sval, rs, re, r, xin);
            }
          }
          const T d = reinterpret_cast<const T*>(sb + L.off_stream[0])[tid];
          T bb[K];
          const T* sbp = reinterpret_cast<const T*>(sb + L.off_stream[1]) + (size_t)tid * K;
          constexpr int bytes = K * (int)sizeof(T);
          if constexpr (bytes >= 16) {
            constexpr int per = 16 / (int)sizeof(T);
#pragma unroll
            for (int i = 0; i < bytes / 16; ++i) {
              const int4 q = *(reinterpret_cast<const int4*>(sbp) + i);
              const T* t = reinterpret_cast<const T*>(&q);
#pragma unroll
              for (int j = 0; j < per; ++j) bb[i * per + j] = t[j];
            }
          } else {
#pragma unroll
            for (int c = 0; c < K; ++c) bb[c] = sbp[c];
          }
#pragma unroll
          for (int c = 0; c < K; ++c) o[c] = xx[c] + (w * (bb[c] - acc[c])) / d;   // JacobiGNN.py:119
        }
        if (r < r1) store_vec<T, K>(xout + (size_t)r * K, o);
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(empty + s);
        push_pending(tile, sw & 1, o[0]);   // also publishes the tile stored kMsDefer tiles ago
        if (++s == S) { s = 0; phase ^= 1u; }
        if constexpr (HALO) {
          if (bnd && has_comm) {  // tell the communication CTA that this boundary tile of this sweep is stored
            publish_all();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid == 0) {
              __threadfence();
              atomicAdd(h.done_counter + 4 * (sw & 1), 1u);
            }
          }
        }
      }
    }
    publish_all();
  }
  // ---- exit: the last CTA advances the epoch (and the push counters) for the next launch
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(m.ticket, 1u);
    if (t == gridDim.x - 1) {
      *m.epoch = base + (uint32_t)(kConsumerWarps * nsweeps);
      *m.ticket = 0u;
      if constexpr (HALO) {
        // sweeps 0, 2, .. produce buffer B (dir[0]); sweeps 1, 3, .. buffer A (dir[1])
        if (h.dir[0].pushed_counter) *h.dir[0].pushed_counter += (uint32_t)((nsweeps + 1) / 2);
        if (h.dir[1].pushed_counter) *h.dir[1].pushed_counter += (uint32_t)(nsweeps / 2);
      }
      __threadfence();
    }
  }
}

}  // namespace glab
