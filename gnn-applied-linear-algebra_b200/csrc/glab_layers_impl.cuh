// glab_layers_impl.cuh -- fused SpMV-bearing layer steps (MatVec, residual, Jacobi, Chebyshev,
// power method, Rayleigh quotient, x^T W x) on the row-tile machinery of glab_tiles.cuh.
// Each extern "C" entry is ONE kernel launch that replaces one or more reference GN blocks
// (gathers + edge update + scatter + vertex update); citations are in include/glab.h.
#include <cstdlib>
#include "glab_pipe.cuh"
#include "glab_multisweep.cuh"

namespace glab {

// ---------------------------------------------------------------- epilogues (row sums in regs)
struct NoState {};

// K contiguous elements of a TMA-staged stream row in shared memory, read with 16-byte loads
// (a per-column scalar walk has an 8-way bank conflict for K = 8: lanes are 32 bytes apart).
template <typename T, int K>
__device__ __forceinline__ void lds_row(T (&d)[K], const T* p) {
  constexpr int bytes = K * (int)sizeof(T);
  if constexpr (bytes >= 16) {
    constexpr int per = 16 / (int)sizeof(T);
#pragma unroll
    for (int i = 0; i < bytes / 16; ++i) {
      const int4 q = *(reinterpret_cast<const int4*>(p) + i);
      const T* t = reinterpret_cast<const T*>(&q);
#pragma unroll
      for (int j = 0; j < per; ++j) d[i * per + j] = t[j];
    }
  } else if constexpr (bytes == 8) {
    const int2 q = *reinterpret_cast<const int2*>(p);
    const T* t = reinterpret_cast<const T*>(&q);
#pragma unroll
    for (int j = 0; j < K; ++j) d[j] = t[j];
  } else {
    d[0] = p[0];
  }
}

template <typename T, int K> struct EpiSpmm {  // y = A x            (MatVecGNN.py:109-114)
  T* y;
  using State = NoState;
  __device__ void init(State&) const {}
  __device__ void row(State&, int r, const T (&acc)[K]) const { store_vec<T, K>(y + (size_t)r * K, acc); }
  static constexpr int kStreams = 0;
  __host__ __device__ const T* stream_ptr(int) const { return nullptr; }
  __host__ __device__ int stream_width(int) const { return 0; }
  __device__ void row_staged(State& s, int r, const T (&acc)[K], const T*, const T*, const T*) const { row(s, r, acc); }
  __device__ void finish(State&) const {}
};

template <typename T, int K> struct EpiResidual {  // r = b - A x  (GNNResidual.py:115)
  const T* b;
  T* out;
  using State = NoState;
  __device__ void init(State&) const {}
  __device__ void row(State&, int r, const T (&acc)[K]) const {
    T bb[K], o[K];
    load_vec<T, K>(bb, b + (size_t)r * K);
#pragma unroll
    for (int c = 0; c < K; ++c) o[c] = bb[c] - acc[c];
    store_vec<T, K>(out + (size_t)r * K, o);
  }
  static constexpr int kStreams = 1;
  __host__ __device__ const T* stream_ptr(int) const { return b; }
  __host__ __device__ int stream_width(int) const { return K; }
  __device__ void row_staged(State&, int r, const T (&acc)[K], const T* sb, const T*, const T*) const {
    T bb[K], o[K];
    lds_row<T, K>(bb, sb);
#pragma unroll
    for (int c = 0; c < K; ++c) o[c] = bb[c] - acc[c];
    store_vec<T, K>(out + (size_t)r * K, o);
  }
  __device__ void finish(State&) const {}
};

template <typename T, int K> struct EpiAdd {  // out = b + A x  (coarse-grid correction x + P xc, VCycle.py:226)
  const T* b;
  T* out;
  using State = NoState;
  __device__ void init(State&) const {}
  __device__ void row(State&, int r, const T (&acc)[K]) const {
    T bb[K], o[K];
    load_vec<T, K>(bb, b + (size_t)r * K);
#pragma unroll
    for (int c = 0; c < K; ++c) o[c] = bb[c] + acc[c];
    store_vec<T, K>(out + (size_t)r * K, o);
  }
  static constexpr int kStreams = 1;
  __host__ __device__ const T* stream_ptr(int) const { return b; }
  __host__ __device__ int stream_width(int) const { return K; }
  __device__ void row_staged(State&, int r, const T (&acc)[K], const T* sb, const T*, const T*) const {
    T bb[K], o[K];
    lds_row<T, K>(bb, sb);
#pragma unroll
    for (int c = 0; c < K; ++c) o[c] = bb[c] + acc[c];
    store_vec<T, K>(out + (size_t)r * K, o);
  }
  __device__ void finish(State&) const {}
};

template <typename T, int K> struct EpiJacobi {  // x + w*(b - Ax)/d  (JacobiGNN.py:119)
  const T* diag;
  const T* b;
  const T* x;
  T* xo;
  const T* omega;
  struct State { T w; };
  __device__ void init(State& s) const { s.w = __ldg(omega); }
  __device__ void row(State& s, int r, const T (&acc)[K]) const {
    T bb[K], xx[K], o[K];
    const T d = __ldg(diag + r);
    load_vec<T, K>(bb, b + (size_t)r * K);
    load_vec<T, K>(xx, x + (size_t)r * K);
#pragma unroll
    for (int c = 0; c < K; ++c) o[c] = xx[c] + (s.w * (bb[c] - acc[c])) / d;
    store_vec<T, K>(xo + (size_t)r * K, o);
  }
  static constexpr int kStreams = 3;
  __host__ __device__ const T* stream_ptr(int i) const { return i == 0 ? diag : (i == 1 ? b : x); }
  __host__ __device__ int stream_width(int i) const { return i == 0 ? 1 : K; }
  __device__ void row_staged(State& s, int r, const T (&acc)[K], const T* sd, const T* sb, const T* sx) const {
    T o[K], bb[K], xx[K];
    const T d = sd[0];
    lds_row<T, K>(bb, sb);
    lds_row<T, K>(xx, sx);
#pragma unroll
    for (int c = 0; c < K; ++c) o[c] = xx[c] + (s.w * (bb[c] - acc[c])) / d;
    store_vec<T, K>(xo + (size_t)r * K, o);
  }
  __device__ void finish(State&) const {}
};

template <typename T, int K> struct EpiChebyFirst {  // ChebyGNN.py:117, :160-161
  const T* b;
  const T* x;
  T* xo;
  T* r_;
  T* p_;
  const T* alpha;
  struct State { T a; };
  __device__ void init(State& s) const { s.a = __ldg(alpha); }
  __device__ void row(State& s, int r, const T (&acc)[K]) const {
    T bb[K], xx[K], rr[K], o[K];
    load_vec<T, K>(bb, b + (size_t)r * K);
    load_vec<T, K>(xx, x + (size_t)r * K);
#pragma unroll
    for (int c = 0; c < K; ++c) {
      rr[c] = bb[c] - acc[c];
      o[c] = xx[c] + s.a * rr[c];
    }
    store_vec<T, K>(r_ + (size_t)r * K, rr);
    store_vec<T, K>(p_ + (size_t)r * K, rr);
    store_vec<T, K>(xo + (size_t)r * K, o);
  }
  static constexpr int kStreams = 2;
  __host__ __device__ const T* stream_ptr(int i) const { return i == 0 ? b : x; }
  __host__ __device__ int stream_width(int) const { return K; }
  __device__ void row_staged(State& s, int r, const T (&acc)[K], const T* sb, const T* sx, const T*) const {
    T rr[K], o[K], bb[K], xx[K];
    lds_row<T, K>(bb, sb);
    lds_row<T, K>(xx, sx);
#pragma unroll
    for (int c = 0; c < K; ++c) {
      rr[c] = bb[c] - acc[c];
      o[c] = xx[c] + s.a * rr[c];
    }
    store_vec<T, K>(r_ + (size_t)r * K, rr);
    store_vec<T, K>(p_ + (size_t)r * K, rr);
    store_vec<T, K>(xo + (size_t)r * K, o);
  }
  __device__ void finish(State&) const {}
};

template <typename T, int K> struct EpiChebyNext {  // ChebyGNN.py:214, :240-241
  const T* p_in;
  T* p_out;
  T* r_;
  T* x_;
  const T* alpha_old;
  const T* alpha;
  const T* beta;
  struct State { T ao, a, b; };
  __device__ void init(State& s) const {
    s.ao = __ldg(alpha_old);
    s.a = __ldg(alpha);
    s.b = __ldg(beta);
  }
  __device__ void row(State& s, int r, const T (&acc)[K]) const {
    T pp[K], rr[K], xx[K];
    load_vec<T, K>(pp, p_in + (size_t)r * K);
    load_vec_rw<T, K>(rr, r_ + (size_t)r * K);
    load_vec_rw<T, K>(xx, x_ + (size_t)r * K);
#pragma unroll
    for (int c = 0; c < K; ++c) {
      rr[c] = rr[c] - s.ao * acc[c];
      pp[c] = rr[c] + s.b * pp[c];
      xx[c] = xx[c] + s.a * pp[c];
    }
    store_vec<T, K>(r_ + (size_t)r * K, rr);
    store_vec<T, K>(p_out + (size_t)r * K, pp);
    store_vec<T, K>(x_ + (size_t)r * K, xx);
  }
  static constexpr int kStreams = 3;
  __host__ __device__ const T* stream_ptr(int i) const { return i == 0 ? p_in : (i == 1 ? r_ : x_); }
  __host__ __device__ int stream_width(int) const { return K; }
  __device__ void row_staged(State& s, int r, const T (&acc)[K], const T* sp, const T* sr, const T* sx) const {
    T pp[K], rr[K], xx[K];
    lds_row<T, K>(pp, sp);
    lds_row<T, K>(rr, sr);
    lds_row<T, K>(xx, sx);
#pragma unroll
    for (int c = 0; c < K; ++c) {
      rr[c] = rr[c] - s.ao * acc[c];
      pp[c] = rr[c] + s.b * pp[c];
      xx[c] = xx[c] + s.a * pp[c];
    }
    store_vec<T, K>(r_ + (size_t)r * K, rr);
    store_vec<T, K>(p_out + (size_t)r * K, pp);
    store_vec<T, K>(x_ + (size_t)r * K, xx);
  }
  __device__ void finish(State&) const {}
};

// Sum over the ranks of the partial sums the previous launch published (glab_peer_reduce): one thread
// per CTA waits for all arrival counters and adds the partials in rank order; the CTA shares the result.
__device__ __forceinline__ void peer_sums(const PeerReduceDev& pr, double& t0, double& t1) {
  __shared__ double sh[2];
  if (threadIdx.x == 0) {
    pdl_wait();                                // the publishing launch precedes this one on the stream
    const uint32_t want = *reinterpret_cast<const volatile uint32_t*>(pr.parity_counter);   // own publishes so far
    const uint32_t par = (want - 1u) & 1u;     // parity of the most recent publish
    for (int q = 0; q < pr.world; ++q) {
      SpinGuard guard(pr.timeout_ns);
      while ((int32_t)(ld_acquire_sys(pr.flag_local + 4 * q) - want) < 0) {
        __nanosleep(32);
        if (guard.expired()) { flag_timeout(pr.status, GLAB_STATUS_TIMEOUT_PEER); break; }
      }
    }
    double a = 0.0, b = 0.0;
    for (int q = 0; q < pr.world; ++q) {
      const volatile double* m = pr.mail_local + ((size_t)par * GLAB_MAX_PEERS + q) * 2;
      a += m[0];
      b += m[1];
    }
    sh[0] = a;
    sh[1] = b;
  }
  __syncthreads();
  t0 = sh[0];
  t1 = sh[1];
}

template <typename T> struct EpiPower {  // PowerMethodGNN.py:156, :124, :183, :205
  T* y;
  const double* sumsq_in;
  double* sumsq_out;
  void* ws;
  PeerReduceDev pr;            // world == 0: rank-local sums (single GPU, or the caller reduces them)
  struct State { T n; bool scale; double s; };
  // called by EVERY thread of the CTA before the roles split (it contains a CTA barrier)
  __device__ void prologue(State& s) const {
    s.scale = sumsq_in != nullptr;
    s.n = T(1);
    if (s.scale && pr.world > 0) {
      double t0, t1;
      peer_sums(pr, t0, t1);
      s.n = (T)sqrt(t0);
    }
  }
  __device__ void init(State& s) const {
    s.scale = sumsq_in != nullptr;
    if (!(s.scale && pr.world > 0)) s.n = s.scale ? (T)sqrt(__ldg(sumsq_in)) : T(1);
    s.s = 0.0;
  }
  __device__ void row(State& s, int r, const T (&acc)[1]) const {
    const T v = s.scale ? acc[0] / s.n : acc[0];
    y[r] = v;
    const T sq = v * v;
    s.s += (double)sq;
  }
  static constexpr int kStreams = 0;
  __host__ __device__ const T* stream_ptr(int) const { return nullptr; }
  __host__ __device__ int stream_width(int) const { return 0; }
  __device__ void row_staged(State& s, int r, const T (&acc)[1], const T*, const T*, const T*) const { row(s, r, acc); }
  __device__ void finish(State& s) const { grid_reduce2(s.s, 0.0, ws, sumsq_out, pr.world > 0 ? &pr : nullptr); }
};
template <typename T> struct has_prologue<EpiPower<T>> { static constexpr bool value = true; };

template <typename T> struct EpiRayleigh {  // PowerMethodGNN.py:205, :235, :264, :124, :292
  const T* b_in;
  T* b_out;
  T* y_out;
  const double* sumsq_in;
  double* sums_out;
  void* ws;
  PeerReduceDev pr;            // consumes the norm the last power step published; its own sums stay rank-local
  struct State { T n; bool scale; double s0, s1; };
  __device__ void prologue(State& s) const {
    s.scale = sumsq_in != nullptr;
    s.n = T(1);
    if (s.scale && pr.world > 0) {
      double t0, t1;
      peer_sums(pr, t0, t1);
      s.n = (T)sqrt(t0);
    }
  }
  __device__ void init(State& s) const {
    s.scale = sumsq_in != nullptr;
    if (!(s.scale && pr.world > 0)) s.n = s.scale ? (T)sqrt(__ldg(sumsq_in)) : T(1);
    s.s0 = s.s1 = 0.0;
  }
  __device__ void row(State& s, int r, const T (&acc)[1]) const {
    const T bi = __ldg(b_in + r);
    const T bn = s.scale ? bi / s.n : bi;
    const T ab = s.scale ? acc[0] / s.n : acc[0];
    const T yA = bn * ab;
    const T sq = bn * bn;
    b_out[r] = bn;
    y_out[r] = sq;
    s.s0 += (double)yA;
    s.s1 += (double)sq;
  }
  static constexpr int kStreams = 1;
  __host__ __device__ const T* stream_ptr(int) const { return b_in; }
  __host__ __device__ int stream_width(int) const { return 1; }
  __device__ void row_staged(State& s, int r, const T (&acc)[1], const T* sbi, const T*, const T*) const {
    const T bi = sbi[0];
    const T bn = s.scale ? bi / s.n : bi;
    const T ab = s.scale ? acc[0] / s.n : acc[0];
    const T yA = bn * ab;
    const T sq = bn * bn;
    b_out[r] = bn;
    y_out[r] = sq;
    s.s0 += (double)yA;
    s.s1 += (double)sq;
  }
  __device__ void finish(State& s) const { grid_reduce2(s.s0, s.s1, ws, sums_out); }
};
template <typename T> struct has_prologue<EpiRayleigh<T>> { static constexpr bool value = true; };

template <typename T> struct EpiXtAx {  // MatrixWeightedNorm.py:107-109
  const T* x;
  double* sums_out;
  void* ws;
  struct State { double s; };
  __device__ void init(State& s) const { s.s = 0.0; }
  __device__ void row(State& s, int r, const T (&acc)[1]) const {
    const T v = __ldg(x + r) * acc[0];
    s.s += (double)v;
  }
  static constexpr int kStreams = 1;
  __host__ __device__ const T* stream_ptr(int) const { return x; }
  __host__ __device__ int stream_width(int) const { return 1; }
  __device__ void row_staged(State& s, int r, const T (&acc)[1], const T* sx, const T*, const T*) const {
    const T v = sx[0] * acc[0];
    s.s += (double)v;
  }
  __device__ void finish(State& s) const { grid_reduce2(s.s, 0.0, ws, sums_out); }
};

// ---------------------------------------------------------------- launch
struct Tuning {
  int rpt;        // rows per thread for K == 1 kernels (1 or 2)
  int cap_max;    // staging capacity upper bound in slots
  int persist;    // CTAs per SM for persistent (reducing) kernels
  int pipe;       // 1 = use the TMA pipeline kernel when the tile fits (default), 0 = never
  int stages;     // 0 = auto, else forced ring depth
  int ctas;       // 0 = auto (occupancy API), else forced CTAs per SM for the pipeline
  int pdl;        // 1 = programmatic dependent launch between consecutive pipeline kernels
  int ms;         // several Jacobi sweeps per launch (glab_jacobi_sweeps_*): 0 = never the multi-sweep kernel,
                  // 1 = for operators of at most kMsAutoTiles tiles (where it wins), 2 = always
};
// Measured on B200 (profiles/r02_multisweep.md): 10 sweeps on 0.26 M rows take 0.0067 ms per sweep in one
// multi-sweep launch vs 0.0147 ms as ten launches; break-even near 1 M rows; above, the per-tile cost of
// publishing completions (an L2 read-back per tile) outweighs the saved launch ramps.
constexpr int kMsAutoTiles = 4096;

// Ring depth chosen automatically: at most 4 stages, 3 for the single-column Jacobi sweep.  Measured on B200
// (profiles/r02_kernel_rooflines_stages{3,4}.jsonl, L4096 fp32, registers allow 4 CTAs per SM either way): Jacobi
// 4 stages 0.1347 ms, 3 stages 0.1285 ms, 2 stages 0.1347 ms per sweep -- the fourth stage buys no more bytes in
// flight than the memory system needs, but its 11 KB x 4 CTAs come out of the L1 that serves the x[col] gathers
// (shared memory and L1 share one 256 KB array).  Capping EVERY epilogue at 3 made the isolated kernels faster too
// (cheby_next 0.154 -> 0.145 ms) but the chained smoothing pass SLOWER (2.10 vs 1.95 ms per step: the Chebyshev
// launches that follow the sweeps lose more than the sweeps gain), so only the sweep is capped: 1.88 ms per step,
// 624 Gnnz/s (bench.py).  GLAB_STAGES overrides everything.
constexpr int kMaxAutoStages = 4;
template <class Epi> struct auto_stage_cap { static constexpr int value = kMaxAutoStages; };
// single-column Jacobi sweep: 3 stages (see the measurements in the comment above)
template <typename T> struct auto_stage_cap<EpiJacobi<T, 1>> { static constexpr int value = 3; };

// Bound of the in-kernel waits: caller's value, else GLAB_SPIN_TIMEOUT_MS, else 20 s; < 0 = forever.
static unsigned long long spin_timeout_ns(int64_t timeout_ms) {
  static const int64_t dflt = [] {
    const char* e = getenv("GLAB_SPIN_TIMEOUT_MS");
    return e ? (int64_t)atoll(e) : (int64_t)20000;
  }();
  const int64_t ms = timeout_ms == 0 ? dflt : timeout_ms;
  return ms < 0 ? 0ull : (unsigned long long)ms * 1000000ull;
}

static const Tuning& tuning() {
  static Tuning t = [] {
    Tuning v{1, 4096, 8, 1, 0, 0, 1, 1};
    if (const char* e = getenv("GLAB_MS")) { int c = atoi(e); v.ms = c < 0 ? 0 : (c > 2 ? 2 : c); }
    if (const char* e = getenv("GLAB_RPT")) v.rpt = atoi(e) == 2 ? 2 : 1;
    if (const char* e = getenv("GLAB_CAP")) { int c = atoi(e); if (c >= 256 && c <= 8192) v.cap_max = c & ~31; }
    if (const char* e = getenv("GLAB_PIPE")) v.pipe = atoi(e) != 0;
    if (const char* e = getenv("GLAB_STAGES")) { int c = atoi(e); if (c >= 2 && c <= 8) v.stages = c; }
    if (const char* e = getenv("GLAB_CTAS")) { int c = atoi(e); if (c >= 1 && c <= 8) v.ctas = c; }
    if (const char* e = getenv("GLAB_PDL")) v.pdl = atoi(e) != 0;
    if (const char* e = getenv("GLAB_PERSIST")) { int c = atoi(e); if (c >= 1 && c <= 16) v.persist = c; }
    return v;
  }();
  return t;
}

template <typename T, int K, int RPT, class Epi>
static int launch_tiles(const glab_plan* p, const T* vals, const T* x, const Epi& epi,
                        int64_t row_begin, int64_t row_end, bool persistent, void* stream) {
  constexpr int R = kThreads * RPT;
  const int64_t nrows = row_end - row_begin;
  const int ntiles = (int)((nrows + R - 1) / R);
  int64_t want = (int64_t)R * (p->max_row_nnz > 0 ? p->max_row_nnz : 1);
  int cap = (int)(want < tuning().cap_max ? want : tuning().cap_max);
  cap = (cap + 31) & ~31;
  const size_t smem = tile_smem_bytes<T>(cap, 1);
  auto kern = k_row_tiles<T, K, RPT, Epi>;
  static bool attr_done[kMaxDevices] = {};  // per template instantiation and device
  if (!attr_done[p->device % kMaxDevices]) {
    GLAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    attr_done[p->device % kMaxDevices] = true;
  }
  int grid = ntiles;
  if (persistent) {
    int g = p->sm_count * tuning().persist;
    if (g > kMaxReduceBlocks) g = kMaxReduceBlocks;
    if (grid > g) grid = g;
  }
  if (grid < 1) grid = 1;
  TileArgs<T> a{p->rowptr, p->colidx, vals, (int)row_begin, (int)row_end, cap};
  kern<<<grid, kThreads, smem, as_stream(stream)>>>(a, x, epi, ntiles);
  return (int)cudaGetLastError();
}

constexpr int kNoPipe = -1000;

// Launch with the programmatic-stream-serialization attribute (PDL) so that consecutive sweeps
// overlap the next kernel's launch + prologue with the previous kernel's tail.
template <typename Kern, typename... Args>
static int launch_pdl(Kern kern, int grid, int block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, kern, args...);
}
// Cooperative launch: the driver starts the grid only when ALL of its CTAs can be resident at once and
// rejects one that can never be (cudaErrorCooperativeLaunchTooLarge).  For the kernels whose CTAs wait on
// each other (the fused halo steps) this turns "a CTA was not scheduled because another context / stream
// holds SMs" from a spin into a queued launch.  It excludes programmatic dependent launch.
template <typename Kern, typename... Args>
static int launch_coop(Kern kern, int grid, int block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, kern, args...);
}
// GLAB_HALO_COOP=1: launch the fused halo steps cooperatively (shared GPUs: MPS, concurrent streams, green
// contexts).  Default 0: on a GPU the process owns, grid = SMs x occupancy is resident by construction and
// the programmatic dependent launch between consecutive steps is worth ~2 us per step; the in-kernel waits
// are bounded either way (GLAB_STATUS_TIMEOUT_*).
static bool halo_coop() {
  static const bool v = [] { const char* e = getenv("GLAB_HALO_COOP"); return e && atoi(e) != 0; }();
  return v;
}
static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
template <typename T, int K, int U, class Epi>
static int launch_pipe_halo(const glab_plan*, const T*, const T*, const Epi&, void*, const glab_halo_step*);
template <typename T, int K, int U, class Epi, int IDX>
static int launch_pipe_impl(const glab_plan* p, const T* vals, const T* x, const Epi& epi, int64_t row_begin,
                            int64_t row_end, void* stream);
template <typename T, int K, int U, class Epi, int IDX>
static int launch_pipe_halo_impl(const glab_plan*, const T*, const T*, const Epi&, void*, const glab_halo_step*);

// How the column indices are streamed (k_row_pipe's IDX): 1 = 16-bit row-relative everywhere,
// 2 = per tile (tile16 flags; needs 256-aligned tiles), 0 = int32.
// The 16-bit index paths are compiled for fp32 and for the fp64 5-point kernel; the other fp64
// instantiations sit at their register cap and would spill, so they keep int32 indices.
template <typename T, int K, int U> constexpr bool idx16_ok() { return sizeof(T) == 4 || (K == 1 && U == 5); }

static inline int index_mode(const glab_plan* p, int64_t row_begin) {
  if (!p->coldelta || !p->tile16) return 0;
  if (p->tiles16 == p->tiles_total) return 1;
  return (row_begin % kThreads == 0) ? 2 : 0;
}

template <typename T, int K, int U, class Epi>
static int launch_pipe_u(const glab_plan*, const T*, const T*, const Epi&, int64_t, int64_t, void*,
                         const glab_halo_step*);

// U (gathers in flight per pass) follows the operator: 5- and 9-point stencils get an exact
// unrolled row; anything else 8 (k = 1), 4 (k = 2) or 2 (k >= 4).
template <typename T, int K, class Epi>
static int launch_pipe(const glab_plan* p, const T* vals, const T* x, const Epi& epi, int64_t row_begin,
                       int64_t row_end, void* stream, const glab_halo_step* h = nullptr) {
  if constexpr (K == 1) {
    if (p->max_row_nnz == 5) return launch_pipe_u<T, 1, 5>(p, vals, x, epi, row_begin, row_end, stream, h);
    if (p->max_row_nnz == 9) return launch_pipe_u<T, 1, 9>(p, vals, x, epi, row_begin, row_end, stream, h);
    return launch_pipe_u<T, 1, 8>(p, vals, x, epi, row_begin, row_end, stream, h);
  } else {
    return launch_pipe_u<T, K, (K == 2 ? 4 : 2)>(p, vals, x, epi, row_begin, row_end, stream, h);
  }
}
constexpr int kNoPipeUnused = 0;  // sentinel: operator does not fit the pipeline, use the generic kernel


template <typename T, class Epi>
static bool make_pipe_layout(const glab_plan* p, const Epi& epi, PipeLayout& L, int64_t& slots) {
  slots = (int64_t)kThreads * (p->max_row_nnz > 0 ? p->max_row_nnz : 1);
  if (slots > 24576) return false;
  int off = 0;
  L.off_row = off; off += round_up((kThreads + 1) * 4 + 32, 128);
  L.off_col = off; off += round_up((int)slots * 4 + 32, 128);
  L.off_val = off; off += round_up((int)slots * (int)sizeof(T) + 32, 128);
  for (int i = 0; i < kMaxStreams; ++i) {
    L.off_stream[i] = off;
    if (i < Epi::kStreams) off += round_up(kThreads * epi.stream_width(i) * (int)sizeof(T) + 32, 128);
  }
  L.stage_bytes = off;
  return true;
}

// Fused step + halo exchange launch (whole row block, boundary tiles last).  No fallback: an
// operator that does not fit the pipeline returns GLAB_E_ARG and the caller uses the
// separate wait / boundary / push kernels instead.
template <typename T, int K, int U, class Epi>
static int launch_pipe_halo(const glab_plan* p, const T* vals, const T* x, const Epi& epi, void* stream,
                            const glab_halo_step* hs) {
  // interior tiles of a row block have local, in-band columns: they stream 16-bit indices (IDX 2, or
  // IDX 1 when even the tiles that read the halo tail are within +-32767 of their rows)
  if constexpr (idx16_ok<T, K, U>()) {
    if (p->idx16_halo) {
      switch (index_mode(p, 0)) {
        case 1: return launch_pipe_halo_impl<T, K, U, Epi, 1>(p, vals, x, epi, stream, hs);
        case 2: return launch_pipe_halo_impl<T, K, U, Epi, 2>(p, vals, x, epi, stream, hs);
        default: break;
      }
    }
  }
  return launch_pipe_halo_impl<T, K, U, Epi, 0>(p, vals, x, epi, stream, hs);
}

template <typename T, int K, int U, class Epi, int IDX>
static int launch_pipe_halo_impl(const glab_plan* p, const T* vals, const T* x, const Epi& epi, void* stream,
                                 const glab_halo_step* hs) {
  if (hs->n_wait < 0 || hs->n_wait > GLAB_MAX_PEERS || hs->n_push < 0 || hs->n_push > GLAB_MAX_PEERS)
    return GLAB_E_ARG;
  if ((hs->n_wait > 0 && (!hs->wait_flags || !hs->wait_target)) || (hs->n_push > 0 && (!hs->push || !hs->push_src)) ||
      !hs->done_counter)
    return GLAB_E_ARG;
  const int64_t n = p->n_rows;
  int64_t ib = hs->interior_begin, ie = hs->interior_end;
  if (ib < 0 || ie < ib || ie > n) return GLAB_E_ARG;
  if (ib == ie) ib = ie = 0;  // no interior rows (tiny blocks): every tile is a boundary tile
  else if ((ib % kThreads) || (ie % kThreads && ie != n)) return GLAB_E_ARG;
  if (reinterpret_cast<uintptr_t>(p->rowptr) & 15) return GLAB_E_ARG;
  for (int i = 0; i < Epi::kStreams; ++i)
    if (reinterpret_cast<uintptr_t>(epi.stream_ptr(i)) & 15) return GLAB_E_ARG;
  PipeLayout L;
  int64_t slots;
  if (!make_pipe_layout<T>(p, epi, L, slots)) return GLAB_E_ARG;
  auto kern = k_row_pipe<T, K, U, Epi, true, IDX>;
  static int max_smem_dev[kMaxDevices] = {};
  int& max_smem = max_smem_dev[p->device % kMaxDevices];
  if (!max_smem) {
    cudaFuncAttributes fa;
    GLAB_CUDA(cudaFuncGetAttributes(&fa, kern));
    const int m = 227 * 1024 - (int)fa.sharedSizeBytes;
    GLAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, m));
    max_smem = m;
  }
  if (2 * L.stage_bytes + 128 > max_smem) return GLAB_E_ARG;
  int want_ctas = tuning().ctas ? tuning().ctas : 4;  // measured: shallow rings + more CTAs win for wide k too
  int stages = tuning().stages ? tuning().stages : (max_smem / want_ctas - 128) / L.stage_bytes;
  if (!tuning().stages && stages > auto_stage_cap<Epi>::value) stages = auto_stage_cap<Epi>::value;
  if (stages > 4) stages = 4;
  while (stages > 2 && (size_t)stages * L.stage_bytes + 128 > (size_t)max_smem) --stages;
  if (stages < 2) stages = 2;
  L.stages = stages;
  const size_t smem = (size_t)stages * L.stage_bytes + 128;
  int occ = 0;
  GLAB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kPipeThreads, smem));
  if (occ < 1) return GLAB_E_ARG;
  if (tuning().ctas && occ > tuning().ctas) occ = tuning().ctas;
  const int ntiles = (int)((n + kThreads - 1) / kThreads);
  HaloCtl h;
  h.int_tile0 = (int)(ib / kThreads);
  h.int_tiles = (int)((ie - ib + kThreads - 1) / kThreads);
  h.lead_tiles = h.int_tile0;
  h.trail_tile0 = h.int_tile0 + h.int_tiles;
  h.n_wait = hs->n_wait;
  h.n_push = hs->n_push;
  h.push_k = K;
  for (int i = 0; i < GLAB_MAX_PEERS; ++i) {
    h.wait_flag[i] = i < hs->n_wait ? hs->wait_flags[i] : nullptr;
    if (i < hs->n_push) h.push[i] = hs->push[i];
    else h.push[i] = glab_push_desc{nullptr, -1, 0, nullptr, 0, nullptr};
  }
  h.wait_target = hs->wait_target;
  h.pushed_counter = hs->pushed_counter;
  h.push_src = hs->push_src;
  h.done_counter = hs->done_counter;
  h.status = hs->status;
  h.timeout_ns = spin_timeout_ns(hs->timeout_ms);
  int grid = p->sm_count * occ;   // all co-resident: the communication CTA (block 0) must run
  if (grid > kMaxReduceBlocks) grid = kMaxReduceBlocks;
  if (grid > ntiles + 1) grid = ntiles + 1;
  if (grid < 1) grid = 1;
  if (grid < 2 && hs->n_push > 0) grid = 2;   // a block without rows still has to publish its arrival
  TileArgs<T> a{p->rowptr, p->colidx, vals, 0, (int)n, (int)slots, IDX ? p->coldelta : nullptr,
                IDX == 2 ? p->tile16 : nullptr};
  if (halo_coop()) return launch_coop(kern, grid, kPipeThreads, smem, as_stream(stream), a, x, epi, ntiles, L, h);
  return launch_pdl(kern, grid, kPipeThreads, smem, as_stream(stream), tuning().pdl != 0, a, x, epi, ntiles, L, h);
}

// TMA pipeline launch.  Returns kNoPipe if the operator does not fit the pipeline (caller falls
// back to the generic chunked kernel), 0 on success, or an error code.

template <typename T, int K, int U, class Epi>
static int launch_pipe_u(const glab_plan* p, const T* vals, const T* x, const Epi& epi, int64_t row_begin,
                         int64_t row_end, void* stream, const glab_halo_step* h) {
  if (h) return launch_pipe_halo<T, K, U>(p, vals, x, epi, stream, h);
  if constexpr (idx16_ok<T, K, U>()) {
    switch (index_mode(p, row_begin)) {
      case 1: return launch_pipe_impl<T, K, U, Epi, 1>(p, vals, x, epi, row_begin, row_end, stream);
      case 2: return launch_pipe_impl<T, K, U, Epi, 2>(p, vals, x, epi, row_begin, row_end, stream);
      default: break;
    }
  }
  return launch_pipe_impl<T, K, U, Epi, 0>(p, vals, x, epi, row_begin, row_end, stream);
}

template <typename T, int K, int U, class Epi, int IDX>
static int launch_pipe_impl(const glab_plan* p, const T* vals, const T* x, const Epi& epi, int64_t row_begin,
                            int64_t row_end, void* stream) {
  if (!tuning().pipe) return kNoPipe;
  // 16-byte granule preconditions of the bulk copies (see k_row_pipe)
  if ((reinterpret_cast<uintptr_t>(p->rowptr) & 15) || ((row_begin * 4) & 15)) return kNoPipe;
  for (int i = 0; i < Epi::kStreams; ++i) {
    if (reinterpret_cast<uintptr_t>(epi.stream_ptr(i)) & 15) return kNoPipe;
    if ((row_begin * epi.stream_width(i) * (int64_t)sizeof(T)) & 15) return kNoPipe;
  }
  const int64_t slots = (int64_t)kThreads * (p->max_row_nnz > 0 ? p->max_row_nnz : 1);
  if (slots > 24576) return kNoPipe;
  PipeLayout L;
  int off = 0;
  L.off_row = off; off += round_up((kThreads + 1) * 4 + 32, 128);
  L.off_col = off; off += round_up((int)slots * (IDX == 1 ? 2 : 4) + 32, 128);
  L.off_val = off; off += round_up((int)slots * (int)sizeof(T) + 32, 128);
  for (int i = 0; i < kMaxStreams; ++i) {
    L.off_stream[i] = off;
    if (i < Epi::kStreams) off += round_up(kThreads * epi.stream_width(i) * (int)sizeof(T) + 32, 128);
  }
  L.stage_bytes = off;
  auto kern = k_row_pipe<T, K, U, Epi, false, IDX>;
  static int max_smem_dev[kMaxDevices] = {};  // per instantiation and device: 227 KB minus static smem
  int& max_smem = max_smem_dev[p->device % kMaxDevices];
  if (!max_smem) {
    cudaFuncAttributes fa;
    GLAB_CUDA(cudaFuncGetAttributes(&fa, kern));
    const int m = 227 * 1024 - (int)fa.sharedSizeBytes;
    GLAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, m));
    max_smem = m;
  }
  if (2 * L.stage_bytes + 128 > max_smem) return kNoPipe;
  // ring depth: enough stages that (CTAs/SM x stages) keeps >= ~128 KB in flight per SM, within smem
  int want_ctas = tuning().ctas ? tuning().ctas : 4;  // measured: shallow rings + more CTAs win for wide k too
  int stages = tuning().stages;
  if (!stages) {
    stages = (max_smem / want_ctas - 128) / L.stage_bytes;
    if (stages > auto_stage_cap<Epi>::value) stages = auto_stage_cap<Epi>::value;
  }
  while (stages > 2 && (size_t)stages * L.stage_bytes + 128 > (size_t)max_smem) --stages;
  if (stages < 2) stages = 2;
  L.stages = stages;
  const size_t smem = (size_t)stages * L.stage_bytes + 128;
  int occ = 0;
  GLAB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kPipeThreads, smem));
  if (occ < 1) return kNoPipe;
  if (tuning().ctas && occ > tuning().ctas) occ = tuning().ctas;
  const int64_t nrows = row_end - row_begin;
  const int ntiles = (int)((nrows + kThreads - 1) / kThreads);
  int grid = p->sm_count * occ;
  if (grid > kMaxReduceBlocks) grid = kMaxReduceBlocks;
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) grid = 1;
  TileArgs<T> a{p->rowptr, p->colidx, vals, (int)row_begin, (int)row_end, (int)slots,
                IDX ? p->coldelta : nullptr, IDX == 2 ? p->tile16 : nullptr};
  return launch_pdl(kern, grid, kPipeThreads, smem, as_stream(stream), tuning().pdl != 0, a, x, epi, ntiles, L,
                    NoHalo{});
}

static int check_common(const glab_plan* p, const void* vals, const void* x, int64_t rb, int64_t re) {
  if (!p || !x) return GLAB_E_ARG;
  if (p->nnz > 0 && !vals) return GLAB_E_ARG;
  if (rb < 0 || re < rb || re > p->n_rows) return GLAB_E_ARG;
  return 0;
}

// dispatch on K (and RPT for K == 1)
template <typename T, template <typename, int> class EpiK, class Make>
static int dispatch_k(const glab_plan* p, const T* vals, const T* x, int k, int64_t rb, int64_t re,
                      void* stream, Make make, const glab_halo_step* h = nullptr) {
  if (rb == re && !h) return 0;
  {
    int rc = kNoPipe;
    switch (k) {
      case 1: rc = launch_pipe<T, 1>(p, vals, x, make(EpiK<T, 1>{}), rb, re, stream, h); break;
      case 2: rc = launch_pipe<T, 2>(p, vals, x, make(EpiK<T, 2>{}), rb, re, stream, h); break;
      case 4: rc = launch_pipe<T, 4>(p, vals, x, make(EpiK<T, 4>{}), rb, re, stream, h); break;
      case 8: rc = launch_pipe<T, 8>(p, vals, x, make(EpiK<T, 8>{}), rb, re, stream, h); break;
      default: return GLAB_E_ARG;
    }
    if (rc != kNoPipe || h) return rc;
  }
  switch (k) {
    case 1:
      if (tuning().rpt == 2) return launch_tiles<T, 1, 2>(p, vals, x, make(EpiK<T, 1>{}), rb, re, false, stream);
      return launch_tiles<T, 1, 1>(p, vals, x, make(EpiK<T, 1>{}), rb, re, false, stream);
    case 2: return launch_tiles<T, 2, 1>(p, vals, x, make(EpiK<T, 2>{}), rb, re, false, stream);
    case 4: return launch_tiles<T, 4, 1>(p, vals, x, make(EpiK<T, 4>{}), rb, re, false, stream);
    case 8: return launch_tiles<T, 8, 1>(p, vals, x, make(EpiK<T, 8>{}), rb, re, false, stream);
    default: return GLAB_E_ARG;
  }
}

template <typename T>
static int spmm(const glab_plan* p, const T* vals, const T* x, int k, T* y, int64_t rb, int64_t re,
                void* stream, const glab_halo_step* h = nullptr) {
  int rc = check_common(p, vals, x, rb, re);
  if (rc) return rc;
  if (!y || (const void*)y == (const void*)x) return GLAB_E_ARG;
  return dispatch_k<T, EpiSpmm>(p, vals, x, k, rb, re, stream, [&](auto e) { e.y = y; return e; }, h);
}

template <typename T>
static int residual(const glab_plan* p, const T* vals, const T* x, const T* b, int k, T* r,
                    int64_t rb, int64_t re, void* stream, const glab_halo_step* h = nullptr) {
  int rc = check_common(p, vals, x, rb, re);
  if (rc) return rc;
  if (!b || !r || (const void*)r == (const void*)x) return GLAB_E_ARG;
  return dispatch_k<T, EpiResidual>(p, vals, x, k, rb, re, stream, [&](auto e) {
    e.b = b; e.out = r; return e; }, h);
}

template <typename T>
static int spmm_add(const glab_plan* p, const T* vals, const T* x, const T* b, int k, T* out,
                    int64_t rb, int64_t re, void* stream) {
  int rc = check_common(p, vals, x, rb, re);
  if (rc) return rc;
  if (!b || !out || (const void*)out == (const void*)x) return GLAB_E_ARG;
  return dispatch_k<T, EpiAdd>(p, vals, x, k, rb, re, stream, [&](auto e) {
    e.b = b; e.out = out; return e; });
}

template <typename T>
static int jacobi(const glab_plan* p, const T* vals, const T* diag, const T* b, const T* x_in,
                  T* x_out, const T* omega, int k, int64_t rb, int64_t re, void* stream,
                  const glab_halo_step* h = nullptr) {
  int rc = check_common(p, vals, x_in, rb, re);
  if (rc) return rc;
  if (!diag || !b || !x_out || !omega || x_out == x_in) return GLAB_E_ARG;
  return dispatch_k<T, EpiJacobi>(p, vals, x_in, k, rb, re, stream, [&](auto e) {
    e.diag = diag; e.b = b; e.x = x_in; e.xo = x_out; e.omega = omega; return e; }, h);
}

template <typename T>
static int cheby_first(const glab_plan* p, const T* vals, const T* b, const T* x_in, T* x_out, T* r,
                       T* pv, const T* alpha, int k, int64_t rb, int64_t re, void* stream,
                       const glab_halo_step* h = nullptr) {
  int rc = check_common(p, vals, x_in, rb, re);
  if (rc) return rc;
  if (!b || !x_out || !r || !pv || !alpha || x_out == x_in) return GLAB_E_ARG;
  return dispatch_k<T, EpiChebyFirst>(p, vals, x_in, k, rb, re, stream, [&](auto e) {
    e.b = b; e.x = x_in; e.xo = x_out; e.r_ = r; e.p_ = pv; e.alpha = alpha; return e; }, h);
}

template <typename T>
static int cheby_next(const glab_plan* p, const T* vals, const T* p_in, T* p_out, T* r, T* x,
                      const T* alpha_old, const T* alpha, const T* beta, int k, int64_t rb,
                      int64_t re, void* stream, const glab_halo_step* h = nullptr) {
  int rc = check_common(p, vals, p_in, rb, re);
  if (rc) return rc;
  if (!p_out || !r || !x || !alpha_old || !alpha || !beta || p_out == p_in) return GLAB_E_ARG;
  return dispatch_k<T, EpiChebyNext>(p, vals, p_in, k, rb, re, stream, [&](auto e) {
    e.p_in = p_in; e.p_out = p_out; e.r_ = r; e.x_ = x;
    e.alpha_old = alpha_old; e.alpha = alpha; e.beta = beta; return e; }, h);
}


// ---------------------------------------------------------------- multi-sweep Jacobi launch
// Returns kNoPipe when the multi-sweep kernel cannot run this operator (caller loops over single sweeps).
template <typename T, int K, int U, int IDX, bool HALO>
static int launch_jacobi_ms_impl(const glab_plan* p, const T* vals, const T* diag, const T* b, T* xa, T* xb,
                                 const T* omega, int nsweeps, void* stream, const glab_halo_step* hab,
                                 const glab_halo_step* hba) {
  if ((reinterpret_cast<uintptr_t>(p->rowptr) & 15) || (reinterpret_cast<uintptr_t>(diag) & 15) ||
      (reinterpret_cast<uintptr_t>(b) & 15) || (reinterpret_cast<uintptr_t>(xa) & 15) ||
      (reinterpret_cast<uintptr_t>(xb) & 15) || !p->ms_state)
    return kNoPipe;
  const int64_t slots = (int64_t)kThreads * (p->max_row_nnz > 0 ? p->max_row_nnz : 1);
  if (slots > 24576) return kNoPipe;
  PipeLayout L;
  int off = 0;
  L.off_row = off; off += round_up((kThreads + 1) * 4 + 32, 128);
  L.off_col = off; off += round_up((int)slots * (IDX == 1 ? 2 : 4) + 32, 128);
  L.off_val = off; off += round_up((int)slots * (int)sizeof(T) + 32, 128);
  L.off_stream[0] = off; off += round_up(kThreads * (int)sizeof(T) + 32, 128);
  L.off_stream[1] = off; off += round_up(kThreads * K * (int)sizeof(T) + 32, 128);
  L.off_stream[2] = off;
  L.stage_bytes = off;
  auto kern = k_jacobi_ms<T, K, U, IDX, HALO>;
  static int max_smem_dev[kMaxDevices] = {};
  static int coop_dev[kMaxDevices] = {};
  int& max_smem = max_smem_dev[p->device % kMaxDevices];
  if (!max_smem) {
    cudaFuncAttributes fa;
    GLAB_CUDA(cudaFuncGetAttributes(&fa, kern));
    const int m = 227 * 1024 - (int)fa.sharedSizeBytes;
    GLAB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, m));
    int coop = 0;
    GLAB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, p->device));
    coop_dev[p->device % kMaxDevices] = coop ? 1 : -1;
    max_smem = m;
  }
  if (coop_dev[p->device % kMaxDevices] < 0) return kNoPipe;
  if (2 * L.stage_bytes + 128 > max_smem) return kNoPipe;
  const int want_ctas = tuning().ctas ? tuning().ctas : 4;
  int stages = tuning().stages ? tuning().stages : (max_smem / want_ctas - 128) / L.stage_bytes;
  if (!tuning().stages && stages > kMaxAutoStages) stages = kMaxAutoStages;
  if (stages > 4) stages = 4;
  while (stages > 2 && (size_t)stages * L.stage_bytes + 128 > (size_t)max_smem) --stages;
  if (stages < 2) stages = 2;
  L.stages = stages;
  const size_t smem = (size_t)stages * L.stage_bytes + 128;
  int occ = 0;
  GLAB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kPipeThreads, smem));
  if (occ < 1) return kNoPipe;
  if (tuning().ctas && occ > tuning().ctas) occ = tuning().ctas;
  const int64_t n = p->n_rows;
  const int ntiles = (int)((n + kThreads - 1) / kThreads);
  MsCtl m;
  m.tile_done = p->ms_state;
  m.epoch = p->ms_state + ntiles;
  m.ticket = reinterpret_cast<unsigned int*>(p->ms_state + ntiles + 4);
  m.dep = (p->band_local + kThreads - 1) / kThreads + 1;
  if (m.dep > ntiles || 2 * m.dep + 1 > 1024) m.dep = ntiles > 0 ? ntiles : 1;   // no useful band: one check per sweep
  {
    // TIMING EXPERIMENTS ONLY: GLAB_MS_NODEP=1 skips the inter-sweep dependency checks (results are then wrong)
    static const int nodep = [] { const char* e = getenv("GLAB_MS_NODEP"); return e ? atoi(e) : 0; }();
    if (nodep) m.dep = -1;
    static const int chunk = [] { const char* e = getenv("GLAB_MS_CHUNK"); int c = e ? atoi(e) : 4; return c < 1 ? 1 : (c > 64 ? 64 : c); }();
    m.chunk = chunk;
    // read per launch so that tests can switch it: formally fenced hand-off (slow) instead of the default
    const char* es = getenv("GLAB_MS_STRICT");
    m.strict = (es && atoi(es) != 0) ? 1 : 0;
  }
  m.status = nullptr;
  m.timeout_ns = spin_timeout_ns(0);
  typename std::conditional<HALO, MsHalo, MsNoHalo>::type h;
  int grid = p->sm_count * occ;    // every CTA co-resident (cooperative launch): they wait on one another
  bool comm = false;
  if constexpr (HALO) {
    const glab_halo_step* st[2] = {hab, hba};
    int64_t ib = hab->interior_begin, ie = hab->interior_end;
    if (hba->interior_begin != ib || hba->interior_end != ie || !hab->done_counter) return GLAB_E_ARG;
    if (ib < 0 || ie < ib || ie > n) return GLAB_E_ARG;
    if (ib == ie) ib = ie = 0;
    else if ((ib % kThreads) || (ie % kThreads && ie != n)) return GLAB_E_ARG;
    h.int_tile0 = (int)(ib / kThreads);
    h.int_tiles = (int)((ie - ib + kThreads - 1) / kThreads);
    h.lead_tiles = h.int_tile0;
    h.trail_tile0 = h.int_tile0 + h.int_tiles;
    h.done_counter = hab->done_counter;
    for (int d = 0; d < 2; ++d) {
      const glab_halo_step* q = st[d];
      if (q->n_wait < 0 || q->n_wait > GLAB_MAX_PEERS || q->n_push < 0 || q->n_push > GLAB_MAX_PEERS) return GLAB_E_ARG;
      if ((q->n_wait > 0 && (!q->wait_flags || !q->wait_target)) || (q->n_push > 0 && !q->push)) return GLAB_E_ARG;
      h.dir[d].n_wait = q->n_wait;
      h.dir[d].n_push = q->n_push;
      static const uint32_t zero_word = 0;
      (void)zero_word;
      for (int i = 0; i < GLAB_MAX_PEERS; ++i) {
        h.dir[d].wait_flag[i] = i < q->n_wait ? q->wait_flags[i] : nullptr;
        if (i < q->n_push) h.dir[d].push[i] = q->push[i];
        else h.dir[d].push[i] = glab_push_desc{nullptr, -1, 0, nullptr, 0, nullptr};
      }
      h.dir[d].wait_target = q->wait_target;
      h.dir[d].pushed_counter = q->pushed_counter;
      if (q->n_push > 0) comm = true;
      if (q->n_wait > 0 && !q->wait_target) return GLAB_E_ARG;
    }
    if (!hab->wait_target || !hba->wait_target) return GLAB_E_ARG;   // read unconditionally by the producer
    m.status = hab->status;
    m.timeout_ns = spin_timeout_ns(hab->timeout_ms);
  }
  if (grid > ntiles + (comm ? 1 : 0)) grid = ntiles + (comm ? 1 : 0);
  if (grid < 1) grid = 1;
  if (comm && grid < 2) grid = 2;
  TileArgs<T> a{p->rowptr, p->colidx, vals, 0, (int)n, (int)slots, IDX ? p->coldelta : nullptr,
                IDX == 2 ? p->tile16 : nullptr};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)kPipeThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, kern, a, xa, xb, diag, b, omega, nsweeps, ntiles, L, m, h);
}

template <typename T, int K, int U, bool HALO>
static int launch_jacobi_ms_u(const glab_plan* p, const T* vals, const T* diag, const T* b, T* xa, T* xb,
                              const T* omega, int nsweeps, void* stream, const glab_halo_step* hab,
                              const glab_halo_step* hba) {
  if constexpr (idx16_ok<T, K, U>()) {
    if (!HALO || p->idx16_halo) {
      switch (index_mode(p, 0)) {
        case 1: return launch_jacobi_ms_impl<T, K, U, 1, HALO>(p, vals, diag, b, xa, xb, omega, nsweeps, stream, hab, hba);
        case 2: return launch_jacobi_ms_impl<T, K, U, 2, HALO>(p, vals, diag, b, xa, xb, omega, nsweeps, stream, hab, hba);
        default: break;
      }
    }
  }
  return launch_jacobi_ms_impl<T, K, U, 0, HALO>(p, vals, diag, b, xa, xb, omega, nsweeps, stream, hab, hba);
}

template <typename T, bool HALO>
static int launch_jacobi_ms(const glab_plan* p, const T* vals, const T* diag, const T* b, T* xa, T* xb,
                            const T* omega, int k, int nsweeps, void* stream, const glab_halo_step* hab,
                            const glab_halo_step* hba) {
  switch (k) {
    case 1:
      if (p->max_row_nnz == 5) return launch_jacobi_ms_u<T, 1, 5, HALO>(p, vals, diag, b, xa, xb, omega, nsweeps, stream, hab, hba);
      if (p->max_row_nnz == 9) return launch_jacobi_ms_u<T, 1, 9, HALO>(p, vals, diag, b, xa, xb, omega, nsweeps, stream, hab, hba);
      return launch_jacobi_ms_u<T, 1, 8, HALO>(p, vals, diag, b, xa, xb, omega, nsweeps, stream, hab, hba);
    case 2: return launch_jacobi_ms_u<T, 2, 4, HALO>(p, vals, diag, b, xa, xb, omega, nsweeps, stream, hab, hba);
    case 4: return launch_jacobi_ms_u<T, 4, 2, HALO>(p, vals, diag, b, xa, xb, omega, nsweeps, stream, hab, hba);
    case 8: return launch_jacobi_ms_u<T, 8, 2, HALO>(p, vals, diag, b, xa, xb, omega, nsweeps, stream, hab, hba);
    default: return GLAB_E_ARG;
  }
}

// n_sweeps weighted-Jacobi sweeps, x ping-ponging xa -> xb -> xa ...: one multi-sweep launch when the
// operator fits that kernel, else one fused launch per sweep.  Same arithmetic either way.
template <typename T>
static int jacobi_sweeps(const glab_plan* p, const T* vals, const T* diag, const T* b, T* xa, T* xb,
                         const T* omega, int k, int nsweeps, void* stream, const glab_halo_step* hab,
                         const glab_halo_step* hba) {
  int rc = check_common(p, vals, xa, 0, p ? p->n_rows : 0);
  if (rc) return rc;
  if (!diag || !b || !xb || !omega || xa == xb || nsweeps < 0) return GLAB_E_ARG;
  if ((hab == nullptr) != (hba == nullptr)) return GLAB_E_ARG;
  if (nsweeps == 0) return 0;
  const int64_t ms_tiles = (p->n_rows + kThreads - 1) / kThreads;
  if (tuning().ms && tuning().pipe && nsweeps > 1 && p->n_rows > 0 && (tuning().ms == 2 || ms_tiles <= kMsAutoTiles)) {
    rc = hab ? launch_jacobi_ms<T, true>(p, vals, diag, b, xa, xb, omega, k, nsweeps, stream, hab, hba)
             : launch_jacobi_ms<T, false>(p, vals, diag, b, xa, xb, omega, k, nsweeps, stream, nullptr, nullptr);
    if (rc != kNoPipe) return rc;
  }
  for (int s = 0; s < nsweeps; ++s) {
    const T* xin = (s & 1) ? xb : xa;
    T* xout = (s & 1) ? xa : xb;
    rc = jacobi<T>(p, vals, diag, b, xin, xout, omega, k, 0, p->n_rows, stream, hab ? ((s & 1) ? hba : hab) : nullptr);
    if (rc) return rc;
  }
  return 0;
}

static PeerReduceDev peer_reduce_dev(const glab_halo_step* h) {
  PeerReduceDev d{};
  if (h && h->reduce && h->reduce->world > 0 && h->reduce->world <= GLAB_MAX_PEERS) {
    const glab_peer_reduce* r = h->reduce;
    d.world = r->world;
    d.rank = r->rank;
    d.mail_local = r->mail_local;
    d.flag_local = r->flag_local;
    for (int q = 0; q < GLAB_MAX_PEERS; ++q) {
      d.mail_peer[q] = q < r->world ? r->mail_peer[q] : nullptr;
      d.flag_peer[q] = q < r->world ? r->flag_peer[q] : nullptr;
    }
    d.parity_counter = r->parity_counter;
    d.status = h->status;
    d.timeout_ns = spin_timeout_ns(h->timeout_ms);
  }
  return d;
}

template <typename T, class Epi>
static int launch_reducing(const glab_plan* p, const T* vals, const T* x, const Epi& epi,
                           int64_t rb, int64_t re, void* stream, const glab_halo_step* h = nullptr) {
  // rb == re still launches one (empty) tile so the output sums are written (as zeros).
  {
    const int rc = launch_pipe<T, 1>(p, vals, x, epi, rb, re, stream, h);
    if (rc != kNoPipe || h) return rc;
  }
  if (tuning().rpt == 2) return launch_tiles<T, 1, 2>(p, vals, x, epi, rb, re, true, stream);
  return launch_tiles<T, 1, 1>(p, vals, x, epi, rb, re, true, stream);
}

template <typename T>
static int power_step(const glab_plan* p, const T* vals, const T* b_in, T* y, const double* ss_in,
                      double* ss_out, void* ws, int64_t rb, int64_t re, void* stream,
                      const glab_halo_step* h = nullptr) {
  int rc = check_common(p, vals, b_in, rb, re);
  if (rc) return rc;
  if (!y || !ss_out || !ws || y == b_in) return GLAB_E_ARG;
  EpiPower<T> e{y, ss_in, ss_out, ws, peer_reduce_dev(h)};
  return launch_reducing<T>(p, vals, b_in, e, rb, re, stream, h);
}

template <typename T>
static int rayleigh(const glab_plan* p, const T* vals, const T* b_in, T* b_out, T* y_out,
                    const double* ss_in, double* sums_out, void* ws, int64_t rb, int64_t re,
                    void* stream, const glab_halo_step* h = nullptr) {
  int rc = check_common(p, vals, b_in, rb, re);
  if (rc) return rc;
  if (!b_out || !y_out || !sums_out || !ws || b_out == b_in) return GLAB_E_ARG;
  EpiRayleigh<T> e{b_in, b_out, y_out, ss_in, sums_out, ws, peer_reduce_dev(h)};
  return launch_reducing<T>(p, vals, b_in, e, rb, re, stream, h);
}

template <typename T>
static int xtax(const glab_plan* p, const T* vals, const T* x, double* sums_out, void* ws,
                int64_t rb, int64_t re, void* stream) {
  int rc = check_common(p, vals, x, rb, re);
  if (rc) return rc;
  if (!sums_out || !ws) return GLAB_E_ARG;
  EpiXtAx<T> e{x, sums_out, ws};
  return launch_reducing<T>(p, vals, x, e, rb, re, stream);
}

}  // namespace glab

using namespace glab;

#ifdef GLAB_LAYERS_F32
extern "C" int64_t glab_reduce_workspace_bytes(void) { return 64 + (int64_t)kMaxReduceBlocks * 2 * 8; }
#endif

#define GLAB_INST(SUF, T)                                                                          \
  extern "C" int glab_spmm_##SUF(const glab_plan* p, const T* v, const T* x, int k, T* y,          \
                                 int64_t rb, int64_t re, void* s) {                                \
    return spmm<T>(p, v, x, k, y, rb, re, s);                                                      \
  }                                                                                                \
  extern "C" int glab_residual_##SUF(const glab_plan* p, const T* v, const T* x, const T* b, int k, \
                                     T* r, int64_t rb, int64_t re, void* s) {                      \
    return residual<T>(p, v, x, b, k, r, rb, re, s);                                               \
  }                                                                                                \
  extern "C" int glab_spmm_add_##SUF(const glab_plan* p, const T* v, const T* x, const T* b, int k, \
                                     T* o, int64_t rb, int64_t re, void* s) {                      \
    return spmm_add<T>(p, v, x, b, k, o, rb, re, s);                                               \
  }                                                                                                \
  extern "C" int glab_jacobi_##SUF(const glab_plan* p, const T* v, const T* d, const T* b,         \
                                   const T* xi, T* xo, const T* w, int k, int64_t rb, int64_t re,  \
                                   void* s) {                                                      \
    return jacobi<T>(p, v, d, b, xi, xo, w, k, rb, re, s);                                         \
  }                                                                                                \
  extern "C" int glab_cheby_first_##SUF(const glab_plan* p, const T* v, const T* b, const T* xi,   \
                                        T* xo, T* r, T* pv, const T* a, int k, int64_t rb,         \
                                        int64_t re, void* s) {                                     \
    return cheby_first<T>(p, v, b, xi, xo, r, pv, a, k, rb, re, s);                                \
  }                                                                                                \
  extern "C" int glab_cheby_next_##SUF(const glab_plan* p, const T* v, const T* pi, T* po, T* r,   \
                                       T* x, const T* ao, const T* a, const T* b, int k,           \
                                       int64_t rb, int64_t re, void* s) {                          \
    return cheby_next<T>(p, v, pi, po, r, x, ao, a, b, k, rb, re, s);                              \
  }                                                                                                \
  extern "C" int glab_power_step_##SUF(const glab_plan* p, const T* v, const T* bi, T* y,          \
                                       const double* si, double* so, void* ws, int64_t rb,         \
                                       int64_t re, void* s) {                                      \
    return power_step<T>(p, v, bi, y, si, so, ws, rb, re, s);                                      \
  }                                                                                                \
  extern "C" int glab_rayleigh_##SUF(const glab_plan* p, const T* v, const T* bi, T* bo, T* yo,    \
                                     const double* si, double* so, void* ws, int64_t rb,           \
                                     int64_t re, void* s) {                                        \
    return rayleigh<T>(p, v, bi, bo, yo, si, so, ws, rb, re, s);                                   \
  }                                                                                                \
  extern "C" int glab_xtax_##SUF(const glab_plan* p, const T* v, const T* x, double* so, void* ws, \
                                 int64_t rb, int64_t re, void* s) {                                \
    return xtax<T>(p, v, x, so, ws, rb, re, s);                                                    \
  }                                                                                                \
  extern "C" int glab_jacobi_sweeps_##SUF(const glab_plan* p, const T* v, const T* d, const T* b,  \
                                          T* xa, T* xb, const T* w, int k, int ns, void* s) {      \
    return jacobi_sweeps<T>(p, v, d, b, xa, xb, w, k, ns, s, nullptr, nullptr);                    \
  }

#define GLAB_HALO_INST(SUF, T)                                                                     \
  extern "C" int glab_spmm_halo_##SUF(const glab_plan* p, const T* v, const T* x, int k, T* y,     \
                                      const glab_halo_step* h, void* s) {                          \
    if (!p || !h) return GLAB_E_ARG;                                                               \
    return spmm<T>(p, v, x, k, y, 0, p->n_rows, s, h);                                             \
  }                                                                                                \
  extern "C" int glab_residual_halo_##SUF(const glab_plan* p, const T* v, const T* x, const T* b,  \
                                          int k, T* r, const glab_halo_step* h, void* s) {         \
    if (!p || !h) return GLAB_E_ARG;                                                               \
    return residual<T>(p, v, x, b, k, r, 0, p->n_rows, s, h);                                      \
  }                                                                                                \
  extern "C" int glab_jacobi_halo_##SUF(const glab_plan* p, const T* v, const T* d, const T* b,    \
                                        const T* xi, T* xo, const T* w, int k,                     \
                                        const glab_halo_step* h, void* s) {                        \
    if (!p || !h) return GLAB_E_ARG;                                                               \
    return jacobi<T>(p, v, d, b, xi, xo, w, k, 0, p->n_rows, s, h);                                \
  }                                                                                                \
  extern "C" int glab_jacobi_sweeps_halo_##SUF(const glab_plan* p, const T* v, const T* d,         \
                                               const T* b, T* xa, T* xb, const T* w, int k, int ns, \
                                               const glab_halo_step* hab,                          \
                                               const glab_halo_step* hba, void* s) {               \
    if (!p || !hab || !hba) return GLAB_E_ARG;                                                     \
    return jacobi_sweeps<T>(p, v, d, b, xa, xb, w, k, ns, s, hab, hba);                            \
  }                                                                                                \
  extern "C" int glab_cheby_first_halo_##SUF(const glab_plan* p, const T* v, const T* b,           \
                                             const T* xi, T* xo, T* r, T* pv, const T* a, int k,   \
                                             const glab_halo_step* h, void* s) {                   \
    if (!p || !h) return GLAB_E_ARG;                                                               \
    return cheby_first<T>(p, v, b, xi, xo, r, pv, a, k, 0, p->n_rows, s, h);                       \
  }                                                                                                \
  extern "C" int glab_cheby_next_halo_##SUF(const glab_plan* p, const T* v, const T* pi, T* po,    \
                                            T* r, T* x, const T* ao, const T* a, const T* b,       \
                                            int k, const glab_halo_step* h, void* s) {             \
    if (!p || !h) return GLAB_E_ARG;                                                               \
    return cheby_next<T>(p, v, pi, po, r, x, ao, a, b, k, 0, p->n_rows, s, h);                     \
  }                                                                                                \
  extern "C" int glab_power_step_halo_##SUF(const glab_plan* p, const T* v, const T* bi, T* y,     \
                                            const double* si, double* so, void* ws,                \
                                            const glab_halo_step* h, void* s) {                    \
    if (!p || !h) return GLAB_E_ARG;                                                               \
    return power_step<T>(p, v, bi, y, si, so, ws, 0, p->n_rows, s, h);                             \
  }                                                                                                \
  extern "C" int glab_rayleigh_halo_##SUF(const glab_plan* p, const T* v, const T* bi, T* bo,      \
                                          T* yo, const double* si, double* so, void* ws,           \
                                          const glab_halo_step* h, void* s) {                      \
    if (!p || !h) return GLAB_E_ARG;                                                               \
    return rayleigh<T>(p, v, bi, bo, yo, si, so, ws, 0, p->n_rows, s, h);                          \
  }

// One translation unit per value type (glab_layers_f32.cu / glab_layers_f64.cu) so that the
// ~700 pipeline-kernel instantiations compile in parallel.
#ifdef GLAB_LAYERS_F32
GLAB_INST(f32, float)
GLAB_HALO_INST(f32, float)
#endif
#ifdef GLAB_LAYERS_F64
GLAB_INST(f64, double)
GLAB_HALO_INST(f64, double)
#endif
