// C ABI of phyclone_b200 (see include/phyclone_b200.h).  No torch types, no CPU fallback.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <chrono>
#include <vector>

#include "../../include/phyclone_b200.h"
#include "pcb_kernels.cuh"
#include "pcb_sweep.cuh"

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define PCB_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return fail(PCB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                \
  } while (0)

#define PCB_LAUNCH_CHECK()                                                                          \
  do {                                                                                              \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                             \
    cudaError_t _e = cudaGetLastError();                                                            \
    if (_e != cudaSuccess) return fail(PCB_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(_e)); \
  } while (0)

// ---- layout -------------------------------------------------------------------
const int kStrips[] = {3, 5, 7, 9, 13, 17};

int jobs_double_b() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PCB_NBUF");
    v = (e && atoi(e) == 3) ? 1 : 0;  // default: one operand buffer (12 instead of 8 warps per SM: +12% measured)
  }
  return v;
}
// A, B (+ a second operand buffer with PCB_NBUF=3, or the output buffer O of the multi-round kernel)
size_t jobs_smem_bytes(int RB, int pitch, bool multi = false) {
  return pcb::kSmemHead + ((jobs_double_b() || multi) ? 3ull : 2ull) * RB * pitch * 8 + pcb::kSmemTail;
}

constexpr int kMaxGrid = 8192;  // three rows of a tile in shared memory

size_t wide_smem_bytes(int P, int pitch) {
  return pcb::kSmemHead + 2ull * 32 * pitch * 8 + (size_t)P * 32 * 8 + 32 * 8 + pcb::kSmemTail;
}

// unit-per-warp mapping (node_jobs_wide_kernel): P = 4 warps per CTA, 32 rows = 32/S whole jobs per CTA
int fill_layout_wide(int S, int G, int L, pcb_layout* out) {
  const int P = 4;
  if (S < 2 || (S & 1) || 32 % S != 0) return fail(PCB_ERR_ARG, "wide layout: S must be an even divisor of 32");
  if (G < 1 || G > 1088) return fail(PCB_ERR_ARG, "layout: need 1 <= G <= 1088");
  bool okL = false;
  for (int l : kStrips) okL |= (l == L);
  if (!okL) return fail(PCB_ERR_ARG, "layout: unsupported strip");
  const int Q = (G + L - 1) / L;
  if ((Q + 1) / 2 > P) return fail(PCB_ERR_ARG, "wide layout: four strip pairs must cover the grid");
  int pitch = L + (Q + 1) * L;
  if ((pitch & 1) == 0) ++pitch;  // lanes of a warp are rows: an odd pitch spreads them over the banks
  if (wide_smem_bytes(P, pitch) > 200 * 1024) return fail(PCB_ERR_ARG, "layout: grid too large for shared memory");
  out->S = S;
  out->G = G;
  out->strip = L;
  out->lanes_per_row = P;
  out->n_strips = Q;
  out->pitch = pitch;
  out->lead = L;
  out->rows_per_cta = 32;
  out->kernel = 1;
  out->reserved = 0;
  out->tile_elems = (int64_t)S * pitch;
  return PCB_OK;
}

int fill_layout(int S, int G, int L, int P, pcb_layout* out) {
  if (S < 1 || G < 1 || G > kMaxGrid) return fail(PCB_ERR_ARG, "layout: need S >= 1 and 1 <= G <= 8192");
  bool okL = false;
  for (int l : kStrips) okL |= (l == L);
  if (!okL || (P != 4 && P != 8 && P != 16 && P != 32)) return fail(PCB_ERR_ARG, "layout: unsupported strip/lanes");
  const int Q = (G + L - 1) / L;
  // more strip pairs than lanes: only the widest configuration walks several pairs per lane (grids beyond 1088)
  const int rounds = ((Q + 1) / 2 + P - 1) / P;
  if (rounds > 1 && !(L == 17 && P == 32))
    return fail(PCB_ERR_ARG, "layout: lanes_per_row must cover ceil(Q/2) strip pairs (several pairs per lane: L=17, P=32 only)");
  // lead zeros | Q strips | one spare strip (the operand refill of a body that never runs stays in the row)
  int pitch = L + (Q + 1) * L;
  int want = (P == 4) ? 4 : 8;  // bank rule: rows sharing a half-warp must not collide
  if (const char* e = getenv("PCB_PITCH_MOD")) want = atoi(e);  // tuning knob
  while (pitch % 16 != want) ++pitch;
  const int RB = 32 / P;  // one warp per CTA
  if (jobs_smem_bytes(RB, pitch, rounds > 1) > 200 * 1024)
    return fail(PCB_ERR_ARG, "layout: grid too large for shared memory");
  out->S = S;
  out->G = G;
  out->strip = L;
  out->lanes_per_row = P;
  out->n_strips = Q;
  out->pitch = pitch;
  out->lead = L;
  out->rows_per_cta = RB;
  out->kernel = 0;
  out->reserved = rounds > 1 ? rounds : 0;  // strip pairs per lane and fold step when more than one
  out->tile_elems = (int64_t)S * pitch;
  return PCB_OK;
}

long long lane_slots(int G, int L, int P) {
  const long long Q = (G + L - 1) / L, U = (Q + 1) / 2;
  return P * ((U + P - 1) / P) * (Q + 1) * L * L;
}

// ---- programmatic dependent launch: inside the sweep's chain of ~300 small dependent kernels every launch is allowed
// to start while its predecessor drains (the kernels hold at pdl_wait() before touching anything the predecessor
// wrote; csrc/pcb_device.cuh).  Off everywhere else.  PCB_SWEEP_PDL=0 disables it.
thread_local bool g_pdl = false;
struct PdlScope {  // launches made while one is alive carry the attribute (PCB_SWEEP_PDL=0: never)
  PdlScope() {
    static const bool on = [] {
      const char* e = getenv("PCB_SWEEP_PDL");
      return !(e && atoi(e) == 0);
    }();
    g_pdl = on;
  }
  ~PdlScope() { g_pdl = false; }
};
template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// CTAs of a launch whose job count only the device knows: a pool that covers the device once at the kernel's occupancy
int pool_ctas() {
  static int v = 0;
  if (!v) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    v = sms * 12;
  }
  return v;
}

template <int L, bool FUSE, bool MULTI = false>
int launch_jobs_LF(const pcb::JobKernelParams& kp, long long n_jobs, int threads, size_t smem, cudaStream_t st) {
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    PCB_CUDA(cudaFuncSetAttribute(pcb::node_jobs_kernel<L, FUSE, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    configured = smem;
  }
  const long long grid = kp.n_jobs_dev ? std::min<long long>(n_jobs * kp.row_blocks, pool_ctas()) : n_jobs * kp.row_blocks;
  if (grid > 2147483647ll) return fail(PCB_ERR_ARG, "too many jobs for one launch");
  PCB_CUDA(launch_k(pcb::node_jobs_kernel<L, FUSE, MULTI>, dim3((unsigned)grid), dim3(threads), smem, st, kp));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PCB_OK;
}

// `fused`: the table may hold PCB_JOB_THEN_SCORE jobs (only the strips the latency / throughput layouts use are built)
template <int L>
int launch_jobs_L(const pcb::JobKernelParams& kp, long long n_jobs, int threads, size_t smem, cudaStream_t st,
                  bool fused = false) {
  if (kp.rounds > 1) {
    if constexpr (L == 17) {
      if (fused) return fail(PCB_ERR_ARG, "node_jobs: no fused jobs on grids beyond 1088");
      return launch_jobs_LF<L, false, true>(kp, n_jobs, threads, smem, st);
    } else {
      return fail(PCB_ERR_ARG, "node_jobs: several strip pairs per lane need the 17-wide strip");
    }
  }
  if (fused) {
    if constexpr (L == 7 || L == 13 || L == 17)
      return launch_jobs_LF<L, true>(kp, n_jobs, threads, smem, st);
    else
      return fail(PCB_ERR_ARG, "node_jobs: fused jobs are not built for this strip width");
  }
  return launch_jobs_LF<L, false>(kp, n_jobs, threads, smem, st);
}

template <int L>
int launch_wide_L(const pcb::JobKernelParams& kp, long long n_jobs, size_t smem, cudaStream_t st) {
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    PCB_CUDA(cudaFuncSetAttribute(pcb::node_jobs_wide_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  long long grid = (n_jobs * kp.S + 31) / 32;
  if (kp.n_jobs_dev) grid = std::min<long long>(grid, pool_ctas() / 4);
  if (grid > 2147483647ll) return fail(PCB_ERR_ARG, "too many jobs for one launch");
  PCB_CUDA(launch_k(pcb::node_jobs_wide_kernel<L>, dim3((unsigned)grid), dim3(32 * kp.P), smem, st, kp));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PCB_OK;
}

// ---- the latency mapping (node_jobs_rows_kernel): one warp per row, the bodies of a strip pair split over H lanes
int rows_rb(int S) {  // rows (warps) per CTA: the largest divisor of S up to 4 (PCB_ROWS_RB: up to that many)
  static int cap = 0;
  if (!cap) {
    const char* e = getenv("PCB_ROWS_RB");
    cap = (e && atoi(e) >= 1 && atoi(e) <= 4) ? atoi(e) : 4;
  }
  for (int rb = cap; rb > 1; --rb)
    if (S % rb == 0) return rb;
  return 1;
}
size_t rows_smem_bytes(const pcb_layout& l) { return pcb::kSmemHead + 3ull * rows_rb(l.S) * l.pitch * 8 + pcb::kSmemTail; }
bool rows_supported(const pcb_layout& l) {
  return (l.kernel & 3) == 0 && l.reserved <= 1 && (l.strip == 7 || l.strip == 13 || l.strip == 17) &&
         l.n_strips >= 1 && l.n_strips <= 32 && l.lead >= l.strip - 1 && l.pitch - l.lead >= l.n_strips * l.strip &&
         rows_smem_bytes(l) <= 200 * 1024;
}
// launches of at most this many jobs go to the latency mapping when the store asked for it (PCB_ROWS_MAX_JOBS)
long long rows_max_jobs() {
  static long long v = -1;
  if (v < 0) {
    const char* env = getenv("PCB_ROWS_MAX_JOBS");
    v = (env && *env) ? atoll(env) : 1024;
  }
  return v;
}
template <int L>
int launch_rows_L(const pcb::JobKernelParams& kp, long long n_jobs, size_t smem, cudaStream_t st) {
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    PCB_CUDA(cudaFuncSetAttribute(pcb::node_jobs_rows_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  // device-counted launches: a pool that covers the device a few times over
  const long long grid = kp.n_jobs_dev ? std::min<long long>(n_jobs * kp.row_blocks, pool_ctas() / 3) : n_jobs * kp.row_blocks;
  if (grid > 2147483647ll) return fail(PCB_ERR_ARG, "too many jobs for one launch");
  PCB_CUDA(launch_k(pcb::node_jobs_rows_kernel<L>, dim3((unsigned)grid), dim3(32 * kp.RB), smem, st, kp));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PCB_OK;
}
int launch_rows(pcb::JobKernelParams kp, const pcb_layout& l, long long n_jobs, cudaStream_t st) {
  kp.RB = rows_rb(l.S);
  kp.row_blocks = l.S / kp.RB;
  // bodies per lane: the fewest for which the strips, split over ceil((q+1)/T) lanes each, fit a warp
  int T = 1;
  for (;; ++T) {
    int n = 0;
    for (int q = 0; q < l.n_strips; ++q) n += (q + T) / T;
    if (n <= 32) break;
  }
  kp.P = T;
  const size_t smem = rows_smem_bytes(l);
  switch (l.strip) {
    case 7:
      return launch_rows_L<7>(kp, n_jobs, smem, st);
    case 13:
      return launch_rows_L<13>(kp, n_jobs, smem, st);
    case 17:
      return launch_rows_L<17>(kp, n_jobs, smem, st);
  }
  return fail(PCB_ERR_ARG, "node_jobs: unsupported strip for the latency mapping");
}

int check_arena(const pcb_arena* a) {
  if (!a || !a->lg || !a->lin || !a->rmax || a->capacity < 1) return fail(PCB_ERR_ARG, "arena: null");
  return PCB_OK;
}

// ---- scratch device memory for the host-buffer entry points --------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  ~DevBuf() {  // per-thread scratch goes away with its thread (a chain = one host thread)
    if (p) cudaFree(p);
  }
  int ensure(size_t bytes) {
    if (bytes <= cap) return PCB_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();  // do not leave the failure for the next launch check to report as its own
      p = nullptr;
      return fail(e == cudaErrorMemoryAllocation ? PCB_ERR_NOMEM : PCB_ERR_CUDA,
                  std::string("cudaMalloc: ") + cudaGetErrorString(e));
    }
    cap = want;
    return PCB_OK;
  }
  template <class T>
  T* as() {
    return static_cast<T*>(p);
  }
};

struct HostScratch {
  DevBuf in, out, lg, lin, rmax, jobs, kids, ids, rows_a, rows_b, misc, red;
  pcb_layout layout{};
  int64_t arena_cap = 0;
  bool arena_zeroed_for = false;
};
thread_local HostScratch g_hs;

int ensure_arena(HostScratch& hs, int S, int G, int64_t tiles, pcb_arena* a) {
  pcb_layout lay;
  int rc = pcb_make_layout(S, G, -1, &lay);
  if (rc) return rc;
  const bool same = hs.arena_cap >= tiles && memcmp(&hs.layout, &lay, sizeof(lay)) == 0;
  if (!same) {
    const size_t tb = (size_t)lay.tile_elems * 8;
    if ((rc = hs.lg.ensure(tb * tiles))) return rc;
    if ((rc = hs.lin.ensure(tb * tiles))) return rc;
    if ((rc = hs.rmax.ensure((size_t)S * 8 * tiles))) return rc;
    PCB_CUDA(cudaMemsetAsync(hs.lg.p, 0, tb * tiles, 0));
    PCB_CUDA(cudaMemsetAsync(hs.lin.p, 0, tb * tiles, 0));
    hs.layout = lay;
    hs.arena_cap = tiles;
  }
  a->layout = hs.layout;
  a->lg = hs.lg.as<double>();
  a->lin = hs.lin.as<double>();
  a->rmax = hs.rmax.as<double>();
  a->capacity = hs.arena_cap;
  return PCB_OK;
}

// Shared driver of the dense host entry points: import `n_in` dense tiles, run B jobs, export B tiles.
// job b folds input tiles kids[koff[b]..koff[b+1]) ; base[b] (input tile index or -1); flags.
int run_dense(const double* in_host, int64_t n_in, const std::vector<int32_t>& kids, const std::vector<int64_t>& koff,
              const std::vector<int32_t>& base, int flags, double prior, int64_t B, int S, int G, double* out_host,
              double* lse_host, double* last_host) {
  HostScratch& hs = g_hs;
  pcb_arena ar;
  int rc = ensure_arena(hs, S, G, n_in + B, &ar);
  if (rc) return rc;
  const size_t dense_tile = (size_t)S * G * 8;
  if ((rc = hs.in.ensure(dense_tile * std::max<int64_t>(n_in, 1)))) return rc;
  if ((rc = hs.out.ensure(dense_tile * std::max<int64_t>(B, 1)))) return rc;
  std::vector<int32_t> ids(n_in + B);
  for (int64_t i = 0; i < n_in + B; ++i) ids[i] = (int32_t)i;
  if ((rc = hs.ids.ensure(ids.size() * 4))) return rc;
  PCB_CUDA(cudaMemcpyAsync(hs.ids.p, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, 0));
  PCB_CUDA(cudaMemcpyAsync(hs.in.p, in_host, dense_tile * n_in, cudaMemcpyHostToDevice, 0));
  if ((rc = pcb_arena_import_dev(&ar, hs.in.as<double>(), hs.ids.as<int32_t>(), n_in, nullptr))) return rc;
  // jobs with >= 1 child go through the kernel; leaves (0 children) copy their base tile
  std::vector<int32_t> jobs;
  std::vector<int32_t> out_ids(B);
  int64_t n_jobs = 0;
  for (int64_t b = 0; b < B; ++b) {
    const int64_t c = koff[b + 1] - koff[b];
    if (c == 0) {
      if (base.empty() || base[b] < 0) return fail(PCB_ERR_ARG, "job without children and without log_p");
      out_ids[b] = base[b];
      continue;
    }
    out_ids[b] = (int32_t)(n_in + b);
    const int32_t rec[8] = {(int32_t)koff[b], (int32_t)c, base.empty() ? -1 : base[b], out_ids[b], flags,
                            (lse_host ? (int32_t)b : -1), 0, 0};
    jobs.insert(jobs.end(), rec, rec + 8);
    ++n_jobs;
  }
  double *rows_l = nullptr, *rows_t = nullptr;
  if (lse_host) {
    if ((rc = hs.rows_a.ensure((size_t)B * S * 8))) return rc;
    if ((rc = hs.rows_b.ensure((size_t)B * S * 8))) return rc;
    rows_l = hs.rows_a.as<double>();
    rows_t = hs.rows_b.as<double>();
  }
  if (n_jobs) {
    if ((rc = hs.jobs.ensure(jobs.size() * 4))) return rc;
    if ((rc = hs.kids.ensure(std::max<size_t>(kids.size(), 1) * 4))) return rc;
    PCB_CUDA(cudaMemcpyAsync(hs.jobs.p, jobs.data(), jobs.size() * 4, cudaMemcpyHostToDevice, 0));
    PCB_CUDA(cudaMemcpyAsync(hs.kids.p, kids.data(), kids.size() * 4, cudaMemcpyHostToDevice, 0));
    if ((rc = pcb_node_jobs_dev(&ar, hs.jobs.as<int32_t>(), hs.kids.as<int32_t>(), n_jobs, prior, rows_l, rows_t,
                                nullptr)))
      return rc;
  }
  if (out_host) {
    if ((rc = hs.misc.ensure((size_t)B * 4))) return rc;
    PCB_CUDA(cudaMemcpyAsync(hs.misc.p, out_ids.data(), (size_t)B * 4, cudaMemcpyHostToDevice, 0));
    if ((rc = pcb_arena_export_dev(&ar, hs.misc.as<int32_t>(), B, hs.out.as<double>(), nullptr))) return rc;
    PCB_CUDA(cudaMemcpyAsync(out_host, hs.out.p, dense_tile * B, cudaMemcpyDeviceToHost, 0));
  }
  if (lse_host) {
    // reduce rows -> [B] on device, in sample order
    if ((rc = hs.red.ensure((size_t)B * 16))) return rc;
    double* red = hs.red.as<double>();
    if ((rc = pcb_reduce_rows_dev(rows_l, B, S, red, nullptr))) return rc;
    if ((rc = pcb_reduce_rows_dev(rows_t, B, S, red + B, nullptr))) return rc;
    PCB_CUDA(cudaMemcpyAsync(lse_host, red, (size_t)B * 8, cudaMemcpyDeviceToHost, 0));
    PCB_CUDA(cudaMemcpyAsync(last_host, red + B, (size_t)B * 8, cudaMemcpyDeviceToHost, 0));
  }
  PCB_CUDA(cudaStreamSynchronize(0));
  return PCB_OK;
}

}  // namespace

extern "C" {

int pcb_version(void) { return 100; }
const char* pcb_last_error(void) { return g_err.c_str(); }
int64_t pcb_launch_count(void) { return g_launches.load(); }

int pcb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int pcb_set_device(int device) {
  PCB_CUDA(cudaSetDevice(device));
  return PCB_OK;
}

int pcb_make_layout(int32_t S, int32_t G, int32_t variant, pcb_layout* out) {
  if (!out) return fail(PCB_ERR_ARG, "layout: null out");
  if (variant < 0) {
    const char* env = getenv("PCB_VARIANT");
    if (env && *env) variant = atoi(env);
  }
  if (variant >= 10000) return fill_layout_wide(S, G, (variant - 10000) / 100, out);
  if (variant >= 0) return fill_layout(S, G, variant / 100, variant % 100, out);
  if (variant < 0 && G > 1088) return fill_layout(S, G, 17, 32, out);  // several strip pairs per lane
  if (variant == -2) {
    // latency-optimised: a single chain launches a handful of jobs at a time, so the device is empty and the
    // time of a launch is the time of ONE lane's strip pair; take the narrowest strip whose pairs fit a row group
    for (int L : {7, 13, 17}) {
      const int U = ((G + L - 1) / L + 1) / 2;
      for (int P : {4, 8, 16, 32})
        if (U <= P) return fill_layout(S, G, L, P, out);
    }
    return fail(PCB_ERR_ARG, "layout: grid too large");
  }
  // tiles with fewer than 8 rows leave lanes of the lane-per-unit kernel idle (or force narrow strips): the
  // unit-per-warp kernel packs 32/S jobs into a CTA instead (S=4, G=101: 48.8 % vs 39.6 % of the FP64 peak)
  if ((S == 2 || S == 4) && ((G + 12) / 13 + 1) / 2 <= 4) return fill_layout_wide(S, G, 13, out);
  // default: fewest executed lane-FMAs among the conflict-free (odd strip) configurations with
  // at least 7 FMAs per shared-memory operand pair
  long long best = -1;
  int bl = 13, bp = 32;
  for (int L : {7, 13, 17})
    for (int P : {4, 8, 16, 32}) {
      const int Q = (G + L - 1) / L;
      if ((Q + 1) / 2 > P) continue;
      if (L == 17 && G <= 13 * 64) continue;  // the widest strip only where 13 cannot cover the grid
      // executed lane-FMAs per row, times the rows a CTA pads when S is not a multiple of its 32/P rows
      // (S=4 with P=4 would leave half of every warp idle: measured 24% vs 33% of peak for L=7,P=8);
      // the long strip is ~1.3x more efficient per executed FMA (13 DFMAs per operand pair, fewer switches)
      const int RB = 32 / P;
      const long long rows = (long long)RB * ((S + RB - 1) / RB);
      const long long c = lane_slots(G, L, P) * (L == 13 ? 10 : 13) * rows / S;
      if (best < 0 || c < best) best = c, bl = L, bp = P;
    }
  return fill_layout(S, G, bl, bp, out);
}

int pcb_arena_import_dev(const pcb_arena* a, const double* dense, const int32_t* ids_dev, int64_t n, void* stream) {
  int rc = check_arena(a);
  if (rc) return rc;
  if (n <= 0) return PCB_OK;
  const pcb_layout& l = a->layout;
  const long long warps = n * l.S;
  const int wpb = 4;
  pcb::import_tiles_kernel<<<(unsigned)((warps + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      dense, ids_dev, n, a->lg, a->lin, a->rmax, l.S, l.G, l.pitch, l.lead, l.tile_elems);
  PCB_LAUNCH_CHECK();
  return PCB_OK;
}

int pcb_arena_export_dev(const pcb_arena* a, const int32_t* ids_dev, int64_t n, double* dense, void* stream) {
  int rc = check_arena(a);
  if (rc) return rc;
  if (n <= 0) return PCB_OK;
  const pcb_layout& l = a->layout;
  const long long warps = n * l.S;
  const int wpb = 4;
  pcb::export_tiles_kernel<<<(unsigned)((warps + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      a->lg, ids_dev, n, dense, l.S, l.G, l.pitch, l.lead, l.tile_elems);
  PCB_LAUNCH_CHECK();
  return PCB_OK;
}

int pcb_tile_add_dev(const pcb_arena* a, const int32_t* a_ids, const int32_t* b_ids, const int32_t* out_ids,
                     int64_t n, double b_sign, double addend, void* stream) {
  int rc = check_arena(a);
  if (rc) return rc;
  if (n <= 0) return PCB_OK;
  const pcb_layout& l = a->layout;
  const long long warps = n * l.S;
  if ((warps + 3) / 4 > 2147483647ll) return fail(PCB_ERR_ARG, "tile_add: too many tiles for one launch");
  const int wpb = 4;
  PCB_CUDA(launch_k(pcb::tile_add_kernel, dim3((unsigned)((warps + wpb - 1) / wpb)), dim3(wpb * 32), 0, (cudaStream_t)stream,
                    a_ids, b_ids, out_ids, (long long)n, b_sign, addend, a->lg, a->lin, a->rmax, l.S, l.G, l.pitch, l.lead,
                    (long long)l.tile_elems));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PCB_OK;
}

// ---- kernel timing for the roofline figure: CUDA event pairs around every node-jobs launch, owned by the library
// so that the instrumented region pays two cudaEventRecord calls per launch and nothing on the Python side
struct JobTiming {
  bool on = false;
  std::vector<cudaEvent_t> ev;  // 2 per recorded launch
  size_t used = 0, cap = 0;
  int stride = 1;       // every stride-th launch is bracketed (pcb_timing_sample)
  int64_t seen = 0;     // node-jobs launches since begin, bracketed or not
};
thread_local JobTiming g_timing;  // a chain is one host thread: its kernel timing is its own

int pcb_timing_begin(int32_t max_launches) {
  if (max_launches < 1) return fail(PCB_ERR_ARG, "timing: max_launches must be positive");
  const size_t want = 2 * (size_t)max_launches;
  while (g_timing.ev.size() < want) {
    cudaEvent_t e;
    PCB_CUDA(cudaEventCreate(&e));
    g_timing.ev.push_back(e);
  }
  g_timing.cap = want;
  g_timing.used = 0;
  g_timing.seen = 0;
  g_timing.on = true;
  return PCB_OK;
}

int pcb_timing_sample(int32_t stride, int64_t* n_seen) {
  if (stride >= 1) g_timing.stride = stride;
  if (n_seen) *n_seen = g_timing.seen;
  return PCB_OK;
}

int pcb_timing_end(double* total_ms, int64_t* n_launches) {
  g_timing.on = false;
  PCB_CUDA(cudaDeviceSynchronize());
  double total = 0.0;
  for (size_t i = 0; i + 1 < g_timing.used; i += 2) {
    float ms = 0.f;
    PCB_CUDA(cudaEventElapsedTime(&ms, g_timing.ev[i], g_timing.ev[i + 1]));
    total += ms;
  }
  if (total_ms) *total_ms = total;
  if (n_launches) *n_launches = (int64_t)(g_timing.used / 2);
  g_timing.used = 0;
  return PCB_OK;
}

// mode bits of a launch: the table may hold PCB_JOB_THEN_SCORE jobs / the caller knows the launch is a small one
constexpr int kJobsFused = 1, kJobsSmall = 2;
static int node_jobs_launch(const pcb_arena* a, const int32_t* jobs_dev, const int32_t* children_dev, int64_t n_jobs,
                            double prior, double* rows_lse, double* rows_last, void* stream,
                            const int* n_jobs_dev = nullptr, int mode = 0);

static int jobs_launch_timed(const pcb_arena* a, const int32_t* jobs_dev, const int32_t* children_dev, int64_t n_jobs,
                             double prior, double* rows_lse, double* rows_last, void* stream, const int* n_jobs_dev,
                             int mode = 0) {
  const bool timed = g_timing.on && n_jobs > 0 && (g_timing.seen++ % g_timing.stride) == 0 && g_timing.used + 2 <= g_timing.cap;
  if (timed) PCB_CUDA(cudaEventRecord(g_timing.ev[g_timing.used], (cudaStream_t)stream));
  const int rc = node_jobs_launch(a, jobs_dev, children_dev, n_jobs, prior, rows_lse, rows_last, stream, n_jobs_dev, mode);
  if (timed && rc == PCB_OK) {
    PCB_CUDA(cudaEventRecord(g_timing.ev[g_timing.used + 1], (cudaStream_t)stream));
    g_timing.used += 2;
  }
  return rc;
}

int pcb_node_jobs_dev(const pcb_arena* a, const int32_t* jobs_dev, const int32_t* children_dev, int64_t n_jobs,
                      double prior, double* rows_lse, double* rows_last, void* stream) {
  return jobs_launch_timed(a, jobs_dev, children_dev, n_jobs, prior, rows_lse, rows_last, stream, nullptr);
}

// n_jobs_dev != NULL: `n_jobs` is only an upper bound, the count is read on the device (tables written by a kernel
// earlier in the stream)
static int node_jobs_launch(const pcb_arena* a, const int32_t* jobs_dev, const int32_t* children_dev, int64_t n_jobs,
                            double prior, double* rows_lse, double* rows_last, void* stream, const int* n_jobs_dev,
                            int mode) {
  const bool fused = mode & kJobsFused;
  int rc = check_arena(a);
  if (rc) return rc;
  if (n_jobs <= 0) return PCB_OK;
  if (!jobs_dev || !children_dev) return fail(PCB_ERR_ARG, "node_jobs: null job table");
  const pcb_layout& l = a->layout;
  pcb::JobKernelParams kp;
  kp.jobs = jobs_dev;
  kp.children = children_dev;
  kp.lg = a->lg;
  kp.lin = a->lin;
  kp.rmax = a->rmax;
  kp.sc_lse = rows_lse;
  kp.sc_last = rows_last;
  kp.tile_elems = l.tile_elems;
  kp.prior = prior;
  kp.S = l.S;
  kp.G = l.G;
  kp.Q = l.n_strips;
  kp.U = (l.n_strips + 1) / 2;
  kp.P = l.lanes_per_row;
  kp.pitch = l.pitch;
  kp.lead = l.lead;
  kp.RB = l.rows_per_cta;
  kp.row_blocks = (l.S + l.rows_per_cta - 1) / l.rows_per_cta;
  kp.rounds = l.reserved > 1 ? l.reserved : 1;
  kp.double_b = kp.rounds > 1 ? 0 : jobs_double_b();
  kp.n_jobs = n_jobs;
  kp.n_jobs_dev = n_jobs_dev;
  if ((l.kernel & PCB_KERNEL_LATENCY) && rows_supported(l) && ((mode & kJobsSmall) || (!n_jobs_dev && n_jobs <= rows_max_jobs()))) {
    return launch_rows(kp, l, n_jobs, (cudaStream_t)stream);
  }
  if ((l.kernel & 3) == 1) {
    if (fused) return fail(PCB_ERR_ARG, "node_jobs: the unit-per-warp kernel has no fused jobs");
    const size_t wsmem = wide_smem_bytes(l.lanes_per_row, l.pitch);
    cudaStream_t wst = (cudaStream_t)stream;
    switch (l.strip) {
      case 7:
        return launch_wide_L<7>(kp, n_jobs, wsmem, wst);
      case 13:
        return launch_wide_L<13>(kp, n_jobs, wsmem, wst);
      case 17:
        return launch_wide_L<17>(kp, n_jobs, wsmem, wst);
    }
    return fail(PCB_ERR_ARG, "node_jobs: unsupported strip for the wide kernel");
  }
  const int threads = 32;
  const size_t smem = jobs_smem_bytes(l.rows_per_cta, l.pitch, kp.rounds > 1);
  cudaStream_t st = (cudaStream_t)stream;
  switch (l.strip) {
    case 3:
      return launch_jobs_L<3>(kp, n_jobs, threads, smem, st, fused);
    case 5:
      return launch_jobs_L<5>(kp, n_jobs, threads, smem, st, fused);
    case 7:
      return launch_jobs_L<7>(kp, n_jobs, threads, smem, st, fused);
    case 9:
      return launch_jobs_L<9>(kp, n_jobs, threads, smem, st, fused);
    case 13:
      return launch_jobs_L<13>(kp, n_jobs, threads, smem, st, fused);
    case 17:
      return launch_jobs_L<17>(kp, n_jobs, threads, smem, st, fused);
  }
  return fail(PCB_ERR_ARG, "node_jobs: unsupported strip");
}

int pcb_reduce_rows_dev(const double* rows_dev, int64_t n_slots, int32_t S, double* out_dev, void* stream) {
  if (n_slots <= 0) return PCB_OK;
  pcb::reduce_rows_kernel<<<(unsigned)((n_slots + 127) / 128), 128, 0, (cudaStream_t)stream>>>(rows_dev, n_slots, S,
                                                                                               out_dev);
  PCB_LAUNCH_CHECK();
  return PCB_OK;
}

thread_local double g_plan_enqueue_s = 0.0, g_plan_sync_s = 0.0;  // per host thread (= per chain)
thread_local int64_t g_plan_flushes = 0;

// One flush of the host's tile store in ONE call: stage the packed id lists / job tables (pinned host -> device),
// run every launch of the plan in dependency order, reduce the score rows and copy the sums back.
// plan: n_rec records of 6 int32: {kind (0 = tile add, 1 = node jobs), offset into packed, n, n_kids, negate, 0}.
// events: NULL, or 2 CUDA events per record (recorded around node-jobs launches only: the bench's kernel timing).
int pcb_run_plan_dev(const pcb_arena* a, const int32_t* packed_host, int64_t n_ints, int32_t* packed_dev,
                     const int32_t* plan, int32_t n_rec, double prior, double* rows_lse, double* rows_last,
                     int64_t n_slots, double* sums_dev, double* sums_host, void* const* events, void* stream) {
  int rc = check_arena(a);
  if (rc) return rc;
  if (n_rec <= 0) return PCB_OK;
  if (!packed_host || !packed_dev || !plan) return fail(PCB_ERR_ARG, "run_plan: null pointer");
  if (n_slots > 0 && (!rows_lse || !rows_last || !sums_dev || !sums_host))
    return fail(PCB_ERR_ARG, "run_plan: score buffers missing");
  cudaStream_t st = (cudaStream_t)stream;
  const auto t_begin = std::chrono::steady_clock::now();
  // every tile id / score slot of the plan must lie inside the arena / the score buffers: a stale plan (ids from before
  // a reset or grow) would otherwise write into neighbouring allocations
  for (int r = 0; r < n_rec; ++r) {
    const int32_t* rec = plan + 6 * r;
    const int64_t off = rec[1], n = rec[2];
    if (off < 0 || n < 0 || rec[3] < 0 || off + (rec[0] == 0 ? 3 * n : 8 * n + rec[3]) > n_ints)
      return fail(PCB_ERR_ARG, "run_plan: record outside the packed buffer");
    const int32_t* q = packed_host + off;
    const int64_t cap = a->capacity;
    if (rec[0] == 0) {
      for (int64_t i = 0; i < 3 * n; ++i)
        if (q[i] < 0 || q[i] >= cap) return fail(PCB_ERR_ARG, "run_plan: tile id outside the arena");
    } else {
      const int32_t* kids = q + 8 * n;
      for (int64_t i = 0; i < n; ++i) {
        const int32_t* j = q + 8 * i;
        if (j[0] < 0 || j[1] < 1 || (int64_t)j[0] + j[1] > rec[3] || j[2] < -1 || j[2] >= cap || j[3] < -1 || j[3] >= cap ||
            j[5] < -1 || j[5] >= n_slots || j[6] < 0 || j[6] > cap)
          return fail(PCB_ERR_ARG, "run_plan: job record outside the arena / score buffers");
      }
      for (int64_t i = 0; i < rec[3]; ++i)
        if (kids[i] < 0 || kids[i] >= cap) return fail(PCB_ERR_ARG, "run_plan: child tile id outside the arena");
    }
  }
  PCB_CUDA(cudaMemcpyAsync(packed_dev, packed_host, (size_t)n_ints * 4, cudaMemcpyHostToDevice, st));
  PdlScope pdl_scope;  // the plan's launches depend on each other level by level: let each start in its predecessor's tail
  for (int r = 0; r < n_rec; ++r) {
    const int32_t* rec = plan + 6 * r;
    const int64_t off = rec[1], n = rec[2];
    const int32_t* p0 = packed_dev + off;
    if (rec[0] == 0) {
      if ((rc = pcb_tile_add_dev(a, p0, p0 + n, p0 + 2 * n, n, rec[4] ? -1.0 : 1.0, 0.0, stream))) return rc;
    } else {
      if (events) PCB_CUDA(cudaEventRecord((cudaEvent_t)events[2 * r], st));
      if ((rc = pcb_node_jobs_dev(a, p0, p0 + 8 * n, n, prior, rows_lse, rows_last, stream))) return rc;
      if (events) PCB_CUDA(cudaEventRecord((cudaEvent_t)events[2 * r + 1], st));
    }
  }
  if (n_slots > 0) {
    // one launch reduces both score kinds over the samples and writes the sums straight into the caller's pinned
    // (device-accessible) host buffer: no separate device-to-host copy; the synchronise below makes them visible
    (void)sums_dev;
    PCB_CUDA(launch_k(pcb::reduce_rows2_kernel, dim3((unsigned)((2 * n_slots + 127) / 128)), dim3(128), 0, st, rows_lse,
                      rows_last, (long long)n_slots, a->layout.S, sums_host));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const auto t_enq = std::chrono::steady_clock::now();
    PCB_CUDA(cudaStreamSynchronize(st));
    const auto t_end = std::chrono::steady_clock::now();
    g_plan_enqueue_s += std::chrono::duration<double>(t_enq - t_begin).count();
    g_plan_sync_s += std::chrono::duration<double>(t_end - t_enq).count();
    ++g_plan_flushes;
  }
  return PCB_OK;
}

// where a synchronising flush spends its time on the host: enqueueing (copies + launches) vs waiting for the stream
int pcb_plan_stats(double* enqueue_s, double* sync_s, int64_t* flushes) {
  if (enqueue_s) *enqueue_s = g_plan_enqueue_s;
  if (sync_s) *sync_s = g_plan_sync_s;
  if (flushes) *flushes = g_plan_flushes;
  return PCB_OK;
}

// ---- host-buffer entry points ---------------------------------------------------
int pcb_np_conv_dims_host(const double* c1, const double* c2, double* out, int64_t B, int32_t S, int32_t G) {
  if (!c1 || !c2 || !out || B < 0) return fail(PCB_ERR_ARG, "np_conv_dims: null pointer");
  if (B == 0) return PCB_OK;
  // inputs interleaved into one dense block: tile 2b = c1[b], tile 2b+1 = c2[b]
  std::vector<double> in((size_t)2 * B * S * G);
  const size_t t = (size_t)S * G;
  for (int64_t b = 0; b < B; ++b) {
    memcpy(&in[(2 * b) * t], c1 + b * t, t * 8);
    memcpy(&in[(2 * b + 1) * t], c2 + b * t, t * 8);
  }
  std::vector<int32_t> kids(2 * B);
  std::vector<int64_t> koff(B + 1);
  for (int64_t b = 0; b < B; ++b) {
    kids[2 * b] = (int32_t)(2 * b);
    kids[2 * b + 1] = (int32_t)(2 * b + 1);
    koff[b] = 2 * b;
  }
  koff[B] = 2 * B;
  return run_dense(in.data(), 2 * B, kids, koff, {}, 0, 0.0, B, S, G, out, nullptr, nullptr);
}

int pcb_sub_compute_S_host(const double* log_D, double* log_S, int64_t B, int32_t S, int32_t G) {
  if (!log_D || !log_S || B < 0) return fail(PCB_ERR_ARG, "sub_compute_S: null pointer");
  if (B == 0) return PCB_OK;
  std::vector<int32_t> kids(B);
  std::vector<int64_t> koff(B + 1);
  for (int64_t b = 0; b < B; ++b) kids[b] = (int32_t)b, koff[b] = b;
  koff[B] = B;
  return run_dense(log_D, B, kids, koff, {}, PCB_JOB_SCAN, 0.0, B, S, G, log_S, nullptr, nullptr);
}

int pcb_compute_log_S_host(const double* children, const int64_t* child_off, double* log_S, int64_t B, int32_t S,
                           int32_t G) {
  if (!children || !child_off || !log_S || B < 0) return fail(PCB_ERR_ARG, "compute_log_S: null pointer");
  if (B == 0) return PCB_OK;
  const int64_t n = child_off[B];
  std::vector<int32_t> kids(n);
  for (int64_t i = 0; i < n; ++i) kids[i] = (int32_t)i;
  std::vector<int64_t> koff(child_off, child_off + B + 1);
  for (int64_t b = 0; b < B; ++b)
    if (koff[b + 1] - koff[b] < 1) return fail(PCB_ERR_ARG, "compute_log_S: every job needs >= 1 child");
  return run_dense(children, n, kids, koff, {}, PCB_JOB_SCAN, 0.0, B, S, G, log_S, nullptr, nullptr);
}

int pcb_node_update_host(const double* log_p, const double* children, const int64_t* child_off, double* log_r,
                         int64_t B, int32_t S, int32_t G) {
  if (!log_p || !child_off || !log_r || B < 0) return fail(PCB_ERR_ARG, "node_update: null pointer");
  if (B == 0) return PCB_OK;
  const int64_t n = child_off[B];
  if (n > 0 && !children) return fail(PCB_ERR_ARG, "node_update: null children");
  const size_t t = (size_t)S * G;
  std::vector<double> in((size_t)(n + B) * t);
  if (n) memcpy(in.data(), children, (size_t)n * t * 8);
  memcpy(in.data() + (size_t)n * t, log_p, (size_t)B * t * 8);
  std::vector<int32_t> kids(n), base(B);
  for (int64_t i = 0; i < n; ++i) kids[i] = (int32_t)i;
  for (int64_t b = 0; b < B; ++b) base[b] = (int32_t)(n + b);
  std::vector<int64_t> koff(child_off, child_off + B + 1);
  return run_dense(in.data(), n + B, kids, koff, base, PCB_JOB_SCAN, 0.0, B, S, G, log_r, nullptr, nullptr);
}

int pcb_root_terms_host(const double* root_tiles, double* lse_sum, double* last_sum, int64_t B, int32_t S,
                        int32_t G) {
  if (!root_tiles || !lse_sum || !last_sum || B < 0) return fail(PCB_ERR_ARG, "root_terms: null pointer");
  if (B == 0) return PCB_OK;
  HostScratch& hs = g_hs;
  int rc;
  const size_t bytes = (size_t)B * S * G * 8;
  if ((rc = hs.in.ensure(bytes))) return rc;
  if ((rc = hs.rows_a.ensure((size_t)B * S * 8))) return rc;
  if ((rc = hs.rows_b.ensure((size_t)B * S * 8))) return rc;
  if ((rc = hs.misc.ensure((size_t)B * 16))) return rc;
  PCB_CUDA(cudaMemcpyAsync(hs.in.p, root_tiles, bytes, cudaMemcpyHostToDevice, 0));
  const long long warps = B * S;
  pcb::root_terms_rows_kernel<<<(unsigned)((warps + 3) / 4), 128>>>(hs.in.as<double>(), B, S, G, hs.rows_a.as<double>(),
                                                                     hs.rows_b.as<double>());
  PCB_LAUNCH_CHECK();
  double* red = hs.misc.as<double>();
  if ((rc = pcb_reduce_rows_dev(hs.rows_a.as<double>(), B, S, red, nullptr))) return rc;
  if ((rc = pcb_reduce_rows_dev(hs.rows_b.as<double>(), B, S, red + B, nullptr))) return rc;
  PCB_CUDA(cudaMemcpyAsync(lse_sum, red, (size_t)B * 8, cudaMemcpyDeviceToHost, 0));
  PCB_CUDA(cudaMemcpyAsync(last_sum, red + B, (size_t)B * 8, cudaMemcpyDeviceToHost, 0));
  PCB_CUDA(cudaStreamSynchronize(0));
  return PCB_OK;
}

int pcb_likelihood_grid_dev(const int64_t* refc, const int64_t* altc, const double* tc, const int64_t* cn,
                            const double* mu, const double* log_pi, const int32_t* n_geno, int64_t M, int32_t S,
                            int32_t Cmax, int32_t G, int32_t density, double precision, double* out, void* stream) {
  if (M <= 0) return PCB_OK;
  if (G < 2 || Cmax < 1 || (density != 0 && density != 1)) return fail(PCB_ERR_ARG, "likelihood_grid: bad arguments");
  const long long cells = (long long)M * S * G;
  pcb::likelihood_grid_kernel<<<(unsigned)((cells + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      (const long long*)refc, (const long long*)altc, tc, (const long long*)cn, mu, log_pi, n_geno, (long long)M * S,
      Cmax, G, density, precision, out);
  PCB_LAUNCH_CHECK();
  return PCB_OK;
}

int pcb_likelihood_grid_host(const int64_t* refc, const int64_t* altc, const double* tc, const int64_t* cn,
                             const double* mu, const double* log_pi, const int32_t* n_geno, int64_t M, int32_t S,
                             int32_t Cmax, int32_t G, int32_t density, double precision, double* out) {
  if (!refc || !altc || !tc || !cn || !mu || !log_pi || !n_geno || !out) return fail(PCB_ERR_ARG, "likelihood_grid: null");
  if (M <= 0) return PCB_OK;
  HostScratch& hs = g_hs;
  const size_t ms = (size_t)M * S;
  const size_t b_counts = ms * 8, b_cn = ms * Cmax * 3 * 8, b_lp = ms * Cmax * 8, b_ng = ms * 4;
  // one packed staging buffer: refc | altc | tc | cn | mu | log_pi | n_geno
  const size_t total = 3 * b_counts + 2 * b_cn + b_lp + b_ng;
  int rc;
  if ((rc = hs.in.ensure(total))) return rc;
  if ((rc = hs.out.ensure(ms * G * 8))) return rc;
  char* d = hs.in.as<char>();
  size_t o = 0;
  auto put = [&](const void* src, size_t n) -> void* {
    void* dst = d + o;
    cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, 0);
    o += n;
    return dst;
  };
  auto* d_ref = (const int64_t*)put(refc, b_counts);
  auto* d_alt = (const int64_t*)put(altc, b_counts);
  auto* d_tc = (const double*)put(tc, b_counts);
  auto* d_cn = (const int64_t*)put(cn, b_cn);
  auto* d_mu = (const double*)put(mu, b_cn);
  auto* d_lp = (const double*)put(log_pi, b_lp);
  auto* d_ng = (const int32_t*)put(n_geno, b_ng);
  PCB_CUDA(cudaGetLastError());
  if ((rc = pcb_likelihood_grid_dev(d_ref, d_alt, d_tc, d_cn, d_mu, d_lp, d_ng, M, S, Cmax, G, density, precision,
                                    hs.out.as<double>(), nullptr)))
    return rc;
  PCB_CUDA(cudaMemcpyAsync(out, hs.out.p, ms * G * 8, cudaMemcpyDeviceToHost, 0));
  PCB_CUDA(cudaStreamSynchronize(0));
  return PCB_OK;
}

int pcb_cluster_sum_host(const double* grids, const int64_t* cluster_off, int64_t K, int32_t S, int32_t G,
                         double* out) {
  if (!grids || !cluster_off || !out || K < 0) return fail(PCB_ERR_ARG, "cluster_sum: null pointer");
  if (K == 0) return PCB_OK;
  HostScratch& hs = g_hs;
  const long long tile = (long long)S * G;
  const int64_t M = cluster_off[K];
  int rc;
  if ((rc = hs.in.ensure((size_t)std::max<int64_t>(M, 1) * tile * 8))) return rc;
  if ((rc = hs.out.ensure((size_t)K * tile * 8))) return rc;
  if ((rc = hs.misc.ensure((size_t)(K + 1) * 8))) return rc;
  PCB_CUDA(cudaMemcpyAsync(hs.in.p, grids, (size_t)M * tile * 8, cudaMemcpyHostToDevice, 0));
  PCB_CUDA(cudaMemcpyAsync(hs.misc.p, cluster_off, (size_t)(K + 1) * 8, cudaMemcpyHostToDevice, 0));
  const long long n = K * tile;
  pcb::cluster_sum_kernel<<<(unsigned)((n + 255) / 256), 256>>>(hs.in.as<double>(), hs.misc.as<long long>(), K, tile,
                                                                hs.out.as<double>());
  PCB_LAUNCH_CHECK();
  PCB_CUDA(cudaMemcpyAsync(out, hs.out.p, (size_t)n * 8, cudaMemcpyDeviceToHost, 0));
  PCB_CUDA(cudaStreamSynchronize(0));
  return PCB_OK;
}

int pcb_outlier_marginal_host(const double* values, int64_t K, int32_t S, int32_t G, double* out) {
  if (!values || !out || K < 0) return fail(PCB_ERR_ARG, "outlier_marginal: null pointer");
  if (K == 0) return PCB_OK;
  HostScratch& hs = g_hs;
  int rc;
  if ((rc = hs.in.ensure((size_t)K * S * G * 8))) return rc;
  if ((rc = hs.rows_a.ensure((size_t)K * S * 8))) return rc;
  if ((rc = hs.misc.ensure((size_t)K * 8))) return rc;
  PCB_CUDA(cudaMemcpyAsync(hs.in.p, values, (size_t)K * S * G * 8, cudaMemcpyHostToDevice, 0));
  const long long warps = K * S;
  const int wpb = 4;
  pcb::outlier_marginal_rows_kernel<<<(unsigned)((warps + wpb - 1) / wpb), wpb * 32, (size_t)wpb * G * 8>>>(
      hs.in.as<double>(), K, S, G, hs.rows_a.as<double>());
  PCB_LAUNCH_CHECK();
  if ((rc = pcb_reduce_rows_dev(hs.rows_a.as<double>(), K, S, hs.misc.as<double>(), nullptr))) return rc;
  PCB_CUDA(cudaMemcpyAsync(out, hs.misc.p, (size_t)K * 8, cudaMemcpyDeviceToHost, 0));
  PCB_CUDA(cudaStreamSynchronize(0));
  return PCB_OK;
}

// ---- MAP prevalences (max-plus fold with back-pointers) ---------------------------------------------------
int pcb_map_trees_dev(const double* node_log_p, const int32_t* tree_off, const int32_t* post_order,
                      const int32_t* child_off, const int32_t* child_idx, int64_t n_trees, int64_t n_nodes,
                      int64_t n_edges, int32_t S, int32_t G, double* log_r_max, int32_t* d_choice, int32_t* s_choice,
                      int32_t* max_idx, void* stream) {
  if (n_trees <= 0 || n_nodes <= 0) return PCB_OK;
  if (!node_log_p || !tree_off || !post_order || !child_off || !log_r_max || !s_choice || !max_idx ||
      (n_edges > 0 && (!child_idx || !d_choice)))
    return fail(PCB_ERR_ARG, "map_trees_dev: null pointer");
  if (S < 1 || G < 1 || G > 4096) return fail(PCB_ERR_ARG, "map_trees_dev: need S >= 1 and 1 <= G <= 4096");
  if (n_trees * S > 2147483647ll) return fail(PCB_ERR_ARG, "map_trees_dev: too many trees for one launch");
  pcb::MapParams p{node_log_p, tree_off, post_order, child_off, child_idx, log_r_max, d_choice, s_choice, max_idx, S, G};
  const int threads = std::min(1024, 32 * ((G + 31) / 32));
  const size_t smem = (size_t)3 * G * 8;
  if (smem > 48 * 1024)
    PCB_CUDA(cudaFuncSetAttribute(pcb::map_trees_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pcb::map_trees_kernel<<<(unsigned)(n_trees * S), threads, smem, (cudaStream_t)stream>>>(p);
  PCB_LAUNCH_CHECK();
  return PCB_OK;
}

int pcb_map_trees_host(const double* node_log_p, const int32_t* tree_off, const int32_t* post_order,
                       const int32_t* child_off, const int32_t* child_idx, int64_t n_trees, int64_t n_nodes,
                       int64_t n_edges, int32_t S, int32_t G, int32_t* max_idx, double* log_r_max) {
  if (n_trees <= 0 || n_nodes <= 0) return PCB_OK;
  if (!node_log_p || !tree_off || !post_order || !child_off || !max_idx || (n_edges > 0 && !child_idx))
    return fail(PCB_ERR_ARG, "map_trees: null pointer");
  // structure checks on the host copy: a malformed tree must not send the kernel out of bounds
  if (tree_off[0] != 0 || tree_off[n_trees] != n_nodes || child_off[0] != 0 || child_off[n_nodes] != n_edges)
    return fail(PCB_ERR_ARG, "map_trees: offsets do not cover the nodes / edges");
  for (int64_t i = 0; i < n_nodes; ++i)
    if (post_order[i] < 0 || post_order[i] >= n_nodes || child_off[i + 1] < child_off[i])
      return fail(PCB_ERR_ARG, "map_trees: bad post-order entry or child offsets");
  for (int64_t e = 0; e < n_edges; ++e)
    if (child_idx[e] < 0 || child_idx[e] >= n_nodes) return fail(PCB_ERR_ARG, "map_trees: child index out of range");
  HostScratch& hs = g_hs;
  const size_t tile = (size_t)S * G;
  const size_t n_int = (size_t)(n_trees + 1) + n_nodes + (n_nodes + 1) + std::max<int64_t>(n_edges, 1);
  int rc;
  if ((rc = hs.in.ensure(n_nodes * tile * 8))) return rc;
  if ((rc = hs.out.ensure(n_nodes * tile * 8))) return rc;
  if ((rc = hs.ids.ensure(n_int * 4))) return rc;
  if ((rc = hs.jobs.ensure(std::max<int64_t>(n_edges, 1) * tile * 4))) return rc;
  if ((rc = hs.kids.ensure(n_nodes * tile * 4))) return rc;
  if ((rc = hs.misc.ensure((size_t)n_nodes * S * 4))) return rc;
  int32_t* d_tree_off = hs.ids.as<int32_t>();
  int32_t* d_post = d_tree_off + (n_trees + 1);
  int32_t* d_child_off = d_post + n_nodes;
  int32_t* d_child_idx = d_child_off + (n_nodes + 1);
  PCB_CUDA(cudaMemcpyAsync(hs.in.p, node_log_p, n_nodes * tile * 8, cudaMemcpyHostToDevice, 0));
  PCB_CUDA(cudaMemcpyAsync(d_tree_off, tree_off, (size_t)(n_trees + 1) * 4, cudaMemcpyHostToDevice, 0));
  PCB_CUDA(cudaMemcpyAsync(d_post, post_order, (size_t)n_nodes * 4, cudaMemcpyHostToDevice, 0));
  PCB_CUDA(cudaMemcpyAsync(d_child_off, child_off, (size_t)(n_nodes + 1) * 4, cudaMemcpyHostToDevice, 0));
  if (n_edges > 0) PCB_CUDA(cudaMemcpyAsync(d_child_idx, child_idx, (size_t)n_edges * 4, cudaMemcpyHostToDevice, 0));
  PCB_CUDA(cudaMemsetAsync(hs.misc.p, 0, (size_t)n_nodes * S * 4, 0));
  if ((rc = pcb_map_trees_dev(hs.in.as<double>(), d_tree_off, d_post, d_child_off, d_child_idx, n_trees, n_nodes, n_edges,
                              S, G, hs.out.as<double>(), hs.jobs.as<int32_t>(), hs.kids.as<int32_t>(),
                              hs.misc.as<int32_t>(), nullptr)))
    return rc;
  PCB_CUDA(cudaMemcpyAsync(max_idx, hs.misc.p, (size_t)n_nodes * S * 4, cudaMemcpyDeviceToHost, 0));
  if (log_r_max) PCB_CUDA(cudaMemcpyAsync(log_r_max, hs.out.p, n_nodes * tile * 8, cudaMemcpyDeviceToHost, 0));
  PCB_CUDA(cudaStreamSynchronize(0));
  return PCB_OK;
}

int pcb_conv_log_host(const double* log_x, const double* log_y, double* out, int64_t B, int32_t n) {
  if (B <= 0 || n <= 0) return PCB_OK;
  if (!log_x || !log_y || !out) return fail(PCB_ERR_ARG, "conv_log: null pointer");
  HostScratch& hs = g_hs;
  const size_t bytes = (size_t)B * n * 8;
  int rc;
  if ((rc = hs.in.ensure(2 * bytes))) return rc;
  if ((rc = hs.out.ensure(bytes))) return rc;
  double* dx = hs.in.as<double>();
  double* dy = dx + (size_t)B * n;
  PCB_CUDA(cudaMemcpyAsync(dx, log_x, bytes, cudaMemcpyHostToDevice, 0));
  PCB_CUDA(cudaMemcpyAsync(dy, log_y, bytes, cudaMemcpyHostToDevice, 0));
  const long long total = (long long)B * n;
  pcb::conv_log_kernel<<<(unsigned)((total + 127) / 128), 128>>>(dx, dy, hs.out.as<double>(), B, n);
  PCB_LAUNCH_CHECK();
  PCB_CUDA(cudaMemcpyAsync(out, hs.out.p, bytes, cudaMemcpyDeviceToHost, 0));
  PCB_CUDA(cudaStreamSynchronize(0));
  return PCB_OK;
}

int pcb_normalize_segments_dev(const double* log_p, const int64_t* seg_off_dev, int64_t U, double* log_q, double* q,
                                double* log_norm, double* ess, void* stream) {
  if (U <= 0) return PCB_OK;
  if (!log_p || !seg_off_dev || !log_q || !q) return fail(PCB_ERR_ARG, "normalize_segments_dev: null pointer");
  pcb::normalize_segments_kernel<<<(unsigned)U, 128, 0, (cudaStream_t)stream>>>(log_p, (const long long*)seg_off_dev,
                                                                               log_q, q, log_norm, ess);
  PCB_LAUNCH_CHECK();
  return PCB_OK;
}

static int normalize_segments_impl(const double* log_p_host, const int64_t* seg_off, int64_t U, double* log_q,
                                   double* q, double* log_norm, double* ess) {
  HostScratch& hs = g_hs;
  const int64_t n = seg_off[U];
  int rc;
  if ((rc = hs.in.ensure((size_t)std::max<int64_t>(n, 1) * 8))) return rc;
  if ((rc = hs.out.ensure((size_t)std::max<int64_t>(n, 1) * 16))) return rc;
  if ((rc = hs.misc.ensure((size_t)(U + 1) * 8 + (size_t)U * 16))) return rc;
  PCB_CUDA(cudaMemcpyAsync(hs.in.p, log_p_host, (size_t)n * 8, cudaMemcpyHostToDevice, 0));
  PCB_CUDA(cudaMemcpyAsync(hs.misc.p, seg_off, (size_t)(U + 1) * 8, cudaMemcpyHostToDevice, 0));
  double* d_lq = hs.out.as<double>();
  double* d_q = d_lq + n;
  double* d_ln = (double*)(hs.misc.as<char>() + (size_t)(U + 1) * 8);
  double* d_ess = d_ln + U;
  pcb::normalize_segments_kernel<<<(unsigned)U, 128>>>(hs.in.as<double>(), hs.misc.as<long long>(), d_lq, d_q, d_ln,
                                                       d_ess);
  PCB_LAUNCH_CHECK();
  if (log_q) PCB_CUDA(cudaMemcpyAsync(log_q, d_lq, (size_t)n * 8, cudaMemcpyDeviceToHost, 0));
  if (q) PCB_CUDA(cudaMemcpyAsync(q, d_q, (size_t)n * 8, cudaMemcpyDeviceToHost, 0));
  if (log_norm) PCB_CUDA(cudaMemcpyAsync(log_norm, d_ln, (size_t)U * 8, cudaMemcpyDeviceToHost, 0));
  if (ess) PCB_CUDA(cudaMemcpyAsync(ess, d_ess, (size_t)U * 8, cudaMemcpyDeviceToHost, 0));
  PCB_CUDA(cudaStreamSynchronize(0));
  return PCB_OK;
}

int pcb_swarm_weights_host(const double* log_w, int64_t N, double* log_Z, double* log_W, double* W, double* ess) {
  if (!log_w || N < 1) return fail(PCB_ERR_ARG, "swarm_weights: need N >= 1");
  const int64_t off[2] = {0, N};
  return normalize_segments_impl(log_w, off, 1, log_W, W, log_Z, ess);
}

int pcb_normalize_segments_host(const double* log_p, const int64_t* seg_off, int64_t U, double* log_q, double* q,
                                double* log_norm) {
  if (!log_p || !seg_off || U < 0) return fail(PCB_ERR_ARG, "normalize_segments: null pointer");
  if (U == 0) return PCB_OK;
  return normalize_segments_impl(log_p, seg_off, U, log_q, q, log_norm, nullptr);
}

int pcb_resample_host(const double* W, int64_t N, const double* uniforms, int64_t n_draws, int64_t* ancestors) {
  if (!W || !uniforms || !ancestors || N < 1 || n_draws < 0) return fail(PCB_ERR_ARG, "resample: bad arguments");
  if (n_draws == 0) return PCB_OK;
  HostScratch& hs = g_hs;
  int rc;
  if ((rc = hs.in.ensure((size_t)(2 * N + n_draws) * 8))) return rc;
  if ((rc = hs.out.ensure((size_t)n_draws * 8))) return rc;
  double* dW = hs.in.as<double>();
  double* dcdf = dW + N;
  double* du = dcdf + N;
  PCB_CUDA(cudaMemcpyAsync(dW, W, (size_t)N * 8, cudaMemcpyHostToDevice, 0));
  PCB_CUDA(cudaMemcpyAsync(du, uniforms, (size_t)n_draws * 8, cudaMemcpyHostToDevice, 0));
  pcb::resample_kernel<<<1, 1024>>>(dW, N, du, n_draws, hs.out.as<long long>(), dcdf);
  PCB_LAUNCH_CHECK();
  PCB_CUDA(cudaMemcpyAsync(ancestors, hs.out.p, (size_t)n_draws * 8, cudaMemcpyDeviceToHost, 0));
  PCB_CUDA(cudaStreamSynchronize(0));
  return PCB_OK;
}

// ---- device-resident SMC sweep ---------------------------------------------------------------------------
extern "C++" {
namespace {
struct SweepWorkspace {
  DevBuf ints, dbls, keys, outs;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
};
thread_local SweepWorkspace g_sw;

struct Carver {
  char* base;
  size_t off = 0;
  template <class T>
  T* take(size_t n) {
    off = (off + 15) & ~(size_t)15;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};
}  // namespace
}  // extern "C++"

int pcb_smc_sweep_dev(const pcb_arena* a, const pcb_sweep_desc* d, void* stream) {
  namespace sw = pcb::sweep;
  int rc = check_arena(a);
  if (rc) return rc;
  if (!d) return fail(PCB_ERR_ARG, "smc_sweep: null descriptor");
  const int N = d->N, T = d->T, Rmax = d->Rmax, S = a->layout.S;
  if (N < 1 || N > 1024 || T < 1 || Rmax < 1 || Rmax > sw::kMaxRoots)
    return fail(PCB_ERR_ARG, "smc_sweep: need 1 <= N <= 1024, T >= 1, 1 <= Rmax <= 64");
  if (d->conditional && (N < 2 || !d->ret_th_int || !d->ret_ba_int || !d->ret_th_dbl || !d->ret_ba_dbl || !d->ret_key ||
                         !d->ret_kind || !d->ret_arg))
    return fail(PCB_ERR_ARG, "smc_sweep: conditional sweep without a retained path (or N < 2)");
  if (!d->dp_idx || !d->value_id || !d->single_id || !d->dp_op || !d->dp_opn || !d->dp_marg || !d->lfact || !d->logn ||
      !d->rterm || d->n_tab < Rmax + 2)
    return fail(PCB_ERR_ARG, "smc_sweep: missing input table");
  if (!d->anc || !d->kind || !d->name || !d->mask || !d->raw || !d->logp || !d->logq || !d->ess || !d->resampled ||
      !d->final_idx || !d->status)
    return fail(PCB_ERR_ARG, "smc_sweep: missing output buffer");
  const long long scratch = (long long)N * Rmax;
  if (d->tile_base < 0 || d->tile_limit > a->capacity || d->tile_base + scratch + 2ll * N > d->tile_limit)
    return fail(PCB_ERR_ARG, "smc_sweep: tile range too small for the candidate scratch");
  for (int t = 0; t < T; ++t)
    if (d->value_id[t] < 0 || d->value_id[t] >= a->capacity || d->single_id[t] < 0 || d->single_id[t] >= a->capacity)
      return fail(PCB_ERR_ARG, "smc_sweep: data tile id outside the arena");
  cudaStream_t st = (cudaStream_t)stream;
  const int irec = (sw::IHEAD + sw::F_COUNT * Rmax + 3) & ~3;  // padded to 16 bytes: records move as int4
  const int n_ret_kids = d->conditional ? d->n_ret_kids : 0;
  const long long maxB = (long long)N * (Rmax + 1);
  const int pool_cap = n_ret_kids + N * T * std::min(Rmax, 16) + 64;
  const size_t TN = (size_t)T * N;

  // ---- carve the workspace (two passes: size, then pointers)
  sw::Params P{};
  size_t need_i = 0, need_d = 0, need_k = 0;
  for (int pass = 0; pass < 2; ++pass) {
    Carver ci{pass ? g_sw.ints.as<char>() : nullptr}, cd{pass ? g_sw.dbls.as<char>() : nullptr},
        ck{pass ? g_sw.keys.as<char>() : nullptr};
    P.dp_idx = ci.take<int>(T);
    P.value_id = ci.take<int>(T);
    P.single_id = ci.take<int>(T);
    P.ret_th_int = ci.take<int>(d->conditional ? (size_t)T * irec : 0);
    P.ret_ba_int = ci.take<int>(d->conditional ? (size_t)T * irec : 0);
    P.cur_dl = ci.take<int>(d->perm_dist ? (size_t)N * Rmax : 0);
    P.ext_dl = ci.take<int>(d->perm_dist ? (size_t)N * Rmax : 0);
    P.ret_th_dl = ci.take<int>(d->perm_dist && d->conditional ? (size_t)T * Rmax : 0);
    P.ret_kind = ci.take<int>(d->conditional ? T : 0);
    P.ret_arg = ci.take<int>(d->conditional ? T : 0);
    P.kid_pool = ci.take<int>(pool_cap);
    P.cur_int = ci.take<int>((size_t)N * irec);
    P.ext_int = ci.take<int>((size_t)N * irec);
    P.use_ret = ci.take<int>(N);
    P.rep = ci.take<int>(N + 1);
    P.job_off = ci.take<int>(N + 1);
    P.jobsB = ci.take<int>((size_t)maxB * 8);
    P.kidsB = ci.take<int>((size_t)maxB * Rmax);
    P.addA = ci.take<int>((size_t)N * Rmax);
    P.addB = ci.take<int>((size_t)N * Rmax);
    P.addO = ci.take<int>((size_t)N * Rmax);
    P.jobsD = ci.take<int>((size_t)N * 8);
    P.kidsD = ci.take<int>((size_t)N * Rmax);
    P.jobsE = ci.take<int>((size_t)2 * N * 8);
    P.kidsE = ci.take<int>((size_t)2 * N * Rmax);
    P.slotE = ci.take<int>(N);
    P.dec_kind = ci.take<int>(N);
    P.dec_name = ci.take<int>(N);
    P.nB = ci.take<int>(8);  // nB, nAdd, nD, nE, pool_top, tile_top, status, final
    P.o_anc = ci.take<int>(TN);
    P.o_kind = ci.take<int>(TN);
    P.o_name = ci.take<int>(TN);
    P.o_resampled = ci.take<int>(T);
    need_i = ci.off;
    P.dp_op = cd.take<double>(T);
    P.dp_opn = cd.take<double>(T);
    P.dp_marg = cd.take<double>(T);
    P.lfact = cd.take<double>(d->n_tab);
    P.logn = cd.take<double>(d->n_tab);
    P.rterm = cd.take<double>(d->n_tab);
    P.ret_th_dbl = cd.take<double>(d->conditional ? (size_t)T * sw::DREC : 0);
    P.ret_ba_dbl = cd.take<double>(d->conditional ? (size_t)T * sw::DREC : 0);
    P.cur_dbl = cd.take<double>((size_t)N * sw::DREC);
    P.ext_dbl = cd.take<double>((size_t)N * sw::DREC);
    P.lw = cd.take<double>(N);
    P.raw = cd.take<double>(N);
    P.scB_lse = cd.take<double>((size_t)maxB * S);
    P.scB_last = cd.take<double>((size_t)maxB * S);
    P.scE_lse = cd.take<double>((size_t)2 * N * S);
    P.scE_last = cd.take<double>((size_t)2 * N * S);
    P.dec_logq = cd.take<double>(N);
    P.dec_lp = cd.take<double>(N);
    P.dec_lp1 = cd.take<double>(N);
    P.dec_lpdf = cd.take<double>(N);
    P.cur_lc = cd.take<double>(d->perm_dist ? (size_t)N * Rmax : 0);
    P.ext_lc = cd.take<double>(d->perm_dist ? (size_t)N * Rmax : 0);
    P.ret_th_lc = cd.take<double>(d->perm_dist && d->conditional ? (size_t)T * Rmax : 0);
    P.o_raw = cd.take<double>(TN);
    P.o_logp = cd.take<double>(TN);
    P.o_logq = cd.take<double>(TN);
    P.o_ess = cd.take<double>(T);
    need_d = cd.off;
    P.ret_key = ck.take<sw::u64>(d->conditional ? T : 0);
    P.cur_key = ck.take<sw::u64>(N);
    P.ext_key = ck.take<sw::u64>(N);
    P.dec_mask = ck.take<sw::u64>(N);
    P.o_mask = ck.take<sw::u64>(TN);
    P.pairs = ck.take<sw::u64>(1);
    need_k = ck.off;
    if (!pass) {
      if ((rc = g_sw.ints.ensure(need_i + 64))) return rc;
      if ((rc = g_sw.dbls.ensure(need_d + 64))) return rc;
      if ((rc = g_sw.keys.ensure(need_k + 64))) return rc;
    }
  }
  P.nAdd = P.nB + 1;
  P.nD = P.nB + 2;
  P.nE = P.nB + 3;
  P.pool_top = P.nB + 4;
  P.tile_top = P.nB + 5;
  P.status = P.nB + 6;
  P.o_final = P.nB + 7;
  P.N = N;
  P.T = T;
  P.Rmax = Rmax;
  P.conditional = d->conditional ? 1 : 0;
  P.outliers = d->outliers ? 1 : 0;
  P.S = S;
  P.irec = irec;
  P.n_tab = d->n_tab;
  P.log_alpha = d->log_alpha;
  P.threshold = d->threshold;
  P.log_half = d->log_half;
  P.perm = d->perm_dist ? 1 : 0;
  P.seed = d->seed;
  P.pool_cap = pool_cap;
  P.tile_limit = (int)d->tile_limit;
  P.scratch0 = (int)d->tile_base;
  bool small_jobs = false;
  {
    const char* e = getenv("PCB_SWEEP_FUSE");
    const char* e2 = getenv("PCB_ROWS_MAX_PARTICLES");
    // a sweep of a few hundred particles launches a few hundred fold chains per step: the latency mapping
    small_jobs = (a->layout.kernel & PCB_KERNEL_LATENCY) && rows_supported(a->layout) && N <= ((e2 && *e2) ? atoi(e2) : 256);
    P.fuse = ((a->layout.kernel & 3) == 0 &&
              (small_jobs || a->layout.strip == 7 || a->layout.strip == 13 || a->layout.strip == 17) &&
              !(e && atoi(e) == 0)) ? 1 : 0;
  }
  const int jobs_mode = small_jobs ? kJobsSmall : 0;

  // ---- inputs: host -> device
  auto up = [&](const void* dst, const void* src, size_t bytes) -> cudaError_t {
    if (!bytes) return cudaSuccess;
    return cudaMemcpyAsync(const_cast<void*>(dst), src, bytes, cudaMemcpyHostToDevice, st);
  };
  PCB_CUDA(up(P.dp_idx, d->dp_idx, (size_t)T * 4));
  PCB_CUDA(up(P.value_id, d->value_id, (size_t)T * 4));
  PCB_CUDA(up(P.single_id, d->single_id, (size_t)T * 4));
  PCB_CUDA(up(P.dp_op, d->dp_op, (size_t)T * 8));
  PCB_CUDA(up(P.dp_opn, d->dp_opn, (size_t)T * 8));
  PCB_CUDA(up(P.dp_marg, d->dp_marg, (size_t)T * 8));
  PCB_CUDA(up(P.lfact, d->lfact, (size_t)d->n_tab * 8));
  PCB_CUDA(up(P.logn, d->logn, (size_t)d->n_tab * 8));
  PCB_CUDA(up(P.rterm, d->rterm, (size_t)d->n_tab * 8));
  if (d->conditional) {
    PCB_CUDA(up(P.ret_th_int, d->ret_th_int, (size_t)T * irec * 4));
    PCB_CUDA(up(P.ret_ba_int, d->ret_ba_int, (size_t)T * irec * 4));
    PCB_CUDA(up(P.ret_th_dbl, d->ret_th_dbl, (size_t)T * sw::DREC * 8));
    PCB_CUDA(up(P.ret_ba_dbl, d->ret_ba_dbl, (size_t)T * sw::DREC * 8));
    PCB_CUDA(up(P.ret_key, d->ret_key, (size_t)T * 8));
    PCB_CUDA(up(P.ret_kind, d->ret_kind, (size_t)T * 4));
    if (d->perm_dist) {
      if (!d->ret_subtree_data || !d->ret_subtree_log_count) return fail(PCB_ERR_ARG, "smc_sweep: perm_dist tables missing");
      PCB_CUDA(up(P.ret_th_dl, d->ret_subtree_data, (size_t)T * Rmax * 4));
      PCB_CUDA(up(P.ret_th_lc, d->ret_subtree_log_count, (size_t)T * Rmax * 8));
    }
    PCB_CUDA(up(P.ret_arg, d->ret_arg, (size_t)T * 4));
    if (n_ret_kids) PCB_CUDA(up(P.kid_pool, d->ret_kids, (size_t)n_ret_kids * 4));
  }
  const int counters[8] = {0, 0, 0, 0, n_ret_kids, (int)(d->tile_base + scratch), 0, -1};
  PCB_CUDA(up(P.nB, counters, sizeof(counters)));
  PCB_CUDA(cudaMemsetAsync(P.pairs, 0, 8, st));

  if (!g_sw.e0) {
    PCB_CUDA(cudaEventCreate(&g_sw.e0));
    PCB_CUDA(cudaEventCreate(&g_sw.e1));
  }
  PCB_CUDA(cudaEventRecord(g_sw.e0, st));
  const pcb_layout& l = a->layout;
  const int threads1 = std::min(1024, 32 * ((N + 31) / 32));
  sw::init_kernel<<<1, threads1, 0, st>>>(P, d->lw_uniform);
  PCB_LAUNCH_CHECK();
  int64_t launches = 1;
  const unsigned grid_build = (unsigned)((N + sw::kBuildWarps - 1) / sw::kBuildWarps);
  const unsigned grid_decide = (unsigned)((N + sw::kDecideWarps - 1) / sw::kDecideWarps);
  {
  PdlScope pdl_scope;
  for (int t = 0; t < T; ++t) {
    PCB_CUDA(launch_k(sw::build_kernel, dim3(grid_build), dim3(32 * sw::kBuildWarps), 0, st, P, t));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (t > 0) {
      const int wpb = 8;
      const long long max_rows = (long long)N * std::min(Rmax, t) * S;
      const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((max_rows + wpb - 1) / wpb, 148 * 8));
      PCB_CUDA(launch_k(sw::add_list_kernel, dim3(grid), dim3(wpb * 32), 0, st, P.addA, P.addB, P.addO, P.nAdd, a->lg, a->lin,
                        a->rmax, l.S, l.G, l.pitch, l.lead, (long long)l.tile_elems));
      g_launches.fetch_add(1, std::memory_order_relaxed);
      ++launches;
    }
    const long long boundB = (long long)N * (std::min(Rmax, t) + 1);
    if ((rc = jobs_launch_timed(a, P.jobsB, P.kidsB, boundB, d->prior, P.scB_lse, P.scB_last, stream, P.nB, jobs_mode)))
      return rc;
    PCB_CUDA(launch_k(sw::decide_kernel, dim3(grid_decide), dim3(32 * sw::kDecideWarps), 0, st, P, t));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    launches += 3;
    if (t > 0) {
      if (!P.fuse) {
        if ((rc = jobs_launch_timed(a, P.jobsD, P.kidsD, N, d->prior, nullptr, nullptr, stream, P.nD, jobs_mode)))
          return rc;
        ++launches;
      }
      if ((rc = jobs_launch_timed(a, P.jobsE, P.kidsE, 2ll * N, d->prior, P.scE_lse, P.scE_last, stream, P.nE,
                                  jobs_mode | (P.fuse ? kJobsFused : 0))))
        return rc;
      ++launches;
    }
    PCB_CUDA(launch_k(sw::weigh_kernel, dim3(1), dim3(threads1), 0, st, P, t, d->log_uniform, d->lw_uniform));  // a thread per particle
    g_launches.fetch_add(1, std::memory_order_relaxed);
    ++launches;
  }
  }
  PCB_CUDA(cudaEventRecord(g_sw.e1, st));

  // ---- outputs: device -> host, one synchronisation for the whole sweep
  auto down = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
  };
  int tail[8];
  PCB_CUDA(down(d->anc, P.o_anc, TN * 4));
  PCB_CUDA(down(d->kind, P.o_kind, TN * 4));
  PCB_CUDA(down(d->name, P.o_name, TN * 4));
  PCB_CUDA(down(d->mask, P.o_mask, TN * 8));
  PCB_CUDA(down(d->raw, P.o_raw, TN * 8));
  PCB_CUDA(down(d->logp, P.o_logp, TN * 8));
  PCB_CUDA(down(d->logq, P.o_logq, TN * 8));
  PCB_CUDA(down(d->ess, P.o_ess, (size_t)T * 8));
  PCB_CUDA(down(d->resampled, P.o_resampled, (size_t)T * 4));
  PCB_CUDA(down(tail, P.nB, sizeof(tail)));
  unsigned long long pairs = 0;
  PCB_CUDA(down(&pairs, P.pairs, 8));
  PCB_CUDA(cudaStreamSynchronize(st));
  if (d->conv_pairs) *d->conv_pairs = (int64_t)pairs;
  *d->final_idx = tail[7];
  *d->status = tail[6];
  if (d->tiles_used) *d->tiles_used = (int64_t)tail[5] - d->tile_base;
  if (d->n_launches) *d->n_launches = launches;
  if (d->device_ms) {
    float ms = 0.f;
    PCB_CUDA(cudaEventElapsedTime(&ms, g_sw.e0, g_sw.e1));
    *d->device_ms = ms;
  }
  return PCB_OK;
}

int pcb_fp64_peak(int32_t iters, double* flops_per_s, double* elapsed_ms) {
  if (!flops_per_s) return fail(PCB_ERR_ARG, "fp64_peak: null");
  int dev = 0, sms = 0;
  PCB_CUDA(cudaGetDevice(&dev));
  PCB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* d = nullptr;
  PCB_CUDA(cudaMalloc(&d, 8));
  cudaEvent_t e0, e1;
  PCB_CUDA(cudaEventCreate(&e0));
  PCB_CUDA(cudaEventCreate(&e1));
  const int blocks = sms * 8, threads = 256;
  pcb::fp64_peak_kernel<<<blocks, threads>>>(d, 16, 1.0);  // warm-up
  PCB_LAUNCH_CHECK();
  PCB_CUDA(cudaEventRecord(e0));
  pcb::fp64_peak_kernel<<<blocks, threads>>>(d, iters, 1.0);
  PCB_LAUNCH_CHECK();
  PCB_CUDA(cudaEventRecord(e1));
  PCB_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  PCB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  const double fmas = (double)blocks * threads * (double)iters * 16.0 * 8.0;
  *flops_per_s = 2.0 * fmas / (ms * 1e-3);
  if (elapsed_ms) *elapsed_ms = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return PCB_OK;
}

}  // extern "C"
