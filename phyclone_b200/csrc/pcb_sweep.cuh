// Device-resident (conditional) SMC sweep: proposal normalisation, draws, particle construction, weights / ESS and
// resampling for all T data points without a host round trip.
//
// Reference path (relative to the reference package phyclone/): smc/samplers/base.py:32-58 (sweep loop, _get_log_w),
// smc/samplers/conditional.py:76-125 and standard.py:20-45 (swarm init / update / resample),
// smc/kernels/semi_adapted.py:36-194 (candidates, log_q, sample), smc/kernels/base.py:38-67 (create_particle),
// smc/swarm/swarm.py:21-61 (weights, ESS), tree/distributions.py:39-133,186-219 (FS-CRP prior terms),
// mcmc/particle_gibbs.py:42-48 (final pick).
//
// A particle is the root level of its forest (what csrc/pcb_hostmod.cpp calls a ForestState): the ordered roots with
// their log_r / log_p tile ids, child tile lists, data counts and subtree sizes plus the running prior statistics
// and the structure hash.  Per step the host-free pipeline is
//     build   candidate tiles (root + value) and candidate score jobs of every particle        [1 CTA]
//     add     tile_add over the candidate list                                                  [grid]
//     jobs    node_jobs kernel: candidate scores (dummy-root folds)                             [persistent pool]
//     decide  log_q / q, the particle's draws, successor state, node + score jobs               [1 CTA]
//     jobs    node_jobs kernel: new node / re-folded root tiles                                 [persistent pool]
//     jobs    node_jobs kernel: scores of the new-node trees                                    [persistent pool]
//     weigh   log_w, log_Z, W, ESS, resampling (ancestors), state permutation                   [1 CTA]
// Draws are inverse-CDF transforms of uniforms keyed by (seed, step, slot, draw number); the CPU definition of this
// mode, decision for decision, is oracle/uniform.py + oracle/sampler.py run with a SweepRNG.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "pcb_device.cuh"

namespace pcb {
namespace sweep {

typedef unsigned long long u64;

constexpr int IHEAD = 5;  // R, n_nodes, n_slots, n_edges, n_out
enum { F_NAME = 0, F_LR, F_LP, F_NDATA, F_SIZE, F_SLOT, F_GIDX, F_KOFF, F_KCNT, F_COUNT };
constexpr int DREC = 8;  // crp, mult, oprior, omarg, log_p, log_p_one, log_w increment (retained path), spare
enum { D_CRP = 0, D_MULT, D_OPRIOR, D_OMARG, D_LOGP, D_LOGP1, D_LOGW };
enum { K_ATTACH = 0, K_NEW = 1, K_OUTLIER = 2 };
constexpr int FLAG_RETAINED_PARENT = 8;  // decision flag: the parent was value-equal to the retained particle
constexpr int JOB_SCAN = 1, JOB_ADD_PRIOR = 2;
constexpr u64 GOLDEN = 0x9E3779B97F4A7C15ull;
constexpr int RESAMPLE_SLOT = (1 << 20) - 1, FINAL_SLOT = (1 << 20) - 2;
enum { ST_OK = 0, ST_ROOTS = 1, ST_TILES = 2, ST_POOL = 4, ST_NUMERIC = 8 };

struct Params {
  int N, T, Rmax, conditional, outliers, S, irec, n_tab;
  int fuse;  // 1: a new node's tile and the score of its tree are ONE two-stage job (PCB_JOB_THEN_SCORE)
  int perm;  // 1: weights carry the root-permutation density (smc/kernels/base.py:49-55, smc/utils.py:18-66)
  // per root: data points in its subtree, log number of compatible orders of the subtree   [slot][Rmax]
  int *cur_dl, *ext_dl;
  double *cur_lc, *ext_lc, *dec_lpdf;
  const int* ret_th_dl;
  const double* ret_th_lc;
  double log_alpha, threshold, log_half;
  u64 seed;
  const int *dp_idx, *value_id, *single_id;
  const double *dp_op, *dp_opn, *dp_marg;
  const double *lfact, *logn, *rterm;
  const int *ret_th_int, *ret_ba_int;
  const double *ret_th_dbl, *ret_ba_dbl;
  const u64* ret_key;
  const int *ret_kind, *ret_arg;  // the retained path's own edit at step t: kind; node name (ATTACH) / #children (NEW)
  int *cur_int, *ext_int;
  double *cur_dbl, *ext_dbl;
  u64 *cur_key, *ext_key;
  double *lw, *raw;
  int *use_ret, *rep, *job_off;
  int *jobsB, *kidsB, *nB, *addA, *addB, *addO, *nAdd;
  int *jobsD, *kidsD, *nD, *jobsE, *kidsE, *nE;
  double *scB_lse, *scB_last, *scE_lse, *scE_last;
  int *slotE, *dec_kind, *dec_name;
  u64* dec_mask;
  double *dec_logq, *dec_lp, *dec_lp1;
  int *kid_pool, *pool_top, *tile_top;
  int pool_cap, tile_limit, scratch0;
  int *o_anc, *o_kind, *o_name;
  u64* o_mask;
  double *o_raw, *o_logp, *o_logq, *o_ess;
  int *o_resampled, *o_final, *status;
  u64* pairs;  // pair convolutions requested so far (sum over jobs of children - 1): the roofline's work count
};

__host__ __device__ inline u64 mix64(u64 z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// oracle/uniform.py::u01
__host__ __device__ inline double u01(u64 seed, int t, int i, int c) {
  const u64 idx = (((u64)t << 40) | ((u64)i << 20) | (u64)c) + 1ull;
  return (double)(mix64(seed + GOLDEN * idx) >> 11) * (1.0 / 9007199254740992.0);
}
// structure hash pieces of csrc/pcb_hostmod.cpp (its mix64 differs from the one above)
__device__ inline u64 hmix(u64 x) {
  x ^= x >> 30;
  x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27;
  x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
__device__ inline u64 edge_hash(long long e, long long s, long long d) {
  return hmix(hmix(hmix((u64)e + 0x9e3779b97f4a7c15ull) ^ (u64)s) ^ (u64)d);
}
__device__ inline u64 data_hash(long long node, long long pos, long long idx) {
  return hmix(hmix(hmix((u64)node + 0x632be59bd9b4e019ull) ^ (u64)pos) ^ (u64)idx) ^ 0x5bd1e995ull;
}

__device__ inline int* fld(int* rec, int f, int Rmax) { return rec + IHEAD + f * Rmax; }
__device__ inline const int* fld(const int* rec, int f, int Rmax) { return rec + IHEAD + f * Rmax; }

// which records describe particle i as a PARENT at step t: its own, or -- if its tree equals the retained particle's
// (structure hash) -- the retained path's thawed state and as-built base (the value-keyed proposal cache of
// smc/kernels/semi_adapted.py:219-237: the retained path's proposal is built first and serves every equal particle)
struct ParentView {
  const int *pI, *bI;
  const double *pD, *bD;
  const int* pDl;     // subtree data counts of the roots (thawed order), perm mode only
  const double* pLc;  // log order counts of the roots' subtrees
  bool ret;
};
__device__ inline ParentView parent_view(const Params& p, int t, int i) {
  ParentView v;
  v.ret = p.conditional && t >= 1 && p.cur_key[i] == p.ret_key[t - 1];
  if (v.ret) {
    v.pI = p.ret_th_int + (size_t)(t - 1) * p.irec;
    v.bI = p.ret_ba_int + (size_t)(t - 1) * p.irec;
    v.pD = p.ret_th_dbl + (size_t)(t - 1) * DREC;
    v.bD = p.ret_ba_dbl + (size_t)(t - 1) * DREC;
    v.pDl = p.perm ? p.ret_th_dl + (size_t)(t - 1) * p.Rmax : nullptr;
    v.pLc = p.perm ? p.ret_th_lc + (size_t)(t - 1) * p.Rmax : nullptr;
  } else {
    v.pI = v.bI = p.cur_int + (size_t)i * p.irec;
    v.pD = v.bD = p.cur_dbl + (size_t)i * DREC;
    v.pDl = p.perm ? p.cur_dl + (size_t)i * p.Rmax : nullptr;
    v.pLc = p.perm ? p.cur_lc + (size_t)i * p.Rmax : nullptr;
  }
  return v;
}

constexpr int JOB_THEN_SCORE = 4;
__device__ inline void put_job(int* jobs, int j, int coff, int C, int base, int out, int flags, int slot, int base2,
                               int more = 0) {
  int* r = jobs + (size_t)j * 8;
  r[0] = coff;
  r[1] = C;
  r[2] = base;
  r[3] = out;
  r[4] = flags;
  r[5] = slot;
  r[6] = base2 + 1;
  r[7] = more;
}

// ------------------------------------------------------------------------------------------------------------
// build: candidate tiles + candidate score jobs of step t (semi_adapted.py:85-141).  One warp per particle.
// Candidates of a parent with R roots, in the reference's order: attach to root 0..R-1 (base order), the outlier
// tree if outliers are modelled, the lone new-root tree if the parent is empty.  Particles with equal trees (equal
// structure hash -- resampled copies, or lineages that met) share ONE set of candidates, built by the first of them
// (`rep`): the value-keyed proposal cache of the reference (semi_adapted.py:219-237), which is also what keeps the
// candidate launch at the number of DISTINCT parents.
// ------------------------------------------------------------------------------------------------------------
constexpr int kBuildWarps = 8;

__global__ void build_kernel(const Params p, const int t) {
  pdl_trigger();
  pdl_wait();
  // slot 0 of a conditional sweep is the retained particle: it builds candidates too (its own proposal probability
  // is needed for its weight, and every particle equal to it shares them), but its state comes from the path tables
  const int first = 0;
  const int Rmax = p.Rmax;
  const int lane = threadIdx.x & 31;
  const int i = first + blockIdx.x * kBuildWarps + (threadIdx.x >> 5);
  if (i >= p.N) return;
  if (t > 0 && !(p.conditional && i == 0)) {
    // this particle's parent state: the record of the particle it was resampled from (ancestors of step t-1; the keys
    // were permuted by the weigh kernel already -- every warp's search below reads all of them)
    const int a = p.o_anc[(size_t)(t - 1) * p.N + i];
    const int4* src = reinterpret_cast<const int4*>(p.ext_int + (size_t)a * p.irec);
    int4* dst = reinterpret_cast<int4*>(p.cur_int + (size_t)i * p.irec);
    const int n4 = p.irec >> 2;
#pragma unroll 4
    for (int k = lane; k < n4; k += 32) dst[k] = src[k];
    if (lane < DREC) p.cur_dbl[(size_t)i * DREC + lane] = p.ext_dbl[(size_t)a * DREC + lane];
    if (p.perm)
      for (int k = lane; k < Rmax; k += 32) {
        p.cur_dl[(size_t)i * Rmax + k] = p.ext_dl[(size_t)a * Rmax + k];
        p.cur_lc[(size_t)i * Rmax + k] = p.ext_lc[(size_t)a * Rmax + k];
      }
    __syncwarp();
  }
  const u64 key = p.cur_key[i];
  int rep = i;
  for (int j = first + lane; j < i; j += 32)
    if (p.cur_key[j] == key) {
      rep = j;
      break;
    }
  rep = __reduce_min_sync(0xffffffffu, rep);
  if (lane == 0) p.rep[i] = rep;
  if (rep != i) return;
  const ParentView v = parent_view(p, t, i);
  const int R = v.bI[0];
  const int* lr = fld(v.bI, F_LR, Rmax);
  const int nj = R + ((p.outliers && R > 0) ? 1 : 0) + (R == 0 ? 1 : 0);
  int j0 = 0, a0 = 0;
  if (lane == 0) {
    j0 = atomicAdd(p.nB, nj);
    a0 = atomicAdd(p.nAdd, R);
    p.job_off[i] = j0;
    if (R > 1) atomicAdd(p.pairs, (u64)(R + (p.outliers ? 1 : 0)) * (u64)(R - 1));
  }
  j0 = __shfl_sync(0xffffffffu, j0, 0);
  a0 = __shfl_sync(0xffffffffu, a0, 0);
  const int vid = p.value_id[t], sid = p.single_id[t];
  for (int r = lane; r < R; r += 32) {
    const int c = p.scratch0 + a0 + r;
    p.addA[a0 + r] = lr[r];
    p.addB[a0 + r] = vid;
    p.addO[a0 + r] = c;
    const int j = j0 + r;
    int* kids = p.kidsB + (size_t)j * Rmax;
    for (int q = 0; q < R; ++q) kids[q] = (q == r) ? c : lr[q];
    put_job(p.jobsB, j, j * Rmax, R, -1, -1, JOB_SCAN | JOB_ADD_PRIOR, j, -1);
  }
  if (R > 0 && p.outliers) {
    const int j = j0 + R;
    int* kids = p.kidsB + (size_t)j * Rmax;
    for (int q = lane; q < R; q += 32) kids[q] = lr[q];
    if (lane == 0) put_job(p.jobsB, j, j * Rmax, R, -1, -1, JOB_SCAN | JOB_ADD_PRIOR, j, -1);
  }
  if (R == 0 && lane == 0) {
    p.kidsB[(size_t)j0 * Rmax] = sid;
    put_job(p.jobsB, j0, j0 * Rmax, 1, -1, -1, JOB_SCAN | JOB_ADD_PRIOR, j0, -1);
  }
}

// out = a + b for a device-counted list (tree/tree_node.py:47-52: log_r += value), one warp per (tile, row)
__global__ void add_list_kernel(const int* __restrict__ a_ids, const int* __restrict__ b_ids,
                                const int* __restrict__ out_ids, const int* __restrict__ n_dev, double* lg, double* lin,
                                double* rmax, int S, int G, int pitch, int lead, long long tile_elems) {
  pdl_trigger();
  pdl_wait();
  const long long n_rows = (long long)(*n_dev) * S;
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); wid < n_rows; wid += warps) {
    const long long tl = wid / S;
    const int s = (int)(wid - tl * S);
    const long long off = (long long)s * pitch + lead;
    const double* pa = lg + (long long)a_ids[tl] * tile_elems + off;
    const double* pb = lg + (long long)b_ids[tl] * tile_elems + off;
    const long long o = (long long)out_ids[tl] * tile_elems + off;
    double m = -INFINITY;
    for (int k = lane; k < G; k += 32) m = fmax(m, pa[k] + pb[k]);
    m = group_max(m, 32);
    for (int k = lane; k < G; k += 32) {
      const double v = pa[k] + pb[k];
      lg[o + k] = v;
      lin[o + k] = (v == -INFINITY) ? 0.0 : exp(v - m);
    }
    if (lane == 0) rmax[(long long)out_ids[tl] * S + s] = m;
  }
}

// sum over samples in sample order (tree/distributions.py:197-200)
__device__ inline double row_sum(const double* rows, int slot, int S) {
  double acc = 0.0;
  const double* r = rows + (size_t)slot * S;
  for (int s = 0; s < S; ++s) acc += r[s];
  return acc;
}

// beyond the host's tables (never in practice: they cover T + 2 entries).  Out of line on purpose: inlined, every one of
// the ~25 look-ups of decide_kernel carried its own lgamma + log expansion -- 12 k static instructions, 189 KB, with
// the hot path scattered between cold blocks (a third of the kernel's stall samples were instruction fetches)
__device__ __noinline__ double tab_beyond(int n, bool lf) { return lf ? lgamma((double)n + 1.0) : log((double)n); }

__device__ inline double tab(const double* t, int n, int n_tab, bool lf) {
  if (n < n_tab) return t[n];
  return tab_beyond(n, lf);
}

// (log_p prior, log_p_one prior) from running statistics: csrc/pcb_hostmod.cpp::priors (tree/distributions.py:39-133);
// `sizes` = subtree sizes of the roots in order
__device__ inline void priors(const Params& p, long n, double crp, double mult, const int* sizes, int n_sizes,
                              double oprior, double omarg, double& lp, double& lp1) {
  const double base = (double)n * p.log_alpha + crp;
  lp = base - (double)(n - 1) * tab(p.logn, (int)n + 1, p.n_tab, false) - mult;
  double ways = 0.0;
  for (int i = 0; i < n_sizes; ++i) ways += (double)(sizes[i] - 1) * tab(p.logn, sizes[i], p.n_tab, false);
  lp1 = base + (-ways + p.rterm[n_sizes]) - mult;
  const double extra = oprior + omarg;
  lp += extra;
  lp1 += extra;
}

// smallest k with q[0] + ... + q[k] > u (left-to-right float64 sums); the last index catches the remainder
__device__ inline int inv_cdf(const double* q, int n, double u) {
  double acc = 0.0;
  for (int k = 0; k < n - 1; ++k) {
    acc += q[k];
    if (acc > u) return k;
  }
  return n - 1;
}

__device__ inline int eI_R_after(int kind, int R, int k_new) { return kind == K_NEW ? R - k_new + 1 : R; }

__device__ inline int alloc_tiles(const Params& p, int n) {
  const int t0 = atomicAdd(p.tile_top, n);
  if (t0 + n > p.tile_limit) {
    atomicOr(p.status, ST_TILES);
    return p.scratch0;  // in bounds; the sweep is void and the host reruns it with a larger arena
  }
  return t0;
}

// -log(number of root permutations) of a forest from its roots' (log order count, subtree data count), n_all data
// points of which n_out are outliers: RootPermutationDistribution.log_pdf (smc/utils.py:18-66)
__device__ inline double perm_log_pdf(const Params& p, const double* lc, const int* dl, int R, int n_all, int n_out) {
  double count = 0.0;
  int total = 0;
  for (int r = 0; r < R; ++r) count += lc[r];
  if (R > 0) {
    for (int r = 0; r < R; ++r) total += dl[r];
    double multi = tab(p.lfact, total, p.n_tab, true);
    for (int r = 0; r < R; ++r) multi -= tab(p.lfact, dl[r], p.n_tab, true);
    count += multi;
  }
  count += tab(p.lfact, n_all, p.n_tab, true) - tab(p.lfact, n_out, p.n_tab, true) -
           tab(p.lfact, n_all - n_out, p.n_tab, true);
  return -count;
}

constexpr int kMaxRoots = 64;
constexpr int kDecideWarps = 4;
constexpr int kRec4 = (IHEAD + F_COUNT * kMaxRoots + 3) / 4;

// ------------------------------------------------------------------------------------------------------------
// decide: proposal normalisation (a17), the particle's draws, successor state, node / score jobs.
// One warp per particle: lane r scores candidate r and copies root r of the state; the scalar logic is executed
// redundantly by all lanes (warp-uniform), side effects by lane 0.
// Jobs: table D = tiles of NEW nodes (the scores of their trees need them), table E = scores of the new-node trees
// plus the refreshed tiles of roots that took the data point (only the next step reads those): two dependent launches.
// ------------------------------------------------------------------------------------------------------------
__global__ void decide_kernel(const Params p, const int t) {
  __shared__ double sh_lp[kDecideWarps][kMaxRoots + 2], sh_lp1[kDecideWarps][kMaxRoots + 2],
      sh_q[kDecideWarps][kMaxRoots + 2];
  __shared__ int4 sh_rec[kDecideWarps][2][kRec4];
  pdl_trigger();
  pdl_wait();
  const int first = 0;  // slot 0 of a conditional sweep: the retained particle, whose edit is given (see below)
  const int Rmax = p.Rmax;
  const int lane = threadIdx.x & 31, wslot = threadIdx.x >> 5;
  const int i = first + blockIdx.x * kDecideWarps + wslot;
  if (i >= p.N) return;
  double* c_lp = sh_lp[wslot];
  double* c_lp1 = sh_lp1[wslot];
  double* q = sh_q[wslot];
  const int vid = p.value_id[t], sid = p.single_id[t], dpidx = p.dp_idx[t];
  const double op = p.dp_op[t], opn = p.dp_opn[t], marg = p.dp_marg[t];
  const ParentView v = parent_view(p, t, i);
  // the parent's record(s) staged in shared memory: the scalar logic below walks them many times
  {
    const int n4 = p.irec >> 2;
    const int4* gb = reinterpret_cast<const int4*>(v.bI);
    const int4* gp = reinterpret_cast<const int4*>(v.pI);
#pragma unroll 4
    for (int k = lane; k < n4; k += 32) sh_rec[wslot][0][k] = gb[k];
    if (v.ret) {
#pragma unroll 4
      for (int k = lane; k < n4; k += 32) sh_rec[wslot][1][k] = gp[k];
    }
    __syncwarp();
  }
  const int* bI = reinterpret_cast<const int*>(sh_rec[wslot][0]);
  const int* pI = v.ret ? reinterpret_cast<const int*>(sh_rec[wslot][1]) : bI;
  const int R = bI[0];
  const int j0 = p.job_off[p.rep[i]];
  // ---- candidates, scored on the base state (semi_adapted.py:85-141; pcb_hostmod.cpp score_attach_all /
  //      score_outlier / score_new)
  const int* b_nd = fld(bI, F_NDATA, Rmax);
  const int* b_sz = fld(bI, F_SIZE, Rmax);
  const int* b_nm = fld(bI, F_NAME, Rmax);
  const long bn = bI[1];
  const double b_crp = v.bD[D_CRP], b_mult = v.bD[D_MULT], b_opr = v.bD[D_OPRIOR], b_om = v.bD[D_OMARG];
  int n_cand = R;
  {
    const double opr = (op != 0.0) ? b_opr + opn : b_opr;
    const double head = (double)bn * p.log_alpha;
    double ways = 0.0;
    for (int r = 0; r < R; ++r) ways += (double)(b_sz[r] - 1) * tab(p.logn, b_sz[r], p.n_tab, false);
    const double tail_p = -(double)(bn - 1) * tab(p.logn, (int)bn + 1, p.n_tab, false) - b_mult + (opr + b_om);
    const double tail_one = (-ways + p.rterm[R]) - b_mult + (opr + b_om);
    for (int r = lane; r < R; r += 32) {
      const int m = b_nd[r];
      const double crp =
          head + (b_crp + (m >= 1 ? tab(p.lfact, m, p.n_tab, true) - tab(p.lfact, m - 1, p.n_tab, true) : 0.0));
      c_lp[r] = (crp + tail_p) + row_sum(p.scB_lse, j0 + r, p.S);
      c_lp1[r] = (crp + tail_one) + row_sum(p.scB_last, j0 + r, p.S);
    }
  }
  int out_idx = -1;
  if (p.outliers) {
    out_idx = n_cand;
    if (lane == 0) {
      const double opr = (op != 0.0) ? b_opr + op : b_opr;
      double lp, lp1;
      priors(p, bn, b_crp, b_mult, b_sz, R, opr, b_om + marg, lp, lp1);
      if (R > 0) {
        lp += row_sum(p.scB_lse, j0 + R, p.S);
        lp1 += row_sum(p.scB_last, j0 + R, p.S);
      }
      c_lp[n_cand] = lp;
      c_lp1[n_cand] = lp1;
    }
    ++n_cand;
  }
  if (R == 0) {
    if (lane == 0) {
      const double opr = (op != 0.0) ? b_opr + opn : b_opr;
      const int one = 1;
      double lp, lp1;
      priors(p, bn + 1, b_crp, b_mult, &one, 1, opr, b_om, lp, lp1);
      lp += row_sum(p.scB_lse, j0, p.S);
      lp1 += row_sum(p.scB_last, j0, p.S);
      c_lp[n_cand] = lp;
      c_lp1[n_cand] = lp1;
    }
    ++n_cand;
  }
  __syncwarp();
  // log_q = log_p - LSE(log_p); q = exp(log_q) / sum   (semi_adapted.py:151-163, utils/math.py:102-121); sums run
  // left to right as in the reference's scalar loop
  double mx = -INFINITY;
  for (int c = 0; c < n_cand; ++c) mx = fmax(mx, c_lp[c]);
  if (!(mx > -INFINITY) || !(mx < INFINITY)) {
    if (lane == 0) atomicOr(p.status, ST_NUMERIC);
  }
  for (int c = lane; c < n_cand; c += 32) q[c] = exp(c_lp[c] - mx);
  __syncwarp();
  double tot = 0.0;
  for (int c = 0; c < n_cand; ++c) tot += q[c];
  const double lse = log(tot) + mx;
  __syncwarp();
  for (int c = lane; c < n_cand; c += 32) q[c] = exp(c_lp[c] - lse);
  __syncwarp();
  double qs = 0.0;
  for (int c = 0; c < n_cand; ++c) qs += q[c];
  __syncwarp();
  for (int c = lane; c < n_cand; c += 32) q[c] /= qs;
  __syncwarp();

  if (p.conditional && i == 0) {
    // the retained particle: its edit is the one the current tree dictates (conditional.py:19-74); only the proposal's
    // log-probability of that edit is needed (semi_adapted.py:36-61), for its weight
    const int rk = p.ret_kind[t], ra = p.ret_arg[t];
    double lq;
    if (rk == K_NEW && R > 0) {
      const double lbin =
          tab(p.lfact, R, p.n_tab, true) - tab(p.lfact, ra, p.n_tab, true) - tab(p.lfact, R - ra, p.n_tab, true);
      lq = p.log_half - (tab(p.logn, R + 1, p.n_tab, false) + lbin);
    } else {
      int pk;
      if (rk == K_OUTLIER) {
        pk = out_idx;
      } else if (rk == K_NEW) {
        pk = n_cand - 1;
      } else {
        pk = 0;
        while (pk < R && b_nm[pk] != ra) ++pk;
      }
      if (pk < 0 || pk >= n_cand) {
        if (lane == 0) atomicOr(p.status, ST_NUMERIC);
        pk = 0;
      }
      lq = ((R == 0) ? 0.0 : p.log_half) + (c_lp[pk] - lse);
    }
    if (lane == 0) p.dec_logq[0] = lq;
    return;
  }
  // ---- the particle's draws (semi_adapted.py:63-83, 165-194; conventions of oracle/uniform.py)
  int kind, pick = -1, k_new = 0;
  u64 taken_base = 0;  // NEW: chosen positions in BASE order (= the parent's tree_roots the reference draws from)
  double logq;
  if (R == 0) {
    pick = inv_cdf(q, n_cand, u01(p.seed, t, i, 0));
    kind = (pick == out_idx) ? K_OUTLIER : K_NEW;
    logq = c_lp[pick] - lse;
  } else if (u01(p.seed, t, i, 0) < 0.5) {
    pick = inv_cdf(q, n_cand, u01(p.seed, t, i, 1));
    kind = (pick == out_idx) ? K_OUTLIER : K_ATTACH;
    logq = p.log_half + (c_lp[pick] - lse);
  } else {
    kind = K_NEW;
    k_new = min(R, (int)(u01(p.seed, t, i, 1) * (double)(R + 1)));
    unsigned char pos[kMaxRoots];
    for (int r = 0; r < R; ++r) pos[r] = (unsigned char)r;
    for (int m = 0; m < k_new; ++m) {
      const int rest = R - m;
      const int j = m + min(rest - 1, (int)(u01(p.seed, t, i, 2 + m) * (double)rest));
      const unsigned char tmp = pos[m];
      pos[m] = pos[j];
      pos[j] = tmp;
      taken_base |= 1ull << pos[m];
    }
    const double lbin =
        tab(p.lfact, R, p.n_tab, true) - tab(p.lfact, k_new, p.n_tab, true) - tab(p.lfact, R - k_new, p.n_tab, true);
    logq = p.log_half - (tab(p.logn, R + 1, p.n_tab, false) + lbin);
  }

  // ---- successor state = extend(thawed parent, edit)   (pcb_hostmod.cpp fstate_extend)
  const int* s_nm = fld(pI, F_NAME, Rmax);
  const int* s_lr = fld(pI, F_LR, Rmax);
  const int* s_lp = fld(pI, F_LP, Rmax);
  const int* s_nd = fld(pI, F_NDATA, Rmax);
  const int* s_sz = fld(pI, F_SIZE, Rmax);
  const int* s_sl = fld(pI, F_SLOT, Rmax);
  const int* s_gi = fld(pI, F_GIDX, Rmax);
  const int* s_ko = fld(pI, F_KOFF, Rmax);
  const int* s_kc = fld(pI, F_KCNT, Rmax);
  const int pn = pI[1], ps = pI[2], pe = pI[3], po = pI[4];
  const double s_crp = v.pD[D_CRP], s_mult = v.pD[D_MULT], s_opr = v.pD[D_OPRIOR], s_om = v.pD[D_OMARG];
  const u64 s_key = v.ret ? p.ret_key[t - 1] : p.cur_key[i];
  int* eI = p.ext_int + (size_t)i * p.irec;
  double* eD = p.ext_dbl + (size_t)i * DREC;
  int* e_nm = fld(eI, F_NAME, Rmax);
  int* e_lr = fld(eI, F_LR, Rmax);
  int* e_lp = fld(eI, F_LP, Rmax);
  int* e_nd = fld(eI, F_NDATA, Rmax);
  int* e_sz = fld(eI, F_SIZE, Rmax);
  int* e_sl = fld(eI, F_SLOT, Rmax);
  int* e_gi = fld(eI, F_GIDX, Rmax);
  int* e_ko = fld(eI, F_KOFF, Rmax);
  int* e_kc = fld(eI, F_KCNT, Rmax);
  double lp_new = 0.0, lp1_new = 0.0;
  int name_out = -1, slot_e = -1;
  u64 mask_out = 0;
  if (kind != K_NEW || R == 0) {
    lp_new = c_lp[pick];
    lp1_new = c_lp1[pick];
  }
  if (kind == K_OUTLIER || kind == K_ATTACH) {
    for (int r = lane; r < R; r += 32) {
      e_nm[r] = s_nm[r];
      e_lr[r] = s_lr[r];
      e_lp[r] = s_lp[r];
      e_nd[r] = s_nd[r];
      e_sz[r] = s_sz[r];
      e_sl[r] = s_sl[r];
      e_gi[r] = s_gi[r];
      e_ko[r] = s_ko[r];
      e_kc[r] = s_kc[r];
    }
    __syncwarp();
  }
  if (kind == K_OUTLIER) {
    if (lane == 0) {
      eI[0] = R;
      eI[1] = pn;
      eI[2] = ps;
      eI[3] = pe;
      eI[4] = po + 1;
      eD[D_MULT] = s_mult;
      eD[D_CRP] = s_crp;
      eD[D_OMARG] = s_om + marg;
      eD[D_OPRIOR] = (op != 0.0) ? s_opr + op : s_opr;
      p.ext_key[i] = s_key ^ data_hash(-1, po, dpidx);
    }
  } else if (kind == K_ATTACH) {
    const int name = b_nm[pick];
    int j = 0;
    while (j < R && s_nm[j] != name) ++j;
    if (j >= R) {
      if (lane == 0) atomicOr(p.status, ST_NUMERIC);
      j = 0;
    }
    const int n = s_nd[j];
    const int kc = s_kc[j];
    const bool leaf = kc == 0;
    int t1 = 0, je = 0;
    if (lane == 0) {
      t1 = alloc_tiles(p, leaf ? 1 : 2);
      je = atomicAdd(p.nE, leaf ? 1 : 2);
    }
    t1 = __shfl_sync(0xffffffffu, t1, 0);
    je = __shfl_sync(0xffffffffu, je, 0);
    if (!leaf) {
      // log_r = log_p + log_S(children) re-derived as Tree.from_dict does (tree/tree_node.py:61-71), with
      // log_p = old log_p + value formed on the fly (base + base2)
      int* kids = p.kidsE + (size_t)(je + 1) * Rmax;
      for (int c = lane; c < kc; c += 32) kids[c] = p.kid_pool[s_ko[j] + c];
    }
    if (lane == 0) {
      // log_p += value (tree/tree_node.py:42-46): copy of the old log_p tile plus the value tile
      p.kidsE[(size_t)je * Rmax] = s_lp[j];
      put_job(p.jobsE, je, je * Rmax, 1, vid, t1, 0, -1, -1);
      e_lp[j] = t1;
      if (leaf) {
        e_lr[j] = t1;
      } else {
        put_job(p.jobsE, je + 1, (je + 1) * Rmax, kc, s_lp[j], t1 + 1, JOB_SCAN, -1, vid);
        e_lr[j] = t1 + 1;
        if (kc > 1) atomicAdd(p.pairs, (u64)(kc - 1));
      }
      e_nd[j] = n + 1;
      eI[0] = R;
      eI[1] = pn;
      eI[2] = ps;
      eI[3] = pe;
      eI[4] = po;
      eD[D_MULT] = s_mult;
      eD[D_CRP] = s_crp + (n >= 1 ? tab(p.lfact, n, p.n_tab, true) - tab(p.lfact, n - 1, p.n_tab, true) : 0.0);
      eD[D_OMARG] = s_om;
      eD[D_OPRIOR] = (op != 0.0) ? s_opr + opn : s_opr;
      p.ext_key[i] = s_key ^ data_hash(name, n, dpidx);
    }
    name_out = name;
  } else {
    // positions taken in the THAWED parent's root order; the new node folds its children in that order
    u64 take = 0;
    for (int r = 0; r < R; ++r)
      if (taken_base >> r & 1ull) {
        const int name = b_nm[r];
        int j = 0;
        while (j < R && s_nm[j] != name) ++j;
        if (j < R) take |= 1ull << j;
      }
    const int k = k_new;
    const int Rn = R - k + 1;
    if (Rn > Rmax) {
      if (lane == 0) atomicOr(p.status, ST_ROOTS);
    } else {
      int lr_new = sid, koff = 0, jd = 0;
      const bool score = R > 0;  // an empty parent's lone candidate already is this tree
      const bool fused = p.fuse && k > 0 && score;  // node tile and tree score as one two-stage job in table E
      if (lane == 0) {
        if (k > 0) {
          lr_new = alloc_tiles(p, 1);
          koff = atomicAdd(p.pool_top, k);
          if (!fused) jd = atomicAdd(p.nD, 1);
          if (koff + k > p.pool_cap) atomicOr(p.status, ST_POOL);
          if (k > 1) atomicAdd(p.pairs, (u64)(k - 1));
        }
        if (score) {
          slot_e = atomicAdd(p.nE, 1);
          if (Rn > 1) atomicAdd(p.pairs, (u64)(Rn - 1));
        }
      }
      lr_new = __shfl_sync(0xffffffffu, lr_new, 0);
      koff = __shfl_sync(0xffffffffu, koff, 0);
      jd = __shfl_sync(0xffffffffu, jd, 0);
      slot_e = __shfl_sync(0xffffffffu, slot_e, 0);
      const bool pool_ok = koff + k <= p.pool_cap;
      int* kidsE = score ? p.kidsE + (size_t)slot_e * Rmax : nullptr;
      // fused: [adopted roots (stage 1) | roots left alone (stage 2)] in one child list; else the node job's list
      int* kidsD = fused ? kidsE : p.kidsD + (size_t)jd * Rmax;
      const int e_shift = fused ? k - 1 : 0;  // stage-2 children start at index k of the fused list
      int sz = 1;
      u64 key = s_key ^ edge_hash(pe, 0, ps) ^ data_hash(pn, 0, dpidx);
      for (int r = 0; r < R; ++r) {
        if (take >> r & 1ull) {
          sz += s_sz[r];
          key ^= edge_hash(s_sl[r], 0, s_gi[r]) ^ edge_hash(s_sl[r], ps, s_gi[r]);
        }
      }
      for (int r = lane; r < R; r += 32) {
        const u64 below = (r == 0) ? 0ull : (take & ((1ull << r) - 1ull));
        if (take >> r & 1ull) {
          const int c = __popcll(below);
          if (k > 0) {
            kidsD[c] = s_lr[r];
            if (pool_ok) p.kid_pool[koff + c] = s_lr[r];
          }
        } else {
          const int w = 1 + r - __popcll(below);
          e_nm[w] = s_nm[r];
          e_lr[w] = s_lr[r];
          e_lp[w] = s_lp[r];
          e_nd[w] = s_nd[r];
          e_sz[w] = s_sz[r];
          e_sl[w] = s_sl[r];
          e_gi[w] = s_gi[r];
          e_ko[w] = s_ko[r];
          e_kc[w] = s_kc[r];
          if (kidsE) kidsE[w + e_shift] = s_lr[r];
        }
      }
      const double mult = s_mult - tab(p.lfact, R, p.n_tab, true) + tab(p.lfact, R - k + 1, p.n_tab, true) +
                          tab(p.lfact, k, p.n_tab, true);
      const double opr = (op != 0.0) ? s_opr + opn : s_opr;
      if (score) {
        // priors of the new tree from its roots in order: the new node first, then the roots it left alone
        const double base = (double)(pn + 1) * p.log_alpha + s_crp;
        double lp = base - (double)pn * tab(p.logn, pn + 2, p.n_tab, false) - mult;
        double ways = (double)(sz - 1) * tab(p.logn, sz, p.n_tab, false);
        for (int r = 0; r < R; ++r)
          if (!(take >> r & 1ull)) ways += (double)(s_sz[r] - 1) * tab(p.logn, s_sz[r], p.n_tab, false);
        double lp1 = base + (-ways + p.rterm[Rn]) - mult;
        const double extra = opr + s_om;
        lp_new = lp + extra;
        lp1_new = lp1 + extra;
      }
      if (lane == 0) {
        if (k > 0 && !fused) put_job(p.jobsD, jd, jd * Rmax, k, sid, lr_new, JOB_SCAN, -1, -1);
        e_nm[0] = pn;
        e_lr[0] = lr_new;
        e_lp[0] = sid;
        e_nd[0] = 1;
        e_sz[0] = sz;
        e_sl[0] = pe;
        e_gi[0] = ps;
        e_ko[0] = koff;
        e_kc[0] = k;
        eI[0] = Rn;
        eI[1] = pn + 1;
        eI[2] = ps + 1;
        eI[3] = pe + 1;
        eI[4] = po;
        eD[D_MULT] = mult;
        eD[D_CRP] = s_crp;
        eD[D_OMARG] = s_om;
        eD[D_OPRIOR] = opr;
        p.ext_key[i] = key;
        if (fused) {
          put_job(p.jobsE, slot_e, slot_e * Rmax, k, sid, lr_new, JOB_SCAN | JOB_THEN_SCORE, slot_e, -1, Rn - 1);
        } else if (score) {
          kidsE[0] = lr_new;
          put_job(p.jobsE, slot_e, slot_e * Rmax, Rn, -1, -1, JOB_SCAN | JOB_ADD_PRIOR, slot_e, -1);
        }
      }
    }
    name_out = pn;
    mask_out = take;
  }
  if (p.perm) {
    // root-permutation density of the new forest minus the parent's (both from the thawed state: the density does not
    // depend on the order of the roots)
    int* e_dl = p.ext_dl + (size_t)i * Rmax;
    double* e_lc = p.ext_lc + (size_t)i * Rmax;
    const int Rn = eI_R_after(kind, R, k_new);
    if (lane == 0) {
      if (kind == K_NEW) {
        int dsum = 0, w = 1;
        double csum = 0.0, multi = 0.0;
        for (int r = 0; r < R; ++r) {
          if (mask_out >> r & 1ull) {
            dsum += v.pDl[r];
            csum += v.pLc[r];
            multi -= tab(p.lfact, v.pDl[r], p.n_tab, true);
          } else {
            e_dl[w] = v.pDl[r];
            e_lc[w] = v.pLc[r];
            ++w;
          }
        }
        if (k_new > 0) multi += tab(p.lfact, dsum, p.n_tab, true);
        e_dl[0] = 1 + dsum;
        e_lc[0] = csum + multi;  // + log 1! for the node's own single data point
      } else {
        for (int r = 0; r < R; ++r) {
          e_dl[r] = v.pDl[r];
          e_lc[r] = v.pLc[r];
        }
        if (kind == K_ATTACH) {
          int j = 0;
          while (j < R && s_nm[j] != name_out) ++j;
          const int own = s_nd[j];
          e_dl[j] = v.pDl[j] + 1;
          e_lc[j] = v.pLc[j] + (tab(p.lfact, own + 1, p.n_tab, true) - tab(p.lfact, own, p.n_tab, true));
        }
      }
      const int out_new = po + (kind == K_OUTLIER ? 1 : 0);
      const double lpdf_new = perm_log_pdf(p, e_lc, e_dl, Rn, t + 1, out_new);
      const double lpdf_par = (t == 0) ? 0.0 : perm_log_pdf(p, v.pLc, v.pDl, R, t, po);
      p.dec_lpdf[i] = lpdf_new - lpdf_par;
    }
  }
  if (lane == 0) {
    p.slotE[i] = slot_e;
    p.dec_kind[i] = kind | (v.ret ? FLAG_RETAINED_PARENT : 0);
    p.dec_name[i] = name_out;
    p.dec_mask[i] = mask_out;
    p.dec_logq[i] = logq;
    p.dec_lp[i] = lp_new;
    p.dec_lp1[i] = lp1_new;
  }
}

// ------------------------------------------------------------------------------------------------------------
// weigh: incremental weights (kernels/base.py:38-57, samplers/base.py:52-58), swarm statistics (swarm.py:21-61),
// adaptive multinomial resampling as inverse-CDF ancestors (conditional.py:91-109, standard.py:20-32), the final
// pick (particle_gibbs.py:42-48) and the permutation of the particle states for the next step.  One CTA.
// ------------------------------------------------------------------------------------------------------------
__device__ inline double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = group_sum(v, 32);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  for (int i = 0; i < nw; ++i) r += sh[i];
  return r;
}
__device__ inline double block_max(double v, double* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = group_max(v, 32);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = -INFINITY;
  for (int i = 0; i < nw; ++i) r = fmax(r, sh[i]);
  return r;
}

__global__ void weigh_kernel(const Params p, const int t, const double log_uniform, const double lw_uniform) {
  __shared__ double sh_red[32];
  __shared__ double sh_w[1024];
  __shared__ int sh_cnt[1024];
  __shared__ int sh_anc[1024];
  pdl_trigger();
  pdl_wait();
  const int first = p.conditional ? 1 : 0;
  const int N = p.N, i = threadIdx.x;
  const bool last = t >= p.T - 1;
  double raw = -INFINITY;
  if (i < N) {
    if (i >= first) {
      double lp = p.dec_lp[i], lp1 = p.dec_lp1[i];
      const int se = p.slotE[i];
      if (se >= 0) {
        lp += row_sum(p.scE_lse, se, p.S);
        lp1 += row_sum(p.scE_last, se, p.S);
      }
      const double plp = p.cur_dbl[(size_t)i * DREC + D_LOGP];  // 0 for the empty start (parent None)
      double w = (lp - plp) - p.dec_logq[i];
      if (p.perm) w = ((lp - plp) + p.dec_lpdf[i]) - p.dec_logq[i];
      if (last) w = w - lp + lp1;
      p.ext_dbl[(size_t)i * DREC + D_LOGP] = lp;
      p.ext_dbl[(size_t)i * DREC + D_LOGP1] = lp1;
      raw = (p.conditional && t == 0) ? log_uniform : p.lw[i] + w;
      p.o_logp[(size_t)t * N + i] = lp;
      p.o_logq[(size_t)t * N + i] = p.dec_logq[i];
      p.o_kind[(size_t)t * N + i] = p.dec_kind[i];
      p.o_name[(size_t)t * N + i] = p.dec_name[i];
      p.o_mask[(size_t)t * N + i] = p.dec_mask[i];
    } else {  // the retained particle of this step
      const double* rd = p.ret_th_dbl + (size_t)t * DREC;
      const double plp = (t == 0) ? 0.0 : p.ret_th_dbl[(size_t)(t - 1) * DREC + D_LOGP];
      double w = (rd[D_LOGP] - plp) - p.dec_logq[0];
      if (p.perm) {
        const double ppdf = (t == 0) ? 0.0 : p.ret_th_dbl[(size_t)(t - 1) * DREC + D_LOGW];
        w = ((rd[D_LOGP] - plp) + (rd[D_LOGW] - ppdf)) - p.dec_logq[0];  // D_LOGW holds the retained tree's log_pdf
      }
      if (last) w = w - rd[D_LOGP] + rd[D_LOGP1];
      raw = (t == 0) ? log_uniform : p.lw[0] + w;
      p.o_logp[(size_t)t * N] = rd[D_LOGP];
      p.o_logq[(size_t)t * N] = p.dec_logq[0];
      p.o_kind[(size_t)t * N] = -1;
      p.o_name[(size_t)t * N] = -1;
      p.o_mask[(size_t)t * N] = 0;
    }
    p.o_raw[(size_t)t * N + i] = raw;
  }
  // retained record of this step into ext slot 0
  if (p.conditional) {
    const int* ri = p.ret_th_int + (size_t)t * p.irec;
    for (int k = threadIdx.x; k < p.irec; k += blockDim.x) p.ext_int[k] = ri[k];
    if (threadIdx.x < DREC) p.ext_dbl[threadIdx.x] = p.ret_th_dbl[(size_t)t * DREC + threadIdx.x];
    if (threadIdx.x == 0) p.ext_key[0] = p.ret_key[t];
    if (p.perm)
      for (int k = threadIdx.x; k < p.Rmax; k += blockDim.x) {
        p.ext_dl[k] = p.ret_th_dl[(size_t)t * p.Rmax + k];
        p.ext_lc[k] = p.ret_th_lc[(size_t)t * p.Rmax + k];
      }
  }
  // ---- log_Z, log_W, W, ESS
  const double m = block_max(raw, sh_red);
  const double e = (i < N) ? exp(raw - m) : 0.0;
  const double tot = block_sum(e, sh_red);
  const double log_z = log(tot) + m;
  const double log_W = raw - log_z;
  double W = (i < N) ? exp(log_W) : 0.0;
  const double wsum = block_sum(W, sh_red);
  W /= wsum;
  const double sq = block_sum(W * W, sh_red);
  const double ess = 1.0 / sq;
  const bool resample = !last && (ess / (double)N) <= p.threshold;
  if (!(m > -INFINITY) || !(wsum > 0.0)) {
    if (threadIdx.x == 0) atomicOr(p.status, ST_NUMERIC);
  }
  sh_w[i] = W;
  sh_cnt[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    p.o_ess[t] = ess;
    p.o_resampled[t] = resample ? 1 : 0;
    // the job counters of the next step's build / decide kernels
    *p.nB = 0;
    *p.nAdd = 0;
    *p.nD = 0;
    *p.nE = 0;
  }
  if (last) {
    // final pick: p = W / sum(W) once more (utils/math.py:19-21 via particle_gibbs.py:42-48)
    const double w2sum = block_sum((i < N) ? W : 0.0, sh_red);
    if (threadIdx.x == 0) {
      const double u = u01(p.seed, 0, FINAL_SLOT, 0);
      double acc = 0.0;
      int pick = N - 1;
      for (int k = 0; k < N - 1; ++k) {
        acc += sh_w[k] / w2sum;
        if (acc > u) {
          pick = k;
          break;
        }
      }
      *p.o_final = pick;
    }
  }
  if (resample) {
    // cdf left to right in shared memory (thread 0), then one binary search per draw
    __syncthreads();
    if (threadIdx.x == 0) {
      double acc = 0.0;
      for (int k = 0; k < N; ++k) {
        acc += sh_w[k];
        sh_w[k] = acc;
      }
    }
    __syncthreads();
    const int n_draw = N - first;
    if (i < n_draw) {
      const double u = u01(p.seed, t, RESAMPLE_SLOT, i);
      int lo = 0, hi = N - 1;  // smallest k with cdf[k] > u; the last index catches the remainder
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sh_w[mid] > u)
          hi = mid;
        else
          lo = mid + 1;
      }
      atomicAdd(&sh_cnt[lo], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // exclusive prefix of the multiplicities: particle k fills slots [start, start + count)
      int run = first;
      for (int k = 0; k < N; ++k) {
        const int c = sh_cnt[k];
        sh_cnt[k] = run;
        run += c;
      }
      if (first) sh_anc[0] = 0;
    }
    __syncthreads();
    if (i < N) {
      const int beg = sh_cnt[i], end = (i + 1 < N) ? sh_cnt[i + 1] : N;
      for (int s = beg; s < end; ++s) sh_anc[s] = i;
    }
  } else if (i < N) {
    sh_anc[i] = i;
  }
  __syncthreads();
  if (i < N) {
    p.o_anc[(size_t)t * N + i] = sh_anc[i];
    p.lw[i] = resample ? lw_uniform : log_W;
  }
  // ---- next step's parents: only the structure hashes move here (the build kernel's search for equal parents reads
  // all of them); every particle's record is fetched by its own warp of the next build kernel
  if (!last && i < N) p.cur_key[i] = p.ext_key[sh_anc[i]];
}

__global__ void init_kernel(const Params p, const double lw0) {
  for (int i = threadIdx.x; i < p.N; i += blockDim.x) {
    int* r = p.cur_int + (size_t)i * p.irec;
    r[0] = 0;
    r[1] = 0;
    r[2] = 1;
    r[3] = 0;
    r[4] = 0;
    for (int k = 0; k < DREC; ++k) p.cur_dbl[(size_t)i * DREC + k] = 0.0;
    p.cur_key[i] = 0;
    p.lw[i] = lw0;
  }
  if (threadIdx.x == 0) {
    *p.nB = 0;
    *p.nAdd = 0;
    *p.nD = 0;
    *p.nE = 0;
  }
}

}  // namespace sweep
}  // namespace pcb
