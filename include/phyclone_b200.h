/*
 * phyclone_b200 -- C ABI of the B200 (sm_100a) implementation of PhyClone's
 * tree-node likelihood hot path.
 *
 * The reference (Roth-Lab/PhyClone 0.5.1, pure Python + Numba) has no FFI layer:
 * its seams for this path are ordinary Python callables.
 * Every entry point below replaces one of those callables; the citation after
 * each declaration is the reference interface it stands in for, relative to
 * /root/reference/phyclone/.  INTEGRATION.md shows the ctypes stub a maintainer
 * would add on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; all floating point is IEEE binary64;
 *   - "tile" = one [S, G] row-major block of doubles (S samples x G grid points);
 *   - functions named *_host take HOST pointers and do the H2D / D2H copies
 *     themselves on the library's stream (synchronous on return);
 *   - functions named *_dev take DEVICE pointers plus a cudaStream_t passed as
 *     void* and only enqueue work (asynchronous);
 *   - every function returns 0 on success and a non-zero pcb_status otherwise;
 *     pcb_last_error() gives the message for the calling thread;
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with PCB_ERR_CUDA.
 */
#ifndef PHYCLONE_B200_H
#define PHYCLONE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  PCB_OK = 0,
  PCB_ERR_ARG = 1,    /* bad shape / null pointer / out-of-range index */
  PCB_ERR_CUDA = 2,   /* CUDA runtime error (message has the cudaError string) */
  PCB_ERR_NOMEM = 3   /* device or host allocation failed */
} pcb_status;

/* ---- library --------------------------------------------------------- */
int pcb_version(void);
const char* pcb_last_error(void);
int pcb_device_count(void);
/* bind the calling thread's work to `device` (cudaSetDevice) */
int pcb_set_device(int device);
/* number of kernels this library has launched in the calling process */
int64_t pcb_launch_count(void);

/* ---- padded arena layout --------------------------------------------- */
/* The kernels keep tiles in HBM in a padded row layout chosen per (S, G):
 * row pitch `pitch` doubles, `lead` zero doubles in front of every row and
 * zeros behind it, so a tile (or a block of its rows) is one contiguous,
 * 16-byte aligned TMA bulk copy into shared memory and the sliding-window
 * convolution never needs bounds checks. */
typedef struct {
  int32_t S, G;
  int32_t strip;         /* L: outputs per register strip            */
  int32_t lanes_per_row; /* P: lanes cooperating on one row          */
  int32_t n_strips;      /* Q = ceil(G / L)                          */
  int32_t pitch;         /* doubles per padded row                   */
  int32_t lead;          /* leading zero doubles per row (== L)      */
  int32_t rows_per_cta;  /* rows of one tile handled by one CTA      */
  int32_t kernel;        /* bits 0-1: 0 lane-per-unit job kernel, 1 unit-per-warp ("wide", 32 rows of 32/S jobs per CTA);
                            PCB_KERNEL_LATENCY (4), set by the caller: small launches (<= 1024 jobs, sweeps of <= 256
                            particles) run on the warp-per-row mapping -- same results to rounding, shorter fold chains */
  int32_t reserved;
  int64_t tile_elems;    /* S * pitch                                */
} pcb_layout;
#define PCB_KERNEL_LATENCY 4

/* variant = 100*L + P fixes the strip/lane configuration (10000 + 100*L: the unit-per-warp kernel, S an even
 * divisor of 32); -1 picks the throughput-optimal one for (S, G) (saturated
 * batches), -2 the latency-optimal one (narrow strips: a single chain launches a handful of jobs at a time). */
int pcb_make_layout(int32_t S, int32_t G, int32_t variant, pcb_layout* out);

/* An arena is three caller-owned device arrays:
 *   lg   [capacity][S][pitch]  log-domain tiles (padded)
 *   lin  [capacity][S][pitch]  exp(lg - rowmax) (padded, pads MUST be zero)
 *   rmax [capacity][S]         per-row maximum of lg
 * The caller zero-fills lg/lin once (cudaMemset); kernels never write pads. */
typedef struct {
  pcb_layout layout;
  double* lg;
  double* lin;
  double* rmax;
  int64_t capacity;
} pcb_arena;

/* dense [n][S][G] log tiles (device) -> arena slots ids[n]; computes lin/rmax */
int pcb_arena_import_dev(const pcb_arena* arena, const double* dense, const int32_t* ids_dev, int64_t n,
                         void* stream);
/* arena slots ids[n] -> dense [n][S][G] log tiles (device) */
int pcb_arena_export_dev(const pcb_arena* arena, const int32_t* ids_dev, int64_t n, double* dense, void* stream);

/* out[i] = a[i] + b_sign * b[i] (if b_ids && b_ids[i] >= 0) + addend, with lin/rmax refreshed;
 * b_sign is +1.0 or -1.0.  Replaces TreeNode.add_data_point / remove_data_point / the log_p
 * initialisation (tree/tree_node.py:12, 42-59): log_p += value; log_r += value; log_p -= value. */
int pcb_tile_add_dev(const pcb_arena* arena, const int32_t* a_ids, const int32_t* b_ids, const int32_t* out_ids,
                     int64_t n, double b_sign, double addend, void* stream);

/* ---- the node-refresh job kernel -------------------------------------- */
/* One job = one TreeNode.update_node_from_child_r_vals (tree/tree_node.py:61-71)
 * i.e. log_r = log_p + compute_log_S(children) (tree/utils.py:7-47), children
 * folded left-to-right in the order given (conv(c0,c1), conv(c2,acc), ...),
 * each pair step with _np_conv_dims semantics (tree/utils.py:60-85: per-row
 * max shift, truncated linear convolution, `<= 0 -> 1e-100` floor), followed by
 * the prefix log-sum-exp of _sub_compute_S (tree/utils.py:25-30).
 * Optionally the job also emits the two per-tree data terms of
 * TreeJointDistribution.compute_both_log_p_and_log_p_one
 * (tree/distributions.py:197-200): sum_s LSE_g out[s,:] and sum_s out[s,G-1].
 *
 * job record: 8 x int32
 *   [0] first index into children[]        [1] number of children (>= 1)
 *   [2] base tile id (log_p) or -1         [3] output tile id or -1 (scores only)
 *   [4] flags (PCB_JOB_*)                  [5] score slot or -1
 *   [6] second base tile id + 1, or 0: log_p = tile(base) + tile(base2) formed on the fly (a node refresh right after
 *       `log_p += value`, tree/tree_node.py:42-52 then :61-71)          [7] number of second-stage children (PCB_JOB_THEN_SCORE), else 0
 */
#define PCB_JOB_SCAN 1      /* apply the prefix log-sum-exp (log_S); without it the job yields log_D */
#define PCB_JOB_ADD_PRIOR 2 /* add the scalar `prior` (the dummy root's log_p = -log G) */
#define PCB_JOB_THEN_SCORE 4 /* device-resident sweep only: after the node tile is written, fold record[7] more children
                              * onto it as a dummy root (prior added) and emit that tree's scores into `score slot` */
#define PCB_JOB_INTS 8

int pcb_node_jobs_dev(const pcb_arena* arena, const int32_t* jobs_dev, const int32_t* children_dev, int64_t n_jobs,
                      double prior, double* score_rows_lse_dev, /* [n_slots][S] scratch, may be NULL */
                      double* score_rows_last_dev,              /* [n_slots][S] scratch, may be NULL */
                      void* stream);

/* sum the per-row partials over s = 0..S-1 in order: out[slot] = sum_s rows[slot][s] */
int pcb_reduce_rows_dev(const double* rows_dev, int64_t n_slots, int32_t S, double* out_dev, void* stream);

/* ---- reference-facing host entry points (HOST buffers) ---------------- */
/* _np_conv_dims(child_1, child_2), batched over B pairs.   tree/utils.py:60-85 */
int pcb_np_conv_dims_host(const double* c1, const double* c2, double* out, int64_t B, int32_t S, int32_t G);
/* _sub_compute_S(log_D), batched.                           tree/utils.py:25-30 */
int pcb_sub_compute_S_host(const double* log_D, double* log_S, int64_t B, int32_t S, int32_t G);
/* compute_log_S.__wrapped__(children) for B ragged child lists; children are
 * concatenated in fold order, child_off[B+1] delimits them.  tree/utils.py:7-22 */
int pcb_compute_log_S_host(const double* children, const int64_t* child_off, double* log_S, int64_t B, int32_t S,
                           int32_t G);
/* TreeNode.update_node_from_child_r_vals: log_r = log_p (+ log_S(children)).
 * child_off[b+1]==child_off[b] means a leaf.               tree/tree_node.py:61-71 */
int pcb_node_update_host(const double* log_p, const double* children, const int64_t* child_off, double* log_r,
                         int64_t B, int32_t S, int32_t G);
/* data terms of compute_both_log_p_and_log_p_one for B root tiles.
 *                                                           tree/distributions.py:197-200 */
int pcb_root_terms_host(const double* root_tiles, double* lse_sum, double* last_sum, int64_t B, int32_t S,
                        int32_t G);

/* _compute_liklihood_grid over M mutations x S samples.    data/pyclone.py:323-436
 * ref_counts/alt_counts = SampleDataPoint.a/.b, cn/mu [M][S][Cmax][3],
 * log_pi [M][S][Cmax], n_geno [M][S]; density 0 = binomial, 1 = beta-binomial. */
int pcb_likelihood_grid_host(const int64_t* ref_counts, const int64_t* alt_counts, const double* tumour_content,
                             const int64_t* cn, const double* mu, const double* log_pi, const int32_t* n_geno,
                             int64_t M, int32_t S, int32_t Cmax, int32_t G, int32_t density, double precision,
                             double* out /* [M][S][G] */);
int pcb_likelihood_grid_dev(const int64_t* ref_counts, const int64_t* alt_counts, const double* tumour_content,
                            const int64_t* cn, const double* mu, const double* log_pi, const int32_t* n_geno,
                            int64_t M, int32_t S, int32_t Cmax, int32_t G, int32_t density, double precision,
                            double* out, void* stream);
/* value[k] = sum of member grids, members cluster_off[k]..cluster_off[k+1]-1.
 *                                                           data/pyclone.py:89-94 */
int pcb_cluster_sum_host(const double* grids, const int64_t* cluster_off, int64_t K, int32_t S, int32_t G,
                         double* out);
/* DataPoint.outlier_marginal_prob for K tiles.              data/base.py:31-37 */
int pcb_outlier_marginal_host(const double* values, int64_t K, int32_t S, int32_t G, double* out);

/* One flush of the lazily recorded tile algebra (phyclone_b200/tiles.py) in a single call: copies the packed id lists /
 * job tables from pinned host memory, runs the plan's launches in dependency order on `stream` -- tile adds
 * (tree/tree_node.py:42-59) and node jobs (tree/tree_node.py:61-71, tree/distributions.py:197-200) -- reduces the score
 * rows over samples straight into `sums_host` ([2][n_slots]; must be pinned, device-accessible host memory; `sums_dev` is
 * unused); synchronises the stream only if n_slots > 0.
 * plan: n_rec x 6 int32 {kind 0=add | 1=jobs, offset into packed, n, n_kids, negate, 0}; add ids are laid out
 * [a | b | out], a job table (n x 8 int32) is followed by its n_kids child ids.  events: NULL or 2 CUDA events per
 * record, recorded around the node-jobs launches. */
int pcb_run_plan_dev(const pcb_arena* arena, const int32_t* packed_host, int64_t n_ints, int32_t* packed_dev,
                     const int32_t* plan, int32_t n_rec, double prior, double* rows_lse, double* rows_last,
                     int64_t n_slots, double* sums_dev, double* sums_host, void* const* events, void* stream);

/* Host-side split of all synchronising pcb_run_plan_dev calls so far: seconds spent enqueueing (copies + launches) and
 * seconds spent waiting for the stream. */
int pcb_plan_stats(double* enqueue_s, double* sync_s, int64_t* flushes);

/* Kernel timing for roofline figures: between begin and end every pcb_node_jobs_dev launch (also inside
 * pcb_run_plan_dev) is bracketed by a CUDA event pair on its stream; end synchronises the device and returns the summed
 * durations and the number of launches recorded (at most max_launches). */
int pcb_timing_begin(int32_t max_launches);
int pcb_timing_end(double* total_ms, int64_t* n_launches);
/* stride >= 1: from now on only every stride-th launch is bracketed (an event pair drains the stream and keeps the next
 * kernel from starting in its predecessor's tail, so bracketing every launch slows what it measures); n_seen (may be
 * NULL) receives the number of node-jobs launches since the last begin, bracketed or not.  Call it before
 * pcb_timing_end to read the count. */
int pcb_timing_sample(int32_t stride, int64_t* n_seen);

/* MAP cellular prevalences of B trees: the max-plus twin of the node fold with back-pointers.
 *   process_trace/map.py:48-64 compute_max_likelihood, :67-93 compute_log_S (prefix max, ties to the earliest
 *   index), :96-134 compute_log_D / _compute_log_D_n (max_j child[j] + D[i-j], ties to the largest j; D starts at 0),
 *   :137-166 set_max_assignment (walk down from index G-1 of the tree's root).
 * Nodes are numbered 0..n_nodes-1 over all trees; post_order[tree_off[t]..tree_off[t+1]) lists tree t's nodes
 * children-first with its root last; child_idx[child_off[v]..child_off[v+1]) are v's children in fold order
 * (ascending edge slot = the networkx successor order the reference folds in).
 * Out: max_idx[n_nodes][S] (ccf = max_idx/(G-1), map.py:169-173) and, if not NULL, log_R_max[n_nodes][S][G].
 * The _dev form takes device pointers and caller-provided scratch (d_choice [E][S][G], s_choice [N][S][G]). */
int pcb_map_trees_host(const double* node_log_p, const int32_t* tree_off, const int32_t* post_order,
                       const int32_t* child_off, const int32_t* child_idx, int64_t n_trees, int64_t n_nodes,
                       int64_t n_edges, int32_t S, int32_t G, int32_t* max_idx, double* log_r_max);
int pcb_map_trees_dev(const double* node_log_p, const int32_t* tree_off, const int32_t* post_order,
                      const int32_t* child_off, const int32_t* child_idx, int64_t n_trees, int64_t n_nodes,
                      int64_t n_edges, int32_t S, int32_t G, double* log_r_max, int32_t* d_choice, int32_t* s_choice,
                      int32_t* max_idx, void* stream);

/* conv_log for B pairs of length-n vectors: per-output max-shifted direct log-space convolution, no floor.
 *                                                           utils/math.py:212-238 (used by the reference's tests) */
int pcb_conv_log_host(const double* log_x, const double* log_y, double* out, int64_t B, int32_t n);

/* ParticleSwarm.log_norm_const / log_weights / weights / ess.  smc/swarm/swarm.py:21-61 */
int pcb_swarm_weights_host(const double* log_w, int64_t N, double* log_Z, double* log_W, double* W, double* ess);
/* log_normalize + exp + renormalise per candidate segment.
 *                 smc/kernels/semi_adapted.py:151-163, utils/math.py:37-59,121 */
int pcb_normalize_segments_host(const double* log_p, const int64_t* seg_off, int64_t U, double* log_q, double* q,
                                double* log_norm /* [U] */);
/* device flavour; additionally writes ess[u] = 1 / sum_i q_i^2 when ess != NULL (a swarm is one segment) */
int pcb_normalize_segments_dev(const double* log_p, const int64_t* seg_off_dev, int64_t U, double* log_q, double* q,
                               double* log_norm, double* ess, void* stream);
/* optional device resampling mode: ancestors[i] = inverse
 * CDF of W at uniforms[i] (NumPy restatement: oracle.numerics.ancestors_from_uniforms) */
int pcb_resample_host(const double* W, int64_t N, const double* uniforms, int64_t n_draws, int64_t* ancestors);

/* ---- device-resident SMC sweep ------------------------------------------ */
/* One whole (conditional) SMC sweep over T data points with N particles on the device, no host round trip between
 * steps: semi-adapted proposal normalisation (smc/kernels/semi_adapted.py:36-194), the particles' draws, particle
 * construction and incremental weights (smc/kernels/base.py:38-67, smc/samplers/base.py:32-58), swarm weights / ESS
 * (smc/swarm/swarm.py:21-61), adaptive multinomial resampling (smc/samplers/conditional.py:76-125,
 * standard.py:20-45) and the final pick (mcmc/particle_gibbs.py:42-48).
 * Draws are inverse-CDF transforms of uniforms u(seed, step, slot, draw number) (splitmix64 counter hash); the CPU
 * definition of this mode is oracle/uniform.py.  A particle is the root level of its forest; record layouts:
 *   int record  [5 + 9*Rmax, padded to a multiple of 4]: R, n_nodes, n_slots, n_edges, n_outliers, then Rmax-strided
 *                fields name, log_r tile, log_p tile, n_data, subtree size, edge slot, graph index, kid offset, kid count
 *   dbl record  [8]: crp sum, multiplicity, outlier prior, outlier marginal, log_p, log_p_one, log_w increment, 0
 * Conditional sweeps take the retained path as T such records per view: `ret_th_*` = the retained particle after
 * step t re-read through Tree.from_dict (what extends it), `ret_ba_*` = the same tree as built (what its proposal's
 * candidates are scored on), `ret_key` = its structure hash, `ret_kids` = the child tile lists its kid offsets index,
 * `ret_kind` / `ret_arg` = the edit that made it (0 attach: node name, 1 new root: number of adopted roots, 2 outlier);
 * of the dbl record only log_p / log_p_one are read for it (its proposal log-probability is computed on the device).
 * Tiles: [tile_base, tile_base + N*Rmax) is per-step candidate scratch, the rest up to tile_limit is handed out to new
 * root / node tiles.  Small control tables are HOST pointers (copied inside); outputs are HOST buffers filled after the
 * one stream synchronisation at the end.  status: 0 ok, else bit 1 = more than Rmax roots, 2 = tile range exhausted,
 * 4 = kid pool exhausted, 8 = non-finite weights (the sweep's outputs are then void). */
typedef struct {
  int32_t N, T, Rmax, conditional, outliers, n_tab, n_ret_kids;
  int32_t perm_dist; /* 1: particle weights carry the root-permutation density (smc/kernels/base.py:49-55,
                      * smc/utils.py:18-66); the dbl record's 7th entry is then the retained tree's log_pdf */
  double log_alpha, threshold, log_half, log_uniform, lw_uniform, prior;
  uint64_t seed;
  int64_t tile_base, tile_limit;
  const int32_t *dp_idx, *value_id, *single_id; /* [T] */
  const double *dp_op, *dp_opn, *dp_marg;        /* [T] outlier log prob, not-outlier log prob, outlier marginal */
  const double *lfact, *logn, *rterm;            /* [n_tab] log n!, log n, FS-CRP root-count term */
  const int32_t *ret_th_int, *ret_ba_int;        /* [T][(5 + 9*Rmax + 3) & ~3] */
  const double *ret_th_dbl, *ret_ba_dbl;         /* [T][8] */
  const uint64_t* ret_key;                       /* [T] */
  const int32_t* ret_kids;                       /* [n_ret_kids] */
  const int32_t *ret_kind, *ret_arg;             /* [T] */
  const int32_t* ret_subtree_data;               /* [T][Rmax] perm_dist: data points under each root (thawed order) */
  const double* ret_subtree_log_count;           /* [T][Rmax] perm_dist: log #compatible orders of each root's subtree */
  int32_t *anc, *kind, *name;                    /* [T][N] ancestor of slot i after step t; decision of particle i */
  uint64_t* mask;                                /* [T][N] adopted roots of a new node (positions in the parent) */
  double *raw, *logp, *logq;                     /* [T][N] unnormalised log weight, log_p, proposal log prob */
  double* ess;                                   /* [T] */
  int32_t* resampled;                            /* [T] */
  int32_t *final_idx, *status;
  int64_t *tiles_used, *n_launches, *conv_pairs; /* may be NULL; conv_pairs = pair convolutions the sweep's jobs folded */
  double* device_ms;                             /* may be NULL: CUDA-event time of the whole sweep */
} pcb_sweep_desc;
int pcb_smc_sweep_dev(const pcb_arena* arena, const pcb_sweep_desc* desc, void* stream);

/* FP64 FMA-chain microbenchmark used as the roofline denominator.
 * Returns achieved FLOP/s (2 flop per FMA) over `iters` dependent-chain rounds. */
int pcb_fp64_peak(int32_t iters, double* flops_per_s, double* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* PHYCLONE_B200_H */
