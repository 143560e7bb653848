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
  int32_t kernel;        /* 0: lane-per-unit job kernel, 1: unit-per-warp ("wide", 32 rows of 32/S jobs per CTA) */
  int32_t reserved;
  int64_t tile_elems;    /* S * pitch                                */
} pcb_layout;

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
 *   [6],[7] reserved (0)
 */
#define PCB_JOB_SCAN 1      /* apply the prefix log-sum-exp (log_S); without it the job yields log_D */
#define PCB_JOB_ADD_PRIOR 2 /* add the scalar `prior` (the dummy root's log_p = -log G) */
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

/* FP64 FMA-chain microbenchmark used as the roofline denominator.
 * Returns achieved FLOP/s (2 flop per FMA) over `iters` dependent-chain rounds. */
int pcb_fp64_peak(int32_t iters, double* flops_per_s, double* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* PHYCLONE_B200_H */
