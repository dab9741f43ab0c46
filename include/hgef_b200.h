/*
 * hgef_b200.h -- C-ABI of libhgef_b200.so: the B200-native (sm_100a) drop-in for
 * HyperGef's fused hypergraph message-passing path.
 *
 * Plain pointers and sizes only (no torch types).  Every entry point names the
 * reference interface it replaces (paths relative to fishmingyu/HyperGef).
 * The reference binds this path through two pybind11 torch extensions
 * (`hgnnaggr`, `unignnaggr`, setup.py:18); INTEGRATION.md shows the stub a
 * maintainer would put in their place.
 *
 * Conventions
 *   - all functions return 0 (HG_OK) or an HG_E* code; hg_last_error() gives the
 *     thread-local message of the last failure on the calling thread.  The
 *     reference aborts through C assert (hgnnaggr_cuda.cu:8-12) and never checks
 *     a launch; here nothing aborts.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream,
 *     which is what the reference always uses, hgnnaggr_cuda.cu:383).
 *   - device pointers are prefixed d_, host pointers h_.  All index arrays are
 *     int32 (the reference's `typedef int Index`, include/util/check.cuh:11);
 *     addresses are formed in 64 bits (the reference overflows at N*F >= 2^31).
 *   - feature matrices are fp32, row-major, contiguous [rows, F].
 *   - entry points are re-entrant; an hgPlan is immutable after creation except
 *     for its scratch buffer, so one plan must not run on two streams at once.
 */
#ifndef HGEF_B200_H_
#define HGEF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HG_API __attribute__((visibility("default")))

enum {
  HG_OK = 0,
  HG_EINVAL = 1,   /* bad argument (Python shim raises ValueError)            */
  HG_ECUDA = 2,    /* CUDA runtime / launch failure (RuntimeError)            */
  HG_ENOMEM = 3,   /* allocation failure (MemoryError)                        */
  HG_EEMPTY = 4,   /* balancer on a matrix with no non-zero: the reference
                      raises IndexError at balancer.py:32 (balan_key[-1])     */
  HG_EGRAPH = 5    /* plan validation: index out of range in the graph/balancer
                      arrays (the reference would read/write out of bounds)   */
};

HG_API const char *hg_last_error(void);
HG_API int hg_abi_version(void);
/* Compute capability of `device` as major*10+minor, or <0 on error. */
HG_API int hg_device_cc(int device);
/* Process-wide override of one launch-geometry default of the aggregation kernels (ring depth, item size,
 * lags, eviction hints ...: the names are listed in DESIGN.md section 3); clear != 0 restores the default.
 * For measurement and tests: the reference fixes its geometry at compile time (hgnnaggr_cuda.cu:376-404). */
HG_API int hg_tune_set(const char *name, int32_t value, int32_t clear);

/* ------------------------------------------------------------------------- *
 * Balancer.  Replaces balance_schedule.balancer, HyperGsys/balancer.py:15-33
 * (C++ twin hgnn_ef_full_balance_cpu, include/taskbalancer/balancer_kernel.cuh:229-259).
 * Output is bit-identical to the reference: nkey = S+1 segment offsets (sentinel
 * included), ngroup = sum_e w_e^2 groups in (row, write-segment, read-segment) order.
 * count -> allocate -> fill.  HG_EEMPTY when no row has a non-zero.
 * ------------------------------------------------------------------------- */
HG_API int hg_balance_count_host(int64_t nrow, const int32_t *h_csrptr, int32_t ngs,
                                 int64_t *nkey, int64_t *ngroup);
HG_API int hg_balance_fill_host(int64_t nrow, const int32_t *h_csrptr, int32_t ngs,
                                int32_t *h_key, int32_t *h_row, int32_t *h_st, int32_t *h_ed);
/* Device twin (the reference left this as a TODO, source/balancer/balancer_kernel.cu:34).
 * count synchronises `stream` once to return the two sizes. */
HG_API int hg_balance_count_dev(int64_t nrow, const int32_t *d_csrptr, int32_t ngs,
                                int64_t *nkey, int64_t *ngroup, int device, void *stream);
HG_API int hg_balance_fill_dev(int64_t nrow, const int32_t *d_csrptr, int32_t ngs,
                               int64_t nkey, int64_t ngroup, int32_t *d_key, int32_t *d_row,
                               int32_t *d_st, int32_t *d_ed, int device, void *stream);

/* ------------------------------------------------------------------------- *
 * Incidence -> CSR of H and of H^T.  Replaces the scipy calls of
 * HyperGraph.__init__, HyperGsys/hypergraph.py:23-25
 *   H = coo_matrix((1,(V,E)),(N,M)).tocsr() ; H_T = H.transpose().tocsr()
 * with scipy's semantics: column indices ascending in a row, duplicate pairs
 * merged with summed value (data = multiplicity).  `nnz_out` <= nnz_in.
 * Output arrays need room for nnz_in entries; indptr arrays nrow+1 / ncol+1.
 * ------------------------------------------------------------------------- */
HG_API int hg_csr_build_host(int64_t nrow, int64_t ncol, int64_t nnz_in, const int64_t *h_rows,
                             const int64_t *h_cols, int32_t *h_indptr, int32_t *h_indices,
                             float *h_data, int32_t *h_t_indptr, int32_t *h_t_indices,
                             float *h_t_data, int64_t *nnz_out);
HG_API int hg_csr_build_dev(int64_t nrow, int64_t ncol, int64_t nnz_in, const int64_t *d_rows,
                            const int64_t *d_cols, int32_t *d_indptr, int32_t *d_indices,
                            float *d_data, int32_t *d_t_indptr, int32_t *d_t_indices,
                            float *d_t_data, int64_t *nnz_out, int device, void *stream);
/* Degree scalings, hypergraph.py:34-45: out[r] = (sum of data in row r)^power, with
 * +inf replaced by 1 when `inf_to_one` (degV: power -0.5, inf_to_one 1; degE: -1, 0). */
HG_API int hg_degree_scale_dev(int64_t nrow, const int32_t *d_indptr, const float *d_data,
                               float power, int inf_to_one, float *d_out, int device,
                               void *stream);

/* ------------------------------------------------------------------------- *
 * MatrixMarket (.mtx) incidence reader (host).  Replaces read_mtx_file
 * (include/dataloader/dataloader.hpp:22-104): coordinate format, values dropped, 0-based,
 * `symmetric` mirrored and de-duplicated, coordinates sorted row-major; rows = vertices,
 * columns = hyperedges.  open -> (sizes) -> fill -> close; the int64 pairs feed hg_csr_build_*.
 * HG_EINVAL for a missing / malformed file (the reference calls exit()), HG_EGRAPH for an
 * entry outside the declared shape (the reference stores it unchecked).
 * ------------------------------------------------------------------------- */
typedef struct hgMtx hgMtx;
HG_API int hg_mtx_open(const char *path, hgMtx **mtx, int64_t *nrow, int64_t *ncol, int64_t *nnz);
HG_API int hg_mtx_fill(const hgMtx *mtx, int64_t *h_rows, int64_t *h_cols);
HG_API int hg_mtx_close(hgMtx *mtx);

/* ------------------------------------------------------------------------- *
 * Aggregation plan: everything derived once from the balancer output (segment -> hyperedge map, heavy
 * hyperedges, the CSR of H, the row programs of the two stages) so that the per-call path is two kernel
 * launches and no allocation.  Borrowed pointers (d_key .. d_t_indices) must outlive the plan.
 * Validates every index (HG_EGRAPH).
 *   nseg   = nkey - 1 (segments), ngroup = len(group_row)
 * ------------------------------------------------------------------------- */
typedef struct hgPlan hgPlan;
HG_API int hg_plan_create(hgPlan **plan, int64_t num_nodes, int64_t num_edges, int64_t nnz,
                          int64_t nseg, int64_t ngroup, const int32_t *d_key,
                          const int32_t *d_row, const int32_t *d_st, const int32_t *d_ed,
                          const int32_t *d_t_indices, int device, void *stream);
HG_API int hg_plan_destroy(hgPlan *plan);
/* what the plan found: S, #heavy hyperedges (w_e > 1), #segments of heavy hyperedges,
 * 1 when the groups are the balancer's canonical full cross product. */
HG_API int hg_plan_info(const hgPlan *plan, int64_t *nseg, int64_t *nheavy_edges,
                        int64_t *nheavy_segs, int32_t *canonical);

/* Number of this library's kernels launched through the plan so far (aggregation calls only). */
HG_API int hg_plan_launches(const hgPlan *plan, int64_t *kernels);

/* Sizes the plan's per-call buffers (hyperedge features [num_edges, F], heavy-hyperedge scratch) for feature
 * lengths up to F_max, so that no later hg_aggr_forward allocates: call it before capturing the op into a
 * CUDA graph or using a new, wider F inside one.  A buffer that has to grow is never freed before the plan
 * is destroyed (a launch in flight or a captured graph may still name it); growing while `stream` is being
 * captured is refused with HG_EINVAL.  The reference allocates its output inside every call
 * (torch::zeros, hgnnaggr_cuda.cu:374) and keeps no state. */
HG_API int hg_plan_reserve(hgPlan *plan, int32_t F_max, void *stream);

/* Diagnostic (lab library only; zeros otherwise): control words of the last experimental-form launch. */
HG_API int hg_plan_debug(hgPlan *plan, int32_t *h_out8, void *stream);

/* Synchronises `stream` and reports (HG_ECUDA) any fault of the launches issued with this plan (and, in the
 * lab library, an experimental form's bounded-wait give-up).  The reference has no equivalent: it never
 * checks a launch (hgnnaggr_cuda.cu:383-404). */
HG_API int hg_plan_check(hgPlan *plan, void *stream);

/* flags for hg_aggr_* */
enum {
  HG_ACCUMULATE = 1,     /* do not zero-fill Y first (reference: torch::zeros, hgnnaggr_cuda.cu:374) */
  HG_FORCE_SCALAR = 4,   /* disable the 128-bit path (testing) */
  HG_TWO_PASS = 8,       /* always memset + segment kernels (the form chosen when Y fits the L2) */
  HG_FORCE_STREAM = 64,  /* always the stream form: both stages as register-only row streams, two launches (the second
                            a programmatic dependent launch), the hyperedge features make one round trip through the
                            L2 / HBM (the form chosen when Y exceeds the L2) */
  /* EXPERIMENTAL forms, built only into libhgef_b200_lab.so (make -C hypergef_b200/csrc lab; HG_EINVAL in the
   * product library).  All of them merge both stages into one launch; all are measured slower than the stream
   * form and are kept as evidence: */
  HG_FORCE_FUSED = 16,   /* single persistent scatter launch (red.v4 into Y; round 1) */
  HG_FORCE_PULL = 32,    /* gather-only two-phase form with shared-memory staging (round 1) */
  HG_FORCE_RING = 128,   /* one persistent launch, rows moved by TMA bulk copies (cp.async.bulk) into a shared-memory
                            ring, hyperedge features handed over through the L2 and discarded there */
  HG_FORCE_FSTREAM = 256 /* one persistent launch of register-only row streams with the same hand-over */
};

/* ------------------------------------------------------------------------- *
 * Fused two-stage aggregation
 *     Y = diag(a_out) . H . diag(s1*s2) . H^T . diag(a_in) . X
 * as one C-ABI call.  Small graphs (Y fits the L2): one warp per balancer segment, the hyperedge feature stays in
 * registers, vector reductions into a zero-filled Y.  Large graphs: stage A (balancer segments -> hyperedge
 * features Xe) and stage B (Xe -> Y, every row written once, no atomics) as two back-to-back launches; Xe is a
 * plan-owned [num_edges, F] buffer that passes through the L2 / HBM between them.
 * Replaces hgnnaggr_fp_cuda (source/hgnnaggr/hgnnaggr_cuda.cu:350-406) with
 * s1=degE, s2=W, a_out=degV; unignnaggrdeg_fp_cuda (source/unignnaggr/unignnaggr_cuda.cu:392-447)
 * with s2=NULL; unignnaggr_fp_cuda (:449-488) with all scales NULL.  Any scale may
 * be NULL (= 1).  a_in is the gather-side scale of the true-transpose backward
 * (dX = H diag(s) H^T diag(degV) dY); the reference's own backward re-runs the
 * forward on dY (hgnnaggr.cc:58-60), which is the same call as the forward.
 * Any F >= 1 (the reference requires F < 32 or F % 32 == 0).
 * ------------------------------------------------------------------------- */
HG_API int hg_aggr_forward(hgPlan *plan, const float *d_X, const float *d_s1, const float *d_s2,
                           const float *d_a_out, const float *d_a_in, float *d_Y, int32_t F,
                           int32_t flags, void *stream);

/* The reference's schedule, literally: one work unit per balancer GROUP (read segment
 * group_st[g], scale row group_row[g], write segment group_ed[g]) exactly as
 * HGNNAggr_forward_kernel (hgnnaggr_cuda.cu:14-47).  Needs no plan and accepts any
 * group arrays; kept for parity tests and as the "reference schedule" ablation. */
HG_API int hg_aggr_groups(int64_t num_nodes, int64_t ngroup, const int32_t *d_key,
                          const int32_t *d_row, const int32_t *d_st, const int32_t *d_ed,
                          const int32_t *d_t_indices, const float *d_X, const float *d_s1,
                          const float *d_s2, const float *d_a_out, const float *d_a_in,
                          float *d_Y, int32_t F, int32_t flags, int device, void *stream);

/* Host <-> device staging of a COLUMN SLAB of a row-major fp32 matrix, asynchronous on `stream`
 * (cudaMemcpy2DAsync; the host side should be pinned).  The aggregation never mixes feature columns
 * (hgnnaggr_cuda.cu:21,34,44), so a host caller with a wide matrix uploads, aggregates and downloads
 * it slab by slab and the PCIe copies of neighbouring slabs overlap in both directions; the reference
 * has no host-buffer entry point (its extension takes device tensors, hgnnaggr.cc:122-129).
 *   to_device != 0:  dst = device [nrow, ncol] (dense),  src = host [nrow, ld_host], columns col0..col0+ncol
 *   to_device == 0:  dst = host [nrow, ld_host] columns col0..,  src = device [nrow, ncol] (dense) */
HG_API int hg_copy_columns(void *dst, const void *src, int64_t nrow, int64_t ncol, int64_t ld_host,
                           int64_t col0, int32_t to_device, int device, void *stream);

/* ------------------------------------------------------------------------- *
 * The two stages as separate launches, for the vertex/hyperedge-PARTITIONED multi-GPU path
 * (new capability; the reference is single-GPU).  Rows of the CSR are the BOUNDARY hyperedges
 * restricted to the caller's vertex block; interior hyperedges stay on hg_aggr_forward.
 *   hg_edge_reduce   P[e,:]  = sum_{u in row e} a_in[u] * X[u,:]            (P is overwritten)
 *   hg_edge_scatter  Y[v,:] += a_out[v] * scale[e] * Q[e,:]  for v in row e   (Y is accumulated)
 * scale / a_in / a_out may be NULL (= 1).  Any F >= 1.
 * ------------------------------------------------------------------------- */
HG_API int hg_edge_reduce(int64_t nrow, const int32_t *d_indptr, const int32_t *d_indices,
                          const float *d_X, const float *d_a_in, float *d_P, int32_t F, int device,
                          void *stream);
HG_API int hg_edge_scatter(int64_t nrow, const int32_t *d_indptr, const int32_t *d_indices,
                           const float *d_Q, const float *d_scale, const float *d_a_out, float *d_Y,
                           int32_t F, int device, void *stream);

/* f1-mean / f1-max first-stage variants over the un-balanced CSR of H^T.
 * Replace hgnnaggr_mean_fp_cuda / _bp_cuda (hgnnaggr_cuda.cu:408-470) and
 * hgnnaggr_max_fp_cuda / _bp_cuda (:472-543).  The hyperedge loop is bounded by
 * num_edges (the reference bounds it by N).  d_record is int32 [M,F]. */
HG_API int hg_aggr_mean(int64_t num_nodes, int64_t num_edges, const int32_t *d_t_indptr,
                        const int32_t *d_t_indices, const float *d_X, const float *d_s1,
                        const float *d_s2, const float *d_a_out, float *d_Y, int32_t F,
                        int32_t flags, int device, void *stream);
HG_API int hg_aggr_max_forward(int64_t num_nodes, int64_t num_edges, const int32_t *d_t_indptr,
                               const int32_t *d_t_indices, const float *d_X, const float *d_s1,
                               const float *d_s2, const float *d_a_out, float *d_Y,
                               int32_t *d_record, int32_t F, int32_t flags, int device,
                               void *stream);
HG_API int hg_aggr_max_backward(int64_t num_nodes, int64_t num_edges, const int32_t *d_t_indptr,
                                const int32_t *d_t_indices, const float *d_G, const float *d_s1,
                                const float *d_s2, const float *d_a_out, const int32_t *d_record,
                                float *d_dX, int32_t F, int32_t flags, int device, void *stream);

/* The same max op over a PLAN (balanced): one warp per balancer segment, the segments of a split hyperedge meet in a
 * packed (value, vertex id) atomicMax, the second stage is the stream form's stage B (every Y row written once).
 * d_t_indptr is only used when the plan cannot serve the call (non-canonical groups, F % 4 != 0): then the
 * un-balanced kernels above run.  Backward = hgnnaggr_max_bp_cuda's formula (hgnnaggr_cuda.cu:165-208). */
HG_API int hg_plan_max_forward(hgPlan *plan, const int32_t *d_t_indptr, const float *d_X, const float *d_s1,
                               const float *d_s2, const float *d_a_out, float *d_Y, int32_t *d_record,
                               int32_t F, void *stream);
HG_API int hg_plan_max_backward(hgPlan *plan, const int32_t *d_t_indptr, const float *d_G, const float *d_s1,
                                const float *d_s2, const float *d_a_out, const int32_t *d_record,
                                float *d_dX, int32_t F, void *stream);

/* ------------------------------------------------------------------------- *
 * The two stages of hg_aggr_forward as separate calls (the balanced stream kernels of the plan), with the
 * hyperedge features in a CALLER-owned [num_edges, F] buffer -- for callers that put work between the stages.
 * The case in point is SURVEY 8(f) N1: H^T (X Theta) = (H^T X) Theta, so a layer can project the E hyperedge
 * rows instead of the N vertex rows (model/ugsys/hgnn.py:22-23 projects the vertices, then aggregates):
 *   hg_plan_edge_reduce   Xe[e,:] = s1[e] * s2[e] * sum_{v in e} a_in[v] * X[v,:]      (Xe overwritten)
 *   hg_plan_edge_scatter  Y[v,:]  = a_out[v] * sum_{e containing v} Xe[e,:]            (Y overwritten, every row)
 * Any scale may be NULL (= 1).  F must be a multiple of 4 and the pointers 16-byte aligned (HG_EINVAL otherwise);
 * the plan must come from the balancer's canonical schedule.  hg_plan_edge_scatter(hg_plan_edge_reduce(X)) runs
 * the same two kernels as hg_aggr_forward's stream form (bit-identical results unless a hyperedge is split over
 * several balancer segments: those are summed with floating-point reductions whose order varies).
 * ------------------------------------------------------------------------- */
HG_API int hg_plan_edge_reduce(hgPlan *plan, const float *d_X, const float *d_s1, const float *d_s2,
                               const float *d_a_in, float *d_Xe, int32_t F, void *stream);
HG_API int hg_plan_edge_scatter(hgPlan *plan, const float *d_Xe, const float *d_a_out, float *d_Y, int32_t F,
                                void *stream);

/* Gradient of the hyperedge weight W (the reference op returns none, hgnnaggr.cc:62-63;
 * un-scaled host reference include/util/check.cuh:116-143):
 *   dW[e] = s1[e] * sum_k (sum_{u in e} a_in[u] X[u,k]) * (sum_{v in e} a_out[v] G[v,k]) */
HG_API int hg_weight_grad(int64_t num_edges, const int32_t *d_t_indptr,
                          const int32_t *d_t_indices, const float *d_X, const float *d_G,
                          const float *d_s1, const float *d_a_out, const float *d_a_in,
                          float *d_dW, int32_t F, int device, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HGEF_B200_H_ */
