/*
 * hg_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded restatement of the reference's algorithm for the
 * fused hypergraph aggregation path.  It exists to CHECK the CUDA product in
 * tests/, __graft_entry__.smoke() and as bench.py's cpu_baseline leg; nothing
 * under hypergef_b200/ may import, link or execute it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here
 * against (a) golden vectors produced by running the reference's own
 * HyperGsys/balancer.py and scipy path in the build container
 * (oracle/make_golden.py -> tests/golden/), and (b) the reference's own C++
 * (hgnn_ef_full_balance_cpu, hyperaggr_reference_host) compiled in place from
 * /root/reference into oracle/_ref/ (oracle/Makefile).
 *
 * All citations are relative to the reference checkout.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ *
 * Balancer.  Follows HyperGsys/balancer.py:15-33 and its C++ twin
 * HyperGsys/include/taskbalancer/balancer_kernel.cuh:229-259.
 *
 *   for every hyperedge row rid of H_T:
 *     workload = ceil(deg / ngs)                       (balancer.py:19)
 *     key    += lb, lb+ngs, ... while < hb              (balancer.py:20-23)
 *     for i in range(workload): for j in range(workload):
 *        group_st += base+j ; group_ed += base+i ; row += rid   (:24-30)
 *     base += workload                                  (:31)
 *   key += csrptr[nrow] unless already the last entry   (:32-33)
 *
 * Two calls: count (sizes) then fill.  nkey includes the sentinel.  The
 * reference raises IndexError (balan_key[-1] on an empty list) when every
 * row is empty; count reports that case as nkey == 0.
 * ------------------------------------------------------------------------ */
ORC_API int orc_balance_count(int64_t nrow, const int32_t *csrptr, int32_t ngs,
                              int64_t *nkey, int64_t *ngroup)
{
    if (ngs <= 0) return 1;
    int64_t s = 0, g = 0;
    for (int64_t r = 0; r < nrow; ++r) {
        int64_t deg = (int64_t)csrptr[r + 1] - csrptr[r];
        int64_t w = (deg + ngs - 1) / ngs;
        s += w;
        g += w * w;
    }
    *nkey = s ? s + 1 : 0;
    *ngroup = g;
    return 0;
}

ORC_API int orc_balance_fill(int64_t nrow, const int32_t *csrptr, int32_t ngs,
                             int32_t *key, int32_t *row, int32_t *st, int32_t *ed)
{
    if (ngs <= 0) return 1;
    int64_t ks = 0, gs = 0;
    int32_t base = 0;
    for (int64_t r = 0; r < nrow; ++r) {
        int32_t lb = csrptr[r], hb = csrptr[r + 1];
        int32_t w = (int32_t)(((int64_t)hb - lb + ngs - 1) / ngs);
        for (int64_t k = lb; k < hb; k += ngs) key[ks++] = (int32_t)k;
        for (int32_t i = 0; i < w; ++i)
            for (int32_t j = 0; j < w; ++j) {
                st[gs] = base + j;
                ed[gs] = base + i;
                row[gs] = (int32_t)r;
                ++gs;
            }
        base += w;
    }
    if (ks && key[ks - 1] != csrptr[nrow]) key[ks++] = csrptr[nrow];
    return 0;
}

/* ------------------------------------------------------------------------ *
 * Incidence / CSR construction.  Follows HyperGsys/hypergraph.py:23-25:
 *     H   = scipy.sparse.coo_matrix((ones, (V, E)), (N, M)).tocsr()
 *     H_T = H.transpose().tocsr()
 * scipy semantics (coo.tocsr -> sum_duplicates): column indices ascending
 * inside a row, duplicate (row, col) pairs merged with their values summed.
 * The reference's C++ loader does the same with a sorted COO
 * (HyperGsys/include/dataloader/dataloader.hpp:85-141).
 *
 * orc_csr_from_coo: rows/cols int64 [nnz_in] in any order -> CSR with int32
 * indptr[nrow+1], indices, data (float, = multiplicity).  Returns nnz_out via
 * *nnz_out; indices/data must have room for nnz_in entries.
 * ------------------------------------------------------------------------ */
static int cmp_i32(const void *a, const void *b)
{
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

ORC_API int orc_csr_from_coo(int64_t nrow, int64_t ncol, int64_t nnz_in,
                             const int64_t *rows, const int64_t *cols,
                             int32_t *indptr, int32_t *indices, float *data,
                             int64_t *nnz_out)
{
    int64_t *cnt = (int64_t *)calloc((size_t)nrow + 1, sizeof(int64_t));
    int32_t *tmp = (int32_t *)malloc((size_t)(nnz_in ? nnz_in : 1) * sizeof(int32_t));
    if (!cnt || !tmp) { free(cnt); free(tmp); return 2; }
    for (int64_t p = 0; p < nnz_in; ++p) {
        if (rows[p] < 0 || rows[p] >= nrow || cols[p] < 0 || cols[p] >= ncol) {
            free(cnt); free(tmp); return 3;
        }
        cnt[rows[p] + 1]++;
    }
    for (int64_t r = 0; r < nrow; ++r) cnt[r + 1] += cnt[r];
    int64_t *fillp = (int64_t *)malloc((size_t)(nrow + 1) * sizeof(int64_t));
    memcpy(fillp, cnt, (size_t)(nrow + 1) * sizeof(int64_t));
    for (int64_t p = 0; p < nnz_in; ++p) tmp[fillp[rows[p]]++] = (int32_t)cols[p];
    int64_t out = 0;
    indptr[0] = 0;
    for (int64_t r = 0; r < nrow; ++r) {
        int64_t b = cnt[r], e = cnt[r + 1];
        qsort(tmp + b, (size_t)(e - b), sizeof(int32_t), cmp_i32);
        for (int64_t p = b; p < e; ++p) {
            if (p > b && tmp[p] == tmp[p - 1]) {
                data[out - 1] += 1.0f;
            } else {
                indices[out] = tmp[p];
                data[out] = 1.0f;
                ++out;
            }
        }
        indptr[r + 1] = (int32_t)out;
    }
    *nnz_out = out;
    free(cnt); free(tmp); free(fillp);
    return 0;
}

/* CSR (nrow x ncol) -> CSR of the transpose; counting sort, so column indices
 * of the result come out ascending (csc_tocsr in scipy; dataloader.hpp:120-141). */
ORC_API int orc_csr_transpose(int64_t nrow, int64_t ncol, const int32_t *indptr,
                              const int32_t *indices, const float *data,
                              int32_t *t_indptr, int32_t *t_indices, float *t_data)
{
    int64_t nnz = indptr[nrow];
    memset(t_indptr, 0, (size_t)(ncol + 1) * sizeof(int32_t));
    for (int64_t p = 0; p < nnz; ++p) t_indptr[indices[p] + 1]++;
    for (int64_t c = 0; c < ncol; ++c) t_indptr[c + 1] += t_indptr[c];
    int32_t *fillp = (int32_t *)malloc((size_t)(ncol + 1) * sizeof(int32_t));
    if (!fillp) return 2;
    memcpy(fillp, t_indptr, (size_t)(ncol + 1) * sizeof(int32_t));
    for (int64_t r = 0; r < nrow; ++r)
        for (int32_t p = indptr[r]; p < indptr[r + 1]; ++p) {
            int32_t q = fillp[indices[p]]++;
            t_indices[q] = (int32_t)r;
            t_data[q] = data ? data[p] : 1.0f;
        }
    free(fillp);
    return 0;
}

/* ------------------------------------------------------------------------ *
 * Fused aggregation, LITERAL group semantics of the production kernel
 * HyperGsys/source/hgnnaggr/hgnnaggr_cuda.cu:14-47 (one "thread-row" per
 * balancer group), executed sequentially so the summation order is fixed:
 *
 *   for g in groups:  e=row[g]; acc = sum_{p in seg(st[g])} a_in[v_p]*X[v_p]
 *                     acc *= s1[e]*s2[e]
 *                     for p in seg(ed[g]): Y[v_p] += acc * a_out[v_p]
 *
 * s1/s2 (degE, W), a_out (degV) and a_in may be NULL (= 1): NULL s2 gives
 * unignnaggrdeg (unignnaggr_cuda.cu:13-45 with the degV[v] indexing of the
 * _shm variant :92-93), all NULL gives unignnaggr
 * (:219-248).  a_in is not in the reference: it is the gather-side scale of
 * the true-transpose backward.  acc_f64 selects a double
 * accumulator (the fp64 oracle the 1e-5 tolerance is measured against).
 * Y must be zero-filled by the caller (hgnnaggr_cuda.cu:374).
 * ------------------------------------------------------------------------ */
#define GROUP_BODY(ACC_T, OUT_T)                                               \
    for (int64_t g = 0; g < ngroup; ++g) {                                     \
        int32_t e = row[g];                                                    \
        int32_t rs = key[st[g]], re = key[st[g] + 1];                          \
        int32_t ws = key[ed[g]], we = key[ed[g] + 1];                          \
        ACC_T se = (ACC_T)(s1 ? s1[e] : 1.0f) * (ACC_T)(s2 ? s2[e] : 1.0f);    \
        for (int64_t k = 0; k < F; ++k) {                                      \
            ACC_T acc = 0;                                                     \
            for (int32_t p = rs; p < re; ++p) {                                \
                int64_t v = colind[p];                                         \
                ACC_T x = (ACC_T)X[v * F + k];                                 \
                acc += a_in ? x * (ACC_T)a_in[v] : x;                          \
            }                                                                  \
            acc *= se;                                                         \
            for (int32_t p = ws; p < we; ++p) {                                \
                int64_t v = colind[p];                                         \
                ((OUT_T *)Y)[v * F + k] +=                                     \
                    (OUT_T)(a_out ? acc * (ACC_T)a_out[v] : acc);              \
            }                                                                  \
        }                                                                      \
    }

ORC_API int orc_aggr_groups(int64_t ngroup, int64_t F, const int32_t *key,
                            const int32_t *row, const int32_t *st, const int32_t *ed,
                            const int32_t *colind, const float *X, const float *s1,
                            const float *s2, const float *a_out, const float *a_in,
                            void *Y, int acc_f64)
{
    if (acc_f64) { GROUP_BODY(double, double) } else { GROUP_BODY(float, float) }
    return 0;
}

/* ------------------------------------------------------------------------ *
 * The same operator through the two-step formula of the reference's PyG
 * back-end, HyperGsys/model/pygnn/hgnn.py:30-37 (== test/hgnn_test.py:56-63):
 *   Xe = scatter_sum(X[V], E) * degE * W ;  Xv = scatter_sum(Xe[E], V) * degV
 * over the CSR of H_T (row = hyperedge), double accumulation, Y double [N,F].
 * reduce: 0 = sum, 1 = mean (Xe /= deg_e, hgnnaggr_cuda.cu:86-113).
 * ------------------------------------------------------------------------ */
ORC_API int orc_aggr_formula_f64(int64_t N, int64_t M, int64_t F, const int32_t *t_indptr,
                                 const int32_t *t_indices, const float *X, const float *s1,
                                 const float *s2, const float *a_out, const float *a_in,
                                 double *Y, int reduce)
{
    double *xe = (double *)malloc((size_t)(F ? F : 1) * sizeof(double));
    if (!xe) return 2;
    memset(Y, 0, (size_t)(N * F) * sizeof(double));
    for (int64_t e = 0; e < M; ++e) {
        int32_t b = t_indptr[e], h = t_indptr[e + 1];
        for (int64_t k = 0; k < F; ++k) xe[k] = 0.0;
        for (int32_t p = b; p < h; ++p) {
            int64_t v = t_indices[p];
            double a = a_in ? (double)a_in[v] : 1.0;
            for (int64_t k = 0; k < F; ++k) xe[k] += a * (double)X[v * F + k];
        }
        double se = (double)(s1 ? s1[e] : 1.0f) * (double)(s2 ? s2[e] : 1.0f);
        if (reduce == 1 && h > b) se /= (double)(h - b);
        for (int32_t p = b; p < h; ++p) {
            int64_t v = t_indices[p];
            double a = a_out ? (double)a_out[v] : 1.0;
            for (int64_t k = 0; k < F; ++k) Y[v * F + k] += xe[k] * se * a;
        }
    }
    free(xe);
    return 0;
}

/* ------------------------------------------------------------------------ *
 * Un-scaled two-hop aggregation exactly as the reference's own host golden,
 * HyperGsys/include/util/check.cuh:82-114 (hyperaggr_reference_host):
 *   Y[v][k] = sum_{e in H[v]} sum_{u in H_T[e]} X[u][k]     (float, that order)
 * ------------------------------------------------------------------------ */
ORC_API int orc_hyperaggr_host(int64_t N, int64_t F, const int32_t *indptr,
                               const int32_t *indices, const int32_t *t_indptr,
                               const int32_t *t_indices, const float *X, float *Y)
{
    for (int64_t v = 0; v < N; ++v)
        for (int64_t k = 0; k < F; ++k) {
            float a_acc = 0;
            for (int32_t q = indptr[v]; q < indptr[v + 1]; ++q) {
                float b_acc = 0;
                int32_t e = indices[q];
                for (int32_t p = t_indptr[e]; p < t_indptr[e + 1]; ++p)
                    b_acc += X[(int64_t)t_indices[p] * F + k];
                a_acc += b_acc;
            }
            Y[v * F + k] = a_acc;
        }
    return 0;
}

/* ------------------------------------------------------------------------ *
 * f1-max first stage, HyperGsys/source/hgnnaggr/hgnnaggr_cuda.cu:144-178:
 * per hyperedge and column the max member feature (init -1e5, first maximum
 * wins, record_table = argmax vertex), scaled by degE*W, scattered * degV.
 * The reference bounds the hyperedge loop by N; the oracle uses
 * the number of hyperedges M, which is what that code means.
 * Backward (:180-208): grad routed only to the recorded vertex.
 * ------------------------------------------------------------------------ */
ORC_API int orc_aggr_max_fwd(int64_t N, int64_t M, int64_t F, const int32_t *t_indptr,
                             const int32_t *t_indices, const float *X, const float *s1,
                             const float *s2, const float *a_out, double *Y, int32_t *record)
{
    memset(Y, 0, (size_t)(N * F) * sizeof(double));
    for (int64_t e = 0; e < M; ++e) {
        int32_t b = t_indptr[e], h = t_indptr[e + 1];
        float se = (s1 ? s1[e] : 1.0f) * (s2 ? s2[e] : 1.0f);
        for (int64_t k = 0; k < F; ++k) {
            float acc = -1e5f;
            int32_t rec = 0;
            for (int32_t p = b; p < h; ++p) {
                float x = X[(int64_t)t_indices[p] * F + k];
                if (x > acc) { acc = x; rec = t_indices[p]; }
            }
            record[e * F + k] = rec;
            double val = (double)acc * (double)se;
            for (int32_t p = b; p < h; ++p) {
                int64_t v = t_indices[p];
                Y[v * F + k] += val * (a_out ? (double)a_out[v] : 1.0);
            }
        }
    }
    return 0;
}

ORC_API int orc_aggr_max_bwd(int64_t N, int64_t M, int64_t F, const int32_t *t_indptr,
                             const int32_t *t_indices, const float *G, const float *s1,
                             const float *s2, const float *a_out, const int32_t *record,
                             double *dX)
{
    memset(dX, 0, (size_t)(N * F) * sizeof(double));
    for (int64_t e = 0; e < M; ++e) {
        int32_t b = t_indptr[e], h = t_indptr[e + 1];
        double se = (double)(s1 ? s1[e] : 1.0f) * (double)(s2 ? s2[e] : 1.0f);
        for (int64_t k = 0; k < F; ++k) {
            double acc = 0;
            for (int32_t p = b; p < h; ++p) acc += (double)G[(int64_t)t_indices[p] * F + k];
            int64_t v = record[e * F + k];
            dX[v * F + k] += acc * se * (a_out ? (double)a_out[v] : 1.0);
        }
    }
    return 0;
}

/* Gradient of the hyperedge weight W (not returned by the reference op;
 * host reference HyperGsys/include/util/check.cuh:116-143 gives the un-scaled
 * core  sum_k (sum_{u in e} X[u,k]) * (sum_{v in e} G[v,k]) ).  With scales:
 *   dW[e] = s1[e] * sum_k (sum_u a_in[u] X[u,k]) * (sum_v a_out[v] G[v,k]). */
ORC_API int orc_weight_grad_f64(int64_t M, int64_t F, const int32_t *t_indptr,
                                const int32_t *t_indices, const float *X, const float *G,
                                const float *s1, const float *a_out, const float *a_in,
                                double *dW)
{
    for (int64_t e = 0; e < M; ++e) {
        int32_t b = t_indptr[e], h = t_indptr[e + 1];
        double tot = 0;
        for (int64_t k = 0; k < F; ++k) {
            double r1 = 0, r2 = 0;
            for (int32_t p = b; p < h; ++p) {
                int64_t v = t_indices[p];
                r1 += (a_in ? (double)a_in[v] : 1.0) * (double)X[v * F + k];
                r2 += (a_out ? (double)a_out[v] : 1.0) * (double)G[v * F + k];
            }
            tot += r1 * r2;
        }
        dW[e] = tot * (double)(s1 ? s1[e] : 1.0f);
    }
    return 0;
}
