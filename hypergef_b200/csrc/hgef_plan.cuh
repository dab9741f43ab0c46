// hgef_plan.cuh -- the aggregation plan (internal layout; opaque in the C-ABI).
#pragma once

#include <vector>

#include "hgef_common.cuh"

struct hgPlan {
  int device = 0;
  int64_t num_nodes = 0, num_edges = 0, nnz = 0, nseg = 0, ngroup = 0;
  // borrowed from the caller (the balancer output and the CSR column indices of H^T)
  const int32_t *key = nullptr, *row = nullptr, *st = nullptr, *ed = nullptr, *colind = nullptr;
  // derived, owned
  int32_t *seg_edge = nullptr;    // [nseg]   hyperedge of each segment
  int32_t *seg_slot = nullptr;    // [nseg]   -1: the segment is its whole hyperedge ("light");
                                  //          else row of `scratch` its hyperedge reduces into
  int32_t *heavy_segs = nullptr;  // [nheavy_segs] the segments with slot >= 0, ascending
  int64_t nheavy_edges = 0, nheavy_segs = 0;
  int32_t canonical = 0;          // groups == full cross product in balancer order
  int32_t max_seg_len = 0;
  // fused persistent kernel (hgef_fused.cu)
  int32_t *cflag = nullptr;       // [nnz] colind with bit31 = first occurrence of the vertex in
                                  //       H_T_colind order, bit30 = the vertex has exactly one
                                  //       occurrence (its Y row has a single writer)
  int32_t *iso_list = nullptr;    // [niso] vertices in no hyperedge (their Y row is just zero)
  int64_t niso = 0, nexcl = 0;
  int32_t *ctrl = nullptr;        // [nseg + 64] per-call tile counter, give-up flag, block counts, tile flags
  // pull form (hgef_fused.cu): CSR of H (vertex -> its hyperedges, ascending), built by transposing
  // the caller's H^T; Xe scratch [num_edges, F] for the hyperedge features between the two phases
  int32_t *h_ptr = nullptr, *h_ind = nullptr;
  int32_t max_vdeg = 0;
  float *xe = nullptr;
  size_t xe_floats = 0;
  // stream form (hgef_stream.cu): row programs of the two stages, base runs, stage-B vertex order
  int32_t *st_srcA = nullptr, *st_dstA = nullptr, *st_runA = nullptr;
  int32_t *st_srcB = nullptr, *st_dstB = nullptr, *st_needB = nullptr, *st_runB = nullptr;
  int32_t *st_ptrB = nullptr;     // [N + 1] first stage-B position of every unit (vertex in st_perm order)
  int32_t *st_perm = nullptr;     // [N] vertices ordered by their last hyperedge; the tail holds the isolated ones
  int32_t *st_ctrl = nullptr;     // ticket counters of the two-launch form
  int64_t st_nrunA = 0, st_nrunB = 0, st_nunitB = 0, st_niso = 0;
  int32_t st_ready = 0;
  static constexpr int kMaxSched = 8;
  // ring form (hgef_ring.cu): hyperedges ordered by their last stage-B read position (the discard order), and
  // merged A / B / discard ticket orders for one item size and pair of lags
  int32_t *rg_dperm = nullptr, *rg_dlast = nullptr;
  int32_t *rg_runA = nullptr, *rg_runB = nullptr;   // unit-aligned runs of kL0f positions
  int64_t rg_nrunA = 0, rg_nrunB = 0;
  int32_t rg_ready = 0;
  struct RingSched {
    int bpi, lagB, lagC, nslab, discard, ksub;
    int32_t GA, GB, GC, nblkA, nblkB, nitem;
    int4 *tabs;                   // the same records per kind and index: A at [0, 2 GA), B, then discard items
    int32_t lead;                 // max over B items of (A items it needs - its own index): the bound of the split-role form
    int4 *items;                  // per ticket two words: {first position, end position, kind | index << 2, blocks needed},
                                  //   {sub-stream split points 1..3, 0}
    int32_t *ctrl;                // kCtrlHdr + nslab * (nblkA + nblkB) words
  };
  RingSched rg_sched[kMaxSched];
  int rg_nsched = 0;
  int32_t *rg_last_ctrl = nullptr;
  // scratch: partial hyperedge features of heavy hyperedges, [nheavy_edges, F]; L2-resident
  float *scratch = nullptr;
  size_t scratch_floats = 0;
  // padded copies of X / Y for feature lengths that are not a multiple of 4 (stream form on padded rows)
  float *pad_x = nullptr, *pad_y = nullptr;
  size_t pad_x_floats = 0, pad_y_floats = 0;
  // buffers replaced by larger ones: a launch in flight or a captured CUDA graph may still name them, so they
  // live until the plan is destroyed
  std::vector<void *> retired;
  int sm_count = 148;
  int64_t kernels_launched = 0;   // this library's own kernels launched through the plan (bench.py reports it)
};

namespace hg {
// Grows a plan-owned device buffer to `need` floats.  Never frees the old buffer before the plan dies and never
// allocates while `s` is being captured (HG_EINVAL: reserve with hg_plan_reserve or one eager call first).
int plan_grow(hgPlan *p, float **buf, size_t *cap, size_t need, cudaStream_t s, const char *what);
}  // namespace hg
