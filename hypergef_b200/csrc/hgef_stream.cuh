// hgef_stream.cuh -- what the stream form (hgef_stream.cu) and the ring form (hgef_ring.cu) share: the
// layout of a row program and of the control words.
#pragma once

#include "hgef_aggr.cuh"

namespace hg {

// dst word of a row program: bit31 = last member of its unit (store now), bit30 = heavy (reduce into the
// pre-zeroed row), low 30 bits = output row
constexpr uint32_t kEnd = 0x80000000u, kHeavy = 0x40000000u, kIdMask = 0x3fffffffu, kRowMask = kIdMask;
constexpr int kL0 = 16;      // positions per base run (unit-aligned)
constexpr int kL0f = 4;      // positions per run of the fused forms (finer items)
constexpr int kBlk = 8;      // items per completion counter
constexpr int kCtrlHdr = 8;  // stream form: ctrl[0] ticket counter, ctrl[1] give-up flag, ctrl[8..] completion counts
// fused forms: the ticket counter (hammered by every warp), the give-up flag and the completion counters (polled
// by waiting warps) live on separate 128-byte lines -- same-line traffic slows the ticket atomics
constexpr int kFlagOff = 32, kDbgOff = 40, kNextB = 64, kNextC = 96, kCntOff = 128;

// process-wide tuning table (hg_tune_set); -1 / absent = the built-in default
int tune_get(const char *name, int dflt);

int ensure_xe(hgPlan *plan, int F, cudaStream_t s);
// stream form, selected stages: 1 = stage A (X -> plan->xe), 2 = stage B (plan->xe -> Y), 3 = both
int launch_stream_stages(hgPlan *p, const dev::Args &a, int stages, cudaStream_t s, float *xe_ext = nullptr);
bool stream_available(const hgPlan *plan, int F, bool force);
int stream_build_runs(hgPlan *p, int L0, int32_t **runA, int64_t *nrunA, int32_t **runB, int64_t *nrunB, cudaStream_t s);

// merged A / B / discard ticket order over items of `bpi` fine runs (hgef_ring.cu), cached in the plan;
// an item is cut into `ksub` sub-streams at run boundaries (bpi % ksub == 0)
enum { kKindA = 0, kKindB = 1, kKindC = 2 };
int fused_get_sched(hgPlan *p, int bpi, int lagB, int lagC, int nslab, int discard, int ksub, cudaStream_t s,
                    hgPlan::RingSched **out);

// fused stream form (hgef_fstream.cu)
bool fstream_available(const hgPlan *plan, int F, bool force);
int launch_fstream(hgPlan *plan, const dev::Args &a, cudaStream_t s);

// ring form (hgef_ring.cu)
bool ring_available(const hgPlan *plan, int F, bool force);
int launch_ring(hgPlan *plan, const dev::Args &a, cudaStream_t s);
void ring_free(hgPlan *plan);
int ring_check(hgPlan *plan, cudaStream_t s);
int ring_debug(hgPlan *plan, int32_t *out8, cudaStream_t s);

}  // namespace hg
