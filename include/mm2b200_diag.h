/* mm2b200_diag.h — DIAGNOSTIC entry points of libmm2b200.so (bench.py, profiling tools, the parity harness).
 * Not part of the drop-in boundary: nothing in the reference corresponds to them and a Rust caller does not need them.
 * They never change results. */
#ifndef MM2B200_DIAG_H
#define MM2B200_DIAG_H

#include "mm2b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* number of kernels launched by this context since creation (bench.py's gpu_launches claim) */
uint64_t mm2_ctx_launch_count(const mm2_ctx_t* ctx);
/* per-stage device time of the last batched call, in milliseconds (CUDA events on the context stream).
 * names: NUL-separated, double-NUL terminated list matching ms[0..n) */
int mm2_ctx_last_timings(const mm2_ctx_t* ctx, const char** names, const float** ms, int* n);
/* on != 0: the chaining kernels also count DP cells = executed iterations of the inner loop of lchain.rs:80 (rescue
 * reruns included), which is the unit of the integer roofline (20 integer ops per cell).  Counting costs
 * kernel time, so bench.py counts in one untimed pass and times with the counter off. */
int mm2_ctx_count_cells(mm2_ctx_t* ctx, int on);
/* DP cells of the last mapping / chaining call of this context made with the counter on */
uint64_t mm2_ctx_last_cells(const mm2_ctx_t* ctx);

/* the share of rank `rank` of `nranks` in the sharded index build (pure host code, no GPU needed):
 * out8 = {tile path?, first tile (or sequence), one past the last, first / one-past-last byte the sketch reads,
 *         first / one-past-last word of the packed sequence array it packs, bytes it uploads} */
int mm2_shard_plan(const uint64_t* offs, size_t nseq, int w, int k, int flag, int nranks, int rank, uint64_t* out8);

#ifdef __cplusplus
}
#endif
#endif /* MM2B200_DIAG_H */
