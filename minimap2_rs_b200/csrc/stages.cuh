// stages.cuh — host launchers of the mapping stages (seeds.cu, lchain.cu), shared with the orchestration in api.cu.
#pragma once
#include "mm2_internal.cuh"

// per-read result of the chaining stage (device -> host), 64 bytes
struct ReadHit {
  u32 rid_rev;      // (x >> 32) of the chain's first anchor: rev<<31 | rid (an odd-rid anchor reads 0xFFFFFFFF, F5); n_anchors == 0 means no hit
  i32 qs, qe, ts, te;  // paf.rs:136-147 (qs/ts already clamped at 0)
  u32 cm;           // chain length
  i32 score;        // v[best] (lchain.rs:170)
  u32 n_anchors;
  u32 n_mini;       // unfiltered query minimizers
  u32 sum_span;     // paf.rs:160 sum_k
  i32 st_rank, en_rank;  // ranks (in the unfiltered minimizer list) of the chain's first/last forward query position; -1 = absent
  u32 flags;        // bit0: rescue rerun (lchain.rs:321-330)
  i32 best;         // index of the chain's last anchor within the read
  u32 pad0, pad1;
};
static_assert(sizeof(ReadHit) == 64, "ReadHit layout");

int seeds_filter(mm2_ctx* ctx, const u64* d_mkey, const u64* d_mini_off, u32 nreads, u64 n_mini, i32 q_occ_max, float q_occ_frac,
                 u8* d_keep, u32* d_sum_span);
// Query side, version 2 (seeds.cu): count sketch + exact filter where needed + Index::get + compact hit lists + per-read anchor
// offsets (ctx->read_aoff); returns the batch's anchor count (synchronises).  Then anchors built and sorted per read.
int seeds_hits(mm2_ctx* ctx, const IndexView& V, const u64* d_mkey, const u64* d_mval, const u64* d_mini_off, u32 nreads, u64 n_mini,
               i32 q_occ_max, float q_occ_frac, i32 mid_occ, bool full_keep, u32* d_sum_span, u64* n_anchors, u64* n_dropped = nullptr);
int seeds_fill_and_sort(mm2_ctx* ctx, const IndexView& V, const u64* d_mini_off, const u64* d_read_off, u32 nreads, ulonglong2* d_anchors);

// lchain.rs:59-176 forward DP + fallback chain (+ rescue rerun, lchain.rs:321-330) for every read; one warp per read.
// d_A/d_B: int4 per anchor ({f, pprev, v, cnt}, {qs_min, ts_min, first, window start}); d_T, d_W: int per anchor
// (skip marks of lchain.rs:85-86; work list of the anchors that have a non-empty predecessor window).
// d_chain (optional, n_anchors ints): the reported chain of each read, last anchor first, at the read's anchor offset.
int chain_batch(mm2_ctx* ctx, const ulonglong2* d_anchors, const u64* d_read_aoff, const u64* d_read_off, const u64* d_mini_off,
                const u64* d_mval, const u32* d_sum_span, u32 nreads, const mm2_chain_params_t& p, int do_rescue,
                int4* d_A, int4* d_B, int* d_T, int* d_W, int* d_chain, ReadHit* d_hits, unsigned long long* d_cells);

// general.cu: the multi-chain tail of main.rs:209-218 (only reachable with -n < 2)
int map_general_finish(mm2_ctx* ctx, const mm2_index* idx, const u64* d_read_off, const u64* h_off, size_t nreads,
                       const mm2_map_opts_t* o, const mm2_chain_params_t& p, const SketchOut& so, const u32* d_sum_span, u64 nm, u64 na,
                       mm2_map_result_t* out);

// per-device kernel attributes (dynamic shared memory opt-ins); called by mm2_ctx_create after cudaSetDevice
int lchain_init_device();
int radix_init_device();
int seeds_init_device();
