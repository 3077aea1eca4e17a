// mm2_internal.cuh — shared host/device declarations of libmm2b200.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#include "mm2b200.h"
#include "mm2b200_diag.h"

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t u8;
typedef int32_t i32;
typedef int64_t i64;

void mm2_set_error(const char* fmt, ...);

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      mm2_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__));     \
      return e__ == cudaErrorMemoryAllocation ? MM2_E_OOM : MM2_E_CUDA;                         \
    }                                                                                           \
  } while (0)
#define MM2_TRY(expr)            \
  do {                           \
    int r__ = (expr);            \
    if (r__ != MM2_OK) return r__; \
  } while (0)

// growable device buffer (cudaMalloc is slow; buffers persist in the context and only ever grow).
// `pooled` buffers (the arrays of an index object, which come and go with every build/load) use the stream-ordered
// allocator on the legacy default stream with an unlimited release threshold, so a rebuilt index reuses the pages of the
// freed one instead of paying cudaMalloc/cudaFree again.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  bool pooled = false;
  int ensure(size_t bytes) {
    if (bytes <= cap) return MM2_OK;
    release();
    size_t want = bytes + std::min<size_t>(bytes / 8, (size_t)64 << 20) + 256;   // growth slack, capped: index tables reach tens of GB
    cudaError_t e = alloc(want);
    if (e == cudaErrorMemoryAllocation) {
      // The default pool keeps the pages of freed indexes for the next build (release threshold: unlimited).  When an
      // allocation of either kind fails, hand that cache back to the driver and try once more.
      cudaGetLastError();
      int dev = 0;
      cudaMemPool_t pool;
      if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        cudaDeviceSynchronize();
        cudaMemPoolTrimTo(pool, 0);
        e = alloc(want);
      }
    }
    if (e != cudaSuccess) { mm2_set_error("device allocation of %zu bytes failed: %s", want, cudaGetErrorString(e)); p = nullptr; cudaGetLastError(); return MM2_E_OOM; }
    cap = want;
    return MM2_OK;
  }
  cudaError_t alloc(size_t want) {
    cudaError_t e;
    if (pooled) {
      e = cudaMallocAsync(&p, want, (cudaStream_t)0);
      if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)0);
    } else e = cudaMalloc(&p, want);
    if (e != cudaSuccess) p = nullptr;
    return e;
  }
  void release() {
    if (p) {
      if (pooled) { cudaDeviceSynchronize(); cudaFreeAsync(p, (cudaStream_t)0); } else cudaFree(p);
    }
    p = nullptr; cap = 0;
  }
  template <class T> T* as() const { return (T*)p; }
};

struct PinBuf {  // growable page-locked host staging buffer
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return MM2_OK;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) { mm2_set_error("cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e)); p = nullptr; return MM2_E_OOM; }
    cap = want;
    return MM2_OK;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return (T*)p; }
};

struct StageTimer {  // CUDA-event stopwatch on the context stream
  std::vector<std::string> names;
  std::vector<cudaEvent_t> ev;
  std::vector<float> ms;
  std::string names_blob;
  void reset() { names.clear(); n_used = 0; }
  size_t n_used = 0;
  void mark(cudaStream_t s, const char* name) {  // marks the START of stage `name` (and the end of the previous one)
    if (n_used == ev.size()) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); }
    cudaEventRecord(ev[n_used++], s);
    names.push_back(name);
  }
  void finish() {  // call after a stream sync; the last mark is a terminator
    ms.clear(); names_blob.clear();
    for (size_t i = 0; i + 1 < n_used; ++i) {
      float t = 0; cudaEventElapsedTime(&t, ev[i], ev[i + 1]);
      ms.push_back(t);
      names_blob += names[i]; names_blob.push_back('\0');
    }
    names_blob.push_back('\0');
    names.resize(n_used ? n_used - 1 : 0);   // drop the terminator's name: names[i] <-> ms[i] from here on (add_host appends to both)
  }
  void add_host(const char* name, float t) {  // a host-side wall-clock entry appended after finish()
    if (!names_blob.empty()) names_blob.pop_back();
    names_blob += name; names_blob.push_back('\0'); names_blob.push_back('\0');
    ms.push_back(t);
    names.push_back(name);
  }
  ~StageTimer() { for (auto e : ev) cudaEventDestroy(e); }
};

struct mm2_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  u64 launches = 0;
  StageTimer timer;
  // scratch arenas (named by first use; all only ever grow)
  DevBuf seq, seq_off, tile_seq, tile_first, tile_status, misc;      // sketch inputs / bookkeeping
  DevBuf packed, packed_n;                                           // mm2_map_batch_packed: 2-bit reads, N positions
  DevBuf mkey, mval, mini_off;                                       // minimizers (SoA) + per-sequence offsets
  DevBuf keep, occ_cnt, occ_loc, anchor_off_m, scan_status;          // filter / lookup
  DevBuf anchors, read_aoff, read_class;                             // anchors
  DevBuf read_flag, read_nhit, read_na, flag_list;                   // per-read seeding state (seeds.cu seed_hits_kernel)
  DevBuf dpA, dpB, dpT, dpW, hits, chain_idx, lut;                        // chaining
  DevBuf sort_tmp, sort_tmp2, sort_keys2, sort_vals2, runidx, run_start, run_gp, rs_counts, rs_offs;  // index build
  u64* sorted_k = nullptr; u64* sorted_v = nullptr;  // where the last index_sort_pairs left its result
  PinBuf pin_in, pin_out, pin_small, pin_scalar;  // pin_scalar: 8-byte device->host reads (pageable targets serialise streams)
  mm2_ctx* worker[4] = {nullptr, nullptr, nullptr, nullptr};  // sub-batch pipeline of mm2_map_batch (host buffers)
  cudaStream_t copy_stream = nullptr;             // uploads of the pipelined host path
  std::vector<cudaEvent_t> copy_events;           // one per sub-batch
  bool block_sync = false;                         // MM2_SYNC=block: sleep in the three per-batch waits instead of spinning
  cudaEvent_t sync_event = nullptr;
  int n_workers = 4;
  bool pipeline = true;
  u64 subbatch_bytes = 64ull << 20;
  int lut_n = 0;        // entries of the chaining log table resident in `lut`
  float lut_gap = -1.0f;   // chn_pen_gap of the integer penalty table behind it
  int n_sm = 148;
  // chaining: reads with >= chain_dense_min anchors and more than chain_dense_ratio5 / 5 anchors per base get a CTA each
  int chain_dense_min = 4096, chain_dense_ratio5 = 2;
  u64 mg_sorted_n = 0;  // records left in sort_keys2/sort_vals2 by mm2_mg_sketch_sort
  bool count_cells = false;  // diagnostics (mm2b200_diag.h): DP-cell counter of the chaining kernels
  u64 last_cells = 0;
  DevBuf diag;
  DevBuf fine_tmp;               // index build: fine-bucket offsets (scratch of index_build_lookup)
  DevBuf mg_recv_k, mg_recv_v, mg_stage;   // sharded index build: records received from the other ranks; all-gather staging
};

// The waits of the mapping path (minimizer total, anchor total, end of the batch).  cudaStreamSynchronize spins a host core; with
// several ranks x 5 threads on one box that oversubscribes the cores, so MM2_SYNC=block waits on a blocking event instead.
inline cudaError_t mm2_stream_wait(mm2_ctx* ctx) {
  if (!ctx->block_sync) return cudaStreamSynchronize(ctx->stream);
  if (!ctx->sync_event) { cudaError_t e = cudaEventCreateWithFlags(&ctx->sync_event, cudaEventBlockingSync | cudaEventDisableTiming); if (e != cudaSuccess) return e; }
  cudaError_t e = cudaEventRecord(ctx->sync_event, ctx->stream);
  return e != cudaSuccess ? e : cudaEventSynchronize(ctx->sync_event);
}

#define MM2_LAUNCH(ctx, kern, grid, block, smem, ...)                         \
  do {                                                                         \
    kern<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);             \
    (ctx)->launches += 1;                                                      \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
// device-side index view (see index.cu for the layout)
struct IndexView {
  int w, k, b, flag;
  u32 n_seq;
  u64 n_keys, n_p;
  const ulonglong2* kv;  // n_keys: {(minier>>b)<<1 | is_single, y (single) or start_in_bucket_p<<32 | n}, ascending key inside each bucket
  const u64* bkt_koff;   // (1<<b)+1
  const u64* bkt_poff;   // (1<<b)+1
  const u64* p;          // n_p
  const u32* seq_len;    // n_seq
  // Seed lookup: the keys are sorted by (bucket, minier>>b), so a FINE bucket = (bucket, top bits of a monotone
  // equalising map of minier>>b, index_fine_id below) is a contiguous run of kv[] of ~0.7 entries on average.  tab[] holds
  // one 32-byte line per fine bucket: the run's first two records, or {first record, {TAB_MORE, (n-1)<<32 | index of the
  // second record in kv}} when the run is longer; unused slots hold TAB_EMPTY.  Index::get is ONE 32-byte access for ~96 %
  // of the keys, and the table is written by a streaming pass over the sorted records (no random claims).
  const ulonglong2* tab; // 2 << (b + fine_j) records
  int fine_j, fine_pw, R;   // fine bits per bucket; squarings of the equalising map; R = bits of minier>>b (max(2k-b, 0))
  // blocked Bloom filter over the keys (one 16-byte block per key, 4 bits), small enough to live in L2: answers most
  // of the ~80 % of query minimizers that are absent from the index without touching HBM.  NULL = disabled.
  const uint4* bloom;
  u64 bloom_mask;
};

constexpr u64 TAB_EMPTY = ~0ULL, TAB_MORE = ~0ULL - 1;   // never a key: keys are (minier>>b)<<1|single < 2^57

// Monotone map of hk = minier>>b (R bits) onto fine_j bits.  Minimizer hashes are window MINIMA, so hk is skewed towards 0
// with density ~ w (1-x)^(w-1); u = 1 - (1-x)^(2^pw) (integer arithmetic: exactly monotone, identical wherever it is
// evaluated) spreads the keys of a bucket roughly evenly over its 2^fine_j fine buckets.  Only balance depends on it.
__host__ __device__ __forceinline__ u32 index_fine_cdf(u64 hk, int R, int j, int pw) {
  if (j <= 0) return 0u;
  const u64 t = (R >= 64 ? ~0ULL : ((1ULL << R) - 1)) - hk;
  u32 t32 = R >= 32 ? (u32)(t >> (R - 32)) : (u32)(t << (32 - R));
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int s = 0; s < 5; ++s)   // pw <= 5 (index.cu); predicated instead of a loop
    if (s < pw) t32 = (u32)(((u64)t32 * (u64)t32) >> 32);
  return (~t32) >> (32 - j);
}
__host__ __device__ __forceinline__ u64 index_fine_id(const IndexView& V, u64 minier) {
  const u64 bkt = minier & ((1ULL << V.b) - 1);
  return (bkt << V.fine_j) | (u64)index_fine_cdf(minier >> V.b, V.R, V.fine_j, V.fine_pw);
}

struct mm2_index {
  int device = 0;
  int w = 0, k = 0, b = 0, flag = 0;
  u32 n_seq = 0;
  std::vector<std::string> names;
  std::vector<u8> has_name;
  std::vector<u32> lens;
  std::vector<u64> seq_offset;  // index.rs:29 IndexSeq.offset
  std::vector<u8> is_alt;
  u64 total_len = 0;
  u64 S_words_alloc = 0;       // kroundup64((total+7)/8) as allocated by build (native format writes all of it); 0 = no sequence array
  u64 n_keys = 0, n_p = 0, n_minimizers = 0;
  DevBuf S, kv, bkt_koff, bkt_poff, p, seq_len, tab, bloom;
  int fine_j = 0, fine_pw = 0;
  u64 bloom_mask = 0;
  bool has_bloom = false;
  // occurrence histogram for calc_mid_occ/stats (index.rs:111-141): hist[c] = #keys with count c (c < 65536)
  std::vector<u64> occ_hist;
  std::vector<u32> occ_big;     // counts >= 65536
  float build_ms[5] = {0, 0, 0, 0, 0};
  mm2_index() { S.pooled = kv.pooled = bkt_koff.pooled = bkt_poff.pooled = p.pooled = seq_len.pooled = tab.pooled = bloom.pooled = true; }
  IndexView view() const;
};

// read one u64 from device memory through a pinned bounce slot and wait for it
int read_scalar_u64(mm2_ctx* ctx, const u64* d_src, u64* out);

// ---- stage entry points shared between translation units (all asynchronous on ctx->stream unless noted) ----
struct SketchOut {  // device SoA minimizers of a batch
  u64* key = nullptr;   // key_span
  u64* val = nullptr;   // rid_pos_strand
  u64* seq_off = nullptr;  // nseq+1 (device)
  u64 total = 0;           // host copy of seq_off[nseq] (valid after sketch_device returns; it synchronises)
};
// Sketch nseq sequences resident on the device (d_cat/d_off) into ctx->mkey/mval/mini_off.  h_off is the host copy
// of the offsets.  Synchronises the stream once (to learn the total / handle capacity overflow).
// feed (optional): d_cat is still being uploaded by another stream; ev[c] fires when bytes [0, chunk_end[c]) of it (relative
// to the first sequence) are resident.  The tile kernel is then launched once per chunk so that it overlaps the upload.
struct SketchFeed { int nchunks; const u64* chunk_end; cudaEvent_t* ev; };
// shard (optional, tile kernels only): sketch just the tiles [tile_lo, tile_hi) of the batch's global tile list (the sharded
// index build splits INSIDE sequences this way: a tile depends only on its own bases plus a 2w+k halo).  Only the bytes
// sketch_tile_bytes() names need to be resident; out->total = minimizers of the shard, in global emission order.
struct SketchShard { u64 tile_lo, tile_hi; };
u64 sketch_tile_count(const u64* h_off, size_t nseq, int w);
void sketch_tile_bytes(const u64* h_off, size_t nseq, int w, int k, int is_hpc, u64 tile_lo, u64 tile_hi, u64* byte_lo, u64* byte_hi);
bool sketch_uses_tiles(int w, int k, int is_hpc);
int sketch_device(mm2_ctx* ctx, const u8* d_cat, const u64* d_off, const u64* h_off, size_t nseq, int w, int k,
                  u32 rid_base, u32 rid_step, int is_hpc, SketchOut* out, const SketchFeed* feed = nullptr,
                  const SketchShard* shard = nullptr);

int index_build_device(mm2_ctx* ctx, const u8* h_cat, const u64* h_off, const char* const* names, size_t nseq, int w,
                       int k, int b, int flag, mm2_index** out);
// seed-lookup structures (fine offsets + Bloom filter) from kv / bkt_koff; after build, load and sharded assembly
int index_build_lookup(mm2_ctx* ctx, mm2_index* idx);

// generic single-pass exclusive scan of u32 counts into u64 offsets (n+1 outputs; out[n] = total)
// wide = true: tile sums are accumulated in 64 bits (inputs with no bound on their sum)
int scan_u32_to_u64(mm2_ctx* ctx, const u32* d_in, u64* d_out, size_t n, bool wide = false);
