// radix_sort.cu — stable LSD radix sort of (u64 key, u64 value) pairs, 10 bits per pass, hand-written for sm_100a.
// Replaces the per-bucket `sort_by_key` of index.rs:79 (the bucket id is folded into the high bits of the key, so one
// global sort yields bucket-major, key-minor order; stability keeps positions ascending inside equal keys).
//
// Each pass is three kernels over tiles of RS_TILE consecutive pairs:
//   rs_hist     per-tile histogram of the current digit            (reads 8 B / pair)
//   scan        exclusive scan of the (digit-major, tile-minor) counts  -> first output slot of every (digit, tile)
//   rs_scatter  stable rank of every pair inside its tile + scatter     (reads 16 B, writes 16 B / pair)
// Ranks inside a tile: every warp owns a contiguous slice of the tile and walks it 32 pairs at a time; lanes holding the
// same digit find each other with __match_any_sync, the lowest of them bumps the warp's private counter for that digit,
// and rank = old count + number of lower lanes with the same digit.  The per-warp counters are then prefix-summed over the
// warps of the CTA, which makes the rank stable over the whole tile.
#include "mm2_internal.cuh"

#include <algorithm>

namespace {

constexpr int RS_NT = 256;             // threads per CTA
constexpr int RS_WARPS = RS_NT / 32;
constexpr int RS_ROWS = 16;            // 32-pair rows per warp
constexpr int RS_TILE = RS_WARPS * RS_ROWS * 32;  // 4096 pairs per CTA
constexpr int RS_BITS = 10;            // digit width: 30 key bits (k = 15) take 3 passes
constexpr int RS_BINS = 1 << RS_BITS;
constexpr int RS_DPT = RS_BINS / RS_NT;  // digits per thread in the per-digit loops

// digit of a key: a bit field (the sort passes), or the rank that owns the key's bucket in the sharded index build
// (owner = floor(bucket * nranks / 2^b): rank r owns buckets [ceil(r 2^b / R), ceil((r + 1) 2^b / R)); the key is bucket << shift | ...)
struct DigitFn { int shift, b, nranks; };   // nranks == 0: bit field
__device__ __forceinline__ u32 digit_of(u64 key, const DigitFn f) {
  return f.nranks == 0 ? ((u32)(key >> f.shift) & (RS_BINS - 1)) : (u32)(((key >> f.shift) * (u64)f.nranks) >> f.b);
}

__global__ void __launch_bounds__(RS_NT) rs_hist_kernel(const u64* __restrict__ keys, u64 n, const DigitFn fn, u32 ntiles, u32* __restrict__ counts) {
  __shared__ u32 s_h[RS_BINS];
  for (u32 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#pragma unroll
    for (int q = 0; q < RS_DPT; ++q) s_h[q * RS_NT + threadIdx.x] = 0;
    __syncthreads();
    const u64 base = (u64)tile * RS_TILE;
#pragma unroll 4
    for (int r = 0; r < RS_TILE / RS_NT; ++r) {
      const u64 i = base + (u64)r * RS_NT + threadIdx.x;
      if (i < n) atomicAdd(&s_h[digit_of(keys[i], fn)], 1u);
    }
    __syncthreads();
    // digit-major so that one scan orders digits first, tiles second
#pragma unroll
    for (int q = 0; q < RS_DPT; ++q) counts[(u64)(q * RS_NT + threadIdx.x) * ntiles + tile] = s_h[q * RS_NT + threadIdx.x];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(RS_NT) rs_scatter_kernel(const u64* __restrict__ keys, const u64* __restrict__ vals, u64 n, const DigitFn fn,
                                                           u32 ntiles, const u64* __restrict__ offs, u64* __restrict__ out_keys,
                                                           u64* __restrict__ out_vals) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  u64* sk = reinterpret_cast<u64*>(rs_smem);          // tile staged in (digit, rank) order: keys
  u64* sv = sk + RS_TILE;                             //                                    values
  __shared__ u16 s_cnt[RS_WARPS][RS_BINS];  // per-warp digit counters, then exclusive prefix over the warps
  __shared__ u64 s_base[RS_BINS];           // first global output slot of each digit for this tile
  __shared__ u16 s_toff[RS_BINS];           // first staged slot of each digit inside the tile
  __shared__ u32 s_wsum[RS_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 lt = (1u << lane) - 1u;
  for (u32 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#pragma unroll
    for (int q = 0; q < RS_DPT; ++q) {
      const int dg = q * RS_NT + threadIdx.x;
#pragma unroll
      for (int w = 0; w < RS_WARPS; ++w) s_cnt[w][dg] = 0;
      s_base[dg] = offs[(u64)dg * ntiles + tile];
    }
    __syncthreads();
    const u64 tbase = (u64)tile * RS_TILE;
    const u64 wbase = tbase + (u64)warp * (RS_ROWS * 32);
    const u32 tile_n = (u32)min((u64)RS_TILE, n - tbase);
    u64 k[RS_ROWS];
    u32 rank[RS_ROWS];   // digit << 16 | rank inside the warp's slice
#pragma unroll
    for (int r = 0; r < RS_ROWS; ++r) {
      const u64 i = wbase + (u64)r * 32 + lane;
      const bool in = i < n;
      k[r] = in ? keys[i] : ~0ULL;
      const u32 d = in ? digit_of(k[r], fn) : (u32)RS_BINS;   // out-of-range lanes form their own group
      const u32 peers = __match_any_sync(0xFFFFFFFFu, d);
      const int leader = __ffs(peers) - 1;
      u32 old = 0;
      if (lane == leader && in) { old = s_cnt[warp][d]; s_cnt[warp][d] = (u16)(old + __popc(peers)); }
      old = __shfl_sync(0xFFFFFFFFu, old, leader);
      rank[r] = (d << 16) | (old + __popc(peers & lt));
      __syncwarp();
    }
    __syncthreads();
    // exclusive prefix of the per-warp counters over the warps (each thread owns RS_DPT CONSECUTIVE digits), the digit
    // totals of the tile, and their exclusive scan -> where each digit's run starts in the staged tile
    u32 tot[RS_DPT], tsum = 0;
#pragma unroll
    for (int q = 0; q < RS_DPT; ++q) {
      const int dg = threadIdx.x * RS_DPT + q;
      u32 run = 0;
#pragma unroll
      for (int w = 0; w < RS_WARPS; ++w) { const u32 c = s_cnt[w][dg]; s_cnt[w][dg] = (u16)run; run += c; }
      tot[q] = run; tsum += run;
    }
    u32 inc = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    u32 wb = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) if (w < warp) wb += s_wsum[w];
    {
      u32 run = wb + inc - tsum;
#pragma unroll
      for (int q = 0; q < RS_DPT; ++q) { s_toff[threadIdx.x * RS_DPT + q] = (u16)run; run += tot[q]; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ROWS; ++r) {
      const u64 i = wbase + (u64)r * 32 + lane;
      if (i < n) {
        const u32 d = rank[r] >> 16;
        const u32 lp = s_toff[d] + s_cnt[warp][d] + (rank[r] & 0xFFFFu);
        sk[lp] = k[r];
        sv[lp] = vals[i];
      }
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < tile_n; i += RS_NT) {   // consecutive staged slots of a digit go to consecutive addresses
      const u64 kk = sk[i];
      const u32 d = digit_of(kk, fn);
      const u64 o = s_base[d] + (u64)(i - s_toff[d]);
      out_keys[o] = kk;
      out_vals[o] = sv[i];
    }
    __syncthreads();
  }
}

}  // namespace

int radix_init_device() {   // per device (see lchain_init_device)
  CUDA_TRY(cudaFuncSetAttribute(rs_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_TILE * 16));
  return MM2_OK;
}

// Sorts n pairs by key bits [0, end_bit).  The passes ping-pong between buffer A = (a_keys, a_vals), which holds the input
// and is overwritten, and buffer B; *res_keys / *res_vals say which one holds the result.
int radix_sort_pairs(mm2_ctx* ctx, u64* a_keys, u64* a_vals, u64* b_keys, u64* b_vals, u64 n, int end_bit, u64** res_keys, u64** res_vals) {
  *res_keys = a_keys; *res_vals = a_vals;
  if (n == 0) return MM2_OK;
  const int npass = std::max(1, (end_bit + RS_BITS - 1) / RS_BITS);
  const u32 ntiles = (u32)((n + RS_TILE - 1) / RS_TILE);
  const size_t ncnt = (size_t)ntiles * RS_BINS;
  MM2_TRY(ctx->rs_counts.ensure(ncnt * 4 + 64));
  MM2_TRY(ctx->rs_offs.ensure((ncnt + 1) * 8 + 64));
  const int grid = (int)std::min<u32>(ntiles, 148u * 3u);
  u64 *src_k = a_keys, *src_v = a_vals, *dst_k = b_keys, *dst_v = b_vals;
  for (int pass = 0; pass < npass; ++pass) {
    const DigitFn fn{pass * RS_BITS, 0, 0};
    MM2_LAUNCH(ctx, rs_hist_kernel, grid, RS_NT, 0, src_k, n, fn, ntiles, ctx->rs_counts.as<u32>());
    MM2_TRY(scan_u32_to_u64(ctx, ctx->rs_counts.as<u32>(), ctx->rs_offs.as<u64>(), ncnt));
    MM2_LAUNCH(ctx, rs_scatter_kernel, grid, RS_NT, RS_TILE * 16, src_k, src_v, n, fn, ntiles, ctx->rs_offs.as<u64>(), dst_k, dst_v);
    std::swap(src_k, dst_k); std::swap(src_v, dst_v);
  }
  CUDA_TRY(cudaGetLastError());
  *res_keys = src_k; *res_vals = src_v;
  return MM2_OK;
}

namespace {
__global__ void owner_starts_kernel(const u64* __restrict__ offs, u32 ntiles, int nranks, u64 n, u64* __restrict__ bounds) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > nranks) return;
  bounds[r] = r == nranks ? n : offs[(u64)r * ntiles];   // first output slot of digit r = records owned by the ranks below r
}
}  // namespace

// Sharded index build: ONE stable partition pass that groups the (bucket-major key, value) pairs by the rank that owns their
// bucket (instead of a full local sort: the owner sorts what it receives anyway).  d_bounds (device, nranks + 1): first record
// of every owner in the output.
int radix_partition_by_owner(mm2_ctx* ctx, const u64* keys, const u64* vals, u64* out_keys, u64* out_vals, u64 n, int key_shift, int b, int nranks,
                             u64* d_bounds) {
  if (nranks < 1 || nranks > RS_BINS) { mm2_set_error("partition: 1 <= nranks <= %d", RS_BINS); return MM2_E_ARG; }
  const u32 ntiles = (u32)std::max<u64>(1, (n + RS_TILE - 1) / RS_TILE);
  const size_t ncnt = (size_t)ntiles * RS_BINS;
  MM2_TRY(ctx->rs_counts.ensure(ncnt * 4 + 64));
  MM2_TRY(ctx->rs_offs.ensure((ncnt + 1) * 8 + 64));
  const int grid = (int)std::min<u32>(ntiles, 148u * 3u);
  const DigitFn fn{key_shift, b, nranks};
  MM2_LAUNCH(ctx, rs_hist_kernel, grid, RS_NT, 0, keys, n, fn, ntiles, ctx->rs_counts.as<u32>());
  MM2_TRY(scan_u32_to_u64(ctx, ctx->rs_counts.as<u32>(), ctx->rs_offs.as<u64>(), ncnt));
  if (n) MM2_LAUNCH(ctx, rs_scatter_kernel, grid, RS_NT, RS_TILE * 16, keys, vals, n, fn, ntiles, ctx->rs_offs.as<u64>(), out_keys, out_vals);
  MM2_LAUNCH(ctx, owner_starts_kernel, (nranks + 1 + 63) / 64, 64, 0, ctx->rs_offs.as<u64>(), ntiles, nranks, n, d_bounds);
  CUDA_TRY(cudaGetLastError());
  return MM2_OK;
}

// =====================================================================================================================
// One-sweep variant of the sort passes (8-bit digits): every pass reads each pair ONCE.  All per-digit totals of all passes
// come from one histogram kernel up front; inside a pass a tile publishes its per-digit counts and obtains its global offsets
// by a decoupled look-back over the tiles before it (thread d follows digit d), so there is no separate histogram + scan per
// pass.  Small tiles (2048 pairs, 256 threads x 8, ~40 KB of shared memory, <= 64 registers) keep 4 CTAs = 32 warps per SM in
// flight; the three-kernel passes above sat at 2 CTAs / SM and were latency-bound (19 % warps active, 10 % issue-active).
// Stable like the passes above: rows are ranked in order inside a warp, warps in order inside a tile, tiles by the look-back.
namespace {
constexpr int OS_NT = 256, OS_WARPS = OS_NT / 32, OS_IPT = 8, OS_TILE = OS_NT * OS_IPT, OS_BINS = 256;
constexpr u32 OS_FLAG_AGG = 1u << 30, OS_FLAG_INCL = 2u << 30, OS_VAL = (1u << 30) - 1u;

__global__ void __launch_bounds__(256) os_hist_kernel(const u64* __restrict__ keys, u64 n, int npass, unsigned long long* __restrict__ ghist) {
  __shared__ u32 s_h[8][OS_BINS];
  for (int x = threadIdx.x; x < 8 * OS_BINS; x += 256) (&s_h[0][0])[x] = 0;
  __syncthreads();
  for (u64 i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (u64)gridDim.x * 256ull) {
    const u64 k = keys[i];
    for (int p = 0; p < npass; ++p) atomicAdd(&s_h[p][(u32)(k >> (8 * p)) & 255u], 1u);
  }
  __syncthreads();
  for (int x = threadIdx.x; x < npass * OS_BINS; x += 256) {
    const u32 c = (&s_h[0][0])[x];
    if (c) atomicAdd(&ghist[x], (unsigned long long)c);
  }
}
// exclusive scan of every pass's 256 totals (one CTA of 256 threads per pass)
__global__ void __launch_bounds__(256) os_scan_kernel(unsigned long long* __restrict__ ghist) {
  __shared__ unsigned long long s_w[8];
  unsigned long long* h = ghist + (u64)blockIdx.x * OS_BINS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long v = h[threadIdx.x];
  unsigned long long inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += t; }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  unsigned long long base = 0;
  for (int w = 0; w < warp; ++w) base += s_w[w];
  h[threadIdx.x] = base + inc - v;
}

__global__ void __launch_bounds__(OS_NT, 4) os_pass_kernel(const u64* __restrict__ keys, const u64* __restrict__ vals, u64 n, int shift, u32 ntiles,
                                                           const unsigned long long* __restrict__ gbase, u32* __restrict__ status, u32* __restrict__ ticket,
                                                           u64* __restrict__ out_keys, u64* __restrict__ out_vals) {
  __shared__ __align__(16) u64 sk[OS_TILE];
  __shared__ __align__(16) u64 sv[OS_TILE];
  __shared__ u16 s_cnt[OS_WARPS][OS_BINS];
  __shared__ u64 s_gbase[OS_BINS];
  __shared__ u16 s_toff[OS_BINS];
  __shared__ u32 s_wsum[OS_WARPS];
  __shared__ u32 s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
  const u32 lt = (1u << lane) - 1u;
  for (;;) {
    __syncthreads();                                   // the previous tile's shared state is no longer in use
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);      // tiles in order: the look-back needs every earlier tile running or done
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w) s_cnt[w][tid] = 0;
    __syncthreads();
    const u32 tile = s_tile;
    if (tile >= ntiles) return;
    const u64 tbase = (u64)tile * OS_TILE;
    const u64 wbase = tbase + (u64)warp * (OS_IPT * 32);
    const u32 tile_n = (u32)min((u64)OS_TILE, n - tbase);
    u64 k[OS_IPT];
    u32 rk[OS_IPT];   // digit << 16 | rank inside the warp's slice
#pragma unroll
    for (int r = 0; r < OS_IPT; ++r) {
      const u64 i = wbase + (u64)r * 32 + lane;
      k[r] = i < n ? keys[i] : ~0ULL;
    }
#pragma unroll
    for (int r = 0; r < OS_IPT; ++r) {
      const bool in = wbase + (u64)r * 32 + lane < n;
      const u32 d = in ? ((u32)(k[r] >> shift) & 255u) : 256u;   // out-of-range lanes form their own group
      const u32 peers = __match_any_sync(0xFFFFFFFFu, d);
      const int leader = __ffs(peers) - 1;
      u32 old = 0;
      if (lane == leader && in) { old = s_cnt[warp][d]; s_cnt[warp][d] = (u16)(old + __popc(peers)); }
      old = __shfl_sync(0xFFFFFFFFu, old, leader);
      rk[r] = (d << 16) | (old + __popc(peers & lt));
      __syncwarp();
    }
    __syncthreads();
    // thread d owns digit d: exclusive prefix of the per-warp counts, the tile's total, its place in the staged tile
    u32 tot = 0;
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w) { const u32 c = s_cnt[w][tid]; s_cnt[w][tid] = (u16)tot; tot += c; }
    volatile u32* st = status;
    st[(u64)tile * OS_BINS + tid] = (tile == 0 ? OS_FLAG_INCL : OS_FLAG_AGG) | tot;
    u32 inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    u32 wb = 0;
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w) if (w < warp) wb += s_wsum[w];
    const u32 toff = wb + inc - tot;
    s_toff[tid] = (u16)toff;
    // decoupled look-back for digit `tid`
    u32 excl = 0;
    if (tile != 0) {
      for (i64 t = (i64)tile - 1;; --t) {
        u32 v;
        do { v = st[(u64)t * OS_BINS + tid]; } while ((v >> 30) == 0u);
        excl += v & OS_VAL;
        if ((v >> 30) == 2u) break;
      }
      st[(u64)tile * OS_BINS + tid] = OS_FLAG_INCL | (excl + tot);
    }
    s_gbase[tid] = (u64)gbase[tid] + (u64)excl - (u64)toff;   // output slot of staged slot i of digit d: s_gbase[d] + i
    __syncthreads();
#pragma unroll
    for (int r = 0; r < OS_IPT; ++r) {
      const u64 i = wbase + (u64)r * 32 + lane;
      if (i < n) {
        const u32 d = rk[r] >> 16;
        const u32 lp = s_toff[d] + s_cnt[warp][d] + (rk[r] & 0xFFFFu);
        sk[lp] = k[r];
        sv[lp] = vals[i];
      }
    }
    __syncthreads();
    for (u32 i = tid; i < tile_n; i += OS_NT) {   // consecutive staged slots of a digit go to consecutive addresses
      const u64 kk = sk[i];
      const u64 o = s_gbase[(u32)(kk >> shift) & 255u] + (u64)i;
      out_keys[o] = kk;
      out_vals[o] = sv[i];
    }
  }
}
}  // namespace

// same contract as radix_sort_pairs; n < 2^30 (the 30-bit look-back counters), otherwise the caller uses radix_sort_pairs
int radix_onesweep_pairs(mm2_ctx* ctx, u64* a_keys, u64* a_vals, u64* b_keys, u64* b_vals, u64 n, int end_bit, u64** res_keys, u64** res_vals) {
  *res_keys = a_keys; *res_vals = a_vals;
  if (n == 0) return MM2_OK;
  const int npass = std::max(1, (end_bit + 7) / 8);
  if (npass > 8 || n >= (1ull << 30)) return radix_sort_pairs(ctx, a_keys, a_vals, b_keys, b_vals, n, end_bit, res_keys, res_vals);
  const u32 ntiles = (u32)((n + OS_TILE - 1) / OS_TILE);
  const size_t status_bytes = (size_t)ntiles * OS_BINS * 4;
  MM2_TRY(ctx->rs_counts.ensure(status_bytes + 64));                 // look-back status of one pass
  MM2_TRY(ctx->rs_offs.ensure((size_t)8 * OS_BINS * 8 + 64 + 64));  // per-pass digit bases + tickets
  unsigned long long* d_hist = ctx->rs_offs.as<unsigned long long>();
  u32* d_ticket = reinterpret_cast<u32*>(d_hist + 8 * OS_BINS);
  CUDA_TRY(cudaMemsetAsync(d_hist, 0, (size_t)8 * OS_BINS * 8 + 64, ctx->stream));
  MM2_LAUNCH(ctx, os_hist_kernel, (int)std::min<u64>((n + 255) / 256, 148ull * 8), 256, 0, a_keys, n, npass, d_hist);
  MM2_LAUNCH(ctx, os_scan_kernel, npass, 256, 0, d_hist);
  const int grid = (int)std::min<u32>(ntiles, 148u * 4u);
  u64 *src_k = a_keys, *src_v = a_vals, *dst_k = b_keys, *dst_v = b_vals;
  for (int pass = 0; pass < npass; ++pass) {
    CUDA_TRY(cudaMemsetAsync(ctx->rs_counts.p, 0, status_bytes, ctx->stream));
    MM2_LAUNCH(ctx, os_pass_kernel, grid, OS_NT, 0, src_k, src_v, n, pass * 8, ntiles, d_hist + (size_t)pass * OS_BINS, ctx->rs_counts.as<u32>(),
               d_ticket + pass, dst_k, dst_v);
    std::swap(src_k, dst_k); std::swap(src_v, dst_v);
  }
  CUDA_TRY(cudaGetLastError());
  *res_keys = src_k; *res_vals = src_v;
  return MM2_OK;
}
