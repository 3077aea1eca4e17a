// index.cu — minimizer index construction, persistence and lookup tables on sm_100a.
// Replaces index.rs:427-475 (build_index_from_fasta), :69-109 (add_minimizers + post_process), :111-154 (stats,
// calc_mid_occ, get), :156-424 (native + MMI save/load).
//
// HBM layout of a built index (all buckets flattened; bucket = low b bits of the minimizer hash):
//   hkeys[n_keys]    (minier>>b)<<1 | is_single, ascending inside each bucket   (index.rs:92-95)
//   hvals[n_keys]    y for a singleton, start_in_bucket_p<<32 | n otherwise     (index.rs:95,100)
//   bkt_koff[2^b+1]  first key of each bucket;  bkt_poff[2^b+1] first p entry of each bucket
//   p[n_p]           positions of multi-occurrence keys, per bucket in ascending key order, each run ascending
//   S[...]           4-bit packed sequence (index.rs:11-19), only needed for .mmi / get_ref_subseq
//   tab[...]         open-addressing table {minier<<1|is_single, val} over all keys for O(1) seed lookup
// The per-bucket Vec<Minimizer> + stable sort + HashMap of the reference become ONE device radix sort on the
// re-keyed minimizers (bucket in the high bits, minier>>b below; stable, so equal keys keep ascending y) followed
// by run-length grouping with exclusive scans.
#include "mm2_internal.cuh"

#include <algorithm>
#include <cub/device/device_radix_sort.cuh>

namespace {

__global__ void pack_seq4_kernel(const u8* __restrict__ seq, u64 total, u32* __restrict__ S, u64 nwords) {
  // index.rs:11-19 mm_seq4_set: base o -> bits 4*(o&7) of word o>>3 ; nt4.rs codes, 4 = other
  for (u64 wi = blockIdx.x * (u64)blockDim.x + threadIdx.x; wi < nwords; wi += (u64)gridDim.x * blockDim.x) {
    const u64 o = wi * 8;
    u32 out = 0;
    if (o < total) {
      u32 lo = 0, hi = 0;
      if (o + 8 <= total) {
        const uint2 v = *reinterpret_cast<const uint2*>(seq + o);
        lo = v.x; hi = v.y;
      } else {
        for (int j = 0; j < 8 && o + j < total; ++j) {
          const u32 b = seq[o + j];
          if (j < 4) lo |= b << (8 * j); else hi |= b << (8 * (j - 4));
        }
      }
      const u32 wd[2] = {lo, hi};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const u32 u = wd[h] & 0xDFDFDFDFu;
        const u32 vm = __vcmpeq4(u, 0x41414141u) | __vcmpeq4(u, 0x43434343u) | __vcmpeq4(u, 0x47474747u) | __vcmpeq4(u, 0x54545454u);
        u32 x = (wd[h] >> 1) & 0x03030303u;
        x = x ^ ((x >> 1) & 0x01010101u);
        x = (x & vm) | (0x04040404u & ~vm);
        // bytes past the end of the genome stay 0 (the reference zero-fills S)
        u32 nib = (x & 0xFu) | ((x >> 4) & 0xF0u) | ((x >> 8) & 0xF00u) | ((x >> 12) & 0xF000u);
        if (o + 4 * h + 4 > total) {
          const int valid = (int)((i64)total - (i64)(o + 4 * h));
          nib = valid <= 0 ? 0u : (nib & ((1u << (4 * valid)) - 1u));
        }
        out |= nib << (16 * h);
      }
    }
    S[wi] = out;
  }
}

// bucket-major sort key: (minier & (2^b-1)) << R | minier >> b   (index.rs:70-71 bucket id, :92 key_top)
__global__ void rekey_kernel(const u64* __restrict__ key_span, u64* __restrict__ ckey, u64 n, int b, int R) {
  const u64 bmask = (1ULL << b) - 1;
  for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    const u64 m = key_span[i] >> 8;
    ckey[i] = ((m & bmask) << R) | (m >> b);
  }
}

__global__ void head_flag_kernel(const u64* __restrict__ ckey, u32* __restrict__ flag, u64 n) {
  for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
    flag[i] = (i == 0 || ckey[i] != ckey[i - 1]) ? 1u : 0u;
}

__global__ void run_start_kernel(const u32* __restrict__ flag, const u64* __restrict__ excl, u64* __restrict__ run_start, u64 n) {
  for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i <= n; i += (u64)gridDim.x * blockDim.x) {
    if (i == n) run_start[excl[n]] = n;
    else if (flag[i]) run_start[excl[i]] = i;
  }
}

__global__ void run_pcount_kernel(const u64* __restrict__ run_start, u32* __restrict__ pc, u64 n_keys) {
  for (u64 r = blockIdx.x * (u64)blockDim.x + threadIdx.x; r < n_keys; r += (u64)gridDim.x * blockDim.x) {
    const u64 n = run_start[r + 1] - run_start[r];
    pc[r] = n > 1 ? (u32)n : 0u;
  }
}

__global__ void bucket_off_kernel(const u64* __restrict__ ckey, const u64* __restrict__ run_start, const u64* __restrict__ gp,
                                  u64 n_keys, int b, int R, u64* __restrict__ bkt_koff, u64* __restrict__ bkt_poff) {
  const u64 nb = 1ULL << b;
  for (u64 bi = blockIdx.x * (u64)blockDim.x + threadIdx.x; bi <= nb; bi += (u64)gridDim.x * blockDim.x) {
    u64 lo = 0, hi = n_keys;  // first run whose bucket >= bi
    while (lo < hi) {
      const u64 mid = (lo + hi) >> 1;
      if ((ckey[run_start[mid]] >> R) < bi) lo = mid + 1; else hi = mid;
    }
    bkt_koff[bi] = lo;
    bkt_poff[bi] = gp[lo];
  }
}

__global__ void run_fill_kernel(const u64* __restrict__ ckey, const u64* __restrict__ y, const u64* __restrict__ run_start,
                                const u64* __restrict__ gp, const u64* __restrict__ bkt_poff, u64 n_keys, int R,
                                u64* __restrict__ hkeys, u64* __restrict__ hvals, unsigned long long* __restrict__ hist,
                                u32* __restrict__ big, u32* __restrict__ n_big, u32 big_cap) {
  // occurrence histogram (index.rs:124-141): counts < 64 (virtually all keys) are accumulated per block in shared memory
  __shared__ u32 s_hist[64];
  if (threadIdx.x < 64) s_hist[threadIdx.x] = 0;
  __syncthreads();
  const u64 rmask = R >= 64 ? ~0ULL : ((1ULL << R) - 1);
  for (u64 r = blockIdx.x * (u64)blockDim.x + threadIdx.x; r < n_keys; r += (u64)gridDim.x * blockDim.x) {
    const u64 first = run_start[r];
    const u64 n = run_start[r + 1] - first;
    const u64 ck = ckey[first];
    const u64 key_top = (ck & rmask) << 1;
    if (n == 1) {
      hkeys[r] = key_top | 1;
      hvals[r] = y[first];
    } else {
      hkeys[r] = key_top;
      hvals[r] = ((gp[r] - bkt_poff[ck >> R]) << 32) | n;
    }
    if (n < 64) atomicAdd(&s_hist[n], 1u);
    else if (n < 65536) atomicAdd(&hist[n], 1ULL);
    else { const u32 s = atomicAdd(n_big, 1u); if (s < big_cap) big[s] = (u32)(n > 0xFFFFFFFFull ? 0xFFFFFFFFull : n); }
  }
  __syncthreads();
  if (threadIdx.x < 64 && s_hist[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
}

__global__ void p_fill_kernel(const u64* __restrict__ y, const u32* __restrict__ flag, const u64* __restrict__ excl,
                              const u64* __restrict__ run_start, const u64* __restrict__ gp, u64 n, u64* __restrict__ p,
                              u32* __restrict__ unsorted) {
  for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    const u64 r = excl[i] + flag[i] - 1;
    const u64 first = run_start[r];
    const u64 cnt = run_start[r + 1] - first;
    if (cnt > 1) {
      p[gp[r] + (i - first)] = y[i];
      if (i > first && y[i] < y[i - 1]) *unsorted = 1u;
    }
  }
}

// rare repair path: positions inside a run were not ascending (only possible when the sketch emitted out of
// position order); index.rs:98 sorts each run, so do the same with one thread per run
__global__ void p_sort_runs_kernel(const u64* __restrict__ run_start, const u64* __restrict__ gp, u64 n_keys, u64* __restrict__ p) {
  for (u64 r = blockIdx.x * (u64)blockDim.x + threadIdx.x; r < n_keys; r += (u64)gridDim.x * blockDim.x) {
    const u64 n = run_start[r + 1] - run_start[r];
    if (n < 2) continue;
    u64* a = p + gp[r];
    for (u64 i = 1; i < n; ++i) {
      const u64 v = a[i];
      u64 j = i;
      while (j > 0 && a[j - 1] > v) { a[j] = a[j - 1]; --j; }
      a[j] = v;
    }
  }
}

// ---- open-addressing lookup table over all keys ---------------------------------------------------------------------
__device__ __forceinline__ u64 tab_hash(u64 minier) {
  u64 x = minier * 0x9E3779B97F4A7C15ULL;
  return x ^ (x >> 29);
}
__global__ void tab_build_kernel(const u64* __restrict__ hkeys, const u64* __restrict__ hvals, const u64* __restrict__ bkt_koff,
                                 u64 n_keys, int b, ulonglong2* __restrict__ tab, u64 tab_mask) {
  const u64 nb = 1ULL << b;
  for (u64 r = blockIdx.x * (u64)blockDim.x + threadIdx.x; r < n_keys; r += (u64)gridDim.x * blockDim.x) {
    // bucket of key r: last bucket whose koff <= r
    u64 lo = 0, hi = nb;
    while (lo < hi) {
      const u64 mid = (lo + hi + 1) >> 1;
      if (bkt_koff[mid] <= r) lo = mid; else hi = mid - 1;
    }
    const u64 hk = hkeys[r];
    const u64 minier = ((hk >> 1) << b) | lo;
    const u64 tag = (minier << 1) | (hk & 1);
    u64 slot = tab_hash(minier) & tab_mask;
    for (;;) {
      const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(&tab[slot].x), ~0ULL, (unsigned long long)tag);
      if (old == ~0ULL) { tab[slot].y = hvals[r]; break; }
      slot = (slot + 1) & tab_mask;
    }
  }
}

__device__ __forceinline__ void bloom_bits(u64 minier, u64 mask, u64& blk, uint4& bits) {
  const u64 h = minier * 0xD6E8FEB86659FD93ULL;
  blk = (h >> 40) & mask;
  u32 w[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; ++i) { const u32 b = (u32)(h >> (7 * i)) & 127u; w[b >> 5] |= 1u << (b & 31); }
  bits = make_uint4(w[0], w[1], w[2], w[3]);
}
__global__ void bloom_build_kernel(const u64* __restrict__ hkeys, const u64* __restrict__ bkt_koff, u64 n_keys, int b, u32* __restrict__ bloom,
                                   u64 bloom_mask) {
  const u64 nb = 1ULL << b;
  for (u64 r = blockIdx.x * (u64)blockDim.x + threadIdx.x; r < n_keys; r += (u64)gridDim.x * blockDim.x) {
    u64 lo = 0, hi = nb;
    while (lo < hi) {
      const u64 mid = (lo + hi + 1) >> 1;
      if (bkt_koff[mid] <= r) lo = mid; else hi = mid - 1;
    }
    const u64 minier = ((hkeys[r] >> 1) << b) | lo;
    u64 blk; uint4 bits;
    bloom_bits(minier, bloom_mask, blk, bits);
    u32* w = bloom + blk * 4;
    if (bits.x) atomicOr(w + 0, bits.x);
    if (bits.y) atomicOr(w + 1, bits.y);
    if (bits.z) atomicOr(w + 2, bits.z);
    if (bits.w) atomicOr(w + 3, bits.w);
  }
}

__global__ void index_get_kernel(IndexView V, u64 minier, u64* out3) {
  // out3 = {kind, val_or_count, p offset}
  const u64 bmask = (1ULL << V.b) - 1;
  const u64 bi = minier & bmask;
  const u64 want = (minier >> V.b);
  u64 lo = V.bkt_koff[bi], hi = V.bkt_koff[bi + 1];
  out3[0] = 0; out3[1] = 0; out3[2] = 0;
  while (lo < hi) {
    const u64 mid = (lo + hi) >> 1;
    const u64 kk = V.hkeys[mid] >> 1;
    if (kk < want) lo = mid + 1; else if (kk > want) hi = mid;
    else {
      const u64 v = V.hvals[mid];
      if (V.hkeys[mid] & 1) { out3[0] = 1; out3[1] = v; }
      else { out3[0] = 2; out3[1] = v & 0xffffffffULL; out3[2] = V.bkt_poff[bi] + (v >> 32); }
      return;
    }
  }
}

inline size_t kroundup64(size_t x) {  // index.rs:10
  x -= 1; x |= x >> 1; x |= x >> 2; x |= x >> 4; x |= x >> 8; x |= x >> 16; x |= x >> 32;
  return x + 1;
}

inline int grid_for(u64 n, int block = 256) { return (int)std::min<u64>((n + block - 1) / block, 148ull * 32); }

}  // namespace

IndexView mm2_index::view() const {
  IndexView v;
  v.w = w; v.k = k; v.b = b; v.flag = flag; v.n_seq = n_seq; v.n_keys = n_keys; v.n_p = n_p;
  v.hkeys = hkeys.as<u64>(); v.hvals = hvals.as<u64>(); v.bkt_koff = bkt_koff.as<u64>(); v.bkt_poff = bkt_poff.as<u64>();
  v.p = p.as<u64>(); v.seq_len = seq_len.as<u32>();
  v.tab = tab.as<ulonglong2>(); v.tab_mask = tab_mask;
  v.bloom = has_bloom ? bloom.as<uint4>() : nullptr; v.bloom_mask = bloom_mask;
  return v;
}

// Build the O(1) lookup table from the flat key arrays (also used after loading an index from disk).
int index_build_table(mm2_ctx* ctx, mm2_index* idx) {
  u64 slots = 1024;
  while (slots < idx->n_keys * 2) slots <<= 1;
  MM2_TRY(idx->tab.ensure(slots * 16));
  idx->tab_mask = slots - 1;
  CUDA_TRY(cudaMemsetAsync(idx->tab.p, 0xFF, slots * 16, ctx->stream));
  if (idx->n_keys) {
    MM2_LAUNCH(ctx, tab_build_kernel, grid_for(idx->n_keys), 256, 0, idx->hkeys.as<u64>(), idx->hvals.as<u64>(),
               idx->bkt_koff.as<u64>(), idx->n_keys, idx->b, idx->tab.as<ulonglong2>(), idx->tab_mask);
    CUDA_TRY(cudaGetLastError());
  }
  // Bloom filter in front of the table, only while it fits comfortably in the 126 MB L2
  idx->has_bloom = false;
  u64 nblk = 1024;
  static const u64 keys_per_blk = [] { const char* e = getenv("MM2_BLOOM_KPB"); return (u64)(e && atoi(e) > 0 ? atoi(e) : 16); }();   // experiment knob
  while (nblk < idx->n_keys / keys_per_blk) nblk <<= 1;  // <= 16 keys x 4 bits per 128-bit block: ~2-3 % false positives
  if (idx->n_keys && nblk * 16 <= (40ull << 20) && !getenv("MM2_NO_BLOOM")) {  // must stay L2-resident next to streaming traffic
    MM2_TRY(idx->bloom.ensure(nblk * 16));
    idx->bloom_mask = nblk - 1;
    CUDA_TRY(cudaMemsetAsync(idx->bloom.p, 0, nblk * 16, ctx->stream));
    MM2_LAUNCH(ctx, bloom_build_kernel, grid_for(idx->n_keys), 256, 0, idx->hkeys.as<u64>(), idx->bkt_koff.as<u64>(), idx->n_keys, idx->b,
               idx->bloom.as<u32>(), idx->bloom_mask);
    CUDA_TRY(cudaGetLastError());
    idx->has_bloom = true;
  }
  return MM2_OK;
}

// (ckey, y) pairs sorted by ckey (stable: equal keys keep their input order).  The input arrays are scratch afterwards;
// ctx->sorted_k / ctx->sorted_v point at the result.  MM2_SORT=cub selects the CUB onesweep sort for comparison.
int radix_sort_pairs(mm2_ctx* ctx, u64* a_keys, u64* a_vals, u64* b_keys, u64* b_vals, u64 n, int end_bit, u64** res_keys, u64** res_vals);
static int index_sort_pairs(mm2_ctx* ctx, u64* ckey_in, u64* y_in, u64 n, int end_bit) {
  cudaStream_t st = ctx->stream;
  MM2_TRY(ctx->sort_keys2.ensure(std::max<u64>(1, n) * 8));
  MM2_TRY(ctx->sort_vals2.ensure(std::max<u64>(1, n) * 8));
  ctx->sorted_k = ctx->sort_keys2.as<u64>(); ctx->sorted_v = ctx->sort_vals2.as<u64>();
  if (n == 0) return MM2_OK;
  static int use_cub = -1;
  if (use_cub < 0) { const char* e = getenv("MM2_SORT"); use_cub = (e && !strcmp(e, "cub")) ? 1 : 0; }
  if (!use_cub) return radix_sort_pairs(ctx, ckey_in, y_in, ctx->sort_keys2.as<u64>(), ctx->sort_vals2.as<u64>(), n, end_bit, &ctx->sorted_k, &ctx->sorted_v);
  size_t tmp_bytes = 0;
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ckey_in, ctx->sort_keys2.as<u64>(), y_in, ctx->sort_vals2.as<u64>(), n, 0,
                                           end_bit, st));
  MM2_TRY(ctx->sort_tmp.ensure(tmp_bytes));
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(ctx->sort_tmp.p, tmp_bytes, ckey_in, ctx->sort_keys2.as<u64>(), y_in, ctx->sort_vals2.as<u64>(),
                                           n, 0, end_bit, st));
  ctx->launches += 2 + (u64)((end_bit + 7) / 8);  // CUB's histogram + onesweep passes (library kernels, not ours)
  return MM2_OK;
}

static int index_finish_from_sorted(mm2_ctx* ctx, mm2_index* idx, const u64* ckey, const u64* y, u64 n, bool build_table);

// device-side part of the build once the minimizers (ctx->mkey/mval, n of them) are resident
static int index_finish_from_minimizers(mm2_ctx* ctx, mm2_index* idx, u64 n) {
  const int b = idx->b, k = idx->k;
  const int R = std::max(2 * k - b, 0);
  const int end_bit = std::max(1, std::min(64, b + R));
  if (n) {
    MM2_TRY(ctx->sort_tmp2.ensure(n * 8));
    MM2_LAUNCH(ctx, rekey_kernel, grid_for(n), 256, 0, ctx->mkey.as<u64>(), ctx->sort_tmp2.as<u64>(), n, b, R);
  }
  MM2_TRY(index_sort_pairs(ctx, ctx->sort_tmp2.as<u64>(), ctx->mval.as<u64>(), n, end_bit));
  return index_finish_from_sorted(ctx, idx, ctx->sorted_k, ctx->sorted_v, n, true);
}

// run-length grouping of the sorted pairs into the flat index arrays (+ occurrence histogram, + lookup table)
static int index_finish_from_sorted(mm2_ctx* ctx, mm2_index* idx, const u64* ckey, const u64* y, u64 n, bool build_table) {
  cudaStream_t st = ctx->stream;
  const int b = idx->b, k = idx->k;
  const int R = std::max(2 * k - b, 0);
  const u64 nb = 1ULL << b;
  idx->n_minimizers = n;
  MM2_TRY(idx->bkt_koff.ensure((nb + 1) * 8));
  MM2_TRY(idx->bkt_poff.ensure((nb + 1) * 8));
  idx->occ_hist.assign(65536, 0);
  idx->occ_big.clear();
  if (n == 0) {
    CUDA_TRY(cudaMemsetAsync(idx->bkt_koff.p, 0, (nb + 1) * 8, st));
    CUDA_TRY(cudaMemsetAsync(idx->bkt_poff.p, 0, (nb + 1) * 8, st));
    idx->n_keys = 0; idx->n_p = 0;
    MM2_TRY(idx->hkeys.ensure(8)); MM2_TRY(idx->hvals.ensure(8)); MM2_TRY(idx->p.ensure(8));
    return build_table ? index_build_table(ctx, idx) : MM2_OK;
  }
  ctx->timer.mark(st, "bucket_build");
  // ---- run-length grouping --------------------------------------------------------------------------------------------
  MM2_TRY(ctx->runidx.ensure(n * 4 + (n + 1) * 8 + 64));
  u32* flag = ctx->runidx.as<u32>();
  u64* excl = (u64*)((u8*)ctx->runidx.p + ((n * 4 + 15) / 16) * 16);
  MM2_LAUNCH(ctx, head_flag_kernel, grid_for(n), 256, 0, ckey, flag, n);
  MM2_TRY(scan_u32_to_u64(ctx, flag, excl, n));
  u64 n_keys = 0;
  MM2_TRY(read_scalar_u64(ctx, excl + n, &n_keys));
  idx->n_keys = n_keys;
  MM2_TRY(ctx->run_start.ensure((n_keys + 1) * 8));
  MM2_TRY(ctx->run_gp.ensure(((n_keys * 4 + 15) / 16) * 16 + (n_keys + 1) * 8 + 64));
  u64* run_start = ctx->run_start.as<u64>();
  u32* pc = ctx->run_gp.as<u32>();
  u64* gp = (u64*)((u8*)ctx->run_gp.p + ((n_keys * 4 + 15) / 16) * 16);
  MM2_LAUNCH(ctx, run_start_kernel, grid_for(n + 1), 256, 0, flag, excl, run_start, n);
  MM2_LAUNCH(ctx, run_pcount_kernel, grid_for(n_keys), 256, 0, run_start, pc, n_keys);
  MM2_TRY(scan_u32_to_u64(ctx, pc, gp, n_keys));
  u64 n_p = 0;
  MM2_TRY(read_scalar_u64(ctx, gp + n_keys, &n_p));
  idx->n_p = n_p;
  MM2_TRY(idx->hkeys.ensure(std::max<u64>(1, n_keys) * 8));
  MM2_TRY(idx->hvals.ensure(std::max<u64>(1, n_keys) * 8));
  MM2_TRY(idx->p.ensure(std::max<u64>(1, n_p) * 8));
  MM2_LAUNCH(ctx, bucket_off_kernel, grid_for(nb + 1), 256, 0, ckey, run_start, gp, n_keys, b, R, idx->bkt_koff.as<u64>(),
             idx->bkt_poff.as<u64>());
  // histogram of occurrence counts (calc_mid_occ / stats)
  const u32 big_cap = (u32)(n / 65536 + 16);
  MM2_TRY(ctx->misc.ensure(65536 * 8 + 16 + (size_t)big_cap * 4));
  unsigned long long* d_hist = ctx->misc.as<unsigned long long>();
  u32* d_nbig = (u32*)(d_hist + 65536);
  u32* d_unsorted = d_nbig + 1;
  u32* d_big = d_nbig + 4;
  CUDA_TRY(cudaMemsetAsync(ctx->misc.p, 0, 65536 * 8 + 16, st));
  MM2_LAUNCH(ctx, run_fill_kernel, grid_for(n_keys), 256, 0, ckey, y, run_start, gp, idx->bkt_poff.as<u64>(), n_keys, R,
             idx->hkeys.as<u64>(), idx->hvals.as<u64>(), d_hist, d_big, d_nbig, big_cap);
  MM2_LAUNCH(ctx, p_fill_kernel, grid_for(n), 256, 0, y, flag, excl, run_start, gp, n, idx->p.as<u64>(), d_unsorted);
  CUDA_TRY(cudaGetLastError());
  u32 h_small[2] = {0, 0};
  CUDA_TRY(cudaMemcpyAsync(idx->occ_hist.data(), d_hist, 65536 * 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(h_small, d_nbig, 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (h_small[0]) {
    idx->occ_big.resize(std::min(h_small[0], big_cap));
    CUDA_TRY(cudaMemcpyAsync(idx->occ_big.data(), d_big, idx->occ_big.size() * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    std::sort(idx->occ_big.begin(), idx->occ_big.end());
  }
  if (h_small[1]) {
    MM2_LAUNCH(ctx, p_sort_runs_kernel, grid_for(n_keys), 256, 0, run_start, gp, n_keys, idx->p.as<u64>());
    CUDA_TRY(cudaGetLastError());
  }
  ctx->timer.mark(st, "lookup_table");
  return build_table ? index_build_table(ctx, idx) : MM2_OK;
}

int index_build_device(mm2_ctx* ctx, const u8* h_cat, const u64* h_off, const char* const* names, size_t nseq, int w, int k,
                       int b, int flag, mm2_index** out) {
  if (!(w > 0 && w < 256) || !(k > 0 && k <= 28)) { mm2_set_error("index: need 0<w<256 and 0<k<=28 (sketch.rs:31-32)"); return MM2_E_ARG; }
  if (b < 0 || b > 28) { mm2_set_error("index: bucket bits out of range"); return MM2_E_ARG; }
  if (nseq > 0x7fffffffull) { mm2_set_error("index: too many sequences"); return MM2_E_ARG; }
  const u64 total = nseq ? h_off[nseq] - h_off[0] : 0;
  for (size_t i = 0; i < nseq; ++i)
    if (h_off[i + 1] - h_off[i] >= (1ull << 31)) { mm2_set_error("index: sequence %zu is >= 2^31 bp (sketch.rs:72 packs pos<<1 in 32 bits)", i); return MM2_E_ARG; }
  cudaStream_t st = ctx->stream;
  mm2_index* idx = new mm2_index();
  idx->device = ctx->device; idx->w = w; idx->k = k; idx->b = b; idx->flag = flag; idx->n_seq = (u32)nseq;
  u64 sum = 0;
  for (size_t i = 0; i < nseq; ++i) {
    const bool hn = names && names[i];
    idx->has_name.push_back(1);  // build_index_from_fasta always stores Some(name) (index.rs:435)
    idx->names.push_back(hn ? std::string(names[i]) : std::string());
    idx->lens.push_back((u32)(h_off[i + 1] - h_off[i]));
    idx->seq_offset.push_back(sum);
    idx->is_alt.push_back(0);
    sum += h_off[i + 1] - h_off[i];
  }
  idx->total_len = total;
  auto fail = [&](int rc) { mm2_index_free(idx); return rc; };
#define IB_TRY(x) do { int r_ = (x); if (r_ != MM2_OK) return fail(r_); } while (0)
#define IB_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { mm2_set_error("%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); return fail(MM2_E_CUDA); } } while (0)
  ctx->timer.reset();
  ctx->timer.mark(st, "h2d");
  // offsets rebased to 0 so that they double as IndexSeq.offset
  std::vector<u64> off0(nseq + 1, 0);
  for (size_t i = 0; i <= nseq && nseq; ++i) off0[i] = h_off[i] - h_off[0];
  IB_TRY(ctx->seq.ensure(total + 64));
  IB_TRY(ctx->seq_off.ensure((nseq + 1) * 8));
  IB_TRY(ctx->pin_in.ensure((nseq + 1) * 8));
  memcpy(ctx->pin_in.p, off0.data(), (nseq + 1) * 8);   // pinned bounce: the copy below must not need a stream sync
  IB_CUDA(cudaMemcpyAsync(ctx->seq_off.p, ctx->pin_in.p, (nseq + 1) * 8, cudaMemcpyHostToDevice, st));
  // The genome goes up in chunks on the copy stream, an event after each; the sketch kernel is launched once per chunk on the
  // tiles that are already resident (sketch_device, SketchFeed), so the 2.6 ms upload of a 145 Mbp genome overlaps the
  // sketch instead of preceding it.
  if (!ctx->copy_stream) IB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  const int nchunks = (int)std::max<u64>(1, std::min<u64>(8, total / (16ull << 20)));
  while (ctx->copy_events.size() < (size_t)nchunks) {
    cudaEvent_t e;
    IB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->copy_events.push_back(e);
  }
  std::vector<u64> chunk_end((size_t)nchunks, total);
  {
    const u8* src = h_cat + (nseq ? h_off[0] : 0);
    u64 b0 = 0;
    for (int c = 0; c < nchunks; ++c) {
      const u64 b1 = c + 1 == nchunks ? total : ((total * (u64)(c + 1) / (u64)nchunks) & ~(u64)15);
      chunk_end[(size_t)c] = b1;
      if (b1 > b0) IB_CUDA(cudaMemcpyAsync(ctx->seq.as<u8>() + b0, src + b0, b1 - b0, cudaMemcpyHostToDevice, ctx->copy_stream));
      IB_CUDA(cudaEventRecord(ctx->copy_events[(size_t)c], ctx->copy_stream));
      b0 = b1;
    }
  }
  IB_TRY(idx->seq_len.ensure(std::max<size_t>(1, nseq) * 4));
  if (nseq) IB_CUDA(cudaMemcpyAsync(idx->seq_len.p, idx->lens.data(), nseq * 4, cudaMemcpyHostToDevice, st));
  const u64 words_used = (total + 7) / 8;
  idx->S_words_alloc = total ? kroundup64((size_t)words_used) : 0;
  IB_TRY(idx->S.ensure(std::max<u64>(1, idx->S_words_alloc) * 4));
  ctx->timer.mark(st, "sketch");
  SketchOut so;
  SketchFeed feed{nchunks, chunk_end.data(), ctx->copy_events.data()};
  IB_TRY(sketch_device(ctx, ctx->seq.as<u8>(), ctx->seq_off.as<u64>(), off0.data(), nseq, w, k, 0, 1, flag & 1, &so, &feed));
  ctx->timer.mark(st, "pack");   // every chunk has been waited for on this stream by now
  if (idx->S_words_alloc) {
    MM2_LAUNCH(ctx, pack_seq4_kernel, grid_for(idx->S_words_alloc), 256, 0, ctx->seq.as<u8>(), total, idx->S.as<u32>(), idx->S_words_alloc);
    IB_CUDA(cudaGetLastError());
  }
  ctx->timer.mark(st, "sort");
  IB_TRY(index_finish_from_minimizers(ctx, idx, so.total));
  ctx->timer.mark(st, "end");
  IB_CUDA(cudaStreamSynchronize(st));
  ctx->timer.finish();
  // h2d, pack, sketch, sort, bucket_build, lookup_table
  float tot = 0;
  for (size_t i = 0; i < ctx->timer.ms.size(); ++i) {
    const std::string& nm = ctx->timer.names[i];
    const float t = ctx->timer.ms[i];
    if (nm == "sketch") idx->build_ms[0] = t;
    else if (nm == "sort") idx->build_ms[1] = t;
    else if (nm == "bucket_build" || nm == "lookup_table") idx->build_ms[2] += t;
    else if (nm == "pack") idx->build_ms[3] = t;
    tot += t;
  }
  idx->build_ms[4] = tot;
#undef IB_TRY
#undef IB_CUDA
  *out = idx;
  return MM2_OK;
}

// ---- public index entry points that need kernels ------------------------------------------------------------------
extern "C" int mm2_index_get(const mm2_index_t* idx, uint64_t minier, uint64_t** occ, size_t* n, int* kind) {
  if (!idx || !occ || !n || !kind) { mm2_set_error("mm2_index_get: NULL argument"); return MM2_E_ARG; }
  CUDA_TRY(cudaSetDevice(idx->device));
  u64* d3 = nullptr;
  CUDA_TRY(cudaMalloc(&d3, 24));
  index_get_kernel<<<1, 1>>>(idx->view(), minier, d3);
  u64 h3[3] = {0, 0, 0};
  cudaError_t e = cudaMemcpy(h3, d3, 24, cudaMemcpyDeviceToHost);
  cudaFree(d3);
  CUDA_TRY(e);
  *kind = (int)h3[0];
  *n = 0; *occ = nullptr;
  if (h3[0] == 1) {
    *occ = (u64*)malloc(8); if (!*occ) return MM2_E_OOM;
    (*occ)[0] = h3[1]; *n = 1;
  } else if (h3[0] == 2) {
    *occ = (u64*)malloc(h3[1] * 8); if (!*occ) return MM2_E_OOM;
    CUDA_TRY(cudaMemcpy(*occ, idx->p.as<u64>() + h3[2], h3[1] * 8, cudaMemcpyDeviceToHost));
    *n = (size_t)h3[1];
  }
  return MM2_OK;
}

// =====================================================================================================================
// Multi-GPU index build building blocks.  One process per GPU; the orchestration and the two exchange
// steps (all-to-all of minimizers by bucket owner, replication of the finished bucket ranges) live in
// minimap2_rs_b200/multi_gpu.py on top of torch.distributed/NCCL.  Rank r owns the contiguous bucket range
// [ceil(r*2^b/R), ceil((r+1)*2^b/R)); because the sort key is bucket-major, each destination's records are one contiguous
// slice of the locally sorted array, and because every rank sketches a contiguous range of sequences (ascending rid) and
// the receive buffer is laid out in source-rank order, the receiver's stable re-sort keeps positions ascending.
namespace {
__global__ void owner_bounds_kernel(const u64* __restrict__ ckey, u64 n, int b, int R, int nranks, u64* __restrict__ bounds) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > nranks) return;
  const u64 nb = 1ULL << b;
  const u64 first_bucket = ((u64)r * nb + (u64)nranks - 1) / (u64)nranks;  // ceil(r * 2^b / nranks)
  u64 lo = 0, hi = n;
  while (lo < hi) {
    const u64 mid = (lo + hi) >> 1;
    if ((ckey[mid] >> R) < first_bucket) lo = mid + 1; else hi = mid;
  }
  bounds[r] = lo;
}
}  // namespace

extern "C" int mm2_mg_sketch_sort(mm2_ctx_t* ctx, const uint8_t* cat, const uint64_t* offs, size_t nseq, size_t seq_lo, size_t seq_hi, int w,
                                  int k, int b, int flag, int nranks, uint64_t* counts) {
  if (!ctx || !offs || !counts || nranks < 1 || seq_lo > seq_hi || seq_hi > nseq) { mm2_set_error("mm2_mg_sketch_sort: bad argument"); return MM2_E_ARG; }
  if (b < 0 || b > 28 || nranks > 1024) { mm2_set_error("mm2_mg_sketch_sort: b / nranks out of range"); return MM2_E_ARG; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t ns = seq_hi - seq_lo;
  const u64 base = offs[seq_lo], total = offs[seq_hi] - base;
  std::vector<u64> off0(ns + 1, 0);
  for (size_t i = 0; i <= ns; ++i) off0[i] = offs[seq_lo + i] - base;
  MM2_TRY(ctx->seq.ensure(total + 64));
  MM2_TRY(ctx->seq_off.ensure((ns + 1) * 8));
  if (total) CUDA_TRY(cudaMemcpyAsync(ctx->seq.p, cat + base, total, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(ctx->seq_off.p, off0.data(), (ns + 1) * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  SketchOut so;
  so.total = 0;
  if (ns) MM2_TRY(sketch_device(ctx, ctx->seq.as<u8>(), ctx->seq_off.as<u64>(), off0.data(), ns, w, k, (u32)seq_lo, 1, flag & 1, &so));
  const u64 n = so.total;
  const int R = std::max(2 * k - b, 0);
  const int end_bit = std::max(1, std::min(64, b + R));
  if (n) {
    MM2_TRY(ctx->sort_tmp2.ensure(n * 8));
    MM2_LAUNCH(ctx, rekey_kernel, grid_for(n), 256, 0, ctx->mkey.as<u64>(), ctx->sort_tmp2.as<u64>(), n, b, R);
  }
  MM2_TRY(index_sort_pairs(ctx, ctx->sort_tmp2.as<u64>(), ctx->mval.as<u64>(), n, end_bit));
  MM2_TRY(ctx->misc.ensure(((size_t)nranks + 2) * 8));
  MM2_LAUNCH(ctx, owner_bounds_kernel, (nranks + 1 + 63) / 64, 64, 0, ctx->sorted_k, n, b, R, nranks, ctx->misc.as<u64>());
  std::vector<u64> bounds((size_t)nranks + 1);
  CUDA_TRY(cudaMemcpyAsync(bounds.data(), ctx->misc.p, ((size_t)nranks + 1) * 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  for (int r = 0; r < nranks; ++r) counts[r] = bounds[(size_t)r + 1] - bounds[(size_t)r];
  ctx->mg_sorted_n = n;
  return MM2_OK;
}

extern "C" int mm2_mg_export_sorted(mm2_ctx_t* ctx, void* d_ckey, void* d_y, size_t n) {
  if (!ctx || (n && (!d_ckey || !d_y)) || n != ctx->mg_sorted_n) { mm2_set_error("mm2_mg_export_sorted: bad argument"); return MM2_E_ARG; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (n) {
    CUDA_TRY(cudaMemcpyAsync(d_ckey, ctx->sorted_k, n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(d_y, ctx->sorted_v, n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return MM2_OK;
}

extern "C" int mm2_mg_build_partial(mm2_ctx_t* ctx, void* d_ckey, void* d_y, size_t n, int w, int k, int b, int flag,
                                    mm2_index_t** out) {
  if (!ctx || !out || (n && (!d_ckey || !d_y))) { mm2_set_error("mm2_mg_build_partial: bad argument"); return MM2_E_ARG; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  mm2_index* idx = new mm2_index();
  idx->device = ctx->device; idx->w = w; idx->k = k; idx->b = b; idx->flag = flag;
  const int R = std::max(2 * k - b, 0);
  const int end_bit = std::max(1, std::min(64, b + R));
  ctx->timer.reset();
  ctx->timer.mark(ctx->stream, "sort");
  int rc = index_sort_pairs(ctx, (u64*)d_ckey, (u64*)d_y, n, end_bit);  // the received buffers are scratch from here on
  if (rc == MM2_OK) rc = index_finish_from_sorted(ctx, idx, ctx->sorted_k, ctx->sorted_v, n, false);
  if (rc == MM2_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) { mm2_set_error("stream sync failed"); rc = MM2_E_CUDA; }
  if (rc != MM2_OK) { mm2_index_free(idx); return rc; }
  *out = idx;
  return MM2_OK;
}

extern "C" int mm2_mg_pack_seq(mm2_ctx_t* ctx, const uint8_t* cat, uint64_t total_len, uint64_t word_lo, uint64_t word_hi, void* d_S) {
  if (!ctx || !d_S || word_lo > word_hi) { mm2_set_error("mm2_mg_pack_seq: bad argument"); return MM2_E_ARG; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  const u64 nwords = word_hi - word_lo;
  if (!nwords) return MM2_OK;
  const u64 b0 = std::min<u64>(total_len, word_lo * 8), b1 = std::min<u64>(total_len, word_hi * 8);
  MM2_TRY(ctx->seq.ensure((b1 - b0) + 64));
  if (b1 > b0) CUDA_TRY(cudaMemcpyAsync(ctx->seq.p, cat + b0, b1 - b0, cudaMemcpyHostToDevice, ctx->stream));
  MM2_LAUNCH(ctx, pack_seq4_kernel, grid_for(nwords), 256, 0, ctx->seq.as<u8>(), b1 - b0, (u32*)d_S + word_lo, nwords);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return MM2_OK;
}

extern "C" int mm2_device_copy(mm2_ctx_t* ctx, void* dst, const void* src, size_t nbytes) {
  if (!ctx || (nbytes && (!dst || !src))) { mm2_set_error("mm2_device_copy: bad argument"); return MM2_E_ARG; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (nbytes) CUDA_TRY(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return MM2_OK;
}

extern "C" int mm2_index_raw(const mm2_index_t* idx, mm2_index_raw_t* out) {
  if (!idx || !out) { mm2_set_error("mm2_index_raw: NULL argument"); return MM2_E_ARG; }
  out->n_keys = idx->n_keys; out->n_p = idx->n_p; out->n_minimizers = idx->n_minimizers; out->S_words = idx->S_words_alloc;
  out->hkeys = idx->hkeys.p; out->hvals = idx->hvals.p; out->p = idx->p.p; out->bkt_koff = idx->bkt_koff.p; out->bkt_poff = idx->bkt_poff.p;
  out->S = idx->S.p; out->occ_hist = idx->occ_hist.data(); out->n_occ_big = idx->occ_big.size(); out->occ_big = idx->occ_big.data();
  return MM2_OK;
}

extern "C" int mm2_index_assemble(mm2_ctx_t* ctx, const uint64_t* offs, const char* const* names, size_t nseq, int w, int k, int b, int flag,
                                  uint64_t n_keys, uint64_t n_p, const void* d_hkeys, const void* d_hvals, const void* d_p,
                                  const void* d_koff, const void* d_poff, const void* d_S, uint64_t S_words, const uint64_t* occ_hist,
                                  const uint32_t* occ_big, size_t n_occ_big, mm2_index_t** out) {
  if (!ctx || !out || !d_koff || !d_poff || !occ_hist || (nseq && !offs)) { mm2_set_error("mm2_index_assemble: NULL argument"); return MM2_E_ARG; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  mm2_index* idx = new mm2_index();
  idx->device = ctx->device; idx->w = w; idx->k = k; idx->b = b; idx->flag = flag; idx->n_seq = (u32)nseq;
  u64 sum = 0;
  for (size_t i = 0; i < nseq; ++i) {
    idx->has_name.push_back(1);
    idx->names.push_back(names && names[i] ? std::string(names[i]) : std::string());
    idx->lens.push_back((u32)(offs[i + 1] - offs[i]));
    idx->seq_offset.push_back(sum);
    idx->is_alt.push_back(0);
    sum += offs[i + 1] - offs[i];
  }
  idx->total_len = sum;
  idx->n_keys = n_keys; idx->n_p = n_p; idx->S_words_alloc = S_words;
  const size_t nb = (size_t)1 << b;
  auto fail = [&](int rc) { mm2_index_free(idx); return rc; };
#define AS_TRY(x) do { int r_ = (x); if (r_ != MM2_OK) return fail(r_); } while (0)
#define AS_CUDA(x) do { if ((x) != cudaSuccess) { mm2_set_error("%s:%d: CUDA error", __FILE__, __LINE__); return fail(MM2_E_CUDA); } } while (0)
  AS_TRY(idx->hkeys.ensure(std::max<u64>(1, n_keys) * 8)); AS_TRY(idx->hvals.ensure(std::max<u64>(1, n_keys) * 8));
  AS_TRY(idx->p.ensure(std::max<u64>(1, n_p) * 8)); AS_TRY(idx->bkt_koff.ensure((nb + 1) * 8)); AS_TRY(idx->bkt_poff.ensure((nb + 1) * 8));
  AS_TRY(idx->S.ensure(std::max<u64>(1, S_words) * 4)); AS_TRY(idx->seq_len.ensure(std::max<size_t>(1, nseq) * 4));
  if (n_keys) { AS_CUDA(cudaMemcpyAsync(idx->hkeys.p, d_hkeys, n_keys * 8, cudaMemcpyDeviceToDevice, st)); AS_CUDA(cudaMemcpyAsync(idx->hvals.p, d_hvals, n_keys * 8, cudaMemcpyDeviceToDevice, st)); }
  if (n_p) AS_CUDA(cudaMemcpyAsync(idx->p.p, d_p, n_p * 8, cudaMemcpyDeviceToDevice, st));
  AS_CUDA(cudaMemcpyAsync(idx->bkt_koff.p, d_koff, (nb + 1) * 8, cudaMemcpyDeviceToDevice, st));
  AS_CUDA(cudaMemcpyAsync(idx->bkt_poff.p, d_poff, (nb + 1) * 8, cudaMemcpyDeviceToDevice, st));
  if (S_words && d_S) AS_CUDA(cudaMemcpyAsync(idx->S.p, d_S, S_words * 4, cudaMemcpyDeviceToDevice, st));
  if (nseq) AS_CUDA(cudaMemcpyAsync(idx->seq_len.p, idx->lens.data(), nseq * 4, cudaMemcpyHostToDevice, st));
  idx->occ_hist.assign(occ_hist, occ_hist + 65536);
  if (n_occ_big && occ_big) { idx->occ_big.assign(occ_big, occ_big + n_occ_big); std::sort(idx->occ_big.begin(), idx->occ_big.end()); }
  u64 so = 0;
  for (size_t c = 0; c < 65536; ++c) so += idx->occ_hist[c] * c;
  for (u32 c : idx->occ_big) so += c;
  idx->n_minimizers = so;
  AS_TRY(index_build_table(ctx, idx));
  AS_CUDA(cudaStreamSynchronize(st));
#undef AS_TRY
#undef AS_CUDA
  *out = idx;
  return MM2_OK;
}
