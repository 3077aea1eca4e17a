// index.cu — minimizer index construction, persistence and lookup tables on sm_100a.
// Replaces index.rs:427-475 (build_index_from_fasta), :69-109 (add_minimizers + post_process), :111-154 (stats,
// calc_mid_occ, get), :156-424 (native + MMI save/load).
//
// HBM layout of a built index (all buckets flattened; bucket = low b bits of the minimizer hash):
//   kv[n_keys]       {(minier>>b)<<1 | is_single, y for a singleton / start_in_bucket_p<<32 | n otherwise}, ascending key
//                    inside each bucket (index.rs:92-100): the hash entries of all buckets, one 16-byte record each
//   bkt_koff[2^b+1]  first key of each bucket;  bkt_poff[2^b+1] first p entry of each bucket
//   p[n_p]           positions of multi-occurrence keys, per bucket in ascending key order, each run ascending
//   S[...]           4-bit packed sequence (index.rs:11-19), only needed for .mmi / get_ref_subseq
//   tab[...]         seed lookup: one 32-byte line per FINE bucket (bucket, equalised top bits of minier>>b) holding the
//                    first two records of the bucket's run in kv (or the first + a pointer to the rest).  kv is sorted by
//                    (bucket, key), so a fine bucket is a contiguous run: the table is written by a streaming pass (no
//                    random claims in a GiB-sized table as in round 1) and Index::get is one 32-byte access
//   bloom[...]       blocked Bloom filter over the keys (only while it fits in L2)
// The per-bucket Vec<Minimizer> + stable sort + HashMap of the reference become ONE device radix sort on the
// re-keyed minimizers (bucket in the high bits, minier>>b below; stable, so equal keys keep ascending y) followed
// by run-length grouping with exclusive scans.
#include "mm2_internal.cuh"

#include <algorithm>

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
                                ulonglong2* __restrict__ kv, unsigned long long* __restrict__ hist,
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
    if (n == 1) kv[r] = make_ulonglong2(key_top | 1, y[first]);
    else kv[r] = make_ulonglong2(key_top, ((gp[r] - bkt_poff[ck >> R]) << 32) | n);
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

// ---- seed-lookup structures: fine offsets + Bloom filter, one warp per bucket ------------------------------------------
__device__ __forceinline__ void bloom_bits(u64 minier, u64 mask, u64& blk, uint4& bits) {
  const u64 h = minier * 0xD6E8FEB86659FD93ULL;
  blk = (h >> 40) & mask;
  u32 w[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; ++i) { const u32 b = (u32)(h >> (7 * i)) & 127u; w[b >> 5] |= 1u << (b & 31); }
  bits = make_uint4(w[0], w[1], w[2], w[3]);
}
// Bucket bi owns fine ids [bi << j, (bi + 1) << j).  Its keys kv[koff[bi] .. koff[bi + 1]) are ascending, and the fine map
// is monotone, so fine_off[f] = index of the first key whose fine id is >= f: every key fills the slots between its
// predecessor's fine id (exclusive) and its own (inclusive).  The last bucket also writes the terminator fine_off[F].
__global__ void __launch_bounds__(256) lookup_build_kernel(const ulonglong2* __restrict__ kv, const u64* __restrict__ bkt_koff, u64 nb, int b, int R,
                                                           int j, int pw, u32* __restrict__ fine_off, u32* __restrict__ bloom, u64 bloom_mask) {
  const int lane = threadIdx.x & 31;
  const u64 nwarps = (u64)gridDim.x * (blockDim.x >> 5);
  for (u64 bi = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); bi < nb; bi += nwarps) {
    const u64 k0 = bkt_koff[bi], k1 = bkt_koff[bi + 1];
    const u64 fbase = bi << j, fend = (bi + 1) << j;
    u64 prev = fbase;             // next fine slot of this bucket that is still unwritten (warp-uniform)
    for (u64 c0 = k0; c0 < k1; c0 += 32) {
      const u64 r = c0 + lane;
      const bool act = r < k1;
      u64 f = fend;               // inactive lanes: past the bucket
      if (act) {
        const u64 hk = kv[r].x >> 1;
        f = fbase | (u64)index_fine_cdf(hk, R, j, pw);
        if (bloom) {
          u64 blk; uint4 bits;
          bloom_bits((hk << b) | bi, bloom_mask, blk, bits);
          u32* wd = bloom + blk * 4;
          if (bits.x) atomicOr(wd + 0, bits.x);
          if (bits.y) atomicOr(wd + 1, bits.y);
          if (bits.z) atomicOr(wd + 2, bits.z);
          if (bits.w) atomicOr(wd + 3, bits.w);
        }
      }
      u64 fp = __shfl_up_sync(0xFFFFFFFFu, f, 1);
      if (lane == 0) fp = prev - 1;             // exclusive predecessor (prev >= fbase; fbase - 1 wraps for bucket 0: handled by +1 below)
      if (act) for (u64 x = fp + 1; x <= f; ++x) fine_off[x] = (u32)r;
      const u64 flast = __shfl_sync(0xFFFFFFFFu, f, 31);
      const u32 am = __ballot_sync(0xFFFFFFFFu, act);
      const u64 fmax = __shfl_sync(0xFFFFFFFFu, f, 31 - __clz(am));   // fine id of the last active lane
      prev = (am == 0xFFFFFFFFu ? flast : fmax) + 1;
    }
    // slots after the bucket's last key point at the first key of the next bucket
    for (u64 x = prev + lane; x < fend; x += 32) fine_off[x] = (u32)k1;
    if (bi + 1 == nb && lane == 0) fine_off[fend] = (u32)k1;
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
    const ulonglong2 e = V.kv[mid];
    const u64 kk = e.x >> 1;
    if (kk < want) lo = mid + 1; else if (kk > want) hi = mid;
    else {
      const u64 v = e.y;
      if (e.x & 1) { out3[0] = 1; out3[1] = v; }
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

// one 32-byte table line per fine bucket from the fine offsets (see IndexView::tab)
__global__ void __launch_bounds__(256) tab_fill_kernel(const ulonglong2* __restrict__ kv, const u32* __restrict__ fine_off, u64 F,
                                                       ulonglong2* __restrict__ tab) {
  for (u64 f = blockIdx.x * (u64)blockDim.x + threadIdx.x; f < F; f += (u64)gridDim.x * blockDim.x) {
    const u32 lo = fine_off[f], hi = fine_off[f + 1];
    const u32 n = hi - lo;
    ulonglong2 e0 = make_ulonglong2(TAB_EMPTY, TAB_EMPTY), e1 = e0;
    if (n >= 1) e0 = kv[lo];
    if (n == 2) e1 = kv[lo + 1];
    else if (n > 2) e1 = make_ulonglong2(TAB_MORE, ((u64)(n - 1) << 32) | (u64)(lo + 1));
    tab[2 * f] = e0;
    tab[2 * f + 1] = e1;
  }
}

IndexView mm2_index::view() const {
  IndexView v;
  v.w = w; v.k = k; v.b = b; v.flag = flag; v.n_seq = n_seq; v.n_keys = n_keys; v.n_p = n_p;
  v.kv = kv.as<ulonglong2>(); v.bkt_koff = bkt_koff.as<u64>(); v.bkt_poff = bkt_poff.as<u64>();
  v.p = p.as<u64>(); v.seq_len = seq_len.as<u32>();
  v.tab = tab.as<ulonglong2>(); v.fine_j = fine_j; v.fine_pw = fine_pw; v.R = std::max(2 * k - b, 0);
  v.bloom = has_bloom ? bloom.as<uint4>() : nullptr; v.bloom_mask = bloom_mask;
  return v;
}

// Seed-lookup structures from kv / bkt_koff (after a build, after loading an index from disk, after the sharded assembly).
int index_build_lookup(mm2_ctx* ctx, mm2_index* idx) {
  if (idx->n_keys >= 0xFFFFFFFFull) { mm2_set_error("index: %llu distinct minimizers do not fit the 32-bit fine offsets", (unsigned long long)idx->n_keys); return MM2_E_UNSUPPORTED; }
  const int R = std::max(2 * idx->k - idx->b, 0);
  // fine buckets: the smallest power of two >= n_keys (mean occupancy in (0.5, 1]: ~4 % of the runs are longer than the two
  // records a table line holds), at most R bits below the bucket
  int j = 0;
  while (j < R && j < 30 && (1ULL << (idx->b + j)) < idx->n_keys) ++j;
  idx->fine_j = j;
  // squarings of the equalising map: (1-x)^(2^pw) with 2^pw the power of two nearest to w (in log scale)
  int pw = 0;
  while (pw < 5 && (1 << pw) * (1 << pw) * 2 <= idx->w * idx->w) ++pw;   // 2^pw <= w / sqrt(2)
  idx->fine_pw = pw;
  const u64 nb = 1ULL << idx->b, F = nb << j;
  MM2_TRY(ctx->fine_tmp.ensure((F + 1) * 4));
  MM2_TRY(idx->tab.ensure(F * 32));
  // Bloom filter in front of the lookup, only while it fits comfortably in the 126 MB L2
  idx->has_bloom = false;
  u64 nblk = 1024;
  while (nblk < idx->n_keys / 16) nblk <<= 1;  // <= 16 keys x 4 bits per 128-bit block: ~2-3 % false positives
  if (idx->n_keys && nblk * 16 <= (40ull << 20)) {  // must stay L2-resident next to streaming traffic
    MM2_TRY(idx->bloom.ensure(nblk * 16));
    idx->bloom_mask = nblk - 1;
    CUDA_TRY(cudaMemsetAsync(idx->bloom.p, 0, nblk * 16, ctx->stream));
    idx->has_bloom = true;
  }
  const int grid = (int)std::max<u64>(1, std::min<u64>((nb + 7) / 8, 148ull * 64));
  MM2_LAUNCH(ctx, lookup_build_kernel, grid, 256, 0, idx->kv.as<ulonglong2>(), idx->bkt_koff.as<u64>(), nb, idx->b, R, j, pw,
             ctx->fine_tmp.as<u32>(), idx->has_bloom ? idx->bloom.as<u32>() : (u32*)nullptr, idx->bloom_mask);
  MM2_LAUNCH(ctx, tab_fill_kernel, (int)std::max<u64>(1, std::min<u64>((F + 255) / 256, 148ull * 32)), 256, 0, idx->kv.as<ulonglong2>(),
             ctx->fine_tmp.as<u32>(), F, idx->tab.as<ulonglong2>());
  CUDA_TRY(cudaGetLastError());
  return MM2_OK;
}

// (ckey, y) pairs sorted by ckey (stable: equal keys keep their input order).  The input arrays are scratch afterwards;
// ctx->sorted_k / ctx->sorted_v point at the result.
int radix_sort_pairs(mm2_ctx* ctx, u64* a_keys, u64* a_vals, u64* b_keys, u64* b_vals, u64 n, int end_bit, u64** res_keys, u64** res_vals);
int radix_onesweep_pairs(mm2_ctx* ctx, u64* a_keys, u64* a_vals, u64* b_keys, u64* b_vals, u64 n, int end_bit, u64** res_keys, u64** res_vals);
int radix_partition_by_owner(mm2_ctx* ctx, const u64* keys, const u64* vals, u64* out_keys, u64* out_vals, u64 n, int key_shift, int b, int nranks,
                             u64* d_bounds);
static int index_sort_pairs(mm2_ctx* ctx, u64* ckey_in, u64* y_in, u64 n, int end_bit) {
  MM2_TRY(ctx->sort_keys2.ensure(std::max<u64>(1, n) * 8));
  MM2_TRY(ctx->sort_vals2.ensure(std::max<u64>(1, n) * 8));
  ctx->sorted_k = ctx->sort_keys2.as<u64>(); ctx->sorted_v = ctx->sort_vals2.as<u64>();
  if (n == 0) return MM2_OK;
  static const bool three_kernel = [] { const char* e = getenv("MM2_SORT"); return e && !strcmp(e, "3k"); }();   // A/B timing of the two pass structures
  if (three_kernel) return radix_sort_pairs(ctx, ckey_in, y_in, ctx->sort_keys2.as<u64>(), ctx->sort_vals2.as<u64>(), n, end_bit, &ctx->sorted_k, &ctx->sorted_v);
  return radix_onesweep_pairs(ctx, ckey_in, y_in, ctx->sort_keys2.as<u64>(), ctx->sort_vals2.as<u64>(), n, end_bit, &ctx->sorted_k, &ctx->sorted_v);
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
    MM2_TRY(idx->kv.ensure(16)); MM2_TRY(idx->p.ensure(8));
    return build_table ? index_build_lookup(ctx, idx) : MM2_OK;
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
  MM2_TRY(idx->kv.ensure(std::max<u64>(1, n_keys) * 16));
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
             idx->kv.as<ulonglong2>(), d_hist, d_big, d_nbig, big_cap);
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
  return build_table ? index_build_lookup(ctx, idx) : MM2_OK;
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
// Bucket-sharded multi-GPU index build (index.rs:427-475 split over the GPUs of one box), inside the
// library.  One rank per GPU (a process under torchrun, or a host thread of mm2_index_build_multi):
//   1. every rank sketches its share of the genome — a contiguous range of TILES of the global tile list, i.e. the split
//      runs inside sequences (a tile depends only on its bases plus a 2w+k halo), so a single-chromosome genome is
//      sketched by all ranks; even k / HPC (literal kernel) fall back to whole sequences — and uploads only those bytes;
//   2. one stable partition pass by bucket owner: rank r owns buckets [ceil(r 2^b / R), ceil((r+1) 2^b / R)), so every
//      destination's records become one contiguous slice, still in emission (= position) order;
//   3. all-to-all (grouped ncclSend / ncclRecv over NVLink);
//   4. stable re-sort + run-length grouping of the owned buckets (index.rs:74-109) — the receive buffer is in source-rank
//      = position order, and p_fill repairs a run should it ever not be;
//   5. replication: ONE grouped all-gather-v (a broadcast per owner inside one NCCL group) of the kv records, the
//      position lists and the packed-sequence words; bucket offset tables, occurrence histogram: all-reduce(sum) — a
//      rank's local offset table is 0 below its range and n_local above it, so the element-wise sum IS the global table;
//   6. every rank derives the seed-lookup structures (fine offsets, Bloom filter) from the replicated arrays — a streaming
//      pass, there is no table to build any more.
// mm2_index_build_sharded_emulated runs the same phases for R virtual ranks on ONE GPU (exchange = device copies): the
// single-GPU tests compare its .mmi with the CPU build byte for byte.
#include "comm.cuh"

#include <thread>

namespace {
__global__ void add_u64_kernel(u64* __restrict__ dst, const u64* __restrict__ src, u64 n) {
  for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) dst[i] += src[i];
}

struct ShardPlan {
  int nranks = 1, rank = 0;
  bool tile_path = true;
  u64 tile_lo = 0, tile_hi = 0;        // tile path: tiles of the global tile list
  size_t seq_lo = 0, seq_hi = 0;       // literal path: whole sequences
  u64 sk_lo = 0, sk_hi = 0;            // bytes the sketch reads
  u64 word_lo = 0, word_hi = 0;        // words of S this rank packs
  u64 up_lo = 0, up_hi = 0;            // bytes uploaded (union of the two)
};

// S words of rank r: equal slots of ceil(used / R) words (so that the replication is one in-place ncclAllGather); the zero
// tail up to the allocated power of two is a local memset on every rank
u64 shard_slot_words(u64 total, int nranks) { const u64 used = (total + 7) / 8; return (used + (u64)nranks - 1) / (u64)nranks; }
void shard_words(u64 total, u64 S_words_alloc, int nranks, int rank, u64* lo, u64* hi) {
  (void)S_words_alloc;
  const u64 slot = shard_slot_words(total, nranks);
  *lo = slot * (u64)rank; *hi = slot * (u64)(rank + 1);
}

void make_plan(const u64* off0, size_t nseq, int w, int k, int flag, u64 S_words_alloc, int nranks, int rank, ShardPlan* P) {
  const u64 total = nseq ? off0[nseq] : 0;
  P->nranks = nranks; P->rank = rank;
  P->tile_path = sketch_uses_tiles(w, k, flag & 1);
  if (P->tile_path) {
    const u64 nt = sketch_tile_count(off0, nseq, w);
    P->tile_lo = nt * (u64)rank / (u64)nranks; P->tile_hi = nt * (u64)(rank + 1) / (u64)nranks;
    sketch_tile_bytes(off0, nseq, w, k, flag & 1, P->tile_lo, P->tile_hi, &P->sk_lo, &P->sk_hi);
  } else {
    // contiguous sequence ranges balanced by bases
    auto cut = [&](int r) -> size_t {
      if (r <= 0) return 0;
      if (r >= nranks) return nseq;
      const u64 target = total * (u64)r / (u64)nranks;
      return (size_t)(std::lower_bound(off0, off0 + nseq + 1, target) - off0);
    };
    P->seq_lo = std::min(cut(rank), nseq); P->seq_hi = std::min(std::max(cut(rank + 1), P->seq_lo), nseq);
    P->sk_lo = off0[P->seq_lo] & ~(u64)15; P->sk_hi = off0[P->seq_hi];
    if (P->seq_hi == P->seq_lo) P->sk_lo = P->sk_hi = 0;
  }
  shard_words(total, S_words_alloc, nranks, rank, &P->word_lo, &P->word_hi);
  const u64 pk_lo = std::min(total, P->word_lo * 8), pk_hi = std::min(total, P->word_hi * 8);
  P->up_lo = total; P->up_hi = 0;
  if (P->sk_hi > P->sk_lo) { P->up_lo = std::min(P->up_lo, P->sk_lo); P->up_hi = std::max(P->up_hi, P->sk_hi); }
  if (pk_hi > pk_lo) { P->up_lo = std::min(P->up_lo, pk_lo); P->up_hi = std::max(P->up_hi, pk_hi); }
  if (P->up_hi <= P->up_lo) P->up_lo = P->up_hi = 0;
}

mm2_index* new_index_skeleton(mm2_ctx* ctx, const u64* h_off, const char* const* names, size_t nseq, int w, int k, int b, int flag) {
  mm2_index* idx = new mm2_index();
  idx->device = ctx->device; idx->w = w; idx->k = k; idx->b = b; idx->flag = flag; idx->n_seq = (u32)nseq;
  u64 sum = 0;
  for (size_t i = 0; i < nseq; ++i) {
    idx->has_name.push_back(1);  // build_index_from_fasta always stores Some(name) (index.rs:435)
    idx->names.push_back(names && names[i] ? std::string(names[i]) : std::string());
    idx->lens.push_back((u32)(h_off[i + 1] - h_off[i]));
    idx->seq_offset.push_back(sum);
    idx->is_alt.push_back(0);
    sum += h_off[i + 1] - h_off[i];
  }
  idx->total_len = sum;
  idx->S_words_alloc = sum ? kroundup64((size_t)((sum + 7) / 8)) : 0;
  return idx;
}

int check_build_args(const u64* h_off, size_t nseq, int w, int k, int b) {
  if (!(w > 0 && w < 256) || !(k > 0 && k <= 28)) { mm2_set_error("index: need 0<w<256 and 0<k<=28 (sketch.rs:31-32)"); return MM2_E_ARG; }
  if (b < 0 || b > 28) { mm2_set_error("index: bucket bits out of range"); return MM2_E_ARG; }
  if (nseq > 0x7fffffffull) { mm2_set_error("index: too many sequences"); return MM2_E_ARG; }
  for (size_t i = 0; i < nseq; ++i)
    if (h_off[i + 1] - h_off[i] >= (1ull << 31)) { mm2_set_error("index: sequence %zu is >= 2^31 bp (sketch.rs:72 packs pos<<1 in 32 bits)", i); return MM2_E_ARG; }
  return MM2_OK;
}

// phases 1-2 of one rank: upload its bytes, sketch, sort bucket-major; bounds[r] = first record owned by rank r (nranks + 1)
int shard_local(mm2_ctx* ctx, const ShardPlan& P, const u8* h_cat, const u64* off0, size_t nseq, int w, int k, int b, int flag,
                std::vector<u64>& bounds) {
  cudaStream_t st = ctx->stream;
  const u64 total = nseq ? off0[nseq] : 0;
  MM2_TRY(ctx->seq.ensure(total + 64));
  MM2_TRY(ctx->seq_off.ensure((nseq + 1) * 8));
  MM2_TRY(ctx->pin_in.ensure((nseq + 1) * 8));
  memcpy(ctx->pin_in.p, off0, (nseq + 1) * 8);
  CUDA_TRY(cudaMemcpyAsync(ctx->seq_off.p, ctx->pin_in.p, (nseq + 1) * 8, cudaMemcpyHostToDevice, st));
  if (P.up_hi > P.up_lo) CUDA_TRY(cudaMemcpyAsync(ctx->seq.as<u8>() + P.up_lo, h_cat + P.up_lo, P.up_hi - P.up_lo, cudaMemcpyHostToDevice, st));
  ctx->timer.mark(st, "sketch");
  SketchOut so;
  so.total = 0;
  if (P.tile_path) {
    const SketchShard sh{P.tile_lo, P.tile_hi};
    if (nseq) MM2_TRY(sketch_device(ctx, ctx->seq.as<u8>(), ctx->seq_off.as<u64>(), off0, nseq, w, k, 0, 1, flag & 1, &so, nullptr, &sh));
  } else if (P.seq_hi > P.seq_lo) {
    MM2_TRY(sketch_device(ctx, ctx->seq.as<u8>(), ctx->seq_off.as<u64>() + P.seq_lo, off0 + P.seq_lo, P.seq_hi - P.seq_lo, w, k, (u32)P.seq_lo, 1,
                          flag & 1, &so));
  }
  ctx->timer.mark(st, "sort");
  const u64 n = so.total;
  const int R = std::max(2 * k - b, 0);
  MM2_TRY(ctx->sort_keys2.ensure(std::max<u64>(1, n) * 8));
  MM2_TRY(ctx->sort_vals2.ensure(std::max<u64>(1, n) * 8));
  MM2_TRY(ctx->misc.ensure(((size_t)P.nranks + 2) * 8));
  if (n) {
    MM2_TRY(ctx->sort_tmp2.ensure(n * 8));
    MM2_LAUNCH(ctx, rekey_kernel, grid_for(n), 256, 0, ctx->mkey.as<u64>(), ctx->sort_tmp2.as<u64>(), n, b, R);
  }
  // one stable partition pass by bucket owner (the owner sorts what it receives; a full local sort would be thrown away)
  MM2_TRY(radix_partition_by_owner(ctx, ctx->sort_tmp2.as<u64>(), ctx->mval.as<u64>(), ctx->sort_keys2.as<u64>(), ctx->sort_vals2.as<u64>(), n, R, b,
                                   P.nranks, ctx->misc.as<u64>()));
  ctx->sorted_k = ctx->sort_keys2.as<u64>(); ctx->sorted_v = ctx->sort_vals2.as<u64>();
  bounds.assign((size_t)P.nranks + 1, 0);
  MM2_TRY(ctx->pin_small.ensure(((size_t)P.nranks + 1) * 8));
  CUDA_TRY(cudaMemcpyAsync(ctx->pin_small.p, ctx->misc.p, ((size_t)P.nranks + 1) * 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  memcpy(bounds.data(), ctx->pin_small.p, ((size_t)P.nranks + 1) * 8);
  return MM2_OK;
}

// phase 4 of one rank: the received (ckey, y) records (scratch afterwards) -> the buckets it owns, no lookup structures
int shard_owned(mm2_ctx* ctx, u64* d_ckey, u64* d_y, u64 n, int w, int k, int b, int flag, mm2_index** out) {
  mm2_index* part = new mm2_index();
  part->device = ctx->device; part->w = w; part->k = k; part->b = b; part->flag = flag;
  const int R = std::max(2 * k - b, 0);
  const int end_bit = std::max(1, std::min(64, b + R));
  int rc = index_sort_pairs(ctx, d_ckey, d_y, n, end_bit);
  if (rc == MM2_OK) rc = index_finish_from_sorted(ctx, part, ctx->sorted_k, ctx->sorted_v, n, false);
  if (rc != MM2_OK) { mm2_index_free(part); return rc; }
  *out = part;
  return MM2_OK;
}

// 4-bit packing of the S words [word_lo, word_hi) from the resident bytes
int shard_pack(mm2_ctx* ctx, mm2_index* idx, const ShardPlan& P) {
  const u64 nwords = P.word_hi - P.word_lo;
  if (!nwords) return MM2_OK;
  const u64 b0 = std::min<u64>(idx->total_len, P.word_lo * 8), b1 = std::min<u64>(idx->total_len, P.word_hi * 8);
  MM2_LAUNCH(ctx, pack_seq4_kernel, grid_for(nwords), 256, 0, ctx->seq.as<u8>() + b0, b1 - b0, idx->S.as<u32>() + P.word_lo, nwords);
  CUDA_TRY(cudaGetLastError());
  return MM2_OK;
}

void finish_build_ms(mm2_ctx* ctx, mm2_index* idx) {
  float tot = 0;
  for (size_t i = 0; i < ctx->timer.ms.size(); ++i) {
    const std::string& nm = ctx->timer.names[i];
    const float t = ctx->timer.ms[i];
    if (nm == "sketch" || nm == "h2d") idx->build_ms[0] += t;
    else if (nm == "sort" || nm == "exchange") idx->build_ms[1] += t;
    else if (nm == "bucket_build" || nm == "lookup_table" || nm == "replicate" || nm == "rep_gather" || nm == "rep_compact") idx->build_ms[2] += t;
    else if (nm == "pack") idx->build_ms[3] += t;
    tot += t;
  }
  idx->build_ms[4] = tot;
}
}  // namespace

extern "C" int mm2_index_build_sharded(mm2_ctx_t* ctx, mm2_comm_t* comm, const uint8_t* cat, const uint64_t* offs, const char* const* names,
                                       size_t nseq, int w, int k, int b, int flag, mm2_index_t** out) {
  if (!ctx || !comm || !out || (nseq && (!cat || !offs))) { mm2_set_error("mm2_index_build_sharded: NULL argument"); return MM2_E_ARG; }
  if (comm->ctx != ctx) { mm2_set_error("mm2_index_build_sharded: the communicator belongs to another context"); return MM2_E_ARG; }
  const NcclApi* N = nccl_api();
  if (!N) return MM2_E_UNSUPPORTED;
  static const u64 zero_off[1] = {0};
  if (!nseq) offs = zero_off;
  MM2_TRY(check_build_args(offs, nseq, w, k, b));
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int R = comm->nranks, me = comm->rank;
  std::vector<u64> off0(nseq + 1, 0);
  for (size_t i = 0; i <= nseq && nseq; ++i) off0[i] = offs[i] - offs[0];
  const u8* h_cat = cat ? cat + offs[0] : nullptr;
  mm2_index* idx = new_index_skeleton(ctx, offs, names, nseq, w, k, b, flag);
  mm2_index* part = nullptr;
  auto fail = [&](int rc) { mm2_index_free(idx); if (part) mm2_index_free(part); return rc; };
#define SB_TRY(x) do { int r_ = (x); if (r_ != MM2_OK) return fail(r_); } while (0)
#define SB_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { mm2_set_error("%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); return fail(MM2_E_CUDA); } } while (0)
#define SB_NCCL(x) do { ncclResult_t e_ = (x); if (e_ != ncclSuccess) { mm2_set_error("%s:%d: %s", __FILE__, __LINE__, N->GetErrorString(e_)); return fail(MM2_E_CUDA); } } while (0)
  ShardPlan P;
  make_plan(off0.data(), nseq, w, k, flag, idx->S_words_alloc, R, me, &P);
  ctx->timer.reset();
  ctx->timer.mark(st, "h2d");
  // ---- phases 1-2
  std::vector<u64> bounds;
  SB_TRY(shard_local(ctx, P, h_cat, off0.data(), nseq, w, k, b, flag, bounds));
  // the packed-sequence words of this rank, while its bytes are resident
  const u64 s_slot = shard_slot_words(idx->total_len, R);
  SB_TRY(idx->S.ensure(std::max<u64>(1, std::max<u64>(idx->S_words_alloc, s_slot * (u64)R)) * 4));
  {
    const u64 used = (idx->total_len + 7) / 8;
    if (idx->S_words_alloc > used) SB_CUDA(cudaMemsetAsync(idx->S.as<u32>() + used, 0, (idx->S_words_alloc - used) * 4, st));   // index.rs:454: zero-filled
  }
  ctx->timer.mark(st, "pack");
  SB_TRY(shard_pack(ctx, idx, P));
  // ---- phase 3: counts matrix, then the records
  ctx->timer.mark(st, "exchange");
  const size_t mat_n = (size_t)R * (size_t)std::max(R, 4);   // R x R counts, later R x 4 sizes
  SB_TRY(ctx->diag.ensure((mat_n + (size_t)R + 32) * 8));
  u64* d_mat = ctx->diag.as<u64>() + 8;          // R x R: row s = what source s sends to every destination
  u64* d_row = ctx->diag.as<u64>() + 8 + mat_n;
  std::vector<u64> row((size_t)R), mat((size_t)R * R);
  for (int r = 0; r < R; ++r) row[(size_t)r] = bounds[(size_t)r + 1] - bounds[(size_t)r];
  SB_TRY(ctx->pin_small.ensure((mat_n + 64) * 8));
  memcpy(ctx->pin_small.p, row.data(), (size_t)R * 8);
  SB_CUDA(cudaMemcpyAsync(d_row, ctx->pin_small.p, (size_t)R * 8, cudaMemcpyHostToDevice, st));
  SB_NCCL(N->AllGather(d_row, d_mat, (size_t)R, ncclUint64, comm->comm, st));
  SB_CUDA(cudaMemcpyAsync(ctx->pin_small.p, d_mat, (size_t)R * R * 8, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  memcpy(mat.data(), ctx->pin_small.p, (size_t)R * R * 8);
  u64 n_recv = 0;
  std::vector<u64> roff((size_t)R + 1, 0);
  for (int s = 0; s < R; ++s) { roff[(size_t)s] = n_recv; n_recv += mat[(size_t)s * R + me]; }
  roff[(size_t)R] = n_recv;
  SB_TRY(ctx->mg_recv_k.ensure(std::max<u64>(1, n_recv) * 8));
  SB_TRY(ctx->mg_recv_v.ensure(std::max<u64>(1, n_recv) * 8));
  SB_NCCL(N->GroupStart());
  for (int p = 0; p < R; ++p) {
    const u64 ns = row[(size_t)p], nr = mat[(size_t)p * R + me];
    if (ns) {
      SB_NCCL(N->Send(ctx->sorted_k + bounds[(size_t)p], ns, ncclUint64, p, comm->comm, st));
      SB_NCCL(N->Send(ctx->sorted_v + bounds[(size_t)p], ns, ncclUint64, p, comm->comm, st));
    }
    if (nr) {
      SB_NCCL(N->Recv(ctx->mg_recv_k.as<u64>() + roff[(size_t)p], nr, ncclUint64, p, comm->comm, st));
      SB_NCCL(N->Recv(ctx->mg_recv_v.as<u64>() + roff[(size_t)p], nr, ncclUint64, p, comm->comm, st));
    }
  }
  SB_NCCL(N->GroupEnd());
  // ---- phase 4
  ctx->timer.mark(st, "sort");
  SB_TRY(shard_owned(ctx, ctx->mg_recv_k.as<u64>(), ctx->mg_recv_v.as<u64>(), n_recv, w, k, b, flag, &part));
  // ---- phase 5: sizes, then one grouped all-gather-v
  ctx->timer.mark(st, "replicate");
  u64 mine[4] = {part->n_keys, part->n_p, (u64)part->occ_big.size(), part->n_minimizers};
  memcpy(ctx->pin_small.p, mine, 32);
  SB_CUDA(cudaMemcpyAsync(d_row, ctx->pin_small.p, 32, cudaMemcpyHostToDevice, st));
  SB_NCCL(N->AllGather(d_row, d_mat, 4, ncclUint64, comm->comm, st));
  SB_CUDA(cudaMemcpyAsync(ctx->pin_small.p, d_mat, (size_t)R * 32, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  std::vector<u64> sz((size_t)R * 4);
  memcpy(sz.data(), ctx->pin_small.p, (size_t)R * 32);
  u64 tk = 0, tp = 0, tbig = 0, tmin = 0;
  std::vector<u64> ko((size_t)R + 1, 0), po((size_t)R + 1, 0), bo((size_t)R + 1, 0);
  for (int r = 0; r < R; ++r) {
    ko[(size_t)r] = tk; po[(size_t)r] = tp; bo[(size_t)r] = tbig;
    tk += sz[(size_t)r * 4]; tp += sz[(size_t)r * 4 + 1]; tbig += sz[(size_t)r * 4 + 2]; tmin += sz[(size_t)r * 4 + 3];
  }
  idx->n_keys = tk; idx->n_p = tp; idx->n_minimizers = tmin;
  const u64 nb = 1ULL << b;
  SB_TRY(idx->kv.ensure(std::max<u64>(1, tk) * 16));
  SB_TRY(idx->p.ensure(std::max<u64>(1, tp) * 8));
  SB_TRY(idx->bkt_koff.ensure((nb + 1) * 8));
  SB_TRY(idx->bkt_poff.ensure((nb + 1) * 8));
  SB_TRY(idx->seq_len.ensure(std::max<size_t>(1, nseq) * 4));
  if (nseq) SB_CUDA(cudaMemcpyAsync(idx->seq_len.p, idx->lens.data(), nseq * 4, cudaMemcpyHostToDevice, st));
  // occurrence histogram (+ the rare counts >= 65536) through device scratch
  SB_TRY(ctx->misc.ensure((65536 + tbig + 16) * 8));
  u64* d_hist = ctx->misc.as<u64>();
  u64* d_big = d_hist + 65536;
  SB_CUDA(cudaMemcpyAsync(d_hist, part->occ_hist.data(), 65536 * 8, cudaMemcpyHostToDevice, st));
  std::vector<u64> big64(part->occ_big.begin(), part->occ_big.end());
  if (!big64.empty()) SB_CUDA(cudaMemcpyAsync(d_big + bo[(size_t)me], big64.data(), big64.size() * 8, cudaMemcpyHostToDevice, st));
  // kv and p: every rank's part goes into an equal-sized slot of a staging buffer, ONE in-place ncclAllGather each replicates
  // the slots (ring / NVLS at full NVLink rate; a broadcast per owner ran at a quarter of it), device copies close the gaps.
  // S: the slots are equal by construction, so the all-gather is in place in the index array itself.
  ctx->timer.mark(st, "rep_gather");
  u64 maxk = 0, maxp = 0;
  for (int r = 0; r < R; ++r) { maxk = std::max(maxk, sz[(size_t)r * 4]); maxp = std::max(maxp, sz[(size_t)r * 4 + 1]); }
  const u64 kslot = maxk * 2, pslot = maxp;                     // u64 words per slot
  SB_TRY(ctx->mg_stage.ensure(std::max<u64>(1, (kslot + pslot) * (u64)R) * 8));
  u64* st_k = ctx->mg_stage.as<u64>();
  u64* st_p = st_k + kslot * (u64)R;
  if (part->n_keys) SB_CUDA(cudaMemcpyAsync(st_k + kslot * (u64)me, part->kv.p, part->n_keys * 16, cudaMemcpyDeviceToDevice, st));
  if (part->n_p) SB_CUDA(cudaMemcpyAsync(st_p + pslot * (u64)me, part->p.p, part->n_p * 8, cudaMemcpyDeviceToDevice, st));
  SB_NCCL(N->GroupStart());
  if (kslot) SB_NCCL(N->AllGather(st_k + kslot * (u64)me, st_k, kslot, ncclUint64, comm->comm, st));
  if (pslot) SB_NCCL(N->AllGather(st_p + pslot * (u64)me, st_p, pslot, ncclUint64, comm->comm, st));
  if (s_slot) SB_NCCL(N->AllGather(idx->S.as<u32>() + s_slot * (u64)me, idx->S.p, s_slot, ncclUint32, comm->comm, st));
  for (int r = 0; r < R; ++r) {
    const u64 ng = sz[(size_t)r * 4 + 2];
    if (ng) SB_NCCL(N->Broadcast(d_big + bo[(size_t)r], d_big + bo[(size_t)r], ng, ncclUint64, r, comm->comm, st));
  }
  SB_NCCL(N->AllReduce(part->bkt_koff.p, idx->bkt_koff.p, nb + 1, ncclUint64, ncclSum, comm->comm, st));
  SB_NCCL(N->AllReduce(part->bkt_poff.p, idx->bkt_poff.p, nb + 1, ncclUint64, ncclSum, comm->comm, st));
  SB_NCCL(N->AllReduce(d_hist, d_hist, 65536, ncclUint64, ncclSum, comm->comm, st));
  SB_NCCL(N->GroupEnd());
  ctx->timer.mark(st, "rep_compact");
  for (int r = 0; r < R; ++r) {
    const u64 nk = sz[(size_t)r * 4], np = sz[(size_t)r * 4 + 1];
    if (nk) SB_CUDA(cudaMemcpyAsync(idx->kv.as<u64>() + 2 * ko[(size_t)r], st_k + kslot * (u64)r, nk * 16, cudaMemcpyDeviceToDevice, st));
    if (np) SB_CUDA(cudaMemcpyAsync(idx->p.as<u64>() + po[(size_t)r], st_p + pslot * (u64)r, np * 8, cudaMemcpyDeviceToDevice, st));
  }
  idx->occ_hist.assign(65536, 0);
  SB_CUDA(cudaMemcpyAsync(idx->occ_hist.data(), d_hist, 65536 * 8, cudaMemcpyDeviceToHost, st));
  big64.assign((size_t)tbig, 0);
  if (tbig) SB_CUDA(cudaMemcpyAsync(big64.data(), d_big, tbig * 8, cudaMemcpyDeviceToHost, st));
  // ---- phase 6
  ctx->timer.mark(st, "lookup_table");
  SB_TRY(index_build_lookup(ctx, idx));
  ctx->timer.mark(st, "end");
  SB_CUDA(cudaStreamSynchronize(st));
  ctx->timer.finish();
  idx->occ_big.clear();
  for (u64 v : big64) idx->occ_big.push_back((u32)v);
  std::sort(idx->occ_big.begin(), idx->occ_big.end());
  finish_build_ms(ctx, idx);
  mm2_index_free(part);
#undef SB_TRY
#undef SB_CUDA
#undef SB_NCCL
  *out = idx;
  return MM2_OK;
}

// The same phases for `nranks` virtual ranks on one GPU / one context; the exchange steps are device copies.
extern "C" int mm2_index_build_sharded_emulated(mm2_ctx_t* ctx, int nranks, const uint8_t* cat, const uint64_t* offs, const char* const* names,
                                                size_t nseq, int w, int k, int b, int flag, mm2_index_t** out) {
  if (!ctx || !out || nranks < 1 || nranks > 1024 || (nseq && (!cat || !offs))) { mm2_set_error("mm2_index_build_sharded_emulated: bad argument"); return MM2_E_ARG; }
  static const u64 zero_off[1] = {0};
  if (!nseq) offs = zero_off;
  MM2_TRY(check_build_args(offs, nseq, w, k, b));
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int R = nranks;
  std::vector<u64> off0(nseq + 1, 0);
  for (size_t i = 0; i <= nseq && nseq; ++i) off0[i] = offs[i] - offs[0];
  const u8* h_cat = cat ? cat + offs[0] : nullptr;
  mm2_index* idx = new_index_skeleton(ctx, offs, names, nseq, w, k, b, flag);
  std::vector<mm2_index*> parts((size_t)R, nullptr);
  std::vector<DevBuf> send_k((size_t)R), send_v((size_t)R);
  auto cleanup = [&]() { for (auto p : parts) if (p) mm2_index_free(p); for (auto& d : send_k) d.release(); for (auto& d : send_v) d.release(); };
  auto fail = [&](int rc) { cleanup(); mm2_index_free(idx); return rc; };
#define SE_TRY(x) do { int r_ = (x); if (r_ != MM2_OK) return fail(r_); } while (0)
#define SE_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { mm2_set_error("%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); return fail(MM2_E_CUDA); } } while (0)
  ctx->timer.reset();
  ctx->timer.mark(st, "h2d");
  SE_TRY(idx->S.ensure(std::max<u64>(1, std::max<u64>(idx->S_words_alloc, shard_slot_words(idx->total_len, R) * (u64)R)) * 4));
  {
    const u64 used = (idx->total_len + 7) / 8;
    if (idx->S_words_alloc > used) SE_CUDA(cudaMemsetAsync(idx->S.as<u32>() + used, 0, (idx->S_words_alloc - used) * 4, st));
  }
  std::vector<std::vector<u64>> bounds((size_t)R);
  for (int r = 0; r < R; ++r) {
    ShardPlan P;
    make_plan(off0.data(), nseq, w, k, flag, idx->S_words_alloc, R, r, &P);
    // poison the sequence buffer so that a rank reading bytes outside its upload range cannot go unnoticed
    if (ctx->seq.cap) SE_CUDA(cudaMemsetAsync(ctx->seq.p, 'N', ctx->seq.cap, st));
    SE_TRY(shard_local(ctx, P, h_cat, off0.data(), nseq, w, k, b, flag, bounds[(size_t)r]));
    SE_TRY(shard_pack(ctx, idx, P));
    const u64 n = bounds[(size_t)r][(size_t)R];
    SE_TRY(send_k[(size_t)r].ensure(std::max<u64>(1, n) * 8));
    SE_TRY(send_v[(size_t)r].ensure(std::max<u64>(1, n) * 8));
    if (n) {
      SE_CUDA(cudaMemcpyAsync(send_k[(size_t)r].p, ctx->sorted_k, n * 8, cudaMemcpyDeviceToDevice, st));
      SE_CUDA(cudaMemcpyAsync(send_v[(size_t)r].p, ctx->sorted_v, n * 8, cudaMemcpyDeviceToDevice, st));
    }
    SE_CUDA(cudaStreamSynchronize(st));
  }
  u64 tk = 0, tp = 0, tmin = 0;
  for (int me = 0; me < R; ++me) {   // "all-to-all": in source-rank order, the slice every source holds for `me`
    u64 n_recv = 0;
    for (int s = 0; s < R; ++s) n_recv += bounds[(size_t)s][(size_t)me + 1] - bounds[(size_t)s][(size_t)me];
    SE_TRY(ctx->mg_recv_k.ensure(std::max<u64>(1, n_recv) * 8));
    SE_TRY(ctx->mg_recv_v.ensure(std::max<u64>(1, n_recv) * 8));
    u64 o = 0;
    for (int s = 0; s < R; ++s) {
      const u64 lo = bounds[(size_t)s][(size_t)me], c = bounds[(size_t)s][(size_t)me + 1] - lo;
      if (c) {
        SE_CUDA(cudaMemcpyAsync(ctx->mg_recv_k.as<u64>() + o, send_k[(size_t)s].as<u64>() + lo, c * 8, cudaMemcpyDeviceToDevice, st));
        SE_CUDA(cudaMemcpyAsync(ctx->mg_recv_v.as<u64>() + o, send_v[(size_t)s].as<u64>() + lo, c * 8, cudaMemcpyDeviceToDevice, st));
      }
      o += c;
    }
    SE_TRY(shard_owned(ctx, ctx->mg_recv_k.as<u64>(), ctx->mg_recv_v.as<u64>(), n_recv, w, k, b, flag, &parts[(size_t)me]));
    SE_CUDA(cudaStreamSynchronize(st));
    tk += parts[(size_t)me]->n_keys; tp += parts[(size_t)me]->n_p; tmin += parts[(size_t)me]->n_minimizers;
  }
  idx->n_keys = tk; idx->n_p = tp; idx->n_minimizers = tmin;
  const u64 nb = 1ULL << b;
  SE_TRY(idx->kv.ensure(std::max<u64>(1, tk) * 16));
  SE_TRY(idx->p.ensure(std::max<u64>(1, tp) * 8));
  SE_TRY(idx->bkt_koff.ensure((nb + 1) * 8));
  SE_TRY(idx->bkt_poff.ensure((nb + 1) * 8));
  SE_TRY(idx->seq_len.ensure(std::max<size_t>(1, nseq) * 4));
  if (nseq) SE_CUDA(cudaMemcpyAsync(idx->seq_len.p, idx->lens.data(), nseq * 4, cudaMemcpyHostToDevice, st));
  SE_CUDA(cudaMemsetAsync(idx->bkt_koff.p, 0, (nb + 1) * 8, st));
  SE_CUDA(cudaMemsetAsync(idx->bkt_poff.p, 0, (nb + 1) * 8, st));
  idx->occ_hist.assign(65536, 0);
  u64 ok = 0, op = 0;
  for (int r = 0; r < R; ++r) {
    mm2_index* pt = parts[(size_t)r];
    if (pt->n_keys) SE_CUDA(cudaMemcpyAsync(idx->kv.as<u64>() + 2 * ok, pt->kv.p, pt->n_keys * 16, cudaMemcpyDeviceToDevice, st));
    if (pt->n_p) SE_CUDA(cudaMemcpyAsync(idx->p.as<u64>() + op, pt->p.p, pt->n_p * 8, cudaMemcpyDeviceToDevice, st));
    MM2_LAUNCH(ctx, add_u64_kernel, grid_for(nb + 1), 256, 0, idx->bkt_koff.as<u64>(), pt->bkt_koff.as<u64>(), nb + 1);
    MM2_LAUNCH(ctx, add_u64_kernel, grid_for(nb + 1), 256, 0, idx->bkt_poff.as<u64>(), pt->bkt_poff.as<u64>(), nb + 1);
    for (size_t c = 0; c < 65536; ++c) idx->occ_hist[c] += pt->occ_hist[c];
    idx->occ_big.insert(idx->occ_big.end(), pt->occ_big.begin(), pt->occ_big.end());
    ok += pt->n_keys; op += pt->n_p;
  }
  std::sort(idx->occ_big.begin(), idx->occ_big.end());
  SE_TRY(index_build_lookup(ctx, idx));
  ctx->timer.mark(st, "end");
  SE_CUDA(cudaStreamSynchronize(st));
  ctx->timer.finish();
  cleanup();
#undef SE_TRY
#undef SE_CUDA
  *out = idx;
  return MM2_OK;
}

// One process, several GPUs: a host thread per context builds the sharded index over a communicator of its own and every
// context gets its replica (what `mm2rs index --gpus N` and a single-process Rust caller use).
extern "C" int mm2_index_build_multi(mm2_ctx_t* const* ctxs, int nctx, const uint8_t* cat, const uint64_t* offs, const char* const* names,
                                     size_t nseq, int w, int k, int b, int flag, mm2_index_t** out) {
  if (!ctxs || nctx < 1 || !out) { mm2_set_error("mm2_index_build_multi: bad argument"); return MM2_E_ARG; }
  for (int i = 0; i < nctx; ++i) { if (!ctxs[i]) { mm2_set_error("mm2_index_build_multi: NULL context"); return MM2_E_ARG; } out[i] = nullptr; }
  if (nctx == 1) return mm2_index_build_seqs(ctxs[0], cat, offs, names, nseq, w, k, b, flag, &out[0]);
  unsigned char id[128];
  MM2_TRY(mm2_comm_get_unique_id(id));
  std::vector<int> rc((size_t)nctx, MM2_OK);
  std::vector<std::string> err((size_t)nctx);
  std::vector<std::thread> th;
  for (int i = 0; i < nctx; ++i)
    th.emplace_back([&, i]() {
      mm2_comm_t* comm = nullptr;
      rc[(size_t)i] = mm2_comm_create(ctxs[i], id, nctx, i, &comm);
      if (rc[(size_t)i] == MM2_OK) rc[(size_t)i] = mm2_index_build_sharded(ctxs[i], comm, cat, offs, names, nseq, w, k, b, flag, &out[i]);
      if (rc[(size_t)i] != MM2_OK) err[(size_t)i] = mm2_last_error();
      if (comm) mm2_comm_destroy(comm);
    });
  for (auto& t : th) t.join();
  for (int i = 0; i < nctx; ++i)
    if (rc[(size_t)i] != MM2_OK) {
      for (int j = 0; j < nctx; ++j) { if (out[j]) mm2_index_free(out[j]); out[j] = nullptr; }
      mm2_set_error("rank %d: %s", i, err[(size_t)i].c_str());
      return rc[(size_t)i];
    }
  return MM2_OK;
}

// diagnostic (mm2b200_diag.h): the byte / tile / word ranges rank `rank` of `nranks` would take in the sharded build; pure host code
extern "C" int mm2_shard_plan(const uint64_t* offs, size_t nseq, int w, int k, int flag, int nranks, int rank, uint64_t* out8) {
  if (!offs || !out8 || nranks < 1 || rank < 0 || rank >= nranks) { mm2_set_error("mm2_shard_plan: bad argument"); return MM2_E_ARG; }
  std::vector<u64> off0(nseq + 1, 0);
  for (size_t i = 0; i <= nseq && nseq; ++i) off0[i] = offs[i] - offs[0];
  const u64 total = nseq ? off0[nseq] : 0;
  ShardPlan P;
  make_plan(off0.data(), nseq, w, k, flag, total ? kroundup64((size_t)((total + 7) / 8)) : 0, nranks, rank, &P);
  out8[0] = P.tile_path ? 1 : 0; out8[1] = P.tile_path ? P.tile_lo : P.seq_lo; out8[2] = P.tile_path ? P.tile_hi : P.seq_hi;
  out8[3] = P.sk_lo; out8[4] = P.sk_hi; out8[5] = P.word_lo; out8[6] = P.word_hi; out8[7] = P.up_hi - P.up_lo;
  return MM2_OK;
}
