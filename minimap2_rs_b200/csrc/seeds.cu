// seeds.cu — query-side seeding on sm_100a for a whole batch of reads.
// Replaces seeds.rs:13-36 (filter_query_minimizers), :42-60 (build_anchors_filtered: Index::get per minimizer,
// mid_occ cut, sort by (x,y)) and :62-79 (push_anchor, including its i32->u64 sign extension).
#include "mm2_internal.cuh"

#include <algorithm>

#include "stages.cuh"

namespace {

// ---- seeds.rs:13-36: drop every minimizer whose key occurs > max(q_occ_max, len*q_occ_frac) times in its read -------
constexpr int FL_NT = 128;
constexpr int FL_SLOTS = 4096;          // shared-memory hash table (keys u64 + counts u32 = 48 KB)
constexpr int FL_PER_PASS = FL_SLOTS / 4;  // expected distinct keys per pass (load <= 0.25)

__global__ void __launch_bounds__(FL_NT) filter_kernel(const u64* __restrict__ mkey, const u64* __restrict__ mini_off, u32 nreads,
                                                       i32 q_occ_max, float q_occ_frac, u8* __restrict__ keep,
                                                       u32* __restrict__ sum_span) {
  extern __shared__ __align__(16) unsigned char fl_smem[];
  unsigned long long* hk = reinterpret_cast<unsigned long long*>(fl_smem);
  u32* hc = reinterpret_cast<u32*>(fl_smem + FL_SLOTS * 8);
  __shared__ u32 s_red[FL_NT / 32];
  const u32 r = blockIdx.x;
  if (r >= nreads) return;
  const u64 m0 = mini_off[r], m1 = mini_off[r + 1];
  const u64 n = m1 - m0;
  const int tid = threadIdx.x;
  // One streaming pass: sum of spans (paf.rs:160 sum_k) and a 2048-bin count sketch of the keys.  A key can only be
  // dropped if it occurs > thr times, and then its bin holds > thr as well, so when no bin exceeds thr (the normal
  // case: ~1 key per bin) the exact hash-table passes below are skipped.
  u32* bins = reinterpret_cast<u32*>(fl_smem);
  for (int s = tid; s < 2048; s += FL_NT) bins[s] = 0;
  __syncthreads();
  u32 ss = 0;
  for (u64 i = m0 + tid; i < m1; i += FL_NT) {
    const u64 ks = mkey[i];
    ss += (u32)(ks & 0xff);
    atomicAdd(&bins[(u32)(((ks >> 8) * 0x9E3779B97F4A7C15ULL) >> 53)], 1u);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, d);
  if ((tid & 31) == 0) s_red[tid >> 5] = ss;
  __syncthreads();
  if (tid == 0) { u32 t = 0; for (int x = 0; x < FL_NT / 32; ++x) t += s_red[x]; sum_span[r] = t; }
  if (n == 0 || q_occ_frac <= 0.0f || q_occ_max <= 0) return;          // seeds.rs:14
  if ((i64)(i32)n <= (i64)q_occ_max) return;                            // seeds.rs:15 (`len as i32`)
  const float cf = __fmul_rn((float)n, q_occ_frac);                     // seeds.rs:23
  const u64 cutoff = cf <= 0.0f ? 0ull : (cf >= 18446744073709551616.0f ? ~0ull : (u64)cf);
  const u64 thr = max((u64)q_occ_max, cutoff);                          // cnt > q_occ_max && cnt > cutoff
  {
    int over = 0;
    for (int s = tid; s < 2048; s += FL_NT) over |= (u64)bins[s] > thr;
    if (!__syncthreads_or(over)) return;
  }
  const u32 npass = (u32)((n + FL_PER_PASS - 1) / FL_PER_PASS);
  for (u32 pass = 0; pass < npass; ++pass) {
    __syncthreads();
    for (int s = tid; s < FL_SLOTS; s += FL_NT) { hk[s] = ~0ULL; hc[s] = 0; }
    __syncthreads();
    for (u64 i = m0 + tid; i < m1; i += FL_NT) {
      const u64 key = mkey[i] >> 8;
      const u64 h = key * 0x9E3779B97F4A7C15ULL;
      if ((u32)((h >> 40) % npass) != pass) continue;
      u32 slot = (u32)(h >> 20) & (FL_SLOTS - 1);
      for (;;) {
        const unsigned long long old = atomicCAS(&hk[slot], ~0ULL, (unsigned long long)key);
        if (old == ~0ULL || old == key) { atomicAdd(&hc[slot], 1u); break; }
        slot = (slot + 1) & (FL_SLOTS - 1);
      }
    }
    __syncthreads();
    for (u64 i = m0 + tid; i < m1; i += FL_NT) {
      const u64 key = mkey[i] >> 8;
      const u64 h = key * 0x9E3779B97F4A7C15ULL;
      if ((u32)((h >> 40) % npass) != pass) continue;
      u32 slot = (u32)(h >> 20) & (FL_SLOTS - 1);
      while (hk[slot] != key) slot = (slot + 1) & (FL_SLOTS - 1);
      if ((u64)hc[slot] > thr) keep[i] = 0;
    }
  }
}

// ---- index.rs:143-154 Index::get for every kept minimizer: occurrence count + where the occurrences live -------------
__device__ __forceinline__ u64 tab_hash(u64 minier) {
  u64 x = minier * 0x9E3779B97F4A7C15ULL;
  return x ^ (x >> 29);
}

__global__ void lookup_count_kernel(IndexView V, const u64* __restrict__ mkey, const u8* __restrict__ keep, u64 n, i32 mid_occ,
                                    u32* __restrict__ occ_cnt, u64* __restrict__ occ_loc) {
  const u64 bmask = (1ULL << V.b) - 1;
  for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    u32 cnt = 0; u64 loc = 0;
    if (keep[i]) {
      const u64 minier = mkey[i] >> 8;
      bool maybe = true;
      if (V.bloom) {  // L2-resident pre-filter: a cleared bit proves the key is not in the index
        const u64 h = minier * 0xD6E8FEB86659FD93ULL;
        const uint4 blk = __ldg(&V.bloom[(h >> 40) & V.bloom_mask]);
        const u32 wd[4] = {blk.x, blk.y, blk.z, blk.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) { const u32 b = (u32)(h >> (7 * q)) & 127u; maybe = maybe && ((wd[b >> 5] >> (b & 31)) & 1u); }
      }
      u64 slot = tab_hash(minier) & V.tab_mask;
      while (maybe) {
        const ulonglong2 e = __ldg(&V.tab[slot]);
        if (e.x == ~0ULL) break;
        if ((e.x >> 1) == minier) {
          if (e.x & 1) { cnt = 1; loc = e.y; }                     // Occurrences::Single
          else {
            const u64 c = e.y & 0xffffffffULL;
            if (!((i32)(u32)c > mid_occ)) {                        // seeds.rs:51 `slice.len() as i32 > mid_occ`
              cnt = (u32)c;
              loc = V.bkt_poff[minier & bmask] + (e.y >> 32);
            }
          }
          break;
        }
        slot = (slot + 1) & V.tab_mask;
      }
    }
    occ_cnt[i] = cnt;
    occ_loc[i] = loc;
  }
}

// ---- seeds.rs:62-79 push_anchor ---------------------------------------------------------------------------------------
__device__ __forceinline__ void make_anchor(u64 r, u64 key_span, u64 rps, i32 qlen, u64& x, u64& y) {
  const u64 rid = (r >> 32) & 0xffffffffULL;
  const i32 rpos = (i32)(u32)((r >> 1) & 0xffffffffULL);
  const i32 rstrand = (i32)(r & 1);
  const i32 qpos = (i32)(u32)((rps >> 1) & 0xffffffffULL);
  const i32 qstrand = (i32)(rps & 1);
  const i32 qspan = (i32)(key_span & 0xff);
  const u64 rpos64 = (u64)(i64)rpos;  // `rpos as u64` sign-extends (F5)
  if (rstrand == qstrand) {
    x = (rid << 32) | rpos64;
    y = ((u64)(i64)qspan << 32) | (u64)(i64)qpos;
  } else {
    x = (1ULL << 63) | (rid << 32) | rpos64;
    const i32 qp = (i32)((u32)qlen - ((u32)qpos + 1u - (u32)qspan) - 1u);
    y = ((u64)(i64)qspan << 32) | (u64)(i64)qp;
  }
}

constexpr int AF_NT = 128;
__global__ void __launch_bounds__(AF_NT) anchor_fill_kernel(IndexView V, const u64* __restrict__ mkey, const u64* __restrict__ mval,
                                                            const u64* __restrict__ mini_off, const u64* __restrict__ read_off,
                                                            u32 nreads, const u32* __restrict__ occ_cnt,
                                                            const u64* __restrict__ occ_loc, const u64* __restrict__ aoff,
                                                            ulonglong2* __restrict__ anchors, u64* __restrict__ read_aoff) {
  const u32 r = blockIdx.x;
  if (r >= nreads) return;
  const u64 m0 = mini_off[r], m1 = mini_off[r + 1];
  const i32 qlen = (i32)(read_off[r + 1] - read_off[r]);
  if (threadIdx.x == 0) {
    read_aoff[r] = aoff[m0];
    if (r == nreads - 1) read_aoff[nreads] = aoff[m1];
  }
  for (u64 i = m0 + threadIdx.x; i < m1; i += AF_NT) {
    const u32 c = occ_cnt[i];
    if (!c) continue;
    const u64 ks = mkey[i], rps = mval[i], loc = occ_loc[i];
    u64 o = aoff[i];
    if (c == 1) {
      u64 x, y;
      make_anchor(loc, ks, rps, qlen, x, y);
      anchors[o] = make_ulonglong2(x, y);
    } else {
      for (u32 t = 0; t < c; ++t) {
        u64 x, y;
        make_anchor(V.p[loc + t], ks, rps, qlen, x, y);
        anchors[o + t] = make_ulonglong2(x, y);
      }
    }
  }
}

// ---- seeds.rs:58: per-read sort by the unsigned 128-bit (x, y) -----------------------------------------------------------
// Bitonic network in the "all comparators ascending" form (first sub-step of a stage pairs i with its mirror inside the
// block, later sub-steps pair i with i^j); elements past n behave as +inf and never move, so any n works unpadded.
__device__ __forceinline__ bool a_less(u64 x1, u64 y1, u64 x2, u64 y2) { return x1 < x2 || (x1 == x2 && y1 < y2); }

// Merge sort for the common sizes (n <= 8 * NT): every thread sorts 8 consecutive anchors in registers (19-comparator
// network), then log2(m / 8) merge passes over shared memory: a thread finds its 8 outputs of the pair of runs it sits in by
// a merge-path binary search and merges them serially.  Moves ~2.5x fewer bytes through shared memory than the bitonic
// network below, which was bound by it (profiles/: short-scoreboard and MIO-throttle stalls).  Slots past n hold the
// all-ones record; a real record with that bit pattern is identical to the padding, so the first n outputs are right.
__device__ __forceinline__ void a_cswap(ulonglong2& p, ulonglong2& q) {
  if (a_less(q.x, q.y, p.x, p.y)) { const ulonglong2 t = p; p = q; q = t; }
}

#define SIDX(i) ((i) + ((i) >> 3))
template <int NT>
__global__ void __launch_bounds__(NT) anchor_msort_kernel(ulonglong2* __restrict__ anchors, const u64* __restrict__ read_aoff, u32 nreads,
                                                          u32 lo_excl, u32 hi_incl) {
  extern __shared__ __align__(16) unsigned char as_smem[];
  ulonglong2* sa = reinterpret_cast<ulonglong2*>(as_smem);
  const u32 r = blockIdx.x;
  if (r >= nreads) return;
  const u64 a0 = read_aoff[r];
  const u64 n64 = read_aoff[r + 1] - a0;
  if (n64 <= lo_excl || n64 > hi_incl) return;
  const int n = (int)n64;
  ulonglong2* a = anchors + a0;
  const int m = (n + 7) & ~7;                 // records + padding up to a multiple of 8; runs need not be a power of two
  const int base = (int)threadIdx.x * 8;
  const bool active = base < m;
  const ulonglong2 PADV = make_ulonglong2(~0ULL, ~0ULL);
  ulonglong2 v[8];
  if (active) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = base + e < n ? a[base + e] : PADV;
    a_cswap(v[0], v[1]); a_cswap(v[2], v[3]); a_cswap(v[4], v[5]); a_cswap(v[6], v[7]);
    a_cswap(v[0], v[2]); a_cswap(v[1], v[3]); a_cswap(v[4], v[6]); a_cswap(v[5], v[7]);
    a_cswap(v[1], v[2]); a_cswap(v[5], v[6]);
    a_cswap(v[0], v[4]); a_cswap(v[1], v[5]); a_cswap(v[2], v[6]); a_cswap(v[3], v[7]);
    a_cswap(v[2], v[4]); a_cswap(v[3], v[5]);
    a_cswap(v[1], v[2]); a_cswap(v[3], v[4]); a_cswap(v[5], v[6]);
  }
  for (int R = 8; R < m; R <<= 1) {
    if (active) {
#pragma unroll
      for (int e = 0; e < 8; ++e) sa[SIDX(base + e)] = v[e];
    }
    __syncthreads();
    if (active) {
      const int pair0 = base & ~(2 * R - 1);
      const int d = base - pair0;                 // this thread's outputs are [d, d + 8) of the merged pair
      const int A0 = pair0, B0 = pair0 + R;   // runs A and B; one pad slot per 8 records keeps 128-bit accesses conflict-free
      const int LA = min(R, m - A0), LB = max(0, min(R, m - B0));   // the last pair of a pass may be short
      int lo = max(0, d - LB), hi = min(d, LA);   // merge path: how many of the first d outputs come from A (ties: A first)
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const ulonglong2 pa = sa[SIDX(A0 + mid)], pb = sa[SIDX(B0 + d - 1 - mid)];
        if (!a_less(pb.x, pb.y, pa.x, pa.y)) lo = mid + 1; else hi = mid;
      }
      int ai = lo, bi = d - lo;
      ulonglong2 ka = ai < LA ? sa[SIDX(A0 + ai)] : PADV, kb = bi < LB ? sa[SIDX(B0 + bi)] : PADV;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const bool take_a = bi >= LB || (ai < LA && !a_less(kb.x, kb.y, ka.x, ka.y));
        v[e] = take_a ? ka : kb;
        if (take_a) { ++ai; if (ai < LA) ka = sa[SIDX(A0 + ai)]; } else { ++bi; if (bi < LB) kb = sa[SIDX(B0 + bi)]; }
      }
    }
    __syncthreads();
  }
  if (active) {
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (base + e < n) a[base + e] = v[e];
  }
}
#undef SIDX

template <int CAP, int NT>
__global__ void __launch_bounds__(NT) anchor_sort_smem_kernel(ulonglong2* __restrict__ anchors, const u64* __restrict__ read_aoff,
                                                              u32 nreads, u32 lo_excl, u32 hi_incl) {
  extern __shared__ __align__(16) unsigned char as_smem[];
  ulonglong2* sa = reinterpret_cast<ulonglong2*>(as_smem);
  const u32 r = blockIdx.x;
  if (r >= nreads) return;
  const u64 a0 = read_aoff[r];
  const u64 n64 = read_aoff[r + 1] - a0;
  if (n64 <= lo_excl || n64 > hi_incl) return;
  const int n = (int)n64;
  ulonglong2* a = anchors + a0;
  for (int i = threadIdx.x; i < n; i += NT) sa[i] = a[i];
  __syncthreads();
  int m = 2;
  while (m < n) m <<= 1;          // network size (elements past n behave as +inf and never move)
  const int half = m >> 1;
  // every thread owns comparators, not elements: comparator t of a sub-step with distance j touches
  // i = (t / j) * 2j + (t % j) and its partner, so no thread idles on the "upper" element of a pair
  for (int k = 2; k <= m; k <<= 1) {
    const int hk = k >> 1;
    for (int t = threadIdx.x; t < half; t += NT) {   // first sub-step of the stage: partner = mirror inside the block of k
      const int i = ((t & ~(hk - 1)) << 1) | (t & (hk - 1));
      const int p = i ^ (k - 1);
      if (p < n) {
        const ulonglong2 v1 = sa[i], v2 = sa[p];
        if (a_less(v2.x, v2.y, v1.x, v1.y)) { sa[i] = v2; sa[p] = v1; }
      }
    }
    __syncthreads();
    for (int j = k >> 2; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < half; t += NT) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int p = i + j;
        if (p < n) {
          const ulonglong2 v1 = sa[i], v2 = sa[p];
          if (a_less(v2.x, v2.y, v1.x, v1.y)) { sa[i] = v2; sa[p] = v1; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n; i += NT) a[i] = sa[i];
}

// reads with more anchors than fit in shared memory: same network over global memory (L2-resident), one CTA per read
template <int NT>
__global__ void __launch_bounds__(NT) anchor_sort_gmem_kernel(ulonglong2* __restrict__ anchors, const u64* __restrict__ read_aoff,
                                                              u32 nreads, u32 lo_excl) {
  const u32 r = blockIdx.x;
  if (r >= nreads) return;
  const u64 a0 = read_aoff[r];
  const u64 n = read_aoff[r + 1] - a0;
  if (n <= lo_excl) return;
  ulonglong2* a = anchors + a0;
  for (u64 k = 2; (k >> 1) < n; k <<= 1) {
    for (u64 i = threadIdx.x; i < n; i += NT) {
      const u64 p = i ^ (k - 1);
      if (p > i && p < n) {
        const ulonglong2 v1 = a[i], v2 = a[p];
        if (a_less(v2.x, v2.y, v1.x, v1.y)) { a[i] = v2; a[p] = v1; }
      }
    }
    __syncthreads();
    for (u64 j = k >> 2; j > 0; j >>= 1) {
      for (u64 i = threadIdx.x; i < n; i += NT) {
        const u64 p = i ^ j;
        if (p > i && p < n) {
          const ulonglong2 v1 = a[i], v2 = a[p];
          if (a_less(v2.x, v2.y, v1.x, v1.y)) { a[i] = v2; a[p] = v1; }
        }
      }
      __syncthreads();
    }
  }
}

inline int grid_for(u64 n, int block = 256) { return (int)std::max<u64>(1, std::min<u64>((n + block - 1) / block, 148ull * 32)); }
bool g_attr_done = false;

}  // namespace

int seeds_filter(mm2_ctx* ctx, const u64* d_mkey, const u64* d_mini_off, u32 nreads, u64 n_mini, i32 q_occ_max, float q_occ_frac,
                 u8* d_keep, u32* d_sum_span) {
  if (!g_attr_done) {
    cudaFuncSetAttribute(filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FL_SLOTS * 12);
    cudaFuncSetAttribute(anchor_sort_smem_kernel<1024, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * 16);
    cudaFuncSetAttribute(anchor_sort_smem_kernel<4096, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 16);
    cudaFuncSetAttribute(anchor_sort_smem_kernel<12288, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12288 * 16);
    cudaFuncSetAttribute(anchor_msort_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4608 * 16);
    g_attr_done = true;
  }
  if (n_mini) CUDA_TRY(cudaMemsetAsync(d_keep, 1, n_mini, ctx->stream));
  if (nreads) MM2_LAUNCH(ctx, filter_kernel, nreads, FL_NT, FL_SLOTS * 12, d_mkey, d_mini_off, nreads, q_occ_max, q_occ_frac, d_keep, d_sum_span);
  CUDA_TRY(cudaGetLastError());
  return MM2_OK;
}

int seeds_lookup_count(mm2_ctx* ctx, const IndexView& V, const u64* d_mkey, const u8* d_keep, u64 n_mini, i32 mid_occ, u32* d_cnt,
                       u64* d_loc) {
  if (n_mini) MM2_LAUNCH(ctx, lookup_count_kernel, grid_for(n_mini), 256, 0, V, d_mkey, d_keep, n_mini, mid_occ, d_cnt, d_loc);
  CUDA_TRY(cudaGetLastError());
  return MM2_OK;
}

int seeds_fill_and_sort(mm2_ctx* ctx, const IndexView& V, const u64* d_mkey, const u64* d_mval, const u64* d_mini_off,
                        const u64* d_read_off, u32 nreads, const u32* d_cnt, const u64* d_loc, const u64* d_aoff,
                        ulonglong2* d_anchors, u64* d_read_aoff) {
  if (!nreads) return MM2_OK;
  MM2_LAUNCH(ctx, anchor_fill_kernel, nreads, AF_NT, 0, V, d_mkey, d_mval, d_mini_off, d_read_off, nreads, d_cnt, d_loc, d_aoff,
             d_anchors, d_read_aoff);
  ctx->timer.mark(ctx->stream, "anchor_sort");
  static const bool use_bitonic = [] { const char* e = getenv("MM2_ANCHOR_SORT"); return e && !strcmp(e, "bitonic"); }();   // comparison arm
  if (use_bitonic) {
    MM2_LAUNCH(ctx, (anchor_sort_smem_kernel<1024, 128>), nreads, 128, 1024 * 16, d_anchors, d_read_aoff, nreads, 1u, 1024u);
    MM2_LAUNCH(ctx, (anchor_sort_smem_kernel<4096, 256>), nreads, 256, 4096 * 16, d_anchors, d_read_aoff, nreads, 1024u, 4096u);
  } else {
    MM2_LAUNCH(ctx, (anchor_msort_kernel<128>), nreads, 128, 1152 * 16, d_anchors, d_read_aoff, nreads, 1u, 1024u);
    MM2_LAUNCH(ctx, (anchor_msort_kernel<512>), nreads, 512, 4608 * 16, d_anchors, d_read_aoff, nreads, 1024u, 4096u);
  }
  MM2_LAUNCH(ctx, (anchor_sort_smem_kernel<12288, 1024>), nreads, 1024, 12288 * 16, d_anchors, d_read_aoff, nreads, 4096u, 12288u);
  MM2_LAUNCH(ctx, (anchor_sort_gmem_kernel<1024>), nreads, 1024, 0, d_anchors, d_read_aoff, nreads, 12288u);
  CUDA_TRY(cudaGetLastError());
  return MM2_OK;
}
