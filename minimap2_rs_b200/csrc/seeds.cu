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

// One read by one CTA.  INIT_KEEP: the read's keep flags are set to 1 here (list mode) instead of by a memset of the batch.
template <bool INIT_KEEP>
__device__ __forceinline__ void filter_read(const u32 r, const u64* __restrict__ mkey, const u64* __restrict__ mini_off,
                                            i32 q_occ_max, float q_occ_frac, u8* __restrict__ keep, u32* __restrict__ sum_span,
                                            unsigned char* fl_smem, u32* s_red, unsigned long long* __restrict__ n_dropped) {
  unsigned long long* hk = reinterpret_cast<unsigned long long*>(fl_smem);
  u32* hc = reinterpret_cast<u32*>(fl_smem + FL_SLOTS * 8);
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
  if (INIT_KEEP)
    for (u64 i = m0 + tid; i < m1; i += FL_NT) keep[i] = 1;
  const u32 npass = (u32)((n + FL_PER_PASS - 1) / FL_PER_PASS);
  u32 dropped = 0;
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
      if ((u64)hc[slot] > thr) { keep[i] = 0; ++dropped; }
    }
  }
  if (n_dropped) {   // work counter of the batch (n_minimizers_kept); only reads that reach the exact passes get here
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) dropped += __shfl_xor_sync(0xFFFFFFFFu, dropped, d);
    if ((tid & 31) == 0 && dropped) atomicAdd(n_dropped, (unsigned long long)dropped);
  }
}

__global__ void __launch_bounds__(FL_NT) filter_kernel(const u64* __restrict__ mkey, const u64* __restrict__ mini_off, u32 nreads,
                                                       i32 q_occ_max, float q_occ_frac, u8* __restrict__ keep,
                                                       u32* __restrict__ sum_span) {
  extern __shared__ __align__(16) unsigned char fl_smem[];
  __shared__ u32 s_red[FL_NT / 32];
  if (blockIdx.x >= nreads) return;
  filter_read<false>(blockIdx.x, mkey, mini_off, q_occ_max, q_occ_frac, keep, sum_span, fl_smem, s_red, nullptr);
}

// the exact passes for the reads seed_hits_kernel<0> put on the list (their count sketch had a bin over the threshold)
__global__ void __launch_bounds__(FL_NT) filter_list_kernel(const u64* __restrict__ mkey, const u64* __restrict__ mini_off,
                                                            const u32* __restrict__ list, const u32* __restrict__ n_list,
                                                            i32 q_occ_max, float q_occ_frac, u8* __restrict__ keep,
                                                            u32* __restrict__ sum_span, unsigned long long* __restrict__ n_dropped) {
  extern __shared__ __align__(16) unsigned char fl_smem[];
  __shared__ u32 s_red[FL_NT / 32];
  const u32 nl = *n_list;
  for (u32 j = blockIdx.x; j < nl; j += gridDim.x) {
    __syncthreads();   // the previous read's shared state is no longer in use
    filter_read<true>(list[j], mkey, mini_off, q_occ_max, q_occ_frac, keep, sum_span, fl_smem, s_red, n_dropped);
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

// ---- version 2 of the query side (default): count sketch + Index::get + hit compaction in one pass, one CTA per read --------
// 80 % of the minimizers of a 10 %-error read are absent from the index, so instead of one (count, location) pair per
// minimizer this kernel leaves a compact list of the read's hits {location, count, query position/strand, span} at the
// start of the read's minimizer range, plus the read's hit and anchor counts.  The order of a read's hits is whatever the
// shared-memory counter hands out: the anchors are sorted by their full 128-bit value afterwards (seeds.rs:58), so it is
// not observable.  L2 policy: the Bloom filter is loaded evict_last (it is the only re-used data), everything else streams.
__device__ __forceinline__ u64 l2_policy_evict_last() { u64 p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ u64 l2_policy_evict_first() { u64 p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint4 ld_hint_v4(const uint4* a, u64 pol) {
  uint4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ ulonglong2 ld_hint_v2(const ulonglong2* a, u64 pol) {
  ulonglong2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(v.x), "=l"(v.y) : "l"(a), "l"(pol));
  return v;
}

constexpr int SH_NT = 128;
struct SeedHitArgs {
  IndexView V;
  const u64* mkey; const u64* mval; const u64* mini_off;
  u32 nreads;
  i32 q_occ_max; float q_occ_frac; i32 mid_occ;
  const u8* keep;        // per minimizer; valid only inside the reads on the list (written by filter_list_kernel)
  u32* list; u32* n_list;   // reads whose count sketch asks for the exact filter (MODE 0 appends, MODE 1 consumes)
  u32* sum_span;         // MODE 0 out: paf.rs:160 sum of spans per read
  u64* hit_loc;          // Single: the occurrence itself; Multi: index of the first occurrence in V.p
  u64* hit_q;            // count << 32 | low 32 bits of the minimizer's rid_pos_strand
  u16* hit_aux;          // span | (bit 32 of rid_pos_strand) << 8 (push_anchor's qpos takes bits 1..32, seeds.rs:66)
  u32* read_nhit; u32* read_na;
  u32* err;              // set to 1 if a read has 2^32 or more anchors
};

__device__ __forceinline__ void seed_hits_read(const SeedHitArgs& G, const u32 r, const bool use_keep, u32* s_nh, unsigned long long* s_na) {
  const int tid = threadIdx.x;
  const u64 m0 = G.mini_off[r], m1 = G.mini_off[r + 1];
  const u64 bmask = (1ULL << G.V.b) - 1;
  const u64 pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
  if (tid == 0) { *s_nh = 0; *s_na = 0; }
  __syncthreads();
  u64 na = 0;
  for (u64 i = m0 + tid; i < m1; i += SH_NT) {
    if (use_keep && !G.keep[i]) continue;
    const u64 ks = __ldcs(G.mkey + i);
    const u64 minier = ks >> 8;
    bool maybe = true;
    if (G.V.bloom) {  // a cleared bit proves the key is not in the index
      const u64 h = minier * 0xD6E8FEB86659FD93ULL;
      const uint4 blk = ld_hint_v4(&G.V.bloom[(h >> 40) & G.V.bloom_mask], pol_keep);
      // bit b of the 128-bit block without an indexed array or a branch per bit (-1.5 % kernel time)
      u32 all = 1u;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const u32 b = (u32)h >> (7 * q);
        const u32 w01 = (b & 32u) ? blk.y : blk.x, w23 = (b & 32u) ? blk.w : blk.z;
        all &= ((b & 64u) ? w23 : w01) >> (b & 31u);
      }
      maybe = all & 1u;
    }
    if (!maybe) continue;
    // Index::get (index.rs:143-154): one 32-byte table line per fine bucket (see IndexView::tab)
    const u64 hk = minier >> G.V.b;
    const u64 f = ((minier & bmask) << G.V.fine_j) | (u64)index_fine_cdf(hk, G.V.R, G.V.fine_j, G.V.fine_pw);
    const ulonglong2* line = &G.V.tab[2 * f];
    ulonglong2 e0, e1;   // ONE 256-bit access for the line (two 128-bit loads cost 4 % more kernel time)
    asm volatile("ld.global.nc.L2::cache_hint.v4.u64 {%0, %1, %2, %3}, [%4], %5;" : "=l"(e0.x), "=l"(e0.y), "=l"(e1.x), "=l"(e1.y) : "l"(line), "l"(pol_stream));
    ulonglong2 hit = make_ulonglong2(TAB_EMPTY, 0);
    if ((e0.x >> 1) == hk) hit = e0;
    else if ((e1.x >> 1) == hk) hit = e1;
    else if (e1.x == TAB_MORE) {                                 // a longer run: the rest of it is in kv[], ascending
      const u32 lo = (u32)e1.y, n = (u32)(e1.y >> 32);
      for (u32 e = 0; e < n; ++e) {
        const ulonglong2 kv = ld_hint_v2(&G.V.kv[lo + e], pol_stream);
        if ((kv.x >> 1) >= hk) { if ((kv.x >> 1) == hk) hit = kv; break; }
      }
    }
    u32 cnt = 0; u64 loc = 0;
    if (hit.x != TAB_EMPTY) {
      if (hit.x & 1) { cnt = 1; loc = hit.y; }                   // Occurrences::Single (never skipped, seeds.rs:47)
      else {
        const u64 c = hit.y & 0xffffffffULL;
        if (!((i32)(u32)c > G.mid_occ)) {                        // seeds.rs:51 `slice.len() as i32 > mid_occ`
          cnt = (u32)c;
          loc = G.V.bkt_poff[minier & bmask] + (hit.y >> 32);
        }
      }
    }
    if (cnt) {
      const u64 rps = __ldcs(G.mval + i);
      const u64 o = m0 + atomicAdd(s_nh, 1u);
      G.hit_loc[o] = loc;
      G.hit_q[o] = ((u64)cnt << 32) | (rps & 0xffffffffULL);
      G.hit_aux[o] = (u16)((ks & 0xffu) | (((rps >> 32) & 1u) << 8));
      na += cnt;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) na += __shfl_xor_sync(0xFFFFFFFFu, na, d);
  if ((tid & 31) == 0 && na) atomicAdd(s_na, (unsigned long long)na);
  __syncthreads();
  if (tid == 0) {
    const u64 t = *s_na;
    G.read_nhit[r] = *s_nh;
    G.read_na[r] = t > 0xFFFFFFFFull ? 0u : (u32)t;
    if (t > 0xFFFFFFFFull) *G.err = 1u;
  }
}

// MODE 0: every read: sum of spans + count sketch (seeds.rs:13-36 can only drop a key whose sketch bin exceeds the
// threshold); reads that need the exact filter go on the list, all others are looked up right away.
// MODE 1: the reads on the list, after filter_list_kernel wrote their keep flags.
template <int MODE>
__global__ void __launch_bounds__(SH_NT, 16) seed_hits_kernel(SeedHitArgs G) {
  __shared__ u32 s_bins[MODE == 0 ? 2048 : 1];
  __shared__ u32 s_red[SH_NT / 32];
  __shared__ u32 s_nh;
  __shared__ unsigned long long s_na;
  const int tid = threadIdx.x;
  if (MODE == 1) {
    const u32 nl = *G.n_list;
    for (u32 j = blockIdx.x; j < nl; j += gridDim.x) {
      __syncthreads();
      seed_hits_read(G, G.list[j], true, &s_nh, &s_na);
    }
    return;
  }
  const u32 r = blockIdx.x;
  if (r >= G.nreads) return;
  const u64 m0 = G.mini_off[r], m1 = G.mini_off[r + 1];
  const u64 n = m1 - m0;
  for (int s = tid; s < 2048; s += SH_NT) s_bins[s] = 0;
  __syncthreads();
  u32 ss = 0;
  for (u64 i = m0 + tid; i < m1; i += SH_NT) {
    const u64 ks = G.mkey[i];
    ss += (u32)(ks & 0xff);
    atomicAdd(&s_bins[(u32)(((ks >> 8) * 0x9E3779B97F4A7C15ULL) >> 53)], 1u);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, d);
  if ((tid & 31) == 0) s_red[tid >> 5] = ss;
  __syncthreads();
  if (tid == 0) { u32 t = 0; for (int x = 0; x < SH_NT / 32; ++x) t += s_red[x]; G.sum_span[r] = t; }
  const bool may_filter = !(n == 0 || G.q_occ_frac <= 0.0f || G.q_occ_max <= 0) && !((i64)(i32)n <= (i64)G.q_occ_max);   // seeds.rs:14-15
  if (may_filter) {
    const float cf = __fmul_rn((float)n, G.q_occ_frac);                   // seeds.rs:23
    const u64 cutoff = cf <= 0.0f ? 0ull : (cf >= 18446744073709551616.0f ? ~0ull : (u64)cf);
    const u64 thr = max((u64)G.q_occ_max, cutoff);
    int over = 0;
    for (int s = tid; s < 2048; s += SH_NT) over |= (u64)s_bins[s] > thr;
    if (__syncthreads_or(over)) {
      if (tid == 0) { G.list[atomicAdd(G.n_list, 1u)] = r; G.read_nhit[r] = 0; G.read_na[r] = 0; }
      return;
    }
  }
  seed_hits_read(G, r, false, &s_nh, &s_na);
}

// seeds.rs:44-57 from the compact hit lists; reads with lo_excl < n <= hi_incl anchors (the others are filled inside
// their sort kernel).  Anchor order inside a read is arbitrary here (see above).
constexpr int AF_NT = 128;
__global__ void __launch_bounds__(AF_NT) anchor_fill_hits_kernel(IndexView V, const u64* __restrict__ hit_loc, const u64* __restrict__ hit_q,
                                                                 const u16* __restrict__ hit_aux, const u64* __restrict__ mini_off,
                                                                 const u64* __restrict__ read_off, u32 nreads,
                                                                 const u32* __restrict__ read_nhit, const u64* __restrict__ read_aoff,
                                                                 ulonglong2* __restrict__ anchors, u64 lo_excl) {
  __shared__ unsigned long long s_o;
  const u32 r = blockIdx.x;
  if (r >= nreads) return;
  const u64 a0 = read_aoff[r];
  const u64 n = read_aoff[r + 1] - a0;
  if (n <= lo_excl) return;
  const u64 m0 = mini_off[r];
  const u32 nh = read_nhit[r];
  const i32 qlen = (i32)(read_off[r + 1] - read_off[r]);
  if (threadIdx.x == 0) s_o = 0;
  __syncthreads();
  for (u32 h = threadIdx.x; h < nh; h += AF_NT) {
    const u64 loc = hit_loc[m0 + h], q = hit_q[m0 + h];
    const u32 aux = hit_aux[m0 + h];
    const u32 c = (u32)(q >> 32);
    const u64 ks = aux & 0xffu, rps = (q & 0xffffffffULL) | ((u64)(aux >> 8) << 32);
    const u64 o = a0 + atomicAdd(&s_o, (unsigned long long)c);
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

#define SIDX(i) ((i) + ((i) >> 3))
// where a sort kernel finds the hit lists when it builds the read's anchors itself (seeds.rs:44-57 fused into the sort)
struct HitSrc {
  IndexView V;
  const u64* hit_loc; const u64* hit_q; const u16* hit_aux;
  const u64* mini_off; const u64* read_off; const u32* read_nhit;
};

// record types of the merge sort: the anchor itself (128-bit lexicographic (x, y)) or an order-preserving 64-bit key
__device__ __forceinline__ bool rec_less(const ulonglong2& p, const ulonglong2& q) { return a_less(p.x, p.y, q.x, q.y); }
__device__ __forceinline__ bool rec_less(const u64& p, const u64& q) { return p < q; }
template <class T> __device__ __forceinline__ T rec_pad();
template <> __device__ __forceinline__ ulonglong2 rec_pad<ulonglong2>() { return make_ulonglong2(~0ULL, ~0ULL); }
template <> __device__ __forceinline__ u64 rec_pad<u64>() { return ~0ULL; }
template <class T> __device__ __forceinline__ void rec_cswap(T& p, T& q) {
  if (rec_less(q, p)) { const T t = p; p = q; q = t; }
}

// v[0..8) = this thread's 8 consecutive records (slots base..base+7 of m, padding included) -> the same slots of the sorted
// sequence.  The caller has made sure nobody still reads `sa`.
template <class T>
__device__ __forceinline__ void msort_core(T (&v)[8], T* sa, const int m, const int base, const bool active) {
  const T PADV = rec_pad<T>();
  if (active) {
    rec_cswap(v[0], v[1]); rec_cswap(v[2], v[3]); rec_cswap(v[4], v[5]); rec_cswap(v[6], v[7]);
    rec_cswap(v[0], v[2]); rec_cswap(v[1], v[3]); rec_cswap(v[4], v[6]); rec_cswap(v[5], v[7]);
    rec_cswap(v[1], v[2]); rec_cswap(v[5], v[6]);
    rec_cswap(v[0], v[4]); rec_cswap(v[1], v[5]); rec_cswap(v[2], v[6]); rec_cswap(v[3], v[7]);
    rec_cswap(v[2], v[4]); rec_cswap(v[3], v[5]);
    rec_cswap(v[1], v[2]); rec_cswap(v[3], v[4]); rec_cswap(v[5], v[6]);
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
      const int A0 = pair0, B0 = pair0 + R;   // runs A and B; one pad slot per 8 records keeps the wide accesses conflict-free
      const int LA = min(R, m - A0), LB = max(0, min(R, m - B0));   // the last pair of a pass may be short
      int lo = max(0, d - LB), hi = min(d, LA);   // merge path: how many of the first d outputs come from A (ties: A first)
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const T pa = sa[SIDX(A0 + mid)], pb = sa[SIDX(B0 + d - 1 - mid)];
        if (!rec_less(pb, pa)) lo = mid + 1; else hi = mid;
      }
      int ai = lo, bi = d - lo;
      T ka = ai < LA ? sa[SIDX(A0 + ai)] : PADV, kb = bi < LB ? sa[SIDX(B0 + bi)] : PADV;
      // one load per output, no divergent paths: the run that gave the output is refilled (an exhausted run keeps its stale head,
      // which the bi >= LB / ai < LA tests of the next step never pick)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const bool take_a = bi >= LB || (ai < LA && !rec_less(kb, ka));
        v[e] = take_a ? ka : kb;
        ai += take_a ? 1 : 0; bi += take_a ? 0 : 1;
        const int nidx = take_a ? A0 + ai : B0 + bi;
        const bool more = take_a ? ai < LA : bi < LB;
        const T nv = sa[SIDX(more ? nidx : pair0)];
        if (take_a) { if (more) ka = nv; } else { if (more) kb = nv; }
      }
    }
    __syncthreads();
  }
}

// The read's anchors made straight into shared memory from its compact hit list (seeds.rs:44-57).  K64: as 64-bit keys
// rev | rid (32 - qb bits) | rpos (31 bits) | qpos (qb bits), which order exactly like (x, y) when every anchor of the read
// has rpos >= 0, 0 <= qpos < 2^qb, rid < 2^(32 - qb) and the same span; returns false if some anchor does not fit.
template <int NT, bool K64>
__device__ __forceinline__ bool msort_fill(const HitSrc& H, const u32 r, void* smem, u32* s_o, const int qb, const u32 span0) {
  ulonglong2* sa = reinterpret_cast<ulonglong2*>(smem);
  u64* sk = reinterpret_cast<u64*>(smem);
  const u64 m0 = H.mini_off[r];
  const u32 nh = H.read_nhit[r];
  const i32 qlen = (i32)(H.read_off[r + 1] - H.read_off[r]);
  bool ok = true;
  auto put = [&](int o, u64 x, u64 y) {
    if (K64) {
      const u64 rid = (x >> 32) & 0x7fffffffULL;
      ok = ok && !(x & 0x80000000ULL) && (y >> 32) == (u64)span0 && ((u32)y >> qb) == 0u && (rid >> (32 - qb)) == 0;
      sk[SIDX(o)] = (x & (1ULL << 63)) | (rid << (31 + qb)) | ((x & 0x7fffffffULL) << qb) | (y & 0xffffffffULL);
    } else {
      sa[SIDX(o)] = make_ulonglong2(x, y);
    }
  };
  if (threadIdx.x == 0) *s_o = 0;
  __syncthreads();
  for (u32 h = threadIdx.x; h < nh; h += NT) {
    const u64 loc = H.hit_loc[m0 + h], q = H.hit_q[m0 + h];
    const u32 aux = H.hit_aux[m0 + h];
    const u32 c = (u32)(q >> 32);
    const u64 ks = aux & 0xffu, rps = (q & 0xffffffffULL) | ((u64)(aux >> 8) << 32);
    const int o = (int)atomicAdd(s_o, c);
    u64 x, y;
    if (c == 1) {
      make_anchor(loc, ks, rps, qlen, x, y);
      put(o, x, y);
    } else {
      for (u32 t = 0; t < c; ++t) {
        make_anchor(H.V.p[loc + t], ks, rps, qlen, x, y);
        put(o + (int)t, x, y);
      }
    }
  }
  return __syncthreads_and(ok) != 0;
}

template <int NT, bool FUSED>
__device__ __forceinline__ void anchor_msort_read(ulonglong2* __restrict__ anchors, const u64* __restrict__ read_aoff, const u32 r,
                                                  const HitSrc& H, ulonglong2* sa, u32* s_o) {
  const u64 a0 = read_aoff[r];
  const int n = (int)(read_aoff[r + 1] - a0);
  ulonglong2* a = anchors + a0;
  const int m = (n + 7) & ~7;                 // records + padding up to a multiple of 8; runs need not be a power of two
  const int base = (int)threadIdx.x * 8;
  const bool active = base < m;
  if (FUSED) {
    const i32 qlen = (i32)(H.read_off[r + 1] - H.read_off[r]);
    const int qb = max(1, 32 - __clz(max(qlen, 1)));       // qpos < qlen < 2^qb for every well-formed anchor
    const u32 span0 = H.hit_aux[H.mini_off[r]] & 0xffu;    // n >= 1, so the read has a first hit
    if (qb <= 31 && msort_fill<NT, true>(H, r, sa, s_o, qb, span0)) {
      u64* sk = reinterpret_cast<u64*>(sa);
      u64 kv[8];
      if (active) {
#pragma unroll
        for (int e = 0; e < 8; ++e) kv[e] = base + e < n ? sk[SIDX(base + e)] : ~0ULL;
      }
      __syncthreads();   // everyone has taken its 8 keys out of shared memory
      msort_core<u64>(kv, sk, m, base, active);
      if (active) {
        const u64 qmask = (1ULL << qb) - 1, rmask = (1ULL << (32 - qb)) - 1;
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (base + e < n) {
            const u64 kx = kv[e];
            a[base + e] = make_ulonglong2((kx & (1ULL << 63)) | (((kx >> (31 + qb)) & rmask) << 32) | ((kx >> qb) & 0x7fffffffULL),
                                          ((u64)span0 << 32) | (kx & qmask));
          }
      }
      return;
    }
    msort_fill<NT, false>(H, r, sa, s_o, 0, 0);   // some anchor does not fit the 64-bit key: sort the anchors themselves
  }
  ulonglong2 v[8];
  if (active) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = base + e < n ? (FUSED ? sa[SIDX(base + e)] : a[base + e]) : rec_pad<ulonglong2>();
  }
  if (FUSED) __syncthreads();   // everyone has taken its 8 records out of shared memory
  msort_core<ulonglong2>(v, sa, m, base, active);
  if (active) {
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (base + e < n) a[base + e] = v[e];
  }
}

// LIST = false: CTA r sorts read r if lo_excl < n <= hi_incl.  LIST = true: the CTAs walk a device-built list of reads.
template <int NT, bool FUSED, bool LIST>
__global__ void __launch_bounds__(NT, NT >= 512 ? 2 : 12) anchor_msort_kernel(ulonglong2* __restrict__ anchors, const u64* __restrict__ read_aoff, u32 nreads,
                                                          u32 lo_excl, u32 hi_incl, HitSrc H, const u32* __restrict__ list,
                                                          const u32* __restrict__ n_list) {
  extern __shared__ __align__(16) unsigned char as_smem[];
  ulonglong2* sa = reinterpret_cast<ulonglong2*>(as_smem);
  __shared__ u32 s_o;
  if (LIST) {
    const u32 nl = *n_list;
    for (u32 j = blockIdx.x; j < nl; j += gridDim.x) {
      __syncthreads();
      anchor_msort_read<NT, FUSED>(anchors, read_aoff, list[j], H, sa, &s_o);
    }
  } else {
    const u32 r = blockIdx.x;
    if (r >= nreads) return;
    const u64 n64 = read_aoff[r + 1] - read_aoff[r];
    if (n64 <= lo_excl || n64 > hi_incl) return;
    anchor_msort_read<NT, FUSED>(anchors, read_aoff, r, H, sa, &s_o);
  }
}
#undef SIDX

// reads by anchor count: (1024, 4096] -> list 0, (4096, 12288] -> list 1, above -> list 2 (lists of nreads entries each,
// counters in cnt[0..2]); the common class n <= 1024 needs no list
__global__ void anchor_class_kernel(const u64* __restrict__ read_aoff, u32 nreads, u32* __restrict__ lists, u32* __restrict__ cnt) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nreads) return;
  const u64 n = read_aoff[r + 1] - read_aoff[r];
  if (n <= 1024) return;
  const int c = n <= 4096 ? 0 : n <= 12288 ? 1 : 2;
  lists[(u64)c * nreads + atomicAdd(&cnt[c], 1u)] = r;
}

template <int CAP, int NT>
__device__ __forceinline__ void anchor_sort_smem_read(ulonglong2* __restrict__ anchors, const u64* __restrict__ read_aoff, const u32 r,
                                                      ulonglong2* sa) {
  const u64 a0 = read_aoff[r];
  const int n = (int)(read_aoff[r + 1] - a0);
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
template <int CAP, int NT>
__global__ void __launch_bounds__(NT) anchor_sort_smem_kernel(ulonglong2* __restrict__ anchors, const u64* __restrict__ read_aoff,
                                                              u32 nreads, u32 lo_excl, u32 hi_incl, const u32* __restrict__ list,
                                                              const u32* __restrict__ n_list) {
  extern __shared__ __align__(16) unsigned char as_smem[];
  ulonglong2* sa = reinterpret_cast<ulonglong2*>(as_smem);
  if (list) {
    const u32 nl = *n_list;
    for (u32 j = blockIdx.x; j < nl; j += gridDim.x) {
      __syncthreads();
      anchor_sort_smem_read<CAP, NT>(anchors, read_aoff, list[j], sa);
    }
    return;
  }
  const u32 r = blockIdx.x;
  if (r >= nreads) return;
  const u64 n64 = read_aoff[r + 1] - read_aoff[r];
  if (n64 <= lo_excl || n64 > hi_incl) return;
  anchor_sort_smem_read<CAP, NT>(anchors, read_aoff, r, sa);
}

// reads with more anchors than fit in shared memory: same network over global memory (L2-resident), one CTA per read
template <int NT>
__device__ __forceinline__ void anchor_sort_gmem_read(ulonglong2* __restrict__ anchors, const u64* __restrict__ read_aoff, const u32 r) {
  const u64 a0 = read_aoff[r];
  const u64 n = read_aoff[r + 1] - a0;
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
template <int NT>
__global__ void __launch_bounds__(NT) anchor_sort_gmem_kernel(ulonglong2* __restrict__ anchors, const u64* __restrict__ read_aoff,
                                                              u32 nreads, u32 lo_excl, const u32* __restrict__ list,
                                                              const u32* __restrict__ n_list) {
  if (list) {
    const u32 nl = *n_list;
    for (u32 j = blockIdx.x; j < nl; j += gridDim.x) {
      __syncthreads();
      anchor_sort_gmem_read<NT>(anchors, read_aoff, list[j]);
    }
    return;
  }
  const u32 r = blockIdx.x;
  if (r >= nreads) return;
  if (read_aoff[r + 1] - read_aoff[r] <= lo_excl) return;
  anchor_sort_gmem_read<NT>(anchors, read_aoff, r);
}

int seeds_set_attrs() {
  cudaFuncSetAttribute(filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FL_SLOTS * 12);
  cudaFuncSetAttribute(filter_list_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FL_SLOTS * 12);
  cudaFuncSetAttribute(anchor_sort_smem_kernel<12288, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12288 * 16);
  CUDA_TRY(cudaFuncSetAttribute(anchor_msort_kernel<512, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4608 * 16));
  return MM2_OK;
}

}  // namespace

int seeds_init_device() { return seeds_set_attrs(); }   // per device (see lchain_init_device)

int seeds_filter(mm2_ctx* ctx, const u64* d_mkey, const u64* d_mini_off, u32 nreads, u64 n_mini, i32 q_occ_max, float q_occ_frac,
                 u8* d_keep, u32* d_sum_span) {
  if (n_mini) CUDA_TRY(cudaMemsetAsync(d_keep, 1, n_mini, ctx->stream));
  if (nreads) MM2_LAUNCH(ctx, filter_kernel, nreads, FL_NT, FL_SLOTS * 12, d_mkey, d_mini_off, nreads, q_occ_max, q_occ_frac, d_keep, d_sum_span);
  CUDA_TRY(cudaGetLastError());
  return MM2_OK;
}

// Version 2 of the query side.  Leaves in ctx: hit lists (occ_loc / anchor_off_m / occ_cnt reused as hit_loc / hit_q /
// hit_aux), read_nhit, read_na and read_aoff (nreads + 1 anchor offsets).  full_keep: keep flags are wanted for every
// minimizer afterwards (stage dump), not only inside the reads that went through the exact filter.  Synchronises the
// stream once to return the batch's anchor count.
int seeds_hits(mm2_ctx* ctx, const IndexView& V, const u64* d_mkey, const u64* d_mval, const u64* d_mini_off, u32 nreads, u64 n_mini,
               i32 q_occ_max, float q_occ_frac, i32 mid_occ, bool full_keep, u32* d_sum_span, u64* n_anchors, u64* n_dropped) {
  cudaStream_t st = ctx->stream;
  *n_anchors = 0;
  if (n_dropped) *n_dropped = 0;
  MM2_TRY(ctx->keep.ensure(n_mini + 16));
  MM2_TRY(ctx->occ_loc.ensure((n_mini + 16) * 8));
  MM2_TRY(ctx->anchor_off_m.ensure((n_mini + 16) * 8));
  MM2_TRY(ctx->occ_cnt.ensure((n_mini + 16) * 4));
  MM2_TRY(ctx->flag_list.ensure(((u64)nreads + 16) * 4));
  MM2_TRY(ctx->read_nhit.ensure(((u64)nreads + 16) * 4));
  MM2_TRY(ctx->read_na.ensure(((u64)nreads + 16) * 4));
  MM2_TRY(ctx->read_aoff.ensure(((u64)nreads + 4) * 8));
  u64* d_aoff = ctx->read_aoff.as<u64>();
  u32* d_hdr = reinterpret_cast<u32*>(d_aoff + nreads + 1);   // {number of listed reads, error flag}, then the u64 count of
  CUDA_TRY(cudaMemsetAsync(d_hdr, 0, 16, st));                  // minimizers dropped by the exact filter; read back with the total
  if (full_keep && n_mini) CUDA_TRY(cudaMemsetAsync(ctx->keep.p, 1, n_mini, st));
  SeedHitArgs G;
  G.V = V; G.mkey = d_mkey; G.mval = d_mval; G.mini_off = d_mini_off; G.nreads = nreads;
  G.q_occ_max = q_occ_max; G.q_occ_frac = q_occ_frac; G.mid_occ = mid_occ;
  G.keep = ctx->keep.as<u8>(); G.list = ctx->flag_list.as<u32>(); G.n_list = d_hdr; G.sum_span = d_sum_span;
  G.hit_loc = ctx->occ_loc.as<u64>(); G.hit_q = ctx->anchor_off_m.as<u64>(); G.hit_aux = ctx->occ_cnt.as<u16>();
  G.read_nhit = ctx->read_nhit.as<u32>(); G.read_na = ctx->read_na.as<u32>(); G.err = d_hdr + 1;
  if (nreads) {
    MM2_LAUNCH(ctx, seed_hits_kernel<0>, nreads, SH_NT, 0, G);
    if (q_occ_frac > 0.0f && q_occ_max > 0) {   // otherwise nothing can be on the list (seeds.rs:14)
      const int lgrid = (int)std::min<u32>(nreads, (u32)ctx->n_sm * 4u);
      MM2_LAUNCH(ctx, filter_list_kernel, lgrid, FL_NT, FL_SLOTS * 12, d_mkey, d_mini_off, G.list, G.n_list, q_occ_max, q_occ_frac,
                 ctx->keep.as<u8>(), d_sum_span, reinterpret_cast<unsigned long long*>(d_aoff + nreads + 2));
      MM2_LAUNCH(ctx, seed_hits_kernel<1>, (int)std::min<u32>(nreads, (u32)ctx->n_sm * 12u), SH_NT, 0, G);
    }
  }
  MM2_TRY(scan_u32_to_u64(ctx, ctx->read_na.as<u32>(), d_aoff, nreads, true));
  MM2_TRY(ctx->pin_scalar.ensure(64));
  CUDA_TRY(cudaMemcpyAsync(ctx->pin_scalar.p, d_aoff + nreads, 24, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(mm2_stream_wait(ctx));
  const u64 total = ctx->pin_scalar.as<u64>()[0];
  const u32 err = (u32)(ctx->pin_scalar.as<u64>()[1] >> 32);
  if (err) { mm2_set_error("a read has 2^32 or more anchors"); return MM2_E_OOM; }
  *n_anchors = total;
  if (n_dropped) *n_dropped = ctx->pin_scalar.as<u64>()[2];
  CUDA_TRY(cudaGetLastError());
  return MM2_OK;
}

// Anchors of every read, sorted by (x, y) (seeds.rs:44-58), from the state seeds_hits left in ctx.  Reads with up to 4096
// anchors (all of a normal batch) are built inside their sort kernel and written once; the larger ones are filled first.
int seeds_fill_and_sort(mm2_ctx* ctx, const IndexView& V, const u64* d_mini_off, const u64* d_read_off, u32 nreads, ulonglong2* d_anchors) {
  if (!nreads) return MM2_OK;
  const u64* d_aoff = ctx->read_aoff.as<u64>();
  HitSrc H;
  H.V = V; H.hit_loc = ctx->occ_loc.as<u64>(); H.hit_q = ctx->anchor_off_m.as<u64>(); H.hit_aux = ctx->occ_cnt.as<u16>();
  H.mini_off = d_mini_off; H.read_off = d_read_off; H.read_nhit = ctx->read_nhit.as<u32>();
  MM2_TRY(ctx->read_flag.ensure((3 * (u64)nreads + 16) * 4));
  u32* d_ccnt = ctx->read_flag.as<u32>();        // 3 counters, then 3 lists of nreads entries
  u32* d_lists = d_ccnt + 4;
  CUDA_TRY(cudaMemsetAsync(d_ccnt, 0, 16, ctx->stream));
  MM2_LAUNCH(ctx, anchor_class_kernel, (nreads + 255) / 256, 256, 0, d_aoff, nreads, d_lists, d_ccnt);
  const int lgrid = (int)std::min<u32>(nreads, (u32)ctx->n_sm * 4u);
  MM2_LAUNCH(ctx, anchor_fill_hits_kernel, nreads, AF_NT, 0, V, H.hit_loc, H.hit_q, H.hit_aux, d_mini_off, d_read_off, nreads, H.read_nhit,
             d_aoff, d_anchors, (u64)4096);
  ctx->timer.mark(ctx->stream, "anchor_sort");
  MM2_LAUNCH(ctx, (anchor_msort_kernel<128, true, false>), nreads, 128, 1152 * 16, d_anchors, d_aoff, nreads, 0u, 1024u, H, (const u32*)nullptr, (const u32*)nullptr);
  MM2_LAUNCH(ctx, (anchor_msort_kernel<512, true, true>), lgrid, 512, 4608 * 16, d_anchors, d_aoff, nreads, 1024u, 4096u, H, (const u32*)d_lists, (const u32*)d_ccnt);
  MM2_LAUNCH(ctx, (anchor_sort_smem_kernel<12288, 1024>), lgrid, 1024, 12288 * 16, d_anchors, d_aoff, nreads, 4096u, 12288u,
             (const u32*)(d_lists + nreads), (const u32*)(d_ccnt + 1));
  MM2_LAUNCH(ctx, (anchor_sort_gmem_kernel<1024>), lgrid, 1024, 0, d_anchors, d_aoff, nreads, 12288u, (const u32*)(d_lists + 2 * (u64)nreads),
             (const u32*)(d_ccnt + 2));
  CUDA_TRY(cudaGetLastError());
  return MM2_OK;
}
