// sketch.cu — minimizer sketching on sm_100a.  Replaces sketch.rs:29-100 (sketch_sequence) for whole batches of
// sequences (a genome's chromosomes, or the reads of a mapping batch) in one launch.
//
// Two kernels:
//  * sketch_tile_kernel  — odd k, non-HPC (every BASELINE config).  Position-parallel restatement of the reference's
//    sliding-window state machine: the (min, min_pos) state after step i is a pure function of the last w infos
//    (newest of the equal minima, sketch.rs:84,90-91), and every emission site of sketch.rs:80-96 is gated by
//    l (consecutive valid bases, saturating at w+k) and by how the window minimum changed between step i-1 and i.
//    A CTA owns a tile of 2048-w consecutive steps of one sequence; per-tile output counts are turned into global
//    offsets by a single-pass decoupled look-back, so the minimizers land in exactly the reference's order.
//    tests/models.py:sketch_model is the CPU model of this restatement (checked against the oracle).
//  * sketch_literal_kernel — any k (even k has palindromic k-mers that stall `l`, sketch.rs:67-69) and HPC mode
//    (sketch.rs:51-61): one thread per sequence runs the reference state machine literally.  Correct for every
//    parameter set, slow for long sequences; used only where the tile kernel does not apply.
#include "mm2_internal.cuh"

#include <algorithm>

namespace {

#ifndef MM2_SK_NT
#define MM2_SK_NT 256
#endif
constexpr int SK_NT = MM2_SK_NT;            // threads per CTA
constexpr int SK_CH = 8;                    // key positions per thread
constexpr int SK_REGION = SK_NT * SK_CH;    // key positions held in shared memory: [s-w, s-w+2048)
constexpr int SK_MAXCHUNK = ((15 + SK_REGION + 255 + 28 + 15 + 15) / 16 + 7) / 8 * 8;   // 16-byte chunks a tile can touch (+ pad)

struct SketchParams {
  const u8* seq;          // concatenated ASCII bases
  const u64* seq_off;     // nseq+1
  u64 buf_len;            // readable bytes in seq
  const u32* tile_seq;    // ntiles: sequence of each tile
  const u64* tile_first;  // nseq+1: first tile of each sequence
  u32 nseq, ntiles;
  int w, k;
  int vec_ok;             // seq is 16-byte aligned
  u32 rid_base, rid_step;
  u64* out_key; u64* out_val; u64 out_cap;
  u64* seq_out_off;       // nseq+1
  u64* tile_status;       // ntiles, zero-initialised; [63:62] 1 = aggregate, 2 = inclusive prefix
  u32* ticket;            // zero-initialised
  u32 tile_base;          // first tile of this launch (tickets count from it)
};

// invertible integer mix of sketch.rs:4-13, in the narrowest type that holds 2k bits
template <class KT>
__device__ __forceinline__ KT hash_mix(KT key, KT mask) {
  key = (~key + (key << 21)) & mask;
  key ^= key >> 24;
  key = (key + (key << 3) + (key << 8)) & mask;
  key ^= key >> 14;
  key = (key + (key << 2) + (key << 4)) & mask;
  key ^= key >> 28;
  if (sizeof(KT) == 8) key = (key + (key << 31)) & mask;  // for 2k <= 31 the shifted term is masked away
  return key;
}
template <>
__device__ __forceinline__ u32 hash_mix<u32>(u32 key, u32 mask) {
  key = (~key + (key << 21)) & mask;
  key ^= key >> 24;
  key = (key + (key << 3) + (key << 8)) & mask;
  key ^= key >> 14;
  key = (key + (key << 2) + (key << 4)) & mask;
  key ^= key >> 28;
  // key + (key << 31): bit 31 is outside every mask of <= 31 bits
  return key & mask;
}

// consecutive non-N positions ending at raw index r (0 if r itself is N), saturated at cap
__device__ __forceinline__ int run_len_at(const u32* nm, int r, int cap) {
  int wi = r >> 5, bi = r & 31;
  u32 m = nm[wi] & (0xFFFFFFFFu >> (31 - bi));
  if (m) return min(cap, bi - (31 - __clz(m)));
  int l = bi + 1;
  while (l < cap && wi > 0) {
    --wi;
    m = nm[wi];
    if (m) return min(cap, l + __clz(m));
    l += 32;
  }
  return min(l, cap);
}

template <class KT> struct KeyTraits;
template <> struct KeyTraits<u32> { static constexpr int PAD = 5; };
template <> struct KeyTraits<u64> { static constexpr int PAD = 4; };

// ASCII word (4 bases) -> 4 nt4 codes (2 bits each, in bytes) + per-byte validity mask (0xFF = ACGTacgt)
__device__ __forceinline__ void nt4x4(u32 wd, u32& code4, u32& vmask) {
  u32 u = wd & 0xDFDFDFDFu;
  vmask = __vcmpeq4(u, 0x41414141u) | __vcmpeq4(u, 0x43434343u) | __vcmpeq4(u, 0x47474747u) | __vcmpeq4(u, 0x54545454u);
  u32 x = (wd >> 1) & 0x03030303u;          // A0 C1 G3 T2
  code4 = x ^ ((x >> 1) & 0x01010101u);     // A0 C1 G2 T3 (nt4.rs:2-10)
}
__device__ __forceinline__ u32 pack4(u32 c) {  // 4 code bytes -> 8 bits
  return (c & 3u) | ((c >> 6) & 0xCu) | ((c >> 12) & 0x30u) | ((c >> 18) & 0xC0u);
}
__device__ __forceinline__ u32 nbits4(u32 vmask) {  // 4 validity bytes -> 4 bits, 1 = not ACGT
  u32 t = (~vmask) & 0x01010101u;
  return (t | (t >> 7) | (t >> 14) | (t >> 21)) & 0xFu;
}

template <class KT>
__global__ void __launch_bounds__(SK_NT) sketch_tile_kernel(SketchParams P) {
  constexpr int PAD = KeyTraits<KT>::PAD;
  constexpr KT KMAX = (KT)~(KT)0;
#define KIDX(u) ((u) + ((u) >> PAD))
  __shared__ __align__(16) KT s_key[SK_REGION + (SK_REGION >> PAD) + 8];
  __shared__ __align__(16) u32 s_pack[SK_MAXCHUNK + 4];
  __shared__ __align__(16) u32 s_nm[SK_MAXCHUNK / 2 + 4];
  __shared__ __align__(16) u16 s_ne[SK_REGION];
  __shared__ __align__(16) u32 s_off[SK_REGION];
  __shared__ u16 s_arg[SK_REGION + 1];
  __shared__ u8 s_cnt[SK_REGION + 1];
  __shared__ u8 s_flag[SK_REGION];
  __shared__ u8 s_z[SK_NT];
  __shared__ u32 s_wsum[SK_NT / 32];
  __shared__ u32 s_tile;
  __shared__ u64 s_base;

  const int tid = threadIdx.x;
  const int w = P.w, k = P.k;
  const int cap = w + k;
  const int T = SK_REGION - w;  // steps per tile
  const KT mask = (KT)((((u64)1) << (2 * k)) - 1);
  const int shift1 = 2 * (k - 1);

  for (;;) {
    __syncthreads();  // protects s_tile / all tile state of the previous iteration
    if (tid == 0) s_tile = P.tile_base + atomicAdd(P.ticket, 1u);
    __syncthreads();
    const u32 tile = s_tile;
    if (tile >= P.ntiles) break;
    const u32 q = P.tile_seq[tile];
    const u64 soff = P.seq_off[q];
    const i64 len = (i64)(P.seq_off[q + 1] - soff);
    const i64 s = (i64)(tile - P.tile_first[q]) * T;  // first step of this tile
    const i64 e = min(len, s + (i64)T);               // one past the last step
    const int nsteps = (int)(e - s);                  // may be 0 (empty sequence)
    const i64 P0 = s - w;                             // position of key-space index 0
    const i64 a = P0 - cap;                           // first loaded position
    const i64 gidx = (i64)soff + a;                   // its byte index in P.seq (may be negative)
    const i64 g0 = (gidx >> 4) << 4;
    const int delta = (int)(gidx - g0);
    const int nchunks = (delta + SK_REGION + cap + 15) >> 4;

    // ---- phase 1: 128-bit loads -> 2-bit packed codes + N mask -------------------------------------------------
    for (int c = tid; c < SK_MAXCHUNK + 4; c += SK_NT) {
      u32 packed = 0, nmask = 0xFFFFu;
      if (c < nchunks) {
        const i64 gi = g0 + 16 * (i64)c;
        u32 wd[4] = {0, 0, 0, 0};
        if (P.vec_ok && gi >= 0 && gi + 16 <= (i64)P.buf_len) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(P.seq + gi));
          wd[0] = v.x; wd[1] = v.y; wd[2] = v.z; wd[3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const i64 g = gi + j;
            u32 b = (g >= 0 && g < (i64)P.buf_len) ? (u32)P.seq[g] : 0u;
            wd[j >> 2] |= b << (8 * (j & 3));
          }
        }
        nmask = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          u32 c4, vm;
          nt4x4(wd[j], c4, vm);
          packed |= pack4(c4) << (8 * j);
          nmask |= nbits4(vm) << (4 * j);
        }
        // bytes outside [0, len) of this sequence count as N
        const i64 pstart = a + 16 * (i64)c - delta;
        const i64 lo = max((i64)0, -pstart), hi = min((i64)16, len - pstart);
        u32 inseq = 0;
        if (hi > lo) inseq = ((hi >= 16 ? 0x10000u : (1u << hi)) - 1u) & ~((1u << lo) - 1u);
        nmask = (nmask | ~inseq) & 0xFFFFu;
      }
      if (c < SK_MAXCHUNK + 4) s_pack[c] = packed;
      reinterpret_cast<u16*>(s_nm)[c] = (u16)nmask;
    }
    __syncthreads();

    // ---- phase 2: each thread rolls 8 consecutive k-mers -> hashed keys ---------------------------------------------
    {
      const int r0 = cap + SK_CH * tid + delta;  // raw index of this thread's first key position
      int l = run_len_at(s_nm, r0 - 1, cap);
      // k-mer ending at raw index r0-1, little-endian digits (earliest base lowest)
      const int rs = r0 - k;
      const int wi = rs >> 4, sh = 2 * (rs & 15);
      const u32 w0 = s_pack[wi], w1 = s_pack[wi + 1], w2 = s_pack[wi + 2];
      const u32 flo = __funnelshift_r(w0, w1, sh), fhi = __funnelshift_r(w1, w2, sh);
      const u64 field = (((u64)fhi << 32) | flo) & (u64)mask;
      KT rev = (KT)((~field) & (u64)mask);                       // sketch.rs:66 kmer[1]
      u64 br = __brevll(field);
      br = ((br & 0x5555555555555555ULL) << 1) | ((br >> 1) & 0x5555555555555555ULL);
      KT fwd = (KT)(br >> (64 - 2 * k));                         // sketch.rs:65 kmer[0]
      const u32 cw = __funnelshift_r(s_pack[r0 >> 4], s_pack[(r0 >> 4) + 1], 2 * (r0 & 15));
      const u32 nb = __funnelshift_r(s_nm[r0 >> 5], s_nm[(r0 >> 5) + 1], r0 & 31);
      u32 zbits = 0;
#pragma unroll
      for (int j = 0; j < SK_CH; ++j) {
        const u32 c = (cw >> (2 * j)) & 3u;
        l = ((nb >> j) & 1u) ? 0 : min(l + 1, cap);
        fwd = (KT)(((fwd << 2) | (KT)c) & mask);
        rev = (KT)((rev >> 2) | ((KT)(3u ^ c) << shift1));
        const bool z = !(fwd < rev);
        KT key = KMAX;
        if (l >= k) key = hash_mix<KT>(z ? rev : fwd, mask);
        const int u = SK_CH * tid + j;
        s_key[KIDX(u)] = key;
        zbits |= (u32)z << j;
      }
      s_z[tid] = (u8)zbits;
    }
    __syncthreads();

    // ---- phase 3: newest argmin + multiplicity of the minimum over every window [j-w+1, j], j = s-1 .. e-1 -------------
    for (int qi = tid; qi <= nsteps; qi += SK_NT) {
      KT mk = KMAX; int mp = qi, cnt = 0;
      for (int u = qi; u < qi + w; ++u) {
        const KT kk = s_key[KIDX(u)];
        if (kk <= mk) { cnt = (kk == mk) ? cnt + 1 : 1; mk = kk; mp = u; }
      }
      s_arg[qi] = (u16)mp;
      s_cnt[qi] = (u8)min(cnt, 255);
    }
    __syncthreads();

    // ---- phase 4a: emissions of step i = s + qi (sketch.rs:80-96) -----------------------------------------------------
    for (int qi = tid; qi < SK_REGION; qi += SK_NT) {
      u32 ne = 0, flag = 0;
      if (qi < nsteps) {
        const int ui = w + qi;
        const int prev = s_arg[qi];
        const KT kp = s_key[KIDX(prev)], ki = s_key[KIDX(ui)];
        const int li = run_len_at(s_nm, cap + ui + delta, cap);
        if (kp != KMAX) {
          if (li == cap - 1) {  // first full window: other copies of the current minimum (sketch.rs:80-83)
            int c1 = (int)s_cnt[qi] - 1;
            if (s_key[KIDX(qi)] == kp && qi != prev) c1 -= 1;  // slot being overwritten is excluded
            if (c1 > 0) { ne += (u32)c1; flag |= 2u; }
          }
          if (ki <= kp) {
            if (li >= cap) { ne += 1; flag |= 1u; }             // sketch.rs:85
          } else if (prev == qi) {                               // the minimum just left the window (sketch.rs:87)
            if (li >= cap - 1) { ne += 1; flag |= 1u; }          // sketch.rs:88
          }
        }
        if (!(ki <= kp) && prev == qi && li >= cap - 1) {        // rescan: other copies of the new minimum (sketch.rs:92-95)
          const int cur = s_arg[qi + 1];
          if (s_key[KIDX(cur)] != KMAX) {
            const int c3 = (int)s_cnt[qi + 1] - 1;
            if (c3 > 0) { ne += (u32)c3; flag |= 4u; }
          }
        }
      }
      s_ne[qi] = (u16)ne;
      s_flag[qi] = (u8)flag;
    }
    __syncthreads();

    // ---- phase 4b: exclusive scan of the per-step counts (8 consecutive steps per thread) ------------------------------
    u32 tile_total;
    {
      const uint4 v = reinterpret_cast<const uint4*>(s_ne)[tid];
      u32 c[8] = {v.x & 0xFFFFu, v.x >> 16, v.y & 0xFFFFu, v.y >> 16, v.z & 0xFFFFu, v.z >> 16, v.w & 0xFFFFu, v.w >> 16};
      u32 sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += c[j];
      u32 inc = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if ((tid & 31) >= d) inc += t;
      }
      if ((tid & 31) == 31) s_wsum[tid >> 5] = inc;
      __syncthreads();
      u32 wbase = 0, tot = 0;
#pragma unroll
      for (int x = 0; x < SK_NT / 32; ++x) {
        const u32 ws = s_wsum[x];
        if (x < (tid >> 5)) wbase += ws;
        tot += ws;
      }
      u32 run = wbase + inc - sum;
      u32 o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { o[j] = run; run += c[j]; }
      reinterpret_cast<uint4*>(s_off)[2 * tid] = make_uint4(o[0], o[1], o[2], o[3]);
      reinterpret_cast<uint4*>(s_off)[2 * tid + 1] = make_uint4(o[4], o[5], o[6], o[7]);
      tile_total = tot;
    }
    // end-of-sequence emission of the last minimum (sketch.rs:99)
    const bool last_tile = (e == len);
    const int cur_last = s_arg[nsteps];
    const bool end_emit = last_tile && s_key[KIDX(cur_last)] != KMAX;
    const u32 tile_count = tile_total + (end_emit ? 1u : 0u);

    // ---- decoupled look-back: exclusive prefix of tile_count over all earlier tiles -------------------------------------
    if (tid < 32) {
      volatile u64* st = P.tile_status;
      u64 excl = 0;
      if (tile == 0) {
        if (tid == 0) st[0] = (2ULL << 62) | (u64)tile_count;
      } else {
        if (tid == 0) st[tile] = (1ULL << 62) | (u64)tile_count;
        i64 look = (i64)tile - 1;
        for (;;) {
          const i64 idx = look - tid;
          u64 v = (3ULL << 62);  // lanes past the beginning behave as an inclusive prefix of 0
          if (idx >= 0) { do { v = st[idx]; } while ((v >> 62) == 0); } else v = (2ULL << 62);
          const u32 incl_mask = __ballot_sync(0xFFFFFFFFu, (v >> 62) == 2);
          const int first_incl = incl_mask ? (__ffs(incl_mask) - 1) : 32;
          u64 contrib = (tid <= first_incl) ? (v & ((1ULL << 62) - 1)) : 0;
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xFFFFFFFFu, contrib, d);
          excl += contrib;
          if (incl_mask) break;
          look -= 32;
        }
        if (tid == 0) st[tile] = (2ULL << 62) | (excl + (u64)tile_count);
      }
      if (tid == 0) {
        s_base = excl;
        if (tile == P.tile_first[q]) P.seq_out_off[q] = excl;
        if (tile == P.ntiles - 1) P.seq_out_off[P.nseq] = excl + (u64)tile_count;
      }
    }
    __syncthreads();
    const u64 base = s_base;
    const u64 rid_hi = (u64)(P.rid_base + q * P.rid_step) << 32;

    // ---- phase 4c: write the minimizers in step order ----------------------------------------------------------------
    auto emit = [&](u64 o, int u) {
      if (o < P.out_cap) {
        const u64 pos = (u64)(P0 + u);
        const u32 z = (s_z[u >> 3] >> (u & 7)) & 1u;
        P.out_key[o] = ((u64)s_key[KIDX(u)] << 8) | (u64)k;
        P.out_val[o] = rid_hi | (pos << 1) | (u64)z;
      }
    };
    for (int qi = tid; qi < nsteps; qi += SK_NT) {
      const u32 flag = s_flag[qi];
      if (!flag) continue;
      u64 o = base + s_off[qi];
      const int ui = w + qi;
      const int prev = s_arg[qi];
      if (flag & 2u) {
        const KT kp = s_key[KIDX(prev)];
        for (int u = qi + 1; u < ui; ++u)
          if (s_key[KIDX(u)] == kp && u != prev) emit(o++, u);
      }
      if (flag & 1u) emit(o++, prev);
      if (flag & 4u) {
        const int cur = s_arg[qi + 1];
        const KT kc = s_key[KIDX(cur)];
        for (int u = qi + 1; u <= ui; ++u)
          if (s_key[KIDX(u)] == kc && u != cur) emit(o++, u);
      }
    }
    if (tid == 0 && end_emit) emit(base + tile_total, cur_last);
  }
#undef KIDX
}


// ---------------------------------------------------------------------------------------------------------------------
// The tile kernel used for w >= 9 (below): same phases 1-2, but the window argmin is computed per thread with
// prefix/suffix minima over its 8 consecutive positions (3 merges per position instead of w compares) and the emission
// decisions are taken in registers.
// Window minima are kept as (key, pos << 1 | dup): among equal keys the NEWEST position wins (sketch.rs:84,90-91) and
// `dup` says whether the minimum occurs at least twice in the range.

// rare paths of version 2 (a window holding the same key twice), kept out of line so they cost no registers in the hot loop
template <class KT, int PAD>
__device__ __noinline__ u32 sk_count_dups(const KT* s_key, int lo, int hi, KT kv, int excl) {
  u32 n = 0;
  for (int x = lo; x <= hi; ++x) n += (s_key[x + (x >> PAD)] == kv && x != excl) ? 1u : 0u;
  return n;
}
template <class KT, int PAD>
__device__ __noinline__ u64 sk_emit_dups(const KT* s_key, const u8* s_z, int lo, int hi, int excl, u64 o, u64 cap, u64* out_key,
                                         u64* out_val, u64 rid_hi, i64 P0, int k) {
  const KT kv = s_key[excl + (excl >> PAD)];
  for (int x = lo; x <= hi; ++x)
    if (s_key[x + (x >> PAD)] == kv && x != excl) {
      if (o < cap) {
        out_key[o] = ((u64)kv << 8) | (u64)k;
        out_val[o] = rid_hi | ((u64)(P0 + x) << 1) | (u64)((s_z[x >> 3] >> (x & 7)) & 1u);
      }
      ++o;
    }
  return o;
}

constexpr int SK_LIST = 1024;   // staged minimizers per tile (a tile of random sequence emits ~380)

// ---------------------------------------------------------------------------------------------------------------------
// The tile kernel for w >= 9: eight compute warps plus a ninth "scanner" warp per CTA (with the look-back done by a compute
// warp, all eight waited at a barrier for it: 40 % of the stall samples in round 1's profile).  The compute warps stage the
// tile's records (key, pos << 1 | strand) in one of two shared buffers, publish the tile's count and go on with the next
// tile; the scanner warp does the look-back, publishes the inclusive prefix and writes the staged records to global
// memory with coalesced stores.  Named barriers: 1 = the 256 compute threads; FULL[b] = compute arrives, scanner
// waits; EMPTY[b] = scanner arrives, compute waits before reusing buffer b; BASE = tiles too big to stage.
#ifndef MM2_SK3_OCC
#define MM2_SK3_OCC 5
#endif
#ifndef MM2_SK3_OCC64
#define MM2_SK3_OCC64 4   // 64-bit keys (k > 15): 56 registers; measured 13.55 / 13.16 ms per Gbase of HiFi reads at 3 / 4
#endif
constexpr int SK3_BAR_COMPUTE = 1, SK3_BAR_FULL = 2, SK3_BAR_EMPTY = 4, SK3_BAR_BASE = 6;
__device__ __forceinline__ void sk3_bar_compute() { asm volatile("bar.sync %0, %1;" ::"n"(SK3_BAR_COMPUTE), "n"(SK_NT) : "memory"); }
__device__ __forceinline__ void sk3_bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(SK_NT + 32) : "memory"); }
__device__ __forceinline__ void sk3_bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(SK_NT + 32) : "memory"); }
struct Sk3Meta { u32 tile, count, q, staged; };
// staged variant of sk_emit_dups: key and pos << 1 | strand go to the shared buffers
template <class KT, int PAD>
__device__ __noinline__ u32 sk3_stage_dups(const KT* s_key, const u8* s_z, int lo, int hi, int excl, KT* okey, u32* opos, u32 idx, i64 P0) {
  const KT kv = s_key[excl + (excl >> PAD)];
  for (int x = lo; x <= hi; ++x)
    if (s_key[x + (x >> PAD)] == kv && x != excl) {
      okey[idx] = kv;
      opos[idx] = ((u32)(P0 + x) << 1) | ((s_z[x >> 3] >> (x & 7)) & 1u);
      ++idx;
    }
  return idx;
}
template <class KT>
__global__ void __launch_bounds__(SK_NT + 32, sizeof(KT) == 4 ? MM2_SK3_OCC : MM2_SK3_OCC64) sketch_tile_kernel_v3(SketchParams P) {
  constexpr int PAD = KeyTraits<KT>::PAD;
  constexpr KT KMAX = (KT)~(KT)0;
#define KIDX(u) ((u) + ((u) >> PAD))
  __shared__ __align__(16) KT s_key[SK_REGION + (SK_REGION >> PAD) + 8];
  __shared__ __align__(16) u32 s_pack[SK_MAXCHUNK + 4];
  __shared__ __align__(16) u32 s_nm[SK_MAXCHUNK / 2 + 4];
  __shared__ u8 s_z[SK_NT];
  __shared__ u32 s_wsum[SK_NT / 32];
  __shared__ u32 s_next[2];
  __shared__ u64 s_base;
  __shared__ __align__(16) KT s_okey[2][SK_LIST];
  __shared__ __align__(16) u32 s_opos[2][SK_LIST];
  __shared__ Sk3Meta s_meta[2];

  const int tid = threadIdx.x;
  const int w = P.w, k = P.k;
  const int cap = w + k;
  const int T = SK_REGION - w;
  const KT mask = (KT)((((u64)1) << (2 * k)) - 1);
  const int shift1 = 2 * (k - 1);

  if (tid >= SK_NT) {
    // ---- scanner warp: look-back, inclusive prefix, record write of the tiles this CTA has finished computing --------------
    const int lane = tid - SK_NT;
    volatile u64* st = P.tile_status;
    for (u32 it = 0;; ++it) {
      const int b = (int)(it & 1u);
      sk3_bar_sync(SK3_BAR_FULL + b);
      const u32 tile = s_meta[b].tile;
      if (tile == 0xFFFFFFFFu) break;
      const u32 tile_count = s_meta[b].count, q = s_meta[b].q;
      const bool staged = s_meta[b].staged != 0;
      u64 excl = 0;
      if (tile != 0) {
        i64 look = (i64)tile - 1;
        for (;;) {
          const i64 idx = look - lane;
          u64 v;
          if (idx >= 0) { do { v = st[idx]; } while ((v >> 62) == 0); } else v = (2ULL << 62);
          const u32 incl_mask = __ballot_sync(0xFFFFFFFFu, (v >> 62) == 2);
          const int first_incl = incl_mask ? (__ffs(incl_mask) - 1) : 32;
          u64 contrib = (lane <= first_incl) ? (v & ((1ULL << 62) - 1)) : 0;
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xFFFFFFFFu, contrib, d);
          excl += contrib;
          if (incl_mask) break;
          look -= 32;
        }
        if (lane == 0) st[tile] = (2ULL << 62) | (excl + (u64)tile_count);
      }
      if (lane == 0) {
        if (tile == P.tile_first[q]) P.seq_out_off[q] = excl;
        if (tile == P.ntiles - 1) P.seq_out_off[P.nseq] = excl + (u64)tile_count;
      }
      if (staged) {
        const u64 rid_hi = (u64)(P.rid_base + q * P.rid_step) << 32;
        for (u32 e2 = (u32)lane; e2 < tile_count; e2 += 32) {
          const u64 o = excl + e2;
          if (o < P.out_cap) {
            P.out_key[o] = ((u64)s_okey[b][e2] << 8) | (u64)k;
            P.out_val[o] = rid_hi | (u64)s_opos[b][e2];
          }
        }
      } else {
        if (lane == 0) s_base = excl;
        __threadfence_block();
        sk3_bar_arrive(SK3_BAR_BASE);
      }
      __threadfence_block();
      sk3_bar_arrive(SK3_BAR_EMPTY + b);
    }
    return;
  }

  // Tiles are handed out in order (the look-back needs every earlier tile to be running or done); the next ticket is
  // taken just before this tile's count is published.
  if (tid == 0) s_next[0] = P.tile_base + atomicAdd(P.ticket, 1u);
  sk3_bar_compute();
  u32 tile = s_next[0];
  u32 it = 0;
  for (int par = 0; tile < P.ntiles; par ^= 1) {
    const u32 q = P.tile_seq[tile];
    const u64 soff = P.seq_off[q];
    const i64 len = (i64)(P.seq_off[q + 1] - soff);
    const i64 s = (i64)(tile - P.tile_first[q]) * T;
    const i64 e = min(len, s + (i64)T);
    const int nsteps = (int)(e - s);
    const i64 P0 = s - w;
    const i64 a = P0 - cap;
    const i64 gidx = (i64)soff + a;
    const i64 g0 = (gidx >> 4) << 4;
    const int delta = (int)(gidx - g0);
    const int nchunks = (delta + SK_REGION + cap + 15) >> 4;

    // ---- phase 1: 128-bit loads -> 2-bit packed codes + N mask (as in version 1) ---------------------------------
    for (int c = tid; c < SK_MAXCHUNK + 4; c += SK_NT) {
      u32 packed = 0, nmask = 0xFFFFu;
      if (c < nchunks) {
        const i64 gi = g0 + 16 * (i64)c;
        u32 wd[4] = {0, 0, 0, 0};
        if (P.vec_ok && gi >= 0 && gi + 16 <= (i64)P.buf_len) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(P.seq + gi));
          wd[0] = v.x; wd[1] = v.y; wd[2] = v.z; wd[3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const i64 g = gi + j;
            u32 b = (g >= 0 && g < (i64)P.buf_len) ? (u32)P.seq[g] : 0u;
            wd[j >> 2] |= b << (8 * (j & 3));
          }
        }
        nmask = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          u32 c4, vm;
          nt4x4(wd[j], c4, vm);
          packed |= pack4(c4) << (8 * j);
          nmask |= nbits4(vm) << (4 * j);
        }
        const i64 pstart = a + 16 * (i64)c - delta;
        const i64 lo = max((i64)0, -pstart), hi = min((i64)16, len - pstart);
        u32 inseq = 0;
        if (hi > lo) inseq = ((hi >= 16 ? 0x10000u : (1u << hi)) - 1u) & ~((1u << lo) - 1u);
        nmask = (nmask | ~inseq) & 0xFFFFu;
      }
      s_pack[c] = packed;
      reinterpret_cast<u16*>(s_nm)[c] = (u16)nmask;
    }
    sk3_bar_compute();

    // ---- phase 2: 8 consecutive k-mers per thread -> keys (registers + shared), strand bits, l gates -----------------
    KT K[SK_CH];
    u32 ge_cap = 0, eq_capm1 = 0;
    {
      const int r0 = cap + SK_CH * tid + delta;
      int l = run_len_at(s_nm, r0 - 1, cap);
      const int rs = r0 - k;
      const int wi = rs >> 4, sh = 2 * (rs & 15);
      const u32 w0 = s_pack[wi], w1 = s_pack[wi + 1], w2 = s_pack[wi + 2];
      const u32 flo = __funnelshift_r(w0, w1, sh), fhi = __funnelshift_r(w1, w2, sh);
      const u64 field = (((u64)fhi << 32) | flo) & (u64)mask;
      KT rev = (KT)((~field) & (u64)mask);
      u64 br = __brevll(field);
      br = ((br & 0x5555555555555555ULL) << 1) | ((br >> 1) & 0x5555555555555555ULL);
      KT fwd = (KT)(br >> (64 - 2 * k));
      const u32 cw = __funnelshift_r(s_pack[r0 >> 4], s_pack[(r0 >> 4) + 1], 2 * (r0 & 15));
      const u32 nb = __funnelshift_r(s_nm[r0 >> 5], s_nm[(r0 >> 5) + 1], r0 & 31);
      u32 zbits = 0;
      if (l >= cap && (nb & 0xFFu) == 0u) {
        // no N among these 8 bases and a full run before them: l stays at its cap, every k-mer is hashed
        ge_cap = 0xFFu;
#pragma unroll
        for (int j = 0; j < SK_CH; ++j) {
          const u32 c = (cw >> (2 * j)) & 3u;
          fwd = (KT)(((fwd << 2) | (KT)c) & mask);
          rev = (KT)((rev >> 2) | ((KT)(3u ^ c) << shift1));
          const bool z = !(fwd < rev);
          const KT key = hash_mix<KT>(z ? rev : fwd, mask);
          K[j] = key;
          s_key[KIDX(SK_CH * tid + j)] = key;
          zbits |= (u32)z << j;
        }
      } else {
#pragma unroll
        for (int j = 0; j < SK_CH; ++j) {
          const u32 c = (cw >> (2 * j)) & 3u;
          l = ((nb >> j) & 1u) ? 0 : min(l + 1, cap);
          fwd = (KT)(((fwd << 2) | (KT)c) & mask);
          rev = (KT)((rev >> 2) | ((KT)(3u ^ c) << shift1));
          const bool z = !(fwd < rev);
          KT key = KMAX;
          if (l >= k) key = hash_mix<KT>(z ? rev : fwd, mask);
          K[j] = key;
          s_key[KIDX(SK_CH * tid + j)] = key;
          zbits |= (u32)z << j;
          ge_cap |= (u32)(l >= cap) << j;
          eq_capm1 |= (u32)(l == cap - 1) << j;
        }
      }
      s_z[tid] = (u8)zbits;
    }
    sk3_bar_compute();

    // ---- phase 3: window minima of this thread's 8 positions + emission decisions (sketch.rs:80-96) -----------------
    const int c0 = SK_CH * tid;
    const int u_last = w + nsteps - 1;
    const bool last_tile = (e == len);
    auto keyat = [&](int t) -> KT { return t >= 0 ? s_key[KIDX(t)] : KMAX; };
    u32 tot = 0, eflags = 0;  // bits 0-7: emit prev; 8-15: first-window duplicates; 16-23: rescan duplicates; 24: end emit
    u32 pp[4] = {0, 0, 0, 0};  // position of the previous minimum for each of the 8 steps (u16 x 8)
    int cur7 = 0;              // position of the window minimum at this thread's last position
    if (c0 + SK_CH > w && c0 < w + nsteps) {
      // suffix minima over the w keys before this thread's chunk, newest position winning ties (sketch.rs:84,90-91).
      // A range is (key, pd) with pd = pos << 1 | dup, dup = "the minimum occurs at least twice in the range".
      KT rk = KMAX; int rpd = 0;
      int t = c0 - 1;
      for (; t > c0 - (w - 1) + 7; --t) {
        const KT kx = keyat(t);
        if (kx < rk) { rk = kx; rpd = t << 1; } else if (kx == rk) rpd |= 1;
      }
      KT Sk[SK_CH]; int Spd[SK_CH];
#pragma unroll
      for (int jj = SK_CH - 1; jj >= 0; --jj) {
        const KT kx = keyat(t);
        if (kx < rk) { rk = kx; rpd = t << 1; } else if (kx == rk) rpd |= 1;
        Sk[jj] = rk; Spd[jj] = rpd;
        --t;
      }
      KT pk_prev; int ppd_prev;  // window [c0-w, c0-1]
      {
        const KT kx = keyat(t);
        pk_prev = rk; ppd_prev = rpd;
        if (kx < rk) { pk_prev = kx; ppd_prev = t << 1; } else if (kx == rk) ppd_prev |= 1;
      }
      // Fast path: all 8 steps are inside the tile, every window involved is full (l >= w + k: the previous minimum is a
      // real k-mer and sketch.rs:84/88 emit unconditionally) and no window holds its minimum twice.  Then step u emits
      // the previous minimum exactly when it is replaced (new key <= old minimum) or slides out (its position is u - w).
      bool slow = !(ge_cap == 0xFFu && c0 >= w && c0 + SK_CH - 1 <= u_last && !(last_tile && c0 + SK_CH - 1 == u_last));
      if (!slow) {
        const KT pk0 = pk_prev; const int ppd0 = ppd_prev;
        KT fk = KMAX; int fpd = 0;
        int odd = ppd_prev;
#pragma unroll
        for (int j = 0; j < SK_CH; ++j) {
          const int u = c0 + j;
          const KT ki = K[j];
          if (ki <= fk) { fpd = (u << 1) | (ki == fk ? 1 : 0); fk = ki; }
          KT ck; int cpd;
          if (fk <= Sk[j]) { ck = fk; cpd = fpd | (fk == Sk[j] ? 1 : 0); } else { ck = Sk[j]; cpd = Spd[j]; }
          const int ppos = ppd_prev >> 1;
          pp[j >> 1] |= (u32)ppos << (16 * (j & 1));
          const u32 em = (ki <= pk_prev || ppos == u - w) ? 1u : 0u;
          tot += em; eflags |= em << j;
          odd |= cpd;
          pk_prev = ck; ppd_prev = cpd;
        }
        cur7 = ppd_prev >> 1;
        if (odd & 1) {   // a repeated minimum somewhere: redo these 8 steps with the full rules
          slow = true; tot = 0; eflags = 0; pp[0] = pp[1] = pp[2] = pp[3] = 0; pk_prev = pk0; ppd_prev = ppd0;
        }
      }
      if (slow) {
        KT fk = KMAX; int fpd = 0;  // prefix minima inside the chunk (newer element wins ties)
  #pragma unroll
        for (int j = 0; j < SK_CH; ++j) {
          const int u = c0 + j;
          const KT ki = K[j];
          if (ki <= fk) { fpd = (u << 1) | (ki == fk ? 1 : 0); fk = ki; }
          // window [u-w+1, u] = older part (suffix) + newer part (prefix); the newer part wins ties
          KT ck; int cpd;
          if (fk <= Sk[j]) { ck = fk; cpd = fpd | (fk == Sk[j] ? 1 : 0); } else { ck = Sk[j]; cpd = Spd[j]; }
          const int ppos = ppd_prev >> 1;
          pp[j >> 1] |= (u32)ppos << (16 * (j & 1));
          if (u >= w && u <= u_last) {
            const KT kp = pk_prev;
            const bool gc = (ge_cap >> j) & 1u, ec1 = (eq_capm1 >> j) & 1u;
            if (kp != KMAX) {
              if (ec1 && (ppd_prev & 1)) {
                const u32 c1 = sk_count_dups<KT, PAD>(s_key, u - w + 1, u - 1, kp, ppos);
                if (c1) { tot += c1; eflags |= 1u << (8 + j); }
              }
              if (ki <= kp) {
                if (gc) { tot += 1; eflags |= 1u << j; }
              } else if (ppos == u - w) {
                if (gc || ec1) {
                  tot += 1; eflags |= 1u << j;
                  if (ck != KMAX && (cpd & 1)) {
                    const u32 c3 = sk_count_dups<KT, PAD>(s_key, u - w + 1, u, ck, cpd >> 1);
                    if (c3) { tot += c3; eflags |= 1u << (16 + j); }
                  }
                }
              }
            }
            if (last_tile && u == u_last && ck != KMAX) { tot += 1; eflags |= 1u << 24; }
          }
          if (u == min(c0 + SK_CH - 1, u_last)) cur7 = cpd >> 1;
          pk_prev = ck; ppd_prev = cpd;
        }
      }
    }

    // ---- exclusive scan of the per-thread counts; the tile's count is published at once --------------------------------------
    u32 inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u32 tt = __shfl_up_sync(0xFFFFFFFFu, inc, d);
      if ((tid & 31) >= d) inc += tt;
    }
    if ((tid & 31) == 31) s_wsum[tid >> 5] = inc;
    if (tid == 0) s_next[par ^ 1] = P.tile_base + atomicAdd(P.ticket, 1u);
    sk3_bar_compute();
    const u32 next_tile = s_next[par ^ 1];
    u32 wbase = 0, tile_count = 0;
#pragma unroll
    for (int x = 0; x < SK_NT / 32; ++x) {
      const u32 ws = s_wsum[x];
      if (x < (tid >> 5)) wbase += ws;
      tile_count += ws;
    }
    const u32 my_off = wbase + inc - tot;
    const int b = (int)(it & 1u);
    const bool staged = tile_count <= (u32)SK_LIST;
    if (tid == 0) {
      volatile u64* st = P.tile_status;
      st[tile] = ((tile == 0 ? 2ULL : 1ULL) << 62) | (u64)tile_count;   // tile 0: its inclusive prefix; others: aggregate
    }
    if (it >= 2) sk3_bar_sync(SK3_BAR_EMPTY + b);   // the scanner is done with the tile that used this buffer before

    // ---- stage the records in step order (shared buffer b); the scanner warp writes them out -----------------------------------
    if (staged && tot) {
      KT* okey = s_okey[b];
      u32* opos = s_opos[b];
      u32 idx = my_off;
      auto put = [&](int x) {
        okey[idx] = s_key[KIDX(x)];
        opos[idx] = ((u32)(P0 + x) << 1) | ((s_z[x >> 3] >> (x & 7)) & 1u);
        ++idx;
      };
      if ((eflags >> 8) == 0u) {
        // only "previous minimum" emissions (every clean run).  A thread emits ~1.5 records for its 8 steps, so the loop runs
        // over the set bits (the fully unrolled, predicated 8-step version cost 15 % of the kernel's instructions)
        const u32 P0u = (u32)P0;
        u32 m = eflags, o = my_off;
        while (m) {
          const int j = __ffs(m) - 1;
          m &= m - 1;
          const u32 pw = j < 4 ? (j < 2 ? pp[0] : pp[1]) : (j < 6 ? pp[2] : pp[3]);
          const int x = (int)((pw >> (16 * (j & 1))) & 0xFFFFu);
          okey[o] = s_key[KIDX(x)];
          opos[o] = ((P0u + (u32)x) << 1) | ((s_z[x >> 3] >> (x & 7)) & 1u);
          ++o;
        }
      } else {
        u32 jm = (eflags | (eflags >> 8) | (eflags >> 16)) & 0xFFu;  // steps of this thread that emit anything
        while (jm) {
          const int j = __ffs(jm) - 1;
          jm &= jm - 1;
          const int u = c0 + j;
          const u32 pw = (j >> 1) == 0 ? pp[0] : (j >> 1) == 1 ? pp[1] : (j >> 1) == 2 ? pp[2] : pp[3];
          const int ppos = (int)((pw >> (16 * (j & 1))) & 0xFFFFu);
          if (eflags & (1u << (8 + j))) idx = sk3_stage_dups<KT, PAD>(s_key, s_z, u - w + 1, u - 1, ppos, okey, opos, idx, P0);
          if (eflags & (1u << j)) put(ppos);
          if (eflags & (1u << (16 + j))) {
            const int j1 = j + 1;
            const u32 pw1 = (j1 >> 1) == 0 ? pp[0] : (j1 >> 1) == 1 ? pp[1] : (j1 >> 1) == 2 ? pp[2] : pp[3];
            const int cpos = (j == SK_CH - 1) ? cur7 : (int)((pw1 >> (16 * (j1 & 1))) & 0xFFFFu);
            idx = sk3_stage_dups<KT, PAD>(s_key, s_z, u - w + 1, u, cpos, okey, opos, idx, P0);
          }
        }
        if (eflags & (1u << 24)) put(cur7);
      }
    }
    if (tid == 0) { s_meta[b].tile = tile; s_meta[b].count = tile_count; s_meta[b].q = q; s_meta[b].staged = staged ? 1u : 0u; }
    __threadfence_block();
    sk3_bar_arrive(SK3_BAR_FULL + b);
    if (!staged) {
      // a tile that emits more than a buffer holds (windows full of repeated minima): wait for the scanner's base and
      // let each thread write its own records
      sk3_bar_sync(SK3_BAR_BASE);
      if (tot) {
        const u64 rid_hi = (u64)(P.rid_base + q * P.rid_step) << 32;
        u64 o = s_base + my_off;
        auto emit = [&](int x) {
          if (o < P.out_cap) {
            const u64 pos = (u64)(P0 + x);
            const u32 z = (s_z[x >> 3] >> (x & 7)) & 1u;
            P.out_key[o] = ((u64)s_key[KIDX(x)] << 8) | (u64)k;
            P.out_val[o] = rid_hi | (pos << 1) | (u64)z;
          }
          ++o;
        };
        u32 jm = (eflags | (eflags >> 8) | (eflags >> 16)) & 0xFFu;
        while (jm) {
          const int j = __ffs(jm) - 1;
          jm &= jm - 1;
          const int u = c0 + j;
          const u32 pw = (j >> 1) == 0 ? pp[0] : (j >> 1) == 1 ? pp[1] : (j >> 1) == 2 ? pp[2] : pp[3];
          const int ppos = (int)((pw >> (16 * (j & 1))) & 0xFFFFu);
          if (eflags & (1u << (8 + j)))
            o = sk_emit_dups<KT, PAD>(s_key, s_z, u - w + 1, u - 1, ppos, o, P.out_cap, P.out_key, P.out_val, rid_hi, P0, k);
          if (eflags & (1u << j)) emit(ppos);
          if (eflags & (1u << (16 + j))) {
            const int j1 = j + 1;
            const u32 pw1 = (j1 >> 1) == 0 ? pp[0] : (j1 >> 1) == 1 ? pp[1] : (j1 >> 1) == 2 ? pp[2] : pp[3];
            const int cpos = (j == SK_CH - 1) ? cur7 : (int)((pw1 >> (16 * (j1 & 1))) & 0xFFFFu);
            o = sk_emit_dups<KT, PAD>(s_key, s_z, u - w + 1, u, cpos, o, P.out_cap, P.out_key, P.out_val, rid_hi, P0, k);
          }
        }
        if (eflags & (1u << 24)) emit(cur7);
      }
    }
    tile = next_tile;
    ++it;
  }
  {  // tell the scanner that this CTA has no more tiles
    const int b = (int)(it & 1u);
    if (it >= 2) sk3_bar_sync(SK3_BAR_EMPTY + b);
    if (tid == 0) s_meta[b].tile = 0xFFFFFFFFu;
    __threadfence_block();
    sk3_bar_arrive(SK3_BAR_FULL + b);
  }
#undef KIDX
}

// ---------------------------------------------------------------------------------------------------------------------
// Literal state machine (sketch.rs:29-100), one thread per sequence.  MODE 0: count only; MODE 1: write.
__device__ __forceinline__ u32 nt4_dev(u8 b) {
  const u32 u = b & 0xDFu;
  return u == 0x41u ? 0u : u == 0x43u ? 1u : u == 0x47u ? 2u : u == 0x54u ? 3u : 4u;
}
__device__ __forceinline__ u64 hash64_dev(u64 key, u64 mask) {
  key = (~key + (key << 21)) & mask;
  key ^= key >> 24;
  key = (key + (key << 3) + (key << 8)) & mask;
  key ^= key >> 14;
  key = (key + (key << 2) + (key << 4)) & mask;
  key ^= key >> 28;
  key = (key + (key << 31)) & mask;
  return key;
}

struct LiteralParams {
  const u8* seq; const u64* seq_off; u32 nseq;
  int w, k, is_hpc;
  u32 rid_base, rid_step;
  u64* counts;          // MODE 0 out: nseq (u64)
  const u64* out_off;   // MODE 1 in: nseq+1
  u64* out_key; u64* out_val; u64 out_cap;
  u64* win;             // nseq * 2 * w u64 scratch for the window (keys then vals)
};

template <int MODE>
__global__ void sketch_literal_kernel(LiteralParams P) {
  const u32 q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= P.nseq) return;
  const u8* seq = P.seq + P.seq_off[q];
  const i64 len = (i64)(P.seq_off[q + 1] - P.seq_off[q]);
  const int w = P.w, k = P.k;
  u64* bk = P.win + (u64)q * 2 * (u64)w;
  u64* bv = bk + w;
  const u64 MAXV = ~0ULL;
  for (int j = 0; j < w; ++j) { bk[j] = MAXV; bv[j] = MAXV; }
  const u64 shift1 = 2 * ((u64)k - 1), mask = (1ULL << (2 * k)) - 1;
  u64 kmer0 = 0, kmer1 = 0, mink = MAXV, minv = MAXV;
  int l = 0, buf_pos = 0, min_pos = 0, kmer_span = 0;
  int tq_front = 0, tq_count = 0; int tq[32];
  u64 o = MODE ? P.out_off[q] : 0, n = 0;
  const u64 rid_hi = (u64)(P.rid_base + q * P.rid_step) << 32;
  auto push = [&](u64 kk, u64 vv) {
    if (MODE) { if (o < P.out_cap) { P.out_key[o] = kk; P.out_val[o] = vv; } ++o; }
    ++n;
  };
  for (i64 i = 0; i < len; ++i) {
    const int c = (int)nt4_dev(seq[i]);
    u64 ik = MAXV, iv = MAXV;
    if (c < 4) {
      if (P.is_hpc) {  // sketch.rs:51-61
        i64 skip_len = 1;
        if (i + 1 < len && (int)nt4_dev(seq[i + 1]) == c) {
          i64 t = i + 2;
          while (t < len && (int)nt4_dev(seq[t]) == c) ++t;
          skip_len = t - i;
        }
        tq[(tq_count + tq_front) & 0x1f] = (int)skip_len; tq_count += 1;
        kmer_span += (int)skip_len;
        if (tq_count > k) { kmer_span -= tq[tq_front]; tq_front = (tq_front + 1) & 0x1f; tq_count -= 1; }
      } else {
        kmer_span = (l + 1 < k) ? l + 1 : k;
      }
      kmer0 = ((kmer0 << 2) | (u64)c) & mask;
      kmer1 = (kmer1 >> 2) | ((u64)(3 ^ c) << shift1);
      if (kmer0 != kmer1) {
        const int z = kmer0 < kmer1 ? 0 : 1;
        l += 1;
        if (l >= k && kmer_span < 256) {
          ik = (hash64_dev(z ? kmer1 : kmer0, mask) << 8) | (u64)kmer_span;
          iv = rid_hi | ((u64)i << 1) | (u64)z;
        }
      }
    } else { l = 0; tq_front = 0; tq_count = 0; kmer_span = 0; }
    bk[buf_pos] = ik; bv[buf_pos] = iv;
    if (l == w + k - 1 && mink != MAXV) {
      for (int j = buf_pos + 1; j < w; ++j) if (mink == bk[j] && bv[j] != minv) push(bk[j], bv[j]);
      for (int j = 0; j < buf_pos; ++j) if (mink == bk[j] && bv[j] != minv) push(bk[j], bv[j]);
    }
    if (ik <= mink) {
      if (l >= w + k && mink != MAXV) push(mink, minv);
      mink = ik; minv = iv; min_pos = buf_pos;
    } else if (buf_pos == min_pos) {
      if (l >= w + k - 1 && mink != MAXV) push(mink, minv);
      mink = MAXV;
      for (int j = buf_pos + 1; j < w; ++j) if (mink >= bk[j]) { mink = bk[j]; minv = bv[j]; min_pos = j; }
      for (int j = 0; j <= buf_pos; ++j) if (mink >= bk[j]) { mink = bk[j]; minv = bv[j]; min_pos = j; }
      if (l >= w + k - 1 && mink != MAXV) {
        for (int j = buf_pos + 1; j < w; ++j) if (mink == bk[j] && minv != bv[j]) push(bk[j], bv[j]);
        for (int j = 0; j <= buf_pos; ++j) if (mink == bk[j] && minv != bv[j]) push(bk[j], bv[j]);
      }
    }
    if (++buf_pos == w) buf_pos = 0;
  }
  if (mink != MAXV) push(mink, minv);
  if (!MODE) P.counts[q] = n;
}

__global__ void excl_scan_u64_small(const u64* in, u64* out, u32 n) {  // single thread; n is small on this path
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    u64 run = 0;
    for (u32 i = 0; i < n; ++i) { out[i] = run; run += in[i]; }
    out[n] = run;
  }
}

// number of tiles of each sequence (at least one, so that empty sequences still get an output offset)
__global__ void tile_count_kernel(const u64* __restrict__ seq_off, u32 nseq, int T, u32* __restrict__ cnt) {
  const u32 q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nseq) return;
  const u64 len = seq_off[q + 1] - seq_off[q];
  const u64 c = (len + (u64)T - 1) / (u64)T;
  cnt[q] = (u32)(c ? c : 1);
}
// sequence of each tile: last q with tile_first[q] <= tile
__global__ void tile_seq_kernel(const u64* __restrict__ tile_first, u32 nseq, u32 ntiles, u32* __restrict__ tile_seq) {
  for (u32 t = blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += gridDim.x * blockDim.x) {
    u32 lo = 0, hi = nseq - 1;
    while (lo < hi) {
      const u32 mid = (lo + hi + 1) >> 1;
      if (tile_first[mid] <= (u64)t) lo = mid; else hi = mid - 1;
    }
    tile_seq[t] = lo;
  }
}

int g_num_sms = 0;

}  // namespace

static int num_sms(int device) {
  if (g_num_sms == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
    g_num_sms = v;
  }
  return g_num_sms;
}

// Tiles of the odd-k / non-HPC kernels: T = SK_REGION - w steps each, at least one per sequence.
u64 sketch_tile_count(const u64* h_off, size_t nseq, int w) {
  const u64 T = (u64)(SK_REGION - w);
  u64 nt = 0;
  for (size_t i = 0; i < nseq; ++i) nt += std::max<u64>(1, (h_off[i + 1] - h_off[i] + T - 1) / T);
  return nt;
}
// bytes of the concatenated sequences (relative to h_off[0]) that the tiles [tile_lo, tile_hi) read, halo included
void sketch_tile_bytes(const u64* h_off, size_t nseq, int w, int k, u64 tile_lo, u64 tile_hi, u64* byte_lo, u64* byte_hi) {
  const u64 T = (u64)(SK_REGION - w), total = h_off[nseq] - h_off[0];
  *byte_lo = *byte_hi = 0;
  if (tile_hi <= tile_lo) return;
  u64 t0 = 0, lo = ~0ULL, hi = 0;
  for (size_t q = 0; q < nseq; ++q) {
    const u64 so = h_off[q] - h_off[0], len = h_off[q + 1] - h_off[q];
    const u64 ntq = std::max<u64>(1, (len + T - 1) / T);
    const u64 a = std::max(t0, tile_lo), b = std::min(t0 + ntq, tile_hi);
    if (a < b) {
      const u64 s = so + (a - t0) * T, e = so + std::min(len, (b - t0) * T);
      const u64 halo = (u64)(2 * w + k) + 32;
      lo = std::min(lo, s > halo ? s - halo : 0);
      hi = std::max(hi, std::min(total, e + 32));
    }
    t0 += ntq;
    if (t0 >= tile_hi) break;
  }
  if (lo == ~0ULL) return;
  *byte_lo = lo & ~(u64)15; *byte_hi = hi;
}

namespace {
// "inclusive prefix 0" in the 32 status slots in front of a shard's first tile (one look-back window: every lane must
// find a published value)
__global__ void sketch_seed_prefix_kernel(u64* st, u32 t_lo) { if (threadIdx.x < t_lo && threadIdx.x < 32) st[t_lo - 1 - threadIdx.x] = 2ULL << 62; }
}

int sketch_device(mm2_ctx* ctx, const u8* d_cat, const u64* d_off, const u64* h_off, size_t nseq, int w, int k,
                  u32 rid_base, u32 rid_step, int is_hpc, SketchOut* out, const SketchFeed* feed, const SketchShard* shard) {
  if (!(w > 0 && w < 256) || !(k > 0 && k <= 28)) { mm2_set_error("sketch: need 0<w<256 and 0<k<=28 (sketch.rs:31-32)"); return MM2_E_ARG; }
  if (nseq == 0) { out->key = out->val = nullptr; out->seq_off = nullptr; out->total = 0; return MM2_OK; }
  if (nseq >= 0xFFFFFFFFull) { mm2_set_error("sketch: too many sequences"); return MM2_E_ARG; }
  const u64 total_len = h_off[nseq] - h_off[0];
  cudaStream_t st = ctx->stream;
  MM2_TRY(ctx->mini_off.ensure((nseq + 1) * 8));
  const bool tile_path = (k & 1) && !is_hpc;
  if (shard && !tile_path) { mm2_set_error("sketch: tile shards need the tile kernels (odd k, no HPC)"); return MM2_E_ARG; }
  if (tile_path) {
    const int T = SK_REGION - w;
    const u64 nt = sketch_tile_count(h_off, nseq, w);
    if (nt >= 0xFFFFFFF0ull) { mm2_set_error("sketch: too many tiles"); return MM2_E_ARG; }
    const u32 ntiles = (u32)nt;
    const u32 t_lo = shard ? (u32)std::min<u64>(shard->tile_lo, ntiles) : 0u, t_hi = shard ? (u32)std::min<u64>(shard->tile_hi, ntiles) : ntiles;
    // capacity guess: random sequence gives 2/(w+1) minimizers per base; retried at the exact size if too small
    const u64 len_eff = shard ? std::min<u64>(total_len, (u64)(t_hi - t_lo) * (u64)T) : total_len;
    u64 cap = (u64)((double)len_eff * 2.0 / (double)(w + 1) * 1.25) + 2 * (shard ? 1 : nseq) + 1024;
    // tile -> sequence tables are derived ON the device from the resident offsets (no small H2D copies: they would
    // queue behind another context's bulk read upload on the copy engine and stall this stream)
    MM2_TRY(ctx->tile_first.ensure((nseq + 4) * 8 + (nseq + 8) * 4));
    MM2_TRY(ctx->tile_seq.ensure((size_t)ntiles * 4));
    MM2_TRY(ctx->tile_status.ensure((size_t)ntiles * 8 + 16 + 4 * 64));   // + one ticket counter per launch of a fed sketch
    u64* d_tf64 = ctx->tile_first.as<u64>();
    u32* d_tcnt = (u32*)((u8*)ctx->tile_first.p + ((((nseq + 2) * 8) + 15) / 16) * 16);  // 16-byte aligned for the scan's vector loads
    MM2_LAUNCH(ctx, tile_count_kernel, (int)((nseq + 255) / 256), 256, 0, d_off, (u32)nseq, T, d_tcnt);
    MM2_TRY(scan_u32_to_u64(ctx, d_tcnt, d_tf64, nseq));
    MM2_LAUNCH(ctx, tile_seq_kernel, (int)std::min<u64>(((u64)ntiles + 255) / 256, 148ull * 32), 256, 0, d_tf64, (u32)nseq, ntiles,
               ctx->tile_seq.as<u32>());
    if (t_hi <= t_lo) {   // an empty shard
      MM2_TRY(ctx->mkey.ensure(64)); MM2_TRY(ctx->mval.ensure(64));
      out->total = 0; out->key = ctx->mkey.as<u64>(); out->val = ctx->mval.as<u64>(); out->seq_off = ctx->mini_off.as<u64>();
      return MM2_OK;
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
      MM2_TRY(ctx->mkey.ensure(cap * 8));
      MM2_TRY(ctx->mval.ensure(cap * 8));
      CUDA_TRY(cudaMemsetAsync(ctx->tile_status.p, 0, (size_t)ntiles * 8 + 16 + 4 * 64, st));
      if (t_lo > 0) MM2_LAUNCH(ctx, sketch_seed_prefix_kernel, 1, 32, 0, ctx->tile_status.as<u64>(), t_lo);
      SketchParams P;
      P.tile_base = t_lo;
      P.seq = d_cat; P.seq_off = d_off; P.buf_len = h_off[nseq];
      P.tile_seq = ctx->tile_seq.as<u32>(); P.tile_first = d_tf64;
      P.nseq = (u32)nseq; P.ntiles = t_hi; P.w = w; P.k = k;
      P.vec_ok = ((uintptr_t)d_cat & 15) == 0;
      P.rid_base = rid_base; P.rid_step = rid_step;
      P.out_key = ctx->mkey.as<u64>(); P.out_val = ctx->mval.as<u64>(); P.out_cap = cap;
      P.seq_out_off = ctx->mini_off.as<u64>();
      P.tile_status = ctx->tile_status.as<u64>();
      P.ticket = (u32*)((u8*)ctx->tile_status.p + (size_t)ntiles * 8);
      const int grid = (int)std::min<u64>(t_hi - t_lo, (u64)num_sms(ctx->device) * (MM2_SK3_OCC + 2));
      if (feed && !shard && attempt == 0 && w >= 9 && feed->nchunks >= 1 && feed->nchunks <= 64) {
        // The sequence is still being uploaded (index build): one launch per uploaded chunk over the tiles that lie entirely
        // inside it.  The launches share the tile status array, so the look-back of a launch's first tile finds the
        // inclusive prefix the previous launch left; every launch has its own ticket counter.
        u32 t_prev = 0;
        for (int c = 0; c < feed->nchunks; ++c) {
          u64 t_end = ntiles;
          if (c + 1 < feed->nchunks) {
            const u64 B = feed->chunk_end[c];
            t_end = 0;
            for (size_t q = 0; q < nseq; ++q) {
              const u64 so = h_off[q] - h_off[0], len = h_off[q + 1] - h_off[q];
              const u64 ntq = std::max<u64>(1, (len + T - 1) / T);
              if (so + len + 16 <= B) { t_end += ntq; continue; }
              if (B > so + 16) t_end += std::min<u64>(ntq - 1, (B - so - 16) / (u64)T);   // tiles that end (with a 16-byte margin) below B
              break;                                                                   // later sequences lie above B
            }
          }
          CUDA_TRY(cudaStreamWaitEvent(st, feed->ev[c], 0));
          if (t_end > t_prev) {
            P.tile_base = t_prev; P.ntiles = (u32)t_end;
            P.ticket = (u32*)((u8*)ctx->tile_status.p + (size_t)ntiles * 8 + 16) + c;
            const int g = (int)std::min<u64>(t_end - t_prev, (u64)num_sms(ctx->device) * (MM2_SK3_OCC + 2));
            if (k <= 15) MM2_LAUNCH(ctx, sketch_tile_kernel_v3<u32>, g, SK_NT + 32, 0, P);
            else MM2_LAUNCH(ctx, sketch_tile_kernel_v3<u64>, g, SK_NT + 32, 0, P);
            t_prev = (u32)t_end;
          }
        }
      } else {
        if (feed && attempt == 0) for (int c = 0; c < feed->nchunks; ++c) CUDA_TRY(cudaStreamWaitEvent(st, feed->ev[c], 0));
        if (w >= 9) {   // per-thread prefix/suffix window minima + scanner warp
          if (k <= 15) MM2_LAUNCH(ctx, sketch_tile_kernel_v3<u32>, grid, SK_NT + 32, 0, P);
          else MM2_LAUNCH(ctx, sketch_tile_kernel_v3<u64>, grid, SK_NT + 32, 0, P);
        } else {        // small windows: the O(w) window scan
          if (k <= 15) MM2_LAUNCH(ctx, sketch_tile_kernel<u32>, grid, SK_NT, 0, P);
          else MM2_LAUNCH(ctx, sketch_tile_kernel<u64>, grid, SK_NT, 0, P);
        }
      }
      CUDA_TRY(cudaGetLastError());
      u64 total = 0;
      MM2_TRY(read_scalar_u64(ctx, ctx->mini_off.as<u64>() + nseq, &total));
      out->total = total;
      if (total <= cap) break;
      if (attempt == 1) { mm2_set_error("sketch: output capacity overflow after retry"); return MM2_E_CUDA; }
      cap = total + 16;
    }
  } else {
    // literal path: count, scan, write
    if (feed) for (int c = 0; c < feed->nchunks; ++c) CUDA_TRY(cudaStreamWaitEvent(st, feed->ev[c], 0));
    MM2_TRY(ctx->misc.ensure((nseq + 1) * 8 + nseq * 2 * (size_t)w * 8));
    u64* d_counts = ctx->misc.as<u64>();
    u64* d_win = d_counts + nseq + 1;
    LiteralParams P;
    P.seq = d_cat; P.seq_off = d_off; P.nseq = (u32)nseq; P.w = w; P.k = k; P.is_hpc = is_hpc;
    P.rid_base = rid_base; P.rid_step = rid_step;
    P.counts = d_counts; P.out_off = ctx->mini_off.as<u64>(); P.out_key = nullptr; P.out_val = nullptr; P.out_cap = 0;
    P.win = d_win;
    const int nb = (int)((nseq + 63) / 64);
    MM2_LAUNCH(ctx, sketch_literal_kernel<0>, nb, 64, 0, P);
    MM2_LAUNCH(ctx, excl_scan_u64_small, 1, 1, 0, d_counts, ctx->mini_off.as<u64>(), (u32)nseq);
    u64 total = 0;
    MM2_TRY(read_scalar_u64(ctx, ctx->mini_off.as<u64>() + nseq, &total));
    const u64 cap = total + 16;
    MM2_TRY(ctx->mkey.ensure(cap * 8));
    MM2_TRY(ctx->mval.ensure(cap * 8));
    P.out_key = ctx->mkey.as<u64>(); P.out_val = ctx->mval.as<u64>(); P.out_cap = cap;
    MM2_LAUNCH(ctx, sketch_literal_kernel<1>, nb, 64, 0, P);
    CUDA_TRY(cudaGetLastError());
    out->total = total;
  }
  out->key = ctx->mkey.as<u64>();
  out->val = ctx->mval.as<u64>();
  out->seq_off = ctx->mini_off.as<u64>();
  return MM2_OK;
}
