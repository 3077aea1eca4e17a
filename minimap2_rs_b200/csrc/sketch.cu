// sketch.cu — minimizer sketching on sm_100a.  Replaces sketch.rs:29-100 (sketch_sequence) for whole batches of
// sequences (a genome's chromosomes, or the reads of a mapping batch) in one launch.
//
// Three kernels:
//  * sketch_tile_kernel_v4 — odd k, w >= 9, with or without -H (every BASELINE config).  Minimizers by POSITION: a position
//    is emitted iff its key is the minimum of some window that holds it (morphological opening of the key sequence); the
//    literal per-step rules of sketch.rs:80-96 run only around N bases and tie-rich sequence starts.  A CTA owns a tile
//    of 2048-2w consecutive positions of one sequence; per-tile output counts are turned into global offsets by a
//    single-pass decoupled look-back, so the minimizers land in exactly the reference's order.
//    tests/models.py:sketch_model_v4 is the CPU model of this restatement (checked against the oracle).
//  * sketch_tile_kernel — the same path for w < 9: the window state after step i as a pure function of the last w keys,
//    by an O(w) scan per step (tests/models.py:sketch_model).
//  * sketch_literal_kernel — even k (palindromic k-mers stall `l`, sketch.rs:67-69) and -H outside 9 <= w <= 64: one thread
//    per sequence runs the reference state machine literally.  Correct for every parameter set, slow for long sequences.
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

template <class T> __device__ __forceinline__ T sk_min(T a, T b) { return a < b ? a : b; }
template <class T> __device__ __forceinline__ T sk_max(T a, T b) { return a < b ? b : a; }
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
constexpr int SK_LIST = 1024;   // staged minimizers per tile (a tile of random sequence emits ~380)
// named barriers of the w >= 9 kernel: 1 = the 256 compute threads; FULL[b] (2, 3) = compute arrives, scanner waits
constexpr int SK3_BAR_COMPUTE = 1, SK3_BAR_FULL = 2;
__device__ __forceinline__ void sk3_bar_compute() { asm volatile("bar.sync %0, %1;" ::"n"(SK3_BAR_COMPUTE), "n"(SK_NT) : "memory"); }
__device__ __forceinline__ void sk3_bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(SK_NT + 32) : "memory"); }
__device__ __forceinline__ void sk3_bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(SK_NT + 32) : "memory"); }

// ---------------------------------------------------------------------------------------------------------------------
// sketch_tile_kernel_v4 (w >= 9): minimizers by POSITION instead of by step.
//
// The reference emits positions in ascending order and never twice, and wherever l >= w + k held for the last w + 1 steps
// (no N in bases [x-w-k+1, x+w]) position x is emitted iff its key equals the minimum of some window
// of w consecutive k-mers that holds it — ties included: sketch.rs:80-96 emits every older duplicate of a minimum when it is
// replaced, slides out or is found by the rescan.  "K[x] == min of some window around x" is the morphological opening of
// the key sequence: M[u] = min K[u-w+1..u], D[x] = max M[x..x+w-1], emit iff D[x] == K[x].  A thread owns 8 consecutive
// positions; prefix / suffix extrema inside its chunk and the published suffix minima / prefix maxima of the neighbouring
// chunks give both passes in ~3 min/max per position with no argmin, no duplicate count and no position bookkeeping
// (round 1 tracked (key, pos, dup) per window: 30 % of its instructions), and the record of a flagged position comes from
// the thread's own registers.  Keys are kept LEFT-ALIGNED (hash << (bits - 2k)): the `& mask` of every hash64 step
// (sketch.rs:4-13) becomes the natural wrap of the register, and the three add-shift steps are single multiplies.
//
// Sequence ends need no special rule: windows that would start before the first k-mer or end after the last one do not exist
// (their M is forced below every key), which also covers the final emission of sketch.rs:99.  Two cases keep the literal
// per-step rules (sk4_by_step), run by the threads of the affected chunk and of the next ceil((w+7)/8) chunks into a
// shared position bitmap: chunks that see a real N in bases [c-w-k+1, c+7+w] (the l counter gates the emissions), and the
// chunks at a sequence start whose first w keys hold an equal pair or a missing k-mer (sketch.rs:80-86 emits the
// duplicates of the first partial minimum at l == w+k-1 and drops a minimum replaced before l reaches w+k).
// A tile owns T = 2048 - 2w POSITIONS [s, s+T) of one sequence; its region holds positions [s-w, s-w+2048).
// tests/models.py:sketch_model_v4 is the CPU model (checked against the oracle, tie-rich inputs included).
template <class KT>
__device__ __forceinline__ int sk4_idx(int c, int e) {   // element e (0..7) of chunk c: 16-byte vectors, one plane per vector slot
  constexpr int PLN = 16 / (int)sizeof(KT);
  return ((e / PLN) * SK_NT + c) * PLN + (e % PLN);
}
template <class KT>
__device__ __forceinline__ void sk4_ldvec(const KT* arr, int c, int v, KT* out) {
  constexpr int PLN = 16 / (int)sizeof(KT);
  const uint4 x = *reinterpret_cast<const uint4*>(arr + (v * SK_NT + c) * PLN);
  if constexpr (sizeof(KT) == 4) { out[0] = (KT)x.x; out[1] = (KT)x.y; out[2] = (KT)x.z; out[3] = (KT)x.w; }
  else { out[0] = (KT)(((u64)x.y << 32) | x.x); out[1] = (KT)(((u64)x.w << 32) | x.z); }
}
template <class KT>
__device__ __forceinline__ void sk4_stvec(KT* arr, int c, int v, const KT* in) {
  constexpr int PLN = 16 / (int)sizeof(KT);
  uint4 x;
  if constexpr (sizeof(KT) == 4) { x.x = (u32)in[0]; x.y = (u32)in[1]; x.z = (u32)in[2]; x.w = (u32)in[3]; }
  else { x.x = (u32)in[0]; x.y = (u32)((u64)in[0] >> 32); x.z = (u32)in[1]; x.w = (u32)((u64)in[1] >> 32); }
  *reinterpret_cast<uint4*>(arr + (v * SK_NT + c) * PLN) = x;
}
// hash64 (sketch.rs:4-13) on a left-aligned key x = key << s, s = bits - 2k: arithmetic modulo 2^(2k) is the wrap of the
// register; ~key + (key << 21) = key * (2^21 - 1) - 1; key + (key << 3) + (key << 8) = key * 265; ... (key << 31) + key.
template <class KT>
__device__ __forceinline__ KT sk4_hash(KT x, int s) {
  const KT hm = (KT)~((((KT)1) << s) - 1);
  x = x * (KT)0x1FFFFFu - (((KT)1) << s);
  x ^= (x >> 24) & hm;
  x = x * (KT)265u;
  x ^= (x >> 14) & hm;
  x = x * (KT)21u;
  x ^= (x >> 28) & hm;
  if constexpr (sizeof(KT) == 8) x = x * (KT)0x80000001u;   // 2k <= 30: the shifted term lies outside the key
  return x;
}
// 16 ASCII bases -> 16 2-bit codes (A0 C1 G2 T3, nt4.rs:2-10) + "not ACGTacgt" bits.  Per word: bits 1-2 of a base are its
// code up to the G/T swap (fixed for all 16 at once); one multiply gathers the four codes of a word into its top byte;
// a base is valid iff the remaining bits are those of the letter its code names.
__device__ __forceinline__ void sk4_convert16(const u32 (&wd)[4], u32& packed, u32& nmask) {
  u32 p[4], bad[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const u32 s1 = wd[j] >> 1;
    p[j] = (s1 & 0x03030303u) * 0x01041040u;
    const u32 t = (wd[j] >> 2) & ~s1 & 0x01010101u;                        // code 2 before the swap: T / t
    bad[j] = ((wd[j] ^ 0x41414141u) & 0xD9D9D9D9u) ^ (t * 0x11u);
  }
  const u32 lo = __byte_perm(p[0], p[1], 0x0073), hi = __byte_perm(p[2], p[3], 0x0073);
  u32 pk = __byte_perm(lo, hi, 0x5410);
  pk ^= (pk >> 1) & 0x55555555u;                                          // A0 C1 T2 G3 -> A0 C1 G2 T3
  packed = pk;
  nmask = 0;
  if ((bad[0] | bad[1] | bad[2] | bad[3]) != 0u) {
#pragma unroll
    for (int j = 0; j < 4; ++j) nmask |= nbits4(__vcmpeq4(bad[j], 0u)) << (4 * j);
  }
}
__device__ __forceinline__ u32 sk4_nt4(u8 b) {   // nt4.rs:2-10
  const u32 u = b & 0xDFu;
  return u == 0x41u ? 0u : u == 0x43u ? 1u : u == 0x47u ? 2u : u == 0x54u ? 3u : 4u;
}
// key_span of the output (sketch.rs:74): 32-bit keys carry the hash only (the span is k); 64-bit keys carry the span in the
// 8 bits below the hash, so that (hash, span) compares like the reference's key_span (HPC: sketch.rs:51-61,72-74)
template <class KT>
__device__ __forceinline__ u64 sk4_out_key(KT key, int sla, int k) {
  if constexpr (sizeof(KT) == 8) return (u64)key >> (sla - 8);
  else return ((u64)(key >> sla) << 8) | (u64)k;
}
// HPC mode of THIS reference (sketch.rs:51-61): the base index is not advanced over a homopolymer run, so the k-mers are
// those of the plain sequence; only kmer_span changes: the sum, over the last k bases, of the length of the run that
// REMAINS from each base on.  s_same: bit t = "base t + 1 is the same letter" (raw indices).  Returns min(run, 256).
__device__ __forceinline__ u32 sk4_run_from(const u32* s_same, int t) {
  int wd = t >> 5;
  const int bit = t & 31;
  u32 x = __funnelshift_r(s_same[wd], s_same[wd + 1], bit);
  u32 ones = 0;
  while (x == 0xFFFFFFFFu && ones < 256u) { ones += 32u; ++wd; x = __funnelshift_r(s_same[wd], s_same[wd + 1], bit); }
  if (x != 0xFFFFFFFFu) ones += (u32)(__ffs(~x) - 1);
  return min(ones + 1u, 256u);
}
// spans of the 8 positions of a chunk: sliding sums of k run lengths (s_skip[32 + region index])
__device__ __forceinline__ void sk4_spans(const u16* s_skip, int c0, int k, int (&span)[SK_CH]) {
  const u16* sk = s_skip + 32 + c0;
  int sp = 0;
  for (int t = -(k - 1); t <= 0; ++t) sp += sk[t];
  span[0] = sp;
#pragma unroll
  for (int j = 1; j < SK_CH; ++j) { sp += (int)sk[j] - (int)sk[j - k]; span[j] = sp; }
}
// keys of a dirty chunk: the literal l counter (sketch.rs:63-76); returns zbits | ge_cap << 8 | eq_capm1 << 16
template <class KT, bool HPC>
__device__ __noinline__ u32 sk4_keys_slow(const u32* s_pack, const u32* s_nm, const u16* s_skip, KT* s_K, int tid, int r0, int k, int cap, int sla) {
  constexpr KT KMAX = (KT)~(KT)0;
  const KT mask = (KT)((((u64)1) << (2 * k)) - 1);
  const int shift1 = 2 * (k - 1);
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
  u32 zbits = 0, ge_cap = 0, eq_capm1 = 0;
  int span[SK_CH];
#pragma unroll
  for (int j = 0; j < SK_CH; ++j) span[j] = k;
  if constexpr (HPC) sk4_spans(s_skip, SK_CH * tid, k, span);
#pragma unroll
  for (int j = 0; j < SK_CH; ++j) {
    const u32 c = (cw >> (2 * j)) & 3u;
    l = ((nb >> j) & 1u) ? 0 : min(l + 1, cap);
    fwd = (KT)(((fwd << 2) | (KT)c) & mask);
    rev = (KT)((rev >> 2) | ((KT)(3u ^ c) << shift1));
    const bool z = !(fwd < rev);
    KT key = KMAX;
    if (l >= k && span[j] < 256) {
      key = sk4_hash<KT>((KT)((z ? rev : fwd) << sla), sla);
      if constexpr (sizeof(KT) == 8) key |= (KT)((u64)span[j] << (sla - 8));
    }
    s_K[sk4_idx<KT>(tid, j)] = key;
    zbits |= (u32)z << j;
    ge_cap |= (u32)(l >= cap) << j;
    eq_capm1 |= (u32)(l == cap - 1) << j;
  }
  return zbits | (ge_cap << 8) | (eq_capm1 << 16);
}
// the per-step rules of sketch.rs:80-99 for the 8 steps of chunk `tid`, every emission marked in the position bitmap
template <class KT>
__device__ __noinline__ void sk4_by_step(const KT* s_K, u32* s_emit, int tid, int w, int u_last, bool has_end, u32 gates) {
  constexpr KT KMAX = (KT)~(KT)0;
  const int c0 = SK_CH * tid;
  if (c0 > u_last || c0 + SK_CH <= w) return;
  const u32 ge_cap = gates & 0xFFu, eq_capm1 = (gates >> 8) & 0xFFu;
  auto keyat = [&](int t) -> KT { return t >= 0 ? s_K[sk4_idx<KT>(t >> 3, t & 7)] : KMAX; };
  auto mark = [&](int p) { atomicOr(&s_emit[p >> 5], 1u << (p & 31)); };
  auto mark_dups = [&](int lo, int hi, KT kv, int excl) {
    for (int x = lo; x <= hi; ++x)
      if (keyat(x) == kv && x != excl) mark(x);
  };
  // suffix minima over the w keys before the chunk, newest position winning ties; pd = pos << 1 | "minimum occurs twice"
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
  KT pk_prev = rk; int ppd_prev = rpd;   // window [c0-w, c0-1]
  {
    const KT kx = keyat(t);
    if (kx < rk) { pk_prev = kx; ppd_prev = t << 1; } else if (kx == rk) ppd_prev |= 1;
  }
  KT fk = KMAX; int fpd = 0;
#pragma unroll
  for (int j = 0; j < SK_CH; ++j) {
    const int u = c0 + j;
    const KT ki = keyat(u);
    if (ki <= fk) { fpd = (u << 1) | (ki == fk ? 1 : 0); fk = ki; }
    KT ck; int cpd;   // window [u-w+1, u] = older part (suffix) + newer part (prefix); the newer part wins ties
    if (fk <= Sk[j]) { ck = fk; cpd = fpd | (fk == Sk[j] ? 1 : 0); } else { ck = Sk[j]; cpd = Spd[j]; }
    const int ppos = ppd_prev >> 1;
    if (u >= w && u <= u_last) {
      const KT kp = pk_prev;
      const bool gc = (ge_cap >> j) & 1u, ec1 = (eq_capm1 >> j) & 1u;
      if (kp != KMAX) {
        if (ec1 && (ppd_prev & 1)) mark_dups(u - w + 1, u - 1, kp, ppos);            // sketch.rs:80-83
        if (ki <= kp) {
          if (gc) mark(ppos);                                                        // sketch.rs:84-86
        } else if (ppos == u - w) {
          if (gc || ec1) {
            mark(ppos);                                                              // sketch.rs:88
            if (ck != KMAX && (cpd & 1)) mark_dups(u - w + 1, u, ck, cpd >> 1);      // sketch.rs:92-95
          }
        }
      }
      if (has_end && u == u_last && ck != KMAX) mark(cpd >> 1);                      // sketch.rs:99
    }
    pk_prev = ck; ppd_prev = cpd;
  }
}

#ifndef MM2_SK4_OCC
#define MM2_SK4_OCC 4
#endif
#ifndef MM2_SK4_OCC64
#define MM2_SK4_OCC64 3
#endif
constexpr int SK4_BAR_INIT = 6;
// one tile, prepared by the scanner warp two to three iterations ahead (ticket + the dependent sequence-table reads)
struct __align__(16) Sk4Info {
  u32 tile, q;            // global tile index (0xFFFFFFFF: no more tiles), sequence
  i64 g0;                 // byte index in P.seq of raw base 0 (16-byte aligned, may be negative)
  i64 ps0;                // sequence position of raw base 0
  i64 len;                // sequence length
  u32 p0lo;               // low 32 bits of the sequence position of region index 0
  int delta, nchunks;     // alignment slack in front of the first needed base; 16-byte chunks to convert
  int nown;               // owned positions: region indices [w, w + nown)
  int u_last;             // region index of the sequence's last step (clamped to the region)
  u32 flags;              // 1: the region holds the last step; 2: first tile of the sequence; 4: fewer than w k-mers
  u64 tfirst;             // first tile of the sequence
};
struct Sk4Meta { u32 tile, count, q, pad; u64 tfirst; };
// The kernel: 8 compute warps + 1 scanner warp per CTA.
//  * compute warps: phases 1-3 of tile `it`; before the tile's records are staged in shared buffer it & 1, the records of
//    tile it - 2 (same buffer; its output offset has had two iterations to arrive) go to global memory with coalesced stores;
//  * scanner warp: decoupled look-back (inclusive prefix of the per-tile counts in tile order = the reference's emission
//    order) and the ticket + sequence-table reads of the tile three iterations ahead, so that no compute warp ever waits for a
//    dependent global load at the top of a tile.  scanner -> compute: a sequence number in shared memory (s_done[b]);
//    compute -> scanner: named barrier FULL[b].
template <class KT, int W, bool HPC>
__global__ void __launch_bounds__(SK_NT + 32, sizeof(KT) == 4 ? MM2_SK4_OCC : MM2_SK4_OCC64) sketch_tile_kernel_v4(SketchParams P) {
  static_assert(!HPC || sizeof(KT) == 8, "the span lives in the 8 bits below a 64-bit key");
  constexpr KT KMAX = (KT)~(KT)0;
  constexpr int PLN = 16 / (int)sizeof(KT), NPL = 8 / PLN;
  constexpr int KB = 8 * (int)sizeof(KT);
  // dynamic shared memory (the 64-bit instantiations need 73 KB)
  extern __shared__ __align__(16) unsigned char sk4_smem[];
  KT* const s_K = reinterpret_cast<KT*>(sk4_smem);              // keys (left-aligned hashes; KMAX = no k-mer)
  KT* const s_G = s_K + SK_REGION;                              // suffix minima inside each chunk
  KT* const s_P = s_G + SK_REGION;                              // prefix maxima of the window minima inside each chunk
  KT (*const s_okey)[SK_LIST] = reinterpret_cast<KT (*)[SK_LIST]>(s_P + SK_REGION);
  u32 (*const s_opos)[SK_LIST] = reinterpret_cast<u32 (*)[SK_LIST]>(s_okey + 2);
  __shared__ __align__(16) u32 s_pack[SK_MAXCHUNK + 4];
  __shared__ __align__(16) u32 s_nm[SK_MAXCHUNK / 2 + 4];
  __shared__ __align__(16) u32 s_nc[SK_NT / 32 + 4];            // per 16-base chunk: a real N inside the sequence
  __shared__ __align__(16) u32 s_same[HPC ? SK_MAXCHUNK / 2 + 12 : 1];   // HPC: bit t = base t + 1 is the same letter as base t
  __shared__ __align__(16) u16 s_skip[HPC ? SK_REGION + 32 : 8];         // HPC: min(remaining run, 256) of region index u at [32 + u]
  __shared__ __align__(16) u32 s_emit[SK_REGION / 32];
  __shared__ __align__(16) u32 s_dirty[SK_NT / 32 + 4];         // [0] = 0: "the warp before warp 0"; [1 + warp]
  __shared__ u32 s_wsum[SK_NT / 32];
  __shared__ Sk4Info s_info[3];
  __shared__ u64 s_excl[2];
  __shared__ u32 s_done[2];                                      // iteration + 1 of the last tile whose offset is in s_excl[b]
  __shared__ u32 s_ninfo;                                        // iterations whose s_info slot is filled
  __shared__ Sk4Meta s_meta[2];

  const int tid = threadIdx.x;
  const int w = W ? W : P.w, k = P.k;
  const int cap = w + k;
  const int T = SK_REGION - 2 * w;
  const int sla = KB - 2 * k;

  if (tid >= SK_NT) {
    // ---- scanner warp ------------------------------------------------------------------------------------------------------
    const int lane = tid - SK_NT;
    volatile u64* st = P.tile_status;
    auto fetch_info = [&](int slot) {                          // lane 0
      Sk4Info inf;
      memset(&inf, 0, sizeof inf);
      inf.tile = P.tile_base + atomicAdd(P.ticket, 1u);
      if (inf.tile < P.ntiles) {
        inf.q = P.tile_seq[inf.tile];
        const u64 soff = P.seq_off[inf.q];
        inf.len = (i64)(P.seq_off[inf.q + 1] - soff);
        inf.tfirst = P.tile_first[inf.q];
        const i64 s = (i64)((u64)inf.tile - inf.tfirst) * T;  // first owned position
        const i64 P0 = s - w;                                 // sequence position of region index 0
        const i64 a = P0 - cap;                               // first needed base
        const i64 gidx = (i64)soff + a;
        inf.g0 = (gidx >> 4) << 4;
        inf.delta = (int)(gidx - inf.g0);
        inf.ps0 = a - inf.delta;
        inf.nchunks = (inf.delta + SK_REGION + cap + 15 + (HPC ? 256 : 0)) >> 4;   // HPC: a run may go on for 255 bases past the region
        inf.nown = (int)max((i64)0, min((i64)T, inf.len - s));
        inf.u_last = (int)max((i64)-1, min((i64)(SK_REGION - 1), inf.len - 1 - P0));
        inf.p0lo = (u32)P0;
        inf.flags = ((inf.len - 1 - P0) <= (i64)(SK_REGION - 1) ? 1u : 0u) | (s == 0 ? 2u : 0u) | (inf.len - (k - 1) < (i64)w ? 4u : 0u);
      } else inf.tile = 0xFFFFFFFFu;
      s_info[slot] = inf;
    };
    // Tickets are taken one per iteration, two iterations ahead of the compute warps, so that ticket order stays close to
    // the order in which tiles are finished (a CTA that grabbed several consecutive tiles at once would publish their
    // counts iterations apart, and every later tile's look-back would wait for them).
    if (lane == 0) { s_done[0] = 0; s_done[1] = 0; fetch_info(0); __threadfence_block(); *(volatile u32*)&s_ninfo = 1u; }
    sk3_bar_arrive(SK4_BAR_INIT);
    if (lane == 0) { fetch_info(1); __threadfence_block(); *(volatile u32*)&s_ninfo = 2u; }
    for (u32 it = 0;; ++it) {
      const int b = (int)(it & 1u);
      sk3_bar_sync(SK3_BAR_FULL + b);
      const u32 tile = s_meta[b].tile;
      if (tile == 0xFFFFFFFFu) break;
      const u32 tile_count = s_meta[b].count, q = s_meta[b].q;
      if (lane == 0) { fetch_info((int)((it + 2u) % 3u)); __threadfence_block(); *(volatile u32*)&s_ninfo = it + 3u; }
      u64 excl = 0;
      if (tile != 0) {
        // look-back, 128 tiles per round (4 per lane, newest first): hundreds of tiles are in flight with only their
        // aggregate published, and a round costs one L2 round trip whatever its width
        i64 look = (i64)tile - 1;
        for (;;) {
          u64 mine = 0; bool incl = false;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const i64 idx = look - 4 * lane - e;
            u64 v = (2ULL << 62);
            if (!incl && idx >= 0) { while (((v = st[idx]) >> 62) == 0) __nanosleep(32); }
            if (!incl) { mine += v & ((1ULL << 62) - 1); incl = (v >> 62) == 2; }
          }
          const u32 incl_mask = __ballot_sync(0xFFFFFFFFu, incl);
          const int first_incl = incl_mask ? (__ffs(incl_mask) - 1) : 32;
          u64 contrib = (lane <= first_incl) ? mine : 0;
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xFFFFFFFFu, contrib, d);
          excl += contrib;
          if (incl_mask) break;
          look -= 128;
        }
        if (lane == 0) st[tile] = (2ULL << 62) | (excl + (u64)tile_count);
      }
      if (lane == 0) {
        s_excl[b] = excl;
        if ((u64)tile == s_meta[b].tfirst) P.seq_out_off[q] = excl;
        if (tile == P.ntiles - 1) P.seq_out_off[P.nseq] = excl + (u64)tile_count;
        __threadfence_block();
        *(volatile u32*)&s_done[b] = it + 1u;
      }
    }
    return;
  }

  const int q8 = (w - 1) >> 3, r8 = (w - 1) & 7;   // w - 1 = 8 q8 + r8
  const int c0 = SK_CH * tid;
  if (tid == 0) s_dirty[0] = 0u;
  sk3_bar_sync(SK4_BAR_INIT);                      // s_done, s_ninfo and s_info[0] are there
  // staged, not yet written records: A = tile it - 2 (buffer it & 1), B = tile it - 1 (the other buffer); count 0 = nothing
  u32 cntA = 0, cntB = 0, qA = 0, qB = 0;
  auto drain = [&](int pb, u32 cnt, u32 qq, u32 seq) {         // staged records of an earlier tile -> global, coalesced
    if (cnt == 0) return;
    const u64 rid_hi = (u64)(P.rid_base + qq * P.rid_step) << 32;
    while (*(volatile u32*)&s_done[pb] < seq) __nanosleep(64);
    __threadfence_block();
    const u64 base = *(volatile u64*)&s_excl[pb];
    const KT* ok = s_okey[pb];
    const u32* op = s_opos[pb];
    if (base + cnt <= P.out_cap) {
      // one pointer per array and thread, the (at most SK_LIST / SK_NT) records of a thread at immediate offsets
      u64* dk = P.out_key + base + tid;
      u64* dv = P.out_val + base + tid;
      const KT* okt = ok + tid;
      const u32* opt = op + tid;
#pragma unroll
      for (int r = 0; r < SK_LIST / SK_NT; ++r) {
        if ((u32)tid + (u32)(SK_NT * r) < cnt) {
          dk[SK_NT * r] = sk4_out_key<KT>(okt[SK_NT * r], sla, k);
          dv[SK_NT * r] = rid_hi | (u64)opt[SK_NT * r];
        }
      }
    } else {
      for (u32 e2 = (u32)tid; e2 < cnt; e2 += SK_NT) {
        const u64 o = base + e2;
        if (o < P.out_cap) { P.out_key[o] = sk4_out_key<KT>(ok[e2], sla, k); P.out_val[o] = rid_hi | (u64)op[e2]; }
      }
    }
  };
  u32 it = 0;
  for (;; ++it) {
    const int b = (int)(it & 1u);
    while (*(volatile u32*)&s_ninfo < it + 1u) __nanosleep(64);
    __threadfence_block();
    const Sk4Info& inf = s_info[it % 3u];
    const u32 tile = *(volatile const u32*)&inf.tile;
    if (tile == 0xFFFFFFFFu) break;
    // (the fields are re-read from shared memory where they are used: the slot stays untouched until FULL[b] is signalled)
    const u32 tflags = inf.flags;
    const bool first_tile = (tflags & 2u) != 0u;

    // ---- phase 1: 128-bit loads -> 2-bit packed codes + N mask -----------------------------------------------------------
    if (tid < SK_REGION / 32) s_emit[tid] = 0u;
    {
      const int c = tid;                                      // SK_MAXCHUNK + 4 <= SK_NT: one chunk per thread
      u32 packed = 0, nmask = 0xFFFFu, nreal = 0;
      const i64 g0 = inf.g0, ps0 = inf.ps0, len = inf.len;
      if (c < inf.nchunks) {
        const i64 gi = g0 + 16 * (i64)c;
        u32 wd[4] = {0, 0, 0, 0};
        if (P.vec_ok && gi >= 0 && gi + 16 <= (i64)P.buf_len) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(P.seq + gi));
          wd[0] = v.x; wd[1] = v.y; wd[2] = v.z; wd[3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const i64 g = gi + j;
            u32 bb = (g >= 0 && g < (i64)P.buf_len) ? (u32)P.seq[g] : 0u;
            wd[j >> 2] |= bb << (8 * (j & 3));
          }
        }
        sk4_convert16(wd, packed, nmask);
        nreal = nmask;
        const i64 pstart = ps0 + 16 * (i64)c;                 // bases outside [0, len) of this sequence count as N for the l counter
        if (pstart < 0 || pstart + 16 > len) {
          const i64 lo = max((i64)0, -pstart), hi = min((i64)16, len - pstart);
          u32 inseq = 0;
          if (hi > lo) inseq = ((hi >= 16 ? 0x10000u : (1u << hi)) - 1u) & ~((1u << lo) - 1u);
          nmask = (nmask | ~inseq) & 0xFFFFu;
          nreal &= inseq;
        }
      }
      if (c < SK_MAXCHUNK + 4) { s_pack[c] = packed; reinterpret_cast<u16*>(s_nm)[c] = (u16)nmask; }
      if constexpr (HPC) {
        // "the next base is the same letter" for the 16 bases of this chunk; the 17th base comes from the next chunk's first byte
        u32 same = 0;
        if (c < inf.nchunks) {
          const i64 gn = g0 + 16 * (i64)c + 16;
          const u32 nx = (gn >= 0 && gn < (i64)P.buf_len && (ps0 + 16 * (i64)c + 16) < len && (ps0 + 16 * (i64)c + 16) >= 0) ? sk4_nt4(P.seq[gn]) : 4u;
          const u32 x = packed ^ ((packed >> 2) | ((nx & 3u) << 30));       // 2-bit field t is zero iff code t == code t + 1
          u32 y = ~(x | (x >> 1)) & 0x55555555u;
          y = (y | (y >> 1)) & 0x33333333u; y = (y | (y >> 2)) & 0x0F0F0F0Fu; y = (y | (y >> 4)) & 0x00FF00FFu; y = (y | (y >> 8)) & 0x0000FFFFu;
          const u32 nnext = (nmask >> 1) | ((nx > 3u ? 1u : 0u) << 15);     // N mask of the bases t + 1
          same = y & ~nmask & ~nnext & 0xFFFFu;
        }
        if (c < SK_MAXCHUNK + 24) reinterpret_cast<u16*>(s_same)[c] = (u16)same;
      }
      const u32 nb = __ballot_sync(0xFFFFFFFFu, nreal != 0u);
      if ((tid & 31) == 0) s_nc[tid >> 5] = nb;
    }
    sk3_bar_compute();
    if constexpr (HPC) {
      // remaining run length of every base of the region (and of the k - 1 bases in front of it)
      const int rr = cap + inf.delta;                         // raw index of region index 0
      {
        u16 sk8[SK_CH];
#pragma unroll
        for (int j = 0; j < SK_CH; ++j) sk8[j] = (u16)sk4_run_from(s_same, rr + c0 + j);
        *reinterpret_cast<uint4*>(&s_skip[32 + c0]) = *reinterpret_cast<const uint4*>(sk8);
      }
      if (tid < 32) { const int t = rr - 32 + tid; s_skip[tid] = (u16)(t >= 0 ? sk4_run_from(s_same, t) : 1u); }
      sk3_bar_compute();
    }

    // ---- phase 2: keys of this thread's 8 positions ----------------------------------------------------------------------
    KT K[SK_CH];
    u32 zbits = 0, gates = 0xFFu;
    const int r0 = cap + c0 + inf.delta;                      // raw index of the base at region index c0
    bool dirty;                                               // a real N in bases [c-w-k+1, c+7+w], at 16-base granularity
    {
      const int clo = (r0 - (cap - 1)) >> 4, chi = (r0 + min(SK_CH - 1 + w, SK_REGION - 1 - c0)) >> 4;
      const int nch = chi - clo + 1;
      if (nch <= 32) {
        const u32 x = __funnelshift_r(s_nc[clo >> 5], s_nc[(clo >> 5) + 1], clo & 31);
        dirty = (x & (0xFFFFFFFFu >> (32 - nch))) != 0u;
      } else {
        dirty = false;
        for (int c = clo; c <= chi; ++c) dirty = dirty || ((s_nc[c >> 5] >> (c & 31)) & 1u);
      }
    }
    if (!dirty) {
      const int rs = r0 - k + 1;                              // first base of the first k-mer
      const int wi = rs >> 4, sh = 2 * (rs & 15);
      const KT hm = (KT)~((((KT)1) << sla) - 1);
      if constexpr (sizeof(KT) == 4) {
        const u32 w0 = s_pack[wi], w1 = s_pack[wi + 1], w2 = s_pack[wi + 2];
        const u32 f0 = __funnelshift_r(w0, w1, sh), f1 = __funnelshift_r(w1, w2, sh);   // 32 bases from rs, base b at bits 2b
        const u32 n0 = ~f0, n1 = ~f1;
        u32 g1 = __brev(f0), g0b = __brev(f1);                                           // pair-reversed: base b at bits 62-2b
        g1 = ((g1 & 0x55555555u) << 1) | ((g1 >> 1) & 0x55555555u);
        g0b = ((g0b & 0x55555555u) << 1) | ((g0b >> 1) & 0x55555555u);
#pragma unroll
        for (int j = 0; j < SK_CH; ++j) {
          const u32 rv = __funnelshift_r(n0, n1, 2 * j) << sla;          // sketch.rs:66 kmer[1], left-aligned
          const u32 fw = __funnelshift_l(g0b, g1, 2 * j) & hm;           // sketch.rs:65 kmer[0], left-aligned
          const bool z = fw > rv;                                       // odd k: never equal
          K[j] = sk4_hash<KT>(z ? rv : fw, sla);
          zbits |= (u32)z << j;
        }
      } else {
        const u32 w0 = s_pack[wi], w1 = s_pack[wi + 1], w2 = s_pack[wi + 2], w3 = s_pack[wi + 3];
        const u32 f0 = __funnelshift_r(w0, w1, sh), f1 = __funnelshift_r(w1, w2, sh), f2 = __funnelshift_r(w2, w3, sh);   // 48 bases
        const u32 n0 = ~f0, n1 = ~f1, n2 = ~f2;
        u32 g2 = __brev(f0), g1 = __brev(f1), g0b = __brev(f2);
        g2 = ((g2 & 0x55555555u) << 1) | ((g2 >> 1) & 0x55555555u);
        g1 = ((g1 & 0x55555555u) << 1) | ((g1 >> 1) & 0x55555555u);
        g0b = ((g0b & 0x55555555u) << 1) | ((g0b >> 1) & 0x55555555u);
#pragma unroll
        for (int j = 0; j < SK_CH; ++j) {
          const u32 vl = __funnelshift_r(n0, n1, 2 * j), vh = __funnelshift_r(n1, n2, 2 * j);
          const u64 rv = (((u64)vh << 32) | (u64)vl) << sla;                                      // sla in [8, 62] (HPC uses 64-bit keys for every k)
          const u64 fw = (((u64)__funnelshift_l(g1, g2, 2 * j) << 32) | (u64)__funnelshift_l(g0b, g1, 2 * j)) & (u64)hm;
          const bool z = fw > rv;
          K[j] = sk4_hash<KT>((KT)(z ? rv : fw), sla);
          zbits |= (u32)z << j;
        }
        if constexpr (HPC) {
          int span[SK_CH];
          sk4_spans(s_skip, c0, k, span);
#pragma unroll
          for (int j = 0; j < SK_CH; ++j) K[j] = span[j] < 256 ? (KT)(K[j] | ((u64)span[j] << (sla - 8))) : KMAX;
        } else {
#pragma unroll
          for (int j = 0; j < SK_CH; ++j) K[j] |= (KT)((u64)k << (sla - 8));
        }
      }
      if (first_tile && c0 < w + k - 1) {                     // no k-mer ends before position k - 1 (region index w + k - 1)
#pragma unroll
        for (int j = 0; j < SK_CH; ++j) if (c0 + j < w + k - 1) K[j] = KMAX;
      }
#pragma unroll
      for (int v = 0; v < NPL; ++v) sk4_stvec<KT>(s_K, tid, v, &K[v * PLN]);
    } else {
      const u32 g = sk4_keys_slow<KT, HPC>(s_pack, s_nm, s_skip, s_K, tid, r0, k, cap, sla);
      zbits = g & 0xFFu; gates = g >> 8;
#pragma unroll
      for (int v = 0; v < NPL; ++v) sk4_ldvec<KT>(s_K, tid, v, &K[v * PLN]);
    }
    KT F[SK_CH];                                              // prefix minima inside the chunk
    {
      KT G[SK_CH];                                            // suffix minima inside the chunk
      F[0] = K[0]; G[SK_CH - 1] = K[SK_CH - 1];
#pragma unroll
      for (int j = 1; j < SK_CH; ++j) { F[j] = sk_min(F[j - 1], K[j]); G[SK_CH - 1 - j] = sk_min(G[SK_CH - j], K[SK_CH - 1 - j]); }
#pragma unroll
      for (int v = 0; v < NPL; ++v) sk4_stvec<KT>(s_G, tid, v, &G[v * PLN]);
    }
    {
      const u32 db = __ballot_sync(0xFFFFFFFFu, dirty);
      if ((tid & 31) == 0) s_dirty[(tid >> 5) + 1] = db;
    }
    sk3_bar_compute();

    // ---- phase 3a: window minima M[u] = min K[u-w+1 .. u] of this thread's 8 windows -----------------------------------------
    if (first_tile && tid < 32) {
      // The first w keys of a sequence: sketch.rs:80-86 treats ties among them specially (duplicates of the first partial
      // minimum are emitted at l == w+k-1; a minimum replaced before l reaches w+k is not).  With an equal pair, a missing
      // k-mer or fewer than w k-mers, the chunks that see the sequence start fall back to the by-step rules.
      bool tie = (tflags & 4u) != 0u || w > 32;
      if (!tie) {
        const int x = w + k - 1 + tid;                        // region index of key number `tid`
        const KT kx = tid < w ? s_K[sk4_idx<KT>(x >> 3, x & 7)] : (KT)0;
        bool t2 = tid < w && kx == KMAX;
        for (int d = 1; d < w; ++d) {
          const KT other = (KT)__shfl_down_sync(0xFFFFFFFFu, kx, d);
          t2 = t2 || (tid + d < w && other == kx);
        }
        tie = __any_sync(0xFFFFFFFFu, t2);
      }
      if (tie && tid == 0) {                                  // chunks with c0 - w - k + 1 < w, i.e. c0 < 2w + k - 1
        const int last = (2 * w + k - 2) >> 3;
        for (int t = 0; t <= last && t < SK_NT; t += 32) s_dirty[(t >> 5) + 1] |= (last - t >= 31) ? 0xFFFFFFFFu : ((2u << (last - t)) - 1u);
      }
    }
    KT M[SK_CH];
    if constexpr (W != 0) {
      KT g[16];
#pragma unroll
      for (int x = 0; x < 16; ++x) g[x] = KMAX;
      const int ca = max(tid - q8 - 1, 0), cb = max(tid - q8, 0);
#pragma unroll
      for (int v = 0; v < NPL; ++v) {
        if (r8 > 0 && (v + 1) * PLN - 1 >= 8 - r8) sk4_ldvec<KT>(s_G, ca, v, &g[v * PLN]);
        if (v == 0 || v * PLN <= 7 - r8) sk4_ldvec<KT>(s_G, cb, v, &g[8 + v * PLN]);
      }
      KT common = KMAX;
#pragma unroll
      for (int d = 1; d < q8; ++d) common = sk_min(common, s_G[sk4_idx<KT>(max(tid - d, 0), 0)]);
#pragma unroll
      for (int j = 0; j < SK_CH; ++j) {
        KT m = sk_min(F[j], g[j + 8 - r8]);
        if (q8 > 1) m = sk_min(m, common);
        if (j < r8) m = sk_min(m, g[8]);
        M[j] = m;
      }
    } else {
#pragma unroll
      for (int j = 0; j < SK_CH; ++j) {
        const int p = max(c0 + j - (w - 1), 0);
        const int ch = p >> 3;
        KT m = sk_min(F[j], s_G[sk4_idx<KT>(ch, p & 7)]);
        for (int c = ch + 1; c < tid; ++c) m = sk_min(m, s_G[sk4_idx<KT>(c, 0)]);
        M[j] = m;
      }
    }
    if (tflags & 3u) {
      // windows that start before the first k-mer or end after the last one do not exist: below every key
      const int u_first = first_tile ? 2 * w + k - 2 : 0, u_last = inf.u_last;
#pragma unroll
      for (int j = 0; j < SK_CH; ++j) if (c0 + j < u_first || c0 + j > u_last) M[j] = 0;
    }
    KT GM[SK_CH];                                             // suffix maxima of M inside the chunk
    {
      KT PM[SK_CH];                                           // prefix maxima of M inside the chunk
      PM[0] = M[0]; GM[SK_CH - 1] = M[SK_CH - 1];
#pragma unroll
      for (int j = 1; j < SK_CH; ++j) { PM[j] = sk_max(PM[j - 1], M[j]); GM[SK_CH - 1 - j] = sk_max(GM[SK_CH - j], M[SK_CH - 1 - j]); }
#pragma unroll
      for (int v = 0; v < NPL; ++v) sk4_stvec<KT>(s_P, tid, v, &PM[v * PLN]);
    }
    sk3_bar_compute();

    // ---- phase 3b: D[x] = max M[x .. x+w-1]; x is a minimizer iff D[x] == K[x] ------------------------------------------------
    u32 flags = 0;
    if constexpr (W != 0) {
      KT h[16];
#pragma unroll
      for (int x = 0; x < 16; ++x) h[x] = 0;
      const int ca = min(tid + q8, SK_NT - 1), cb = min(tid + q8 + 1, SK_NT - 1);
#pragma unroll
      for (int v = 0; v < NPL; ++v) {
        if (v == NPL - 1 || (v + 1) * PLN - 1 >= r8) sk4_ldvec<KT>(s_P, ca, v, &h[v * PLN]);
        if (r8 > 0 && v * PLN <= r8 - 1) sk4_ldvec<KT>(s_P, cb, v, &h[8 + v * PLN]);
      }
      KT common = 0;
#pragma unroll
      for (int d = 1; d < q8; ++d) common = sk_max(common, s_P[sk4_idx<KT>(min(tid + d, SK_NT - 1), 7)]);
#pragma unroll
      for (int j = 0; j < SK_CH; ++j) {
        KT d = sk_max(GM[j], h[j + r8]);
        if (q8 > 1) d = sk_max(d, common);
        if (j + r8 >= 8) d = sk_max(d, h[7]);
        flags |= (u32)(d == K[j] && (!HPC || K[j] != KMAX)) << j;   // HPC: a k-mer whose span reaches 256 is no k-mer, with l unchanged
      }
    } else {
#pragma unroll
      for (int j = 0; j < SK_CH; ++j) {
        const int p = min(c0 + j + (w - 1), SK_REGION - 1);
        const int ch = p >> 3;
        KT d = sk_max(GM[j], s_P[sk4_idx<KT>(ch, p & 7)]);
        for (int c = tid + 1; c < ch; ++c) d = sk_max(d, s_P[sk4_idx<KT>(c, 7)]);
        flags |= (u32)(d == K[j] && (!HPC || K[j] != KMAX)) << j;   // HPC: a k-mer whose span reaches 256 is no k-mer, with l unchanged
      }
    }
    {
      const uint4 d0 = *reinterpret_cast<const uint4*>(&s_dirty[0]), d1 = *reinterpret_cast<const uint4*>(&s_dirty[4]);
      if ((d0.y | d0.z | d0.w | d1.x | d1.y | d1.z | d1.w | s_dirty[8]) != 0u) {   // uniform over the CTA; rare
        // by-step marking by every thread within dn = ceil((w+7)/8) <= 32 chunks after a dirty one (a position is emitted at
        // most w steps after its own): dirty bits of threads [tid - dn, tid] in the 64-bit window (previous warp : own warp)
        const int dn = (w + 7) >> 3;
        const u64 win = ((u64)s_dirty[(tid >> 5) + 1] << 32) | (u64)s_dirty[tid >> 5];
        const u64 sel = (dn >= 32 ? 0x1FFFFFFFFULL : ((1ULL << (dn + 1)) - 1ULL)) << ((tid & 31) + 32 - dn);
        const bool mydirty = ((win >> ((tid & 31) + 32)) & 1ULL) != 0ULL;
        if (win & sel) {
          if (mydirty && !dirty) {
            // start-of-sequence fallback of a chunk without N: l = position + 1 at every step (no N since the sequence start)
            u32 ge = 0, eq = 0;
#pragma unroll
            for (int j = 0; j < SK_CH; ++j) {
              const int l = (int)(inf.p0lo + (u32)(c0 + j)) + 1;   // first tile: small numbers
              ge |= (u32)(l >= cap) << j; eq |= (u32)(l == cap - 1) << j;
            }
            gates = ge | (eq << 8);
          }
          sk4_by_step<KT>(s_K, s_emit, tid, w, inf.u_last, (tflags & 1u) != 0u, gates);
        }
        sk3_bar_compute();
        if (mydirty) flags = (s_emit[c0 >> 5] >> (c0 & 31)) & 0xFFu;
      }
    }
    const int nown = inf.nown;
    if (c0 < w || c0 + SK_CH > w + nown) {   // positions this tile owns: region indices [w, w + nown)
      const int lo = w - c0, hi = w + nown - c0;
      u32 om = 0xFFu;
      if (lo > 0) om &= lo >= 8 ? 0u : (0xFFu << lo);
      if (hi < 8) om &= hi <= 0 ? 0u : ((1u << hi) - 1u);
      flags &= om & 0xFFu;
    }
    const u32 tot = (u32)__popc(flags);

    // ---- the records of tile it - 2 leave buffer b (its offset has had two iterations to arrive) ------------------------------
    drain(b, cntA, qA, it - 1u);

    // ---- exclusive scan of the per-thread counts; the tile's count is published at once --------------------------------------
    u32 inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u32 tt = __shfl_up_sync(0xFFFFFFFFu, inc, d);
      if ((tid & 31) >= d) inc += tt;
    }
    if ((tid & 31) == 31) s_wsum[tid >> 5] = inc;
    sk3_bar_compute();
    u32 wv = (tid & 31) < SK_NT / 32 ? s_wsum[tid & 31] : 0u;    // inclusive scan of the warp sums, in every warp
#pragma unroll
    for (int d = 1; d < SK_NT / 32; d <<= 1) {
      const u32 tt = __shfl_up_sync(0xFFFFFFFFu, wv, d);
      if ((tid & 31) >= d) wv += tt;
    }
    const u32 tile_count = __shfl_sync(0xFFFFFFFFu, wv, SK_NT / 32 - 1);
    const u32 wprev = __shfl_sync(0xFFFFFFFFu, wv, ((tid >> 5) + 31) & 31);
    const u32 my_off = ((tid >> 5) ? wprev : 0u) + inc - tot;
    const bool staged = tile_count <= (u32)SK_LIST;
    const u32 q = inf.q;
    if (tid == 0) {
      volatile u64* st = P.tile_status;
      st[tile] = ((tile == 0 ? 2ULL : 1ULL) << 62) | (u64)tile_count;   // tile 0: its inclusive prefix; others: aggregate
      s_meta[b].tile = tile; s_meta[b].count = tile_count; s_meta[b].q = q; s_meta[b].tfirst = inf.tfirst;
    }

    // ---- records of this thread's flagged positions, in position order, into buffer b ------------------------------------------
    const u32 pbase = inf.p0lo + (u32)c0;
    if (staged && flags) {
      KT* pk = s_okey[b] + my_off;
      u32* pp = s_opos[b] + my_off;
      const u32 pb2 = pbase << 1;
#pragma unroll
      for (int j = 0; j < SK_CH; ++j) {
        if ((flags >> j) & 1u) {
          *pk++ = K[j];
          *pp++ = (pb2 + 2u * (u32)j) | ((zbits >> j) & 1u);
        }
      }
    }
    __threadfence_block();
    sk3_bar_arrive(SK3_BAR_FULL + b);
    cntA = cntB; qA = qB;
    if (staged) { cntB = tile_count; qB = q; }
    else {
      // a tile that emits more than a buffer holds (windows full of repeated minima): wait for its own offset and let each
      // thread write its records
      cntB = 0;
      const u64 rid_hi = (u64)(P.rid_base + q * P.rid_step) << 32;
      while (*(volatile u32*)&s_done[b] < it + 1u) __nanosleep(64);
      __threadfence_block();
      u64 o = *(volatile u64*)&s_excl[b] + my_off;
#pragma unroll
      for (int j = 0; j < SK_CH; ++j) {
        if ((flags >> j) & 1u) {
          if (o < P.out_cap) {
            P.out_key[o] = sk4_out_key<KT>(K[j], sla, k);
            P.out_val[o] = rid_hi | (u64)(((pbase + (u32)j) << 1) | ((zbits >> j) & 1u));
          }
          ++o;
        }
      }
    }
  }
  {  // the records still staged (tile it - 2 in buffer it & 1, tile it - 1 in the other), then tell the scanner to stop
    const int b = (int)(it & 1u);
    drain(b, cntA, qA, it - 1u);
    drain(b ^ 1, cntB, qB, it);
    if (tid == 0) s_meta[b].tile = 0xFFFFFFFFu;
    __threadfence_block();
    sk3_bar_arrive(SK3_BAR_FULL + b);
  }
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

// Tiles of the odd-k / non-HPC kernels, at least one per sequence: T = SK_REGION - 2w positions each (v4, w >= 9) or
// SK_REGION - w steps (small windows).
static inline int sk_tile_T(int w) { return w >= 9 ? SK_REGION - 2 * w : SK_REGION - w; }
// which sequences go through the tile kernels: odd k (even k has palindromic k-mers that stall l); HPC only in the w >= 9 kernel
// and while the 255-base run look-ahead fits the chunk table
bool sketch_uses_tiles(int w, int k, int is_hpc) { return (k & 1) && (!is_hpc || (w >= 9 && w <= 64)); }
u64 sketch_tile_count(const u64* h_off, size_t nseq, int w) {
  const u64 T = (u64)sk_tile_T(w);
  u64 nt = 0;
  for (size_t i = 0; i < nseq; ++i) nt += std::max<u64>(1, (h_off[i + 1] - h_off[i] + T - 1) / T);
  return nt;
}
// bytes of the concatenated sequences (relative to h_off[0]) that the tiles [tile_lo, tile_hi) read, halo included
void sketch_tile_bytes(const u64* h_off, size_t nseq, int w, int k, int is_hpc, u64 tile_lo, u64 tile_hi, u64* byte_lo, u64* byte_hi) {
  const u64 T = (u64)sk_tile_T(w), total = h_off[nseq] - h_off[0];
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
      hi = std::max(hi, std::min(total, e + (u64)w + 32 + (is_hpc ? 272 : 0)));   // a v4 region extends w positions above the owned range (HPC: + the run look-ahead)
    }
    t0 += ntq;
    if (t0 >= tile_hi) break;
  }
  if (lo == ~0ULL) return;
  *byte_lo = lo & ~(u64)15; *byte_hi = hi;
}

namespace {
// "inclusive prefix 0" in the 128 status slots in front of a shard's first tile (one look-back window: every lane must
// find a published value; the w < 9 kernel looks at 32 of them)
__global__ void sketch_seed_prefix_kernel(u64* st, u32 t_lo) { if (threadIdx.x < t_lo && threadIdx.x < 128) st[t_lo - 1 - threadIdx.x] = 2ULL << 62; }
}

// w >= 9: the window sizes of the usual presets get their own instantiation (every shared-memory offset an immediate)
template <class KT, int W, bool HPC>
static int launch_sketch_v4_t(mm2_ctx* ctx, int grid, const SketchParams& P) {
  const int smem = (int)(3 * SK_REGION * sizeof(KT) + 2 * SK_LIST * (sizeof(KT) + 4));
  // the opt-in is per device and cheap: set it on every launch (a context may live on any GPU of the process)
  CUDA_TRY(cudaFuncSetAttribute(sketch_tile_kernel_v4<KT, W, HPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CUDA_TRY(cudaFuncSetAttribute(sketch_tile_kernel_v4<KT, W, HPC>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  MM2_LAUNCH(ctx, (sketch_tile_kernel_v4<KT, W, HPC>), grid, SK_NT + 32, smem, P);
  return MM2_OK;
}
static int launch_sketch_v4(mm2_ctx* ctx, int grid, const SketchParams& P, int is_hpc) {
  const bool k32 = P.k <= 15;
  if (is_hpc) return P.w == 10 ? launch_sketch_v4_t<u64, 10, true>(ctx, grid, P) : launch_sketch_v4_t<u64, 0, true>(ctx, grid, P);
  if (P.w == 10) return k32 ? launch_sketch_v4_t<u32, 10, false>(ctx, grid, P) : launch_sketch_v4_t<u64, 10, false>(ctx, grid, P);
  if (P.w == 19) return k32 ? launch_sketch_v4_t<u32, 19, false>(ctx, grid, P) : launch_sketch_v4_t<u64, 19, false>(ctx, grid, P);
  return k32 ? launch_sketch_v4_t<u32, 0, false>(ctx, grid, P) : launch_sketch_v4_t<u64, 0, false>(ctx, grid, P);
}

int sketch_device(mm2_ctx* ctx, const u8* d_cat, const u64* d_off, const u64* h_off, size_t nseq, int w, int k,
                  u32 rid_base, u32 rid_step, int is_hpc, SketchOut* out, const SketchFeed* feed, const SketchShard* shard) {
  if (!(w > 0 && w < 256) || !(k > 0 && k <= 28)) { mm2_set_error("sketch: need 0<w<256 and 0<k<=28 (sketch.rs:31-32)"); return MM2_E_ARG; }
  if (nseq == 0) { out->key = out->val = nullptr; out->seq_off = nullptr; out->total = 0; return MM2_OK; }
  if (nseq >= 0xFFFFFFFFull) { mm2_set_error("sketch: too many sequences"); return MM2_E_ARG; }
  const u64 total_len = h_off[nseq] - h_off[0];
  cudaStream_t st = ctx->stream;
  MM2_TRY(ctx->mini_off.ensure((nseq + 1) * 8));
  const bool tile_path = sketch_uses_tiles(w, k, is_hpc);
  if (shard && !tile_path) { mm2_set_error("sketch: tile shards need the tile kernels (odd k; HPC: 9 <= w <= 64)"); return MM2_E_ARG; }
  if (tile_path) {
    const int T = sk_tile_T(w);
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
      if (t_lo > 0) MM2_LAUNCH(ctx, sketch_seed_prefix_kernel, 1, 128, 0, ctx->tile_status.as<u64>(), t_lo);
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
      const int grid = (int)std::min<u64>(t_hi - t_lo, (u64)num_sms(ctx->device) * (MM2_SK4_OCC + 2));
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
              const u64 mg = 16 + (u64)w + (is_hpc ? 272 : 0);                      // region end + alignment (+ the run look-ahead of HPC)
              if (so + len + mg <= B) { t_end += ntq; continue; }
              if (B > so + mg) t_end += std::min<u64>(ntq - 1, (B - so - mg) / (u64)T);   // tiles whose region ends (with the margin) below B
              break;                                                                   // later sequences lie above B
            }
          }
          CUDA_TRY(cudaStreamWaitEvent(st, feed->ev[c], 0));
          if (t_end > t_prev) {
            P.tile_base = t_prev; P.ntiles = (u32)t_end;
            P.ticket = (u32*)((u8*)ctx->tile_status.p + (size_t)ntiles * 8 + 16) + c;
            const int g = (int)std::min<u64>(t_end - t_prev, (u64)num_sms(ctx->device) * (MM2_SK4_OCC + 2));
            MM2_TRY(launch_sketch_v4(ctx, g, P, is_hpc));
            t_prev = (u32)t_end;
          }
        }
      } else {
        if (feed && attempt == 0) for (int c = 0; c < feed->nchunks; ++c) CUDA_TRY(cudaStreamWaitEvent(st, feed->ev[c], 0));
        if (w >= 9) {   // per-thread prefix/suffix window minima + scanner warp
          MM2_TRY(launch_sketch_v4(ctx, grid, P, is_hpc));
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
