// lchain.cu — minimap2-style chaining DP on sm_100a, one warp per read.
// Replaces lchain.rs:59-92 (forward DP with max_dist/bw/max_chain_iter/max_chain_skip), :162-173 (the fallback chain,
// which under `-n >= 2` is the only chain chain_dp_all ever returns — SURVEY.md F3), :316-330 (rescue_long_join's
// rerun with bw_long) and the per-chain reductions of paf.rs:133-147.
//
// The j-loop of lchain.rs:80-88 is executed 32 predecessors at a time, highest j in lane 0:
//  * every lane evaluates comput_sc (lchain.rs:17-34) for its j;
//  * all lanes with a score publish their mark t[pprev[j]] = i first (marks are only ever compared with the current i
//    and pprev[j] < j, so marks by lanes past the break point are unobservable), then read t[j];
//  * "sc > max_f" in sequence order is "sc > exclusive prefix max" (warp max-scan);
//  * n_skip evolves by x -> max(x-1,0) on a new maximum and x -> x+1 on a marked non-maximum; both are maps
//    x -> max(x+a, b), closed under composition, so a 5-step warp scan gives every lane its n_skip; the first lane
//    whose n_skip exceeds max_chain_skip is the `break`.
// tests/models.py:chain_fwd_model is the CPU model of this restatement (checked against the oracle).
// The gap penalty (lchain.rs:28-32) is computed with explicitly rounded f32 ops (no FMA contraction) and a host-built
// table of 0.5*log2(dd+1) (glibc logf, as Rust's f32::ln), so scores are bit-identical to the CPU.
#include "mm2_internal.cuh"

#include <algorithm>
#include <cmath>

#include "stages.cuh"

namespace {

struct ChainArgs {
  const ulonglong2* anchors;
  const u64* read_aoff;
  const u64* read_off;   // read offsets in bases (qlen)
  const u64* mini_off;   // unfiltered minimizer offsets
  const u64* mval;       // rid_pos_strand of the minimizers
  const u32* sum_span;
  u32 nreads;
  mm2_chain_params_t p;
  int do_rescue;
  const float* half_log;  // 0.5 * mg_log2(dd + 1), dd = 0 .. max(bw, bw_long)
  const int* pen_int;     // chn_pen_skip == 0: the whole gap penalty of lchain.rs:28-32 as a function of dd (same float operations, done once per dd on the host)
  int4* A; int4* B; int* T; int* W; int* chain;
  ReadHit* hits;
  unsigned long long* cells;
  u32* dense;        // [0] number of dense reads, [1] ticket, [2..9] reads per size class, [16 + class * nreads ..] their indices
  int dense_min;     // a read is dense when it has >= dense_min anchors and more than dense_ratio5 / 5 anchors per base
  int dense_ratio5;
};

constexpr int CH_WARPS = 4;
constexpr int NEG_INF = -(1 << 28);

__device__ __forceinline__ int wadd(int a, int b) { return (int)((u32)a + (u32)b); }
__device__ __forceinline__ int wsub(int a, int b) { return (int)((u32)a - (u32)b); }

// The reported chain (lchain.rs:162-173), its coordinates (paf.rs:133-147) and the dv inputs (paf.rs:174-191) of one read.
__device__ __forceinline__ void chain_finish(const ChainArgs& G, u32 r, int lane, const ulonglong2* an, const int4* A, u64 a0, i32 qlen,
                                             u64 m0, u64 m1, int best, int4 bestA, int4 bestB, ReadHit& hit, unsigned long long cells) {
  // ---- the reported chain: fallback chain ending at `best` (lchain.rs:164-172), score v[best] -----------------------------
  const ulonglong2 ab = an[best];
  const ulonglong2 af = an[bestB.z];
  hit.rid_rev = (u32)(af.x >> 32);  // paf.rs:132,148 use the chain's FIRST anchor
  hit.qs = max(bestB.x, 0); hit.qe = wadd((int)(u32)ab.y, 1);
  hit.ts = max(bestB.y, 0); hit.te = wadd((int)(u32)ab.x, 1);
  hit.cm = (u32)bestA.w; hit.score = bestA.z; hit.best = best;
  if (G.chain) {  // stage dump: walk the links (serial; parity harness only)
    if (lane == 0) {
      int* ch = G.chain + a0;
      int i = best, c = 0;
      while (i >= 0) { ch[c++] = i; i = A[i].y; }
    }
  }
  // ---- dv inputs (paf.rs:174-191): ranks of the chain's first/last forward query positions among the read's minimizers ----
  {
    const bool rev = (af.x >> 63) != 0;
    auto qpos_fwd = [&](const ulonglong2& a) -> int {
      const int qp = (int)(u32)a.y, qsp = (int)((a.y >> 32) & 0xff);
      return rev ? wsub(wsub(qlen, 1), wsub(wadd(qp, 1), qsp)) : qp;
    };
    // '+': chain order; '-': reversed chain order (paf.rs:170-173)
    const int first_q = rev ? qpos_fwd(ab) : qpos_fwd(af);
    const int last_q = rev ? qpos_fwd(af) : qpos_fwd(ab);
    const int want = lane == 0 ? first_q : last_q;
    if (lane < 2) {
      // first index whose position == want (positions ascend for odd k, see sketch.cu)
      u64 lo = m0, hi = m1;
      while (lo < hi) {
        const u64 mid = (lo + hi) >> 1;
        const int pos = (int)(u32)((G.mval[mid] >> 1) & 0xffffffffULL);
        if (pos < want) lo = mid + 1; else hi = mid;
      }
      int rank = -1;
      if (lo < m1 && (int)(u32)((G.mval[lo] >> 1) & 0xffffffffULL) == want) rank = (int)(lo - m0);
      if (lane == 0) hit.st_rank = rank; else hit.en_rank = rank;
    }
    hit.en_rank = __shfl_sync(0xFFFFFFFFu, hit.en_rank, 1);
  }
  if (lane == 0) {  // `cells` is warp-uniform
    G.hits[r] = hit;
    if (G.cells) atomicAdd(G.cells, cells);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// chain_ring_kernel — the default.  One warp per read; the DP state of the 32 most recent anchors lives in REGISTERS:
// lane l owns the ring slot of every anchor j with j % 32 == l (static x/y fields and f, pprev, v, cnt, qs_min, ts_min,
// first).  Anchors are read 32 at a time with one coalesced 128-bit load per lane, so the lane that loads anchor i is
// the one that owns its slot; the state is flushed to A/B with one coalesced store per 32 anchors.
//  * Window membership (lchain.rs:75-78) is evaluated per lane on its own slot: the predicate "other rid/strand or
//    rpos(i) > rpos(j) + max_dist_x" is true on a prefix of [0, i) (anchors are sorted by x), so "j >= st" is "predicate
//    false for j", and the window of i is empty exactly when the predicate is true for i - 1.  Anchors with an empty
//    window (about half of them on ONT-like reads) cost nothing: their state is (span, no predecessor) and is committed
//    to the ring in bulk by the next anchor that has a window.
//  * The ring is visited first, in the reference's order (i-1, i-2, ...).  Marks (lchain.rs:86) among ring slots are one
//    warp OR-reduction of 1 << (pprev & 31).  When at most max_chain_skip slots are marked no `break` is possible
//    (n_skip starts at 0 and grows only on marked slots), so max_f is the warp maximum and max_j the first slot in
//    visiting order that reaches it (records are strict prefix maxima).  Otherwise the records are enumerated one
//    ballot at a time and the n_skip walk of lchain.rs:84-85 is done on the ballot masks (chain_tile_walk).
//  * Only when the ring was visited without a break AND the slot that anchor i overwrites (j = i - 32) is still inside
//    the window does the loop continue, 32 predecessors at a time, on A/T in global memory.  On ONT-like reads the loop
//    breaks inside the ring for nearly every anchor (about 27 visited predecessors: max_chain_skip = 25).
// tests/ compares it (and the CTA-per-read variant below) with the CPU restatement, state by state.

__device__ __forceinline__ u32 low_mask(int n) { return n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u); }
// position of the n-th (1-based) set bit of m; popc(m) >= n
__device__ __forceinline__ int nth_set_bit(u32 m, int n) {
  int pos = 0;
#pragma unroll
  for (int s = 16; s; s >>= 1) {
    const int c = __popc(m & (((1u << s) - 1u) << pos));
    if (c < n) { n -= c; pos += s; }
  }
  return pos;
}

// One tile of <= 32 predecessors, lane p = p-th visited (lchain.rs:80-88).  V/M/ACT: ballots of "has a score", "t[j] == i",
// "inside the window".  Updates max_f / n_skip, returns the lane of the last new maximum in rec_last (-1: none) and
// true when the reference's loop breaks inside this tile.
__device__ __forceinline__ bool chain_tile_walk(int lane, u32 V, u32 M, u32 ACT, int sc, int max_skip, int& max_f, int& n_skip,
                                                int& rec_last, unsigned long long& cells) {
  const bool valid = (V >> lane) & 1u;
  // records: strict prefix maxima of the scores, seeded with max_f (lchain.rs:84)
  u32 R = 0;
  int cur = max_f;
  u32 cand = __ballot_sync(0xFFFFFFFFu, valid && sc > cur);
  while (cand) {
    const int l = __ffs(cand) - 1;
    R |= 1u << l;
    cur = __shfl_sync(0xFFFFFFFFu, sc, l);
    cand = __ballot_sync(0xFFFFFFFFu, valid && sc > cur) & ~low_mask(l + 1);
  }
  // n_skip: -1 (floored at 0) on a record, +1 on a marked non-record, break when it exceeds max_skip (lchain.rs:84-85)
  const u32 Mnr = M & V & ~R;
  int x = n_skip, pos = 0, brk = -1;
  u32 rem = R;
  for (;;) {
    const int seg_end = rem ? (__ffs(rem) - 1) : 32;
    const u32 seg = Mnr & low_mask(seg_end) & ~low_mask(pos);
    const int cnt = __popc(seg);
    const int need = max(max_skip + 1 - x, 1);
    if (cnt >= need) { brk = nth_set_bit(seg, need); break; }
    x += cnt;
    if (!rem) break;
    x = max(x - 1, 0);
    pos = seg_end + 1;
    rem &= rem - 1;
  }
  if (brk >= 0) R &= low_mask(brk);
  rec_last = R ? (31 - __clz(R)) : -1;
  if (rec_last >= 0) {
    const int top = __shfl_sync(0xFFFFFFFFu, sc, rec_last);   // == cur unless later records were cut off by the break
    max_f = top;
  }
  if (brk >= 0) { cells += (unsigned)(brk + 1); return true; }
  cells += (unsigned)__popc(ACT);
  n_skip = x;
  return false;
}

// comput_sc (lchain.rs:17-34) of anchor i = (ri, qi) against a predecessor (rj, qj, span_j) of the same rid/strand
template <bool LUT>
__device__ __forceinline__ bool chain_sc(int ri, int qi, int rj, int qj, int span_j, int mdx, int mdy, int bw, float pen_gap,
                                         float pen_skip, const float* __restrict__ half_log, const int* __restrict__ pen_int, int& s0) {
  const int dq = wsub(qi, qj);
  if (dq <= 0 || dq > mdx) return false;
  const int dr = wsub(ri, rj);
  if (dr == 0 || dq > mdy) return false;
  int dd = wsub(dr, dq); if (dd < 0) dd = wsub(0, dd);
  if (dd > bw || dd < 0) return false;
  const int dg = min(dr, dq);
  s0 = min(span_j, dg);
  if (dd != 0 || dg > span_j) {
    if constexpr (LUT) s0 = wsub(s0, pen_int[dd]);
    else {
      const float lin = __fadd_rn(__fmul_rn(pen_gap, (float)dd), __fmul_rn(pen_skip, (float)dg));
      s0 = wsub(s0, __float2int_rz(__fadd_rn(lin, half_log[dd])));
    }
  }
  return true;
}

// The same without branches (one predicate instead of the early returns): used where every lane of the warp evaluates a cell,
// so the warp executes the whole body anyway and the branch / reconvergence instructions are pure overhead.  `act` = the lane
// has a cell at all; the table index is clamped for the lanes that do not.
template <bool LUT = false, bool TRIM = false>
__device__ __forceinline__ bool chain_sc_flat(bool act, int ri, int qi, int rj, int qj, int span_j, int mdx, int mdy, int bw, float pen_gap,
                                              float pen_skip, const float* __restrict__ half_log, int& s0, const int* __restrict__ pen_int = nullptr) {
  const int dq = wsub(qi, qj), dr = wsub(ri, rj);
  int dd = wsub(dr, dq); if (dd < 0) dd = wsub(0, dd);
  bool ok;
  // TRIM (the caller has checked bw >= 0, hence mdx, mdy >= bw >= 0): 1 <= dq <= min(mdx, mdy) and 0 <= dd <= bw as one unsigned
  // compare each
  if constexpr (TRIM) ok = act && (u32)wsub(dq, 1) < (u32)min(mdx, mdy) && dr != 0 && (u32)dd <= (u32)bw;
  else ok = act && dq > 0 && dq <= mdx && dr != 0 && dq <= mdy && dd <= bw && dd >= 0;
  const int dg = min(dr, dq);
  s0 = min(span_j, dg);
  const int ddc = ok ? dd : 0;
  int pen;
  if constexpr (LUT) pen = pen_int[ddc];   // chn_pen_skip == 0: pen_gap * dd + 0 * dg == pen_gap * dd exactly (both products are >= +0)
  else {
    const float lin = __fadd_rn(__fmul_rn(pen_gap, (float)ddc), __fmul_rn(pen_skip, (float)dg));
    pen = __float2int_rz(__fadd_rn(lin, half_log[ddc]));
  }
  if (ddc != 0 || dg > span_j) s0 = wsub(s0, pen);
  return ok;
}

__device__ __forceinline__ bool chain_is_dense(const ChainArgs& G, u32 r) {
  const i64 n = (i64)(G.read_aoff[r + 1] - G.read_aoff[r]);
  const i64 qlen = (i64)(G.read_off[r + 1] - G.read_off[r]);
  return n >= (i64)G.dense_min && n <= 0x7fffffff && n * 5 > qlen * (i64)G.dense_ratio5;
}

// Reads with a dense anchor set (repeat-rich ultra-long reads: predecessor windows of thousands of anchors) are chained by
// one CTA each (chain_dense_kernel); this kernel only lists them.
// The list is kept in 8 size classes (anchor count 2^12 .. >= 2^19) and handed out largest class first: a dense read occupies
// one CTA for its whole length, so starting the long reads first keeps the batch from waiting for a straggler at the end.
constexpr int DENSE_BINS = 8;
__device__ __forceinline__ int dense_bin(i64 n) { return min(max(63 - __clzll((unsigned long long)max(n, (i64)1)) - 12, 0), DENSE_BINS - 1); }
__global__ void chain_classify_kernel(ChainArgs G) {
  const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= G.nreads) return;
  if (chain_is_dense(G, r)) {
    const int b = dense_bin((i64)(G.read_aoff[r + 1] - G.read_aoff[r]));
    atomicAdd(&G.dense[0], 1u);
    G.dense[16 + (u64)b * G.nreads + atomicAdd(&G.dense[2 + b], 1u)] = r;
  }
}
// ticket k (0 .. ndense-1) -> read: classes from the largest down
__device__ __forceinline__ u32 dense_read_of_ticket(const ChainArgs& G, u32 k) {
  for (int b = DENSE_BINS - 1; b >= 0; --b) {
    const u32 c = G.dense[2 + b];
    if (k < c) return G.dense[16 + (u64)b * G.nreads + k];
    k -= c;
  }
  return 0;
}

// one warp per read (chain_ring_kernel).  COUNT: also count DP cells (diagnostic; a compile-time switch, the bookkeeping costs ~5 %)
// FAST (chosen by chain_batch): chn_pen_skip == 0 (integer penalty table), max_chain_iter >= 32 (every filled ring slot
// j >= i - 32 passes the max_chain_iter bound of lchain.rs:78) and bw, bw_long >= 0 (unsigned range compares).  Same results,
// ~5 instructions fewer per anchor with a window.
template <bool COUNT, bool FAST>
__device__ __forceinline__ void chain_read(const ChainArgs& G, const u32 r, const int lane) {
  const u64 a0 = G.read_aoff[r];
  const i64 n64 = (i64)(G.read_aoff[r + 1] - a0);
  const i32 qlen = (i32)(G.read_off[r + 1] - G.read_off[r]);
  const u64 m0 = G.mini_off[r], m1 = G.mini_off[r + 1];
  ReadHit hit;
  hit.rid_rev = 0xFFFFFFFFu; hit.qs = hit.qe = hit.ts = hit.te = 0; hit.cm = 0; hit.score = 0;
  hit.n_anchors = (u32)n64; hit.n_mini = (u32)(m1 - m0); hit.sum_span = G.sum_span[r];
  hit.st_rank = hit.en_rank = -1; hit.flags = 0; hit.best = -1; hit.pad0 = hit.pad1 = 0;
  if (n64 <= 0 || n64 > 0x7fffffff) {
    if (lane == 0) G.hits[r] = hit;
    return;
  }
  const int n = (int)n64;
  const ulonglong2* __restrict__ an = G.anchors + a0;
  int4* A = G.A + a0;
  int4* B = G.B + a0;
  int* T = G.T + a0;
  const mm2_chain_params_t& p = G.p;
  // n_skip never exceeds the number of anchors (< 2^31) and every max_chain_skip < 0 breaks on the first marked slot
  const int max_skip = min(max(p.max_chain_skip, -1), 0x3fffffff);
  const int max_iter = p.max_chain_iter;
  unsigned long long cells = 0;
  constexpr bool count_cells = COUNT;            // a statistic (lchain.rs:80 iterations), not part of the result
  int best = 0;
  int4 bestA = make_int4(0, -1, 0, 0), bestB = make_int4(0, 0, 0, 0);
  bool t_init = false;
  for (int pass = 0; pass < 2; ++pass) {
    const int bw = pass == 0 ? p.bw : p.bw_long;                 // lchain.rs:327-328
    const int mdx = max(p.max_dist_x, bw), mdy = max(p.max_dist_y, bw);  // lchain.rs:63-66
    const int mark_base = pass * n;
    // ring slot of this lane
    int rj = -1, rx = 0, rq = 0, rsp = 0; u32 rhi = 0;
    int rf = 0, rpp = -1;
    int bf = NEG_INF * 4, bi = 0;
    int st = 0;   // lower bound of the window start, used only by windows longer than the ring
    ulonglong2 nxt = make_ulonglong2(0, 0);
    if (lane < n) nxt = an[lane];
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int tile_n = min(32, n - i0);
      const ulonglong2 cur = nxt;
      if (i0 + 32 + lane < n) nxt = an[i0 + 32 + lane];          // next tile's anchors are in flight during this tile
      const int cx = (int)(u32)cur.x, cq = (int)(u32)cur.y, csp = (int)((cur.y >> 32) & 0xff);
      const u32 chi = (u32)(cur.x >> 32);
      const int own_qs = wsub(cq, csp - 1), own_ts = wsub(cx, csp - 1);
      // anchors whose window is not empty: the predicate of lchain.rs:75 is false for i - 1, and max_chain_iter >= 1
      u32 workmask;
      {
        int px = __shfl_up_sync(0xFFFFFFFFu, cx, 1);
        u32 phi = __shfl_up_sync(0xFFFFFFFFu, chi, 1);
        const int rx31 = __shfl_sync(0xFFFFFFFFu, rx, 31);
        const u32 rhi31 = __shfl_sync(0xFFFFFFFFu, rhi, 31);
        if (lane == 0) { px = rx31; phi = rhi31; }
        const int i = i0 + lane;
        const bool work = lane < tile_n && i > 0 && (i - 1) >= wsub(i, max_iter) && phi == chi && !(cx > wadd(px, mdx));
        workmask = __ballot_sync(0xFFFFFFFFu, work);
      }
      int done = 0;   // lanes [0, done) of this tile are committed to the ring
      while (workmask) {
        const int c = __ffs(workmask) - 1;
        workmask &= workmask - 1;
        const int i = i0 + c;
        if (done < c) {
          if (lane >= done && lane < c) {                        // anchors before i with an empty window (lchain.rs:77,89-90)
            rj = i0 + lane; rx = cx; rq = cq; rsp = csp; rhi = chi;
            rf = csp; rpp = -1;
          }
        }
        done = c + 1;
        const int ri = __shfl_sync(0xFFFFFFFFu, cx, c), qi = __shfl_sync(0xFFFFFFFFu, cq, c), spi = __shfl_sync(0xFFFFFFFFu, csp, c);
        const u32 hi_i = __shfl_sync(0xFFFFFFFFu, chi, c);
        const int low_iter = wsub(i, max_iter);                  // lchain.rs:78
        const bool inwin = (FAST ? rj >= 0 : rj >= max(low_iter, 0)) && rhi == hi_i && !(ri > wadd(rx, mdx));   // empty slots hold rj = -1
        int s0;
        const bool valid = chain_sc_flat<FAST, FAST>(inwin, ri, qi, rx, rq, rsp, mdx, mdy, bw, p.chn_pen_gap, p.chn_pen_skip, G.half_log, s0, G.pen_int);
        const int sc = valid ? wadd(s0, rf) : NEG_INF;
        const u32 inmask = __ballot_sync(0xFFFFFFFFu, inwin);
        const u32 vmask = __ballot_sync(0xFFFFFFFFu, valid);
        int max_f = spi, max_j = -1, n_skip = 0;
        bool more = (inmask >> c) & 1u;                          // slot c holds j = i - 32: the window may go on beyond the ring
        if (vmask) {
          const int low_ring = max(i - 32, 0);
          const u32 markbits = __reduce_or_sync(0xFFFFFFFFu, (valid && rpp >= low_ring) ? (1u << (rpp & 31)) : 0u) & vmask;
          // visiting order p <-> lane (c - 1 - p) & 31, i.e. bit 31 - p of rotr(mask, c)
          const u32 Mr = __funnelshift_r(markbits, markbits, c);
          const int m = __reduce_max_sync(0xFFFFFFFFu, sc);      // NEG_INF on the slots without a score
          int hb = 32;                                           // 31 - (first visited slot that reaches m)
          if (m > spi) {
            const u32 eq = __ballot_sync(0xFFFFFFFFu, valid && sc == m);
            hb = 31 - __clz(__funnelshift_r(eq, eq, c));
          }
          if (hb == 32 || (Mr & (0xFFFFFFFEu << hb)) == 0u) {
            // No record at all, or nothing marked before the slot that sets the final maximum: n_skip is still 0 there
            // (records only decrement it), that slot is the last record, and everything after it is a non-record.
            u32 after = Mr;
            if (hb < 32) { max_f = m; max_j = i - 32 + hb; after = Mr & ((1u << hb) - 1u); }
            const int cnt = __popc(after), need = max(max_skip + 1, 1);
            if (cnt >= need) {                                   // lchain.rs:85: the need-th marked slot breaks the loop
              if (count_cells) cells += (unsigned)(nth_set_bit(__brev(after), need) + 1);
              more = false;
            } else {
              n_skip = cnt;
              if (count_cells) cells += (unsigned)__popc(inmask);
            }
          } else {
            const int sc_s = __shfl_sync(0xFFFFFFFFu, sc, (c - 1 - lane) & 31);
            const u32 Vs = __brev(__funnelshift_r(vmask, vmask, c)), As = __brev(__funnelshift_r(inmask, inmask, c));
            int rec_last;
            const bool brk = chain_tile_walk(lane, Vs, __brev(Mr), As, sc_s, max_skip, max_f, n_skip, rec_last, cells);
            if (rec_last >= 0) max_j = i - 1 - rec_last;
            if (brk) more = false;
          }
        } else if (count_cells) {
          cells += (unsigned)__popc(inmask);
        }
        if (more) {
          // ---- the window goes on beyond the ring: 32 predecessors at a time from global memory -------------------------
          if (!t_init) {
            for (int x = lane; x < n; x += 32) T[x] = -1;
            t_init = true;
            __syncwarp();
          }
          const int mark = mark_base + i;
          const int hi_known = i - 32;                           // inside the window
          int lo = st;
          if (valid && rpp >= 0 && rpp < hi_known) T[rpp] = mark;   // marks of the ring slots on older anchors
          // window start: first j in [st, i - 32] for which the predicate of lchain.rs:75 is false
          {
            int rounds = 0;
            for (;;) {
              const int j = lo + lane;
              bool adv = false;
              if (j < hi_known) {
                const u64 xm = an[j].x;
                adv = ((u32)(xm >> 32) != hi_i) || (ri > wadd((int)(u32)xm, mdx));
              }
              const u32 b = __ballot_sync(0xFFFFFFFFu, !adv);
              if (b) { lo += __ffs(b) - 1; break; }
              lo += 32;
              if (++rounds == 2) {
                int hi = hi_known;
                while (lo < hi) {
                  const int mid = (lo + hi) >> 1;
                  const u64 xm = an[mid].x;
                  const bool adv2 = ((u32)(xm >> 32) != hi_i) || (ri > wadd((int)(u32)xm, mdx));
                  if (adv2) lo = mid + 1; else hi = mid;
                }
                break;
              }
            }
          }
          st = lo;
          const int start_j = low_iter > lo ? low_iter : lo;
          {
            for (int jb = i - 33; jb >= start_j; jb -= 32) {
              const int j = jb - lane;
              const bool act = j >= start_j;
              bool v2 = false;
              int sc2 = NEG_INF, ppj = -1;
              if (act) {
                const ulonglong2 v = an[j];
                if ((u32)(v.x >> 32) == hi_i) {                     // lchain.rs:81
                  int s0;
                  if (chain_sc<FAST>(ri, qi, (int)(u32)v.x, (int)(u32)v.y, (int)((v.y >> 32) & 0xff), mdx, mdy, bw, p.chn_pen_gap,
                                    p.chn_pen_skip, G.half_log, G.pen_int, s0)) {
                    const int2 fj = *reinterpret_cast<const int2*>(A + j);
                    sc2 = wadd(s0, fj.x);
                    ppj = fj.y;
                    v2 = true;
                  }
                }
              }
              if (v2 && ppj >= 0) T[ppj] = mark;                    // lchain.rs:86 (all lanes first, see the header)
              __syncwarp();
              const bool tm = v2 ? (T[j] == mark) : false;
              const u32 V2 = __ballot_sync(0xFFFFFFFFu, v2), M2 = __ballot_sync(0xFFFFFFFFu, tm), A2 = __ballot_sync(0xFFFFFFFFu, act);
              int rec_last;
              const bool brk = chain_tile_walk(lane, V2, M2, A2, sc2, max_skip, max_f, n_skip, rec_last, cells);
              if (rec_last >= 0) max_j = jb - rec_last;
              if (brk) break;
            }
          }
        }
        // ---- lchain.rs:89-90.  v, the chain length and the chain extents (paf.rs:136-147) feed no DP decision: they are derived
        //      for the whole tile at once below, by pointer jumping over the best-predecessor links ------------------------------
        if (lane == c) {                                         // anchor i takes over its ring slot
          rj = i; rx = cx; rq = cq; rsp = csp; rhi = chi;
          rf = max_f; rpp = max_j;
        }
      }
      if (lane >= done && lane < tile_n) {
        rj = i0 + lane; rx = cx; rq = cq; rsp = csp; rhi = chi;
        rf = csp; rpp = -1;
      }
      {
        // Every lane now holds its own anchor of this tile (f, best predecessor).  v = the maximum of f along the chain
        // (lchain.rs:90), its length, the minima of the anchor starts and its first anchor (paf.rs:136-147) are aggregates
        // over the predecessor path: a predecessor in an earlier tile contributes its finished values from A / B, and the
        // links inside the tile are followed by pointer jumping (5 rounds for 32 anchors) instead of 5 shuffles per anchor.
        int pv = rf, pcnt = 1, pqs = own_qs, pts = own_ts, pfirst = i0 + lane, ptr = -1;
        if (lane < tile_n && rpp >= 0) {
          if (rpp >= i0) ptr = rpp - i0;
          else {
            const int4 aj = A[rpp], bj = B[rpp];
            pv = max(pv, aj.z); pcnt = aj.w + 1; pqs = min(pqs, bj.x); pts = min(pts, bj.y); pfirst = bj.z;
          }
        }
#pragma unroll
        for (int round = 0; round < 5; ++round) {
          const int src = ptr >= 0 ? ptr : lane;
          const int qv = __shfl_sync(0xFFFFFFFFu, pv, src), qcnt = __shfl_sync(0xFFFFFFFFu, pcnt, src);
          const int qqs = __shfl_sync(0xFFFFFFFFu, pqs, src), qts = __shfl_sync(0xFFFFFFFFu, pts, src);
          const int qfirst = __shfl_sync(0xFFFFFFFFu, pfirst, src), qptr = __shfl_sync(0xFFFFFFFFu, ptr, src);
          if (ptr >= 0) { pv = max(pv, qv); pcnt += qcnt; pqs = min(pqs, qqs); pts = min(pts, qts); pfirst = qfirst; ptr = qptr; }
        }
        if (lane < tile_n) {
          A[i0 + lane] = make_int4(rf, rpp, pv, pcnt);
          B[i0 + lane] = make_int4(pqs, pts, pfirst, 0);
        }
      }
      {  // lchain.rs:163: the LAST maximum of f
        const int fl = lane < tile_n ? rf : NEG_INF * 4;
        const int m = __reduce_max_sync(0xFFFFFFFFu, fl);
        if (m >= bf) {
          const u32 eq = __ballot_sync(0xFFFFFFFFu, fl == m);
          bf = m;
          bi = i0 + 31 - __clz(eq);
        }
      }
      __syncwarp();
    }
    best = bi;
    bestA = A[best]; bestB = B[best];
    if (pass == 1 || !G.do_rescue) break;
    // lchain.rs:321-330 rescue_long_join on the single (fallback) chain
    const ulonglong2 ab = an[best];
    const int qe = wadd((int)(u32)ab.y, 1);
    const int qs = max(bestB.x, 0);
    const int best_cov = max(wsub(qe, qs), 0);
    const int uncovered = max(wsub(qlen, best_cov), 0);
    const bool rescue = uncovered > p.rmq_rescue_size ||
                        (float)best_cov < __fmul_rn((float)qlen, __fsub_rn(1.0f, p.rmq_rescue_ratio));
    if (!rescue) break;
    hit.flags |= 1u;
    __syncwarp();
  }
  chain_finish(G, r, lane, an, A, a0, qlen, m0, m1, best, bestA, bestB, hit, COUNT ? cells : 0ull);
}

#ifndef MM2_CH_OCC
#define MM2_CH_OCC 8   // CTAs per SM (64 registers; measured 7.5 / 6.8 / 8.0 ms per 100k reads at 6 / 8 / 10)
#endif
// ---------------------------------------------------------------------------------------------------------------------
// chain_dense_kernel — one CTA of NW warps per dense read (repeat-rich ultra-long reads: predecessor windows of thousands of
// anchors, lchain.rs:74-91 at its max_chain_iter bound).  The read is processed in tiles of 32 anchors, pipelined:
//   * warp 0 runs the sequential part of tile T (the ring algorithm of chain_read for the last 32 anchors, then the rest of
//     tile T-1, then the FAR part of the window through per-tile summaries);
//   * warps 1.. meanwhile evaluate the far part of the windows of tile T+1: every predecessor j < i0(T) is final by then
//     (tile T-1 was committed before the previous barrier), so the scores sc(i, j) = comput_sc + f[j] of ALL 32 anchors i of
//     the tile against a 32-aligned tile of predecessors are computed by one warp with the predecessors' state held in
//     registers (lane = predecessor), and reduced to {valid bits, best score} per (anchor, predecessor tile), plus the
//     anchor's mark bitmask (lchain.rs:86: every valid predecessor marks ITS predecessor; marks do not depend on where the
//     loop later breaks, because a mark is only ever read at a lower j than the one that set it).
//   * the walk of a far tile needs its cells one by one only when the tile's best score beats the running maximum (a new
//     record, lchain.rs:84); otherwise n_skip just grows by the number of marked valid cells (lchain.rs:85).
// One barrier per 32 anchors instead of two per 1024 cells; the per-anchor latency chain (round 1: ~6,500 cycles per anchor,
// profiles/r01_ncu_v11.md) shrinks to the ring step plus a few lane-parallel rounds over tile summaries.
// tests/models.py:chain_dense_model is the CPU model of this pipeline (checked against the oracle, cell counts included).
// State of the last DENSE_CAP anchors in shared memory (slot = j % DENSE_CAP): {x, q}, {f, pprev}, span.  A window reaches at
// most max_chain_iter <= DENSE_CAP - 64 anchors back, and a slot is overwritten only by anchor j + DENSE_CAP.
#ifdef MM2_DENSE_PROF
// build with -DMM2_DENSE_PROF: cycle counters of the dense pipeline, printed by chain_batch (diagnostic only)
// 0 ring step, 1 rest of tile T-1, 2 far walk, 3 warp 0 waiting at the tile barrier, 4 anchors with a window, 5 far rounds,
// 6 record tiles, 7 total (warp 0), 8 phase A (helper warp 1), 9 helper waiting at the tile barrier, 10 tiles, 11 far (anchor, tile) pairs
__device__ unsigned long long g_dense_prof[12];
#define DPROF_T(v) const long long v = clock64()
#define DPROF_ADD(k, a, b) do { if (lane == 0) dprof[k] += (unsigned long long)((b) - (a)); } while (0)
#define DPROF_INC(k, x) do { if (lane == 0) dprof[k] += (unsigned long long)(x); } while (0)
#else
#define DPROF_T(v)
#define DPROF_ADD(k, a, b)
#define DPROF_INC(k, x)
#endif
constexpr int DENSE_CAP = 5120;
constexpr int DENSE_JT = 160;          // far predecessor tiles per window: ceil((max_chain_iter + 31) / 32) + 1 <= 160
constexpr int DENSE_MKW = DENSE_JT + 2;   // mark words per anchor: [0] = tile T, [1] = tile T-1, [2 + r] = far tile r
struct DenseBuf {                      // summaries of one tile of 32 anchors (double-buffered); odd row strides: lane = anchor
  u32 V[32][DENSE_JT + 1];             // valid cells of (anchor, far tile r), bit p = p-th visited (j descending)
  int TM[32][DENSE_JT + 1];            // best score of the tile's valid cells (NEG_INF: none)
  u32 MK[32][DENSE_MKW + 1];           // marks, bit = j & 31 (ascending)
  int far_lo[32];                      // first far predecessor of the anchor; >= far_hi: no far part
};
constexpr int DENSE_LUT = 2048;        // gap penalties (lchain.rs:28-32) of dd < DENSE_LUT as 16-bit integers in shared memory
struct DenseSh {
  u16* pen16;                          // (int)(chn_pen_gap * dd + 0.5 * log2(dd + 1)): a function of dd alone when chn_pen_skip == 0
  int rescue_flag;                     // pass 0 -> pass 1 decision of warp 0 (lchain.rs:321-330)
  int2* sxq; int2* sfp; u8* ss;        // DENSE_CAP entries each
  DenseBuf* buf;                       // [2]
  u32 next;                            // next dense read of this CTA
  int cm_blk[2];                       // per tile parity: first anchor of the rid/strand block of the tile's last anchor
  u32 cm_hi[2];                        // ... and that anchor's rev|rid
};
constexpr int DENSE_DYN = DENSE_CAP * (8 + 8 + 1) + 2 * (int)sizeof(DenseBuf) + DENSE_LUT * 2 + 64;
constexpr int DENSE_NW = 16;           // warps per dense read: warp 0 walks alone on its scheduler, 12 evaluate, 3 only keep the barriers

template <int NW>
__device__ __forceinline__ void dense_bar_all() { asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory"); }
template <int NW>
__device__ __forceinline__ void dense_bar_helpers() { asm volatile("bar.sync 2, %0;" ::"n"((NW - NW / 4) * 32) : "memory"); }

// Phase A of tile `tile` (anchors i0 .. i0 + 31), run by the helper warps while warp 0 walks tile - 1.  A helper warp takes
// one 32-aligned tile of far predecessors at a time and evaluates it against all 32 anchors with LANE = ANCHOR: the
// predecessor's state is a broadcast shared-memory load, the anchor's a register, so the valid bits and the best score of
// (anchor, tile) accumulate in lane-private registers without a ballot or a reduction per cell, and the predecessor's mark
// target (lchain.rs:86) is warp-uniform: every lane ORs into its own anchor's mark row (conflict-free, odd row stride).
template <int NW>
__device__ __forceinline__ void dense_phase_a(const ChainArgs& G, DenseSh* sh, const ulonglong2* __restrict__ an, const int n, const int tile,
                                              const int bw, const int mdx, const int mdy, const int max_iter, const int wid, const int lane) {
  DenseBuf& B = sh->buf[tile & 1];
  const int i0 = tile * 32;
  const int far_hi = i0 - 32;                       // predecessors below it are final (tiles <= tile - 2 are committed)
  const int hw = wid - 1 - (wid >> 2), nh = NW - NW / 4;   // helper index / count (warps 4, 8, 12 leave warp 0's scheduler alone)
  // zero the mark words of this buffer; helper 0 finds the anchors' far window starts
  for (int x = hw * 32 + lane; x < 32 * (DENSE_MKW + 1); x += nh * 32) (&B.MK[0][0])[x] = 0u;
  int my_x = 0, my_q = 0;                            // anchor i0 + lane of the tile (every helper warp holds a copy)
  u32 my_hi = 0;
  if (i0 + lane < n) { const ulonglong2 a = an[i0 + lane]; my_x = (int)(u32)a.x; my_q = (int)(u32)a.y; my_hi = (u32)(a.x >> 32); }
  if (hw == 0) {
    int lo = far_hi;                                 // "no far part"
    const int i = i0 + lane;
    if (i < n && far_hi > 0) {
      // the far part is non-empty only if the anchor lies in the rid/strand block of the last committed anchor (far_hi - 1):
      // windows never cross a block (lchain.rs:75), and blocks are contiguous because the anchors are sorted by x
      if (my_hi == sh->cm_hi[tile & 1]) {
        int a = max(max(sh->cm_blk[tile & 1], wsub(i, max_iter)), 0), b = far_hi;   // first j in [a, far_hi) with !(ri > x_j + mdx)
        const int2* __restrict__ sxq = sh->sxq;
        while (a < b) {
          const int mid = (a + b) >> 1;
          if (my_x > wadd(sxq[mid % DENSE_CAP].x, mdx)) a = mid + 1; else b = mid;
        }
        lo = a;
      }
    }
    B.far_lo[lane] = lo;
  }
  dense_bar_helpers<NW>();
  if (far_hi <= 0) return;
  const int jt_top = (far_hi >> 5) - 1;              // far_hi is a multiple of 32: the first far tile visited
  const int my_lo = B.far_lo[lane];                  // far_lo of this lane's anchor
  const int lo_min = __reduce_min_sync(0xFFFFFFFFu, my_lo);
  if (lo_min >= far_hi) return;
  const int nr = jt_top - (lo_min >> 5) + 1;         // far tiles that some anchor of the tile needs
  const float pen_gap = G.p.chn_pen_gap, pen_skip = G.p.chn_pen_skip;
  const float* __restrict__ half_log = G.half_log;
  const int2* __restrict__ sxq = sh->sxq; const int2* __restrict__ sfp = sh->sfp; const u8* __restrict__ ss = sh->ss;
  const int ri = my_x, qi = my_q;
  const bool use_lut = pen_skip == 0.0f;
  const u16* __restrict__ pen16 = sh->pen16;
  u32* mk_row = &B.MK[lane][0];
  const int slot_top = (jt_top * 32) % DENSE_CAP;                 // slots go down by 32 per tile and wrap at most once
  for (int r = hw; r < nr; r += nh) {
    const int jt = jt_top - r;
    int slot0 = slot_top - 32 * r; if (slot0 < 0) slot0 += DENSE_CAP;
    const int j0 = jt * 32;
    const bool a_on = my_lo < far_hi && (my_lo >> 5) <= jt;       // this anchor's window reaches the tile
    const int t_min = a_on ? max(my_lo - j0, 0) : 32;             // cells t >= t_min of the tile are inside this anchor's window
    // this lane's predecessor of the tile, for the marks below
    const int pp_l = sfp[slot0 + lane].y;
    const int tw_l = pp_l >= 0 ? 2 + jt_top - (pp_l >> 5) : DENSE_MKW;
    const u32 tb_l = 1u << (pp_l & 31);
    u32 Va = 0;
    int TMa = NEG_INF;
    // (1) the 32 cells of this lane's anchor: straight-line code, no branch, so that the iterations overlap
#pragma unroll 8
    for (int t = 31; t >= 0; --t) {                               // visiting order: j descending, bit p = 31 - t
      const int2 xq = sxq[slot0 + t];                             // broadcast loads
      const int fj = sfp[slot0 + t].x;
      const int span_j = (int)ss[slot0 + t];
      const int dq = wsub(qi, xq.y), dr = wsub(ri, xq.x);
      int dd = wsub(dr, dq); if (dd < 0) dd = wsub(0, dd);
      const bool ok = t >= t_min && dq > 0 && dq <= mdx && dr != 0 && dq <= mdy && dd <= bw && dd >= 0;
      const int dg = min(dr, dq);
      int s0 = min(span_j, dg);
      const int ddc = ok ? dd : 0;
      int pen;
      if (use_lut) {   // chn_pen_skip == 0: the penalty depends on dd only (same float operations, done once per dd)
        pen = ddc < DENSE_LUT ? (int)pen16[ddc] : __float2int_rz(__fadd_rn(__fmul_rn(pen_gap, (float)ddc), half_log[ddc]));
      } else {
        const float lin = __fadd_rn(__fmul_rn(pen_gap, (float)ddc), __fmul_rn(pen_skip, (float)dg));
        pen = __float2int_rz(__fadd_rn(lin, half_log[ddc]));
      }
      if (ddc != 0 || dg > span_j) s0 = wsub(s0, pen);
      const int sc = ok ? wadd(s0, fj) : NEG_INF;
      Va |= (ok ? 1u : 0u) << (31 - t);
      TMa = max(TMa, sc);
    }
    B.V[lane][r] = Va;
    B.TM[lane][r] = TMa;
    // (2) lchain.rs:86: predecessor t marks pprev[t] for every anchor it scores with.  The target word is warp-uniform per t,
    //     so each lane accumulates the bits of its own anchor and flushes when the word changes (lane-private rows: no conflict)
    if (__ballot_sync(0xFFFFFFFFu, Va != 0u)) {
      int mw = DENSE_MKW;                                         // word being accumulated; DENSE_MKW = none
      u32 macc = 0;
#pragma unroll 4
      for (int t = 31; t >= 0; --t) {
        const int tw = __shfl_sync(0xFFFFFFFFu, tw_l, t);
        const u32 tb = __shfl_sync(0xFFFFFFFFu, tb_l, t);
        if (tw != mw) {
          if (mw < DENSE_MKW && macc) atomicOr(&mk_row[mw], macc);
          mw = tw; macc = 0;
        }
        macc |= ((Va >> (31 - t)) & 1u) ? tb : 0u;
      }
      if (mw < DENSE_MKW && macc) atomicOr(&mk_row[mw], macc);
    }
#ifdef MM2_DENSE_PROF
    if (lane == 0) atomicAdd(&g_dense_prof[11], 32ull);
#endif
  }
}

// lchain.rs:90 (v) and the chain reductions of paf.rs:136-147 (length, extents, first anchor) of one committed tile, carried
// along the best-predecessor links.  None of them feeds the DP of later anchors, so they are taken off warp 0's critical path:
// a helper warp derives them one tile behind (lane = anchor; a parent outside the tile is final in global memory, a parent
// inside it sits in a lower lane and is resolved in lane order).
__device__ __forceinline__ void dense_reduce_tile(const ulonglong2* __restrict__ an, int4* A, int4* B, const int n, const int tile, const int lane) {
  const int i0 = tile * 32, i = i0 + lane;
  const bool in = i < n;
  int f = 0, pp = -1, own_qs = 0, own_ts = 0;
  if (in) {
    const int2 fp = *reinterpret_cast<const int2*>(&A[i]);
    f = fp.x; pp = fp.y;
    const ulonglong2 a = an[i];
    const int sp = (int)((a.y >> 32) & 0xff);
    own_qs = wsub((int)(u32)a.y, sp - 1); own_ts = wsub((int)(u32)a.x, sp - 1);
  }
  int v = f, cnt = 1, qs = own_qs, ts = own_ts, first = i;
  if (in && pp >= 0 && pp < i0) {
    const int4 pa = A[pp], pb = B[pp];
    v = pa.z > f ? pa.z : f; cnt = pa.w + 1; qs = min(pb.x, own_qs); ts = min(pb.y, own_ts); first = pb.z;
  }
  const bool intile = in && pp >= i0;
  const int pl = pp & 31;
  u32 todo = __ballot_sync(0xFFFFFFFFu, intile);
  while (todo) {                                                 // ascending lanes: a parent is always final before its children
    const int c = __ffs(todo) - 1;
    todo &= todo - 1;
    const int pv = __shfl_sync(0xFFFFFFFFu, v, pl), pc = __shfl_sync(0xFFFFFFFFu, cnt, pl), pq = __shfl_sync(0xFFFFFFFFu, qs, pl);
    const int pt = __shfl_sync(0xFFFFFFFFu, ts, pl), pf = __shfl_sync(0xFFFFFFFFu, first, pl);
    if (lane == c) { v = pv > f ? pv : f; cnt = pc + 1; qs = min(pq, own_qs); ts = min(pt, own_ts); first = pf; }
  }
  if (in) {
    reinterpret_cast<int2*>(&A[i])[1] = make_int2(v, cnt);
    B[i] = make_int4(qs, ts, first, 0);
  }
}

// One CTA per dense read.  wid == 0: the sequential walk; wid >= 1: phase A of the next tile.
template <int NW>
__device__ __forceinline__ void chain_read_dense(const ChainArgs& G, const u32 r, const int lane, const int wid, DenseSh* sh) {
  const u64 a0 = G.read_aoff[r];
  const i64 n64 = (i64)(G.read_aoff[r + 1] - a0);
  const i32 qlen = (i32)(G.read_off[r + 1] - G.read_off[r]);
  const u64 m0 = G.mini_off[r], m1 = G.mini_off[r + 1];
  ReadHit hit;
  hit.rid_rev = 0xFFFFFFFFu; hit.qs = hit.qe = hit.ts = hit.te = 0; hit.cm = 0; hit.score = 0;
  hit.n_anchors = (u32)n64; hit.n_mini = (u32)(m1 - m0); hit.sum_span = G.sum_span[r];
  hit.st_rank = hit.en_rank = -1; hit.flags = 0; hit.best = -1; hit.pad0 = hit.pad1 = 0;
  const int n = (int)n64;                                        // chain_is_dense: 0 < n <= 0x7fffffff
  const ulonglong2* __restrict__ an = G.anchors + a0;
  int4* A = G.A + a0;
  int4* B = G.B + a0;
  const mm2_chain_params_t& p = G.p;
  const int max_skip = min(max(p.max_chain_skip, -1), 0x3fffffff);
  const int max_iter = p.max_chain_iter;
  const int ntile = (n + 31) >> 5;
  unsigned long long cells = 0;
  const bool count_cells = G.cells != nullptr;
  int best = 0;
  int4 bestA = make_int4(0, -1, 0, 0), bestB = make_int4(0, 0, 0, 0);
  const int2* __restrict__ sxq = sh->sxq; const int2* __restrict__ sfp = sh->sfp; const u8* __restrict__ ss = sh->ss;
#ifdef MM2_DENSE_PROF
  unsigned long long dprof[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const long long dprof_t0 = clock64();
#endif

  for (int pass = 0; pass < 2; ++pass) {
    const int bw = pass == 0 ? p.bw : p.bw_long;                 // lchain.rs:327-328
    const int mdx = max(p.max_dist_x, bw), mdy = max(p.max_dist_y, bw);  // lchain.rs:63-66
    for (int x = wid * 32 + lane; x < DENSE_LUT; x += NW * 32)     // first read after the barrier of tile 0
      sh->pen16[x] = (u16)__float2int_rz(__fadd_rn(__fmul_rn(p.chn_pen_gap, (float)x), G.half_log[min(x, max(bw, 0) + 1)]));
    if (wid > 0) {
      // ---- helpers: phase A of tile T + 1 while warp 0 walks tile T (tiles 0 and 1 have no far part) -----------------------
      for (int T = 0; T < ntile; ++T) {
        DPROF_T(ha0);
        if (wid == 2 && T >= 1) dense_reduce_tile(an, A, B, n, T - 1, lane);   // tile T - 1 was committed before the last barrier
        if (T + 1 < ntile && (wid & 3) != 0) dense_phase_a<NW>(G, sh, an, n, T + 1, bw, mdx, mdy, max_iter, wid, lane);
        DPROF_T(ha1);
        dense_bar_all<NW>();
        DPROF_T(ha2);
        if (wid == 1) { DPROF_ADD(8, ha0, ha1); DPROF_ADD(9, ha1, ha2); }
      }
#ifdef MM2_DENSE_PROF
      if (wid == 1 && lane == 0 && (pass == 1 || true)) { atomicAdd(&g_dense_prof[8], dprof[8]); atomicAdd(&g_dense_prof[9], dprof[9]); dprof[8] = dprof[9] = 0; }
#endif
      if (wid == 2) dense_reduce_tile(an, A, B, n, ntile - 1, lane);
      __threadfence_block();
      dense_bar_all<NW>();                           // every anchor's v / chain length / extents are in global memory
      // the rescue decision (lchain.rs:321-330) is taken by warp 0 and published through shared memory
      dense_bar_all<NW>();
      if (pass == 1 || !sh->rescue_flag) break;
      dense_bar_all<NW>();
      continue;
    }
    // ---- warp 0: the sequential walk -----------------------------------------------------------------------------------------
    int rj = -1, rx = 0, rq = 0, rsp = 0; u32 rhi = 0;           // ring slot of this lane (anchor j with j % 32 == lane)
    int rf = 0, rpp = -1;   // DP result of the slot; v / chain length / chain extents are derived by a helper warp (dense_reduce_tile)
    int bf = NEG_INF * 4, bi = 0;
    int blk_start = 0;   // first anchor of the rid/strand block that holds the last anchor of the previous tile
    ulonglong2 nxt = make_ulonglong2(0, 0);
    if (lane < n) nxt = an[lane];
    for (int T = 0; T < ntile; ++T) {
      const int i0 = T * 32;
      const int tile_n = min(32, n - i0);
      DenseBuf& SB = sh->buf[T & 1];
      const int far_hi = i0 - 32;
      const int jt_top = (far_hi >> 5) - 1;
      const ulonglong2 cur = nxt;
      if (i0 + 32 + lane < n) nxt = an[i0 + 32 + lane];
      const int cx = (int)(u32)cur.x, cq = (int)(u32)cur.y, csp = (int)((cur.y >> 32) & 0xff);
      const u32 chi = (u32)(cur.x >> 32);
      u32 workmask, bnd;
      {
        int px = __shfl_up_sync(0xFFFFFFFFu, cx, 1);
        u32 phi = __shfl_up_sync(0xFFFFFFFFu, chi, 1);
        const int rx31 = __shfl_sync(0xFFFFFFFFu, rx, 31);
        const u32 rhi31 = __shfl_sync(0xFFFFFFFFu, rhi, 31);
        if (lane == 0) { px = rx31; phi = rhi31; }
        const int i = i0 + lane;
        const bool work = lane < tile_n && i > 0 && (i - 1) >= wsub(i, max_iter) && phi == chi && !(cx > wadd(px, mdx));
        workmask = __ballot_sync(0xFFFFFFFFu, work);
        bnd = __ballot_sync(0xFFFFFFFFu, lane < tile_n && (i == 0 || phi != chi));   // anchors that start a rid/strand block
      }
      int done = 0;   // lanes [0, done) of this tile are committed to the ring
      while (workmask) {
        const int c = __ffs(workmask) - 1;
        workmask &= workmask - 1;
        const int i = i0 + c;
        DPROF_T(tp0);
        DPROF_INC(4, 1);
        if (done < c) {
          if (lane >= done && lane < c) {                        // anchors before i with an empty window (lchain.rs:77,89-90)
            rj = i0 + lane; rx = cx; rq = cq; rsp = csp; rhi = chi;
            rf = csp; rpp = -1;
          }
        }
        done = c + 1;
        const int ri = __shfl_sync(0xFFFFFFFFu, cx, c), qi = __shfl_sync(0xFFFFFFFFu, cq, c), spi = __shfl_sync(0xFFFFFFFFu, csp, c);
        const u32 hi_i = __shfl_sync(0xFFFFFFFFu, chi, c);
        const int low_iter = wsub(i, max_iter);                  // lchain.rs:78
        const bool inwin = rj >= max(low_iter, 0) && rhi == hi_i && !(ri > wadd(rx, mdx));   // empty slots hold rj = -1
        int s0;
        const bool valid = chain_sc_flat(inwin, ri, qi, rx, rq, rsp, mdx, mdy, bw, p.chn_pen_gap, p.chn_pen_skip, G.half_log, s0);
        const int sc = valid ? wadd(s0, rf) : NEG_INF;
        const u32 inmask = __ballot_sync(0xFFFFFFFFu, inwin);
        const u32 vmask = __ballot_sync(0xFFFFFFFFu, valid);
        int max_f = spi, max_j = -1, n_skip = 0;
        bool more = (inmask >> c) & 1u;                          // slot c holds j = i - 32: the window may go on beyond the ring
        if (vmask) {
          const int low_ring = max(i - 32, 0);
          const u32 markbits = __reduce_or_sync(0xFFFFFFFFu, (valid && rpp >= low_ring) ? (1u << (rpp & 31)) : 0u) & vmask;
          const u32 Mr = __funnelshift_r(markbits, markbits, c);
          const int m = __reduce_max_sync(0xFFFFFFFFu, sc);
          int hb = 32;
          if (m > spi) {
            const u32 eq = __ballot_sync(0xFFFFFFFFu, valid && sc == m);
            hb = 31 - __clz(__funnelshift_r(eq, eq, c));
          }
          if (hb == 32 || (Mr & (0xFFFFFFFEu << hb)) == 0u) {
            u32 after = Mr;
            if (hb < 32) { max_f = m; max_j = i - 32 + hb; after = Mr & ((1u << hb) - 1u); }
            const int cnt = __popc(after), need = max(max_skip + 1, 1);
            if (cnt >= need) {
              if (count_cells) cells += (unsigned)(nth_set_bit(__brev(after), need) + 1);
              more = false;
            } else {
              n_skip = cnt;
              if (count_cells) cells += (unsigned)__popc(inmask);
            }
          } else {
            const int sc_s = __shfl_sync(0xFFFFFFFFu, sc, (c - 1 - lane) & 31);
            const u32 Vs = __brev(__funnelshift_r(vmask, vmask, c)), As = __brev(__funnelshift_r(inmask, inmask, c));
            int rec_last;
            const bool brk = chain_tile_walk(lane, Vs, __brev(Mr), As, sc_s, max_skip, max_f, n_skip, rec_last, cells);
            if (rec_last >= 0) max_j = i - 1 - rec_last;
            if (brk) more = false;
          }
        } else if (count_cells) {
          cells += (unsigned)__popc(inmask);
        }
        DPROF_T(tp1);
        DPROF_ADD(0, tp0, tp1);
        if (more) {
          // ---- the window goes on beyond the ring --------------------------------------------------------------------------
          // first anchor of the rid/strand block of i: the last block boundary at or before lane c, else the previous tile's
          const u32 bm = bnd & low_mask(c + 1);
          const int blk0 = bm ? i0 + 31 - __clz(bm) : blk_start;
          const int fl = (far_hi > 0) ? SB.far_lo[c] : far_hi;   // first far predecessor (>= far_hi: none)
          u32* MK = &SB.MK[c][0];
          // (1) the ring slots mark their predecessors below the ring (lchain.rs:86): tile T-1's marks stay in a register
          //     (they are only read by the cells of step 2), the far tiles' go to the anchor's mark row in shared memory
          u32 m1;
          {
            const bool rm = valid && rpp >= 0 && rpp < i - 32;
            const int wdx = (i0 >> 5) - (rpp >> 5);              // 1: tile T-1, 2 + r: far tile r (0 cannot be: rpp < i - 32 < i0)
            m1 = __reduce_or_sync(0xFFFFFFFFu, (rm && wdx == 1) ? (1u << (rpp & 31)) : 0u);
            if (rm && wdx >= 2 && wdx < DENSE_MKW) atomicOr(&MK[wdx], 1u << (rpp & 31));
          }
          // (2) the rest of tile T-1: j = i - 33 - lane, down to the window start; its anchors are committed to shared memory
          bool brk = false;
          if (c > 0) {
            const int late_lo = max(max(far_hi, blk0), max(low_iter, 0));   // not below tile T-1, the block, the iteration bound
            const int j = i - 33 - lane;
            const bool inr = j >= late_lo;                       // inside tile T-1 (and the block); the distance test is per cell
            const int slot = ((j % DENSE_CAP) + DENSE_CAP) % DENSE_CAP;
            const int2 xq = sxq[slot], fp = sfp[slot];
            const bool act = inr && !(ri > wadd(xq.x, mdx));     // lchain.rs:75: monotone, so these are the cells down to start_j
            int s1;
            const bool v2 = chain_sc_flat(act, ri, qi, xq.x, xq.y, (int)ss[slot], mdx, mdy, bw, p.chn_pen_gap, p.chn_pen_skip, G.half_log, s1);
            const int sc2 = v2 ? wadd(s1, fp.x) : NEG_INF;
            {                                                    // lchain.rs:86 (all lanes first, see the header of this file)
              const bool lm = v2 && fp.y >= 0;
              const int wdx = (i0 >> 5) - (fp.y >> 5);
              m1 |= __reduce_or_sync(0xFFFFFFFFu, (lm && wdx == 1) ? (1u << (fp.y & 31)) : 0u);
              if (lm && wdx >= 2 && wdx < DENSE_MKW) atomicOr(&MK[wdx], 1u << (fp.y & 31));
            }
            const bool tm = v2 && ((m1 >> (j & 31)) & 1u);
            const u32 V2 = __ballot_sync(0xFFFFFFFFu, v2), M2 = __ballot_sync(0xFFFFFFFFu, tm), A2 = __ballot_sync(0xFFFFFFFFu, act);
            if (__ballot_sync(0xFFFFFFFFu, v2 && sc2 > max_f) == 0u) {
              // no record among these cells (lchain.rs:84): n_skip only grows, by one per marked cell
              const int cnt = __popc(M2), need = max(max_skip + 1 - n_skip, 1);
              if (cnt >= need) { brk = true; if (count_cells) cells += (unsigned)(nth_set_bit(M2, need) + 1); }
              else { n_skip += cnt; if (count_cells) cells += (unsigned)__popc(A2); }
            } else {
              int rec_last;
              brk = chain_tile_walk(lane, V2, M2, A2, sc2, max_skip, max_f, n_skip, rec_last, cells);
              if (rec_last >= 0) max_j = i - 33 - rec_last;
            }
            // the window ends inside tile T-1 when some cell of it fails the distance test (or the bounds above cut it)
            if (A2 != low_mask(c)) more = false;
          }
          __syncwarp();                                          // the far marks of steps 1 and 2 are in shared memory
          DPROF_T(tp2);
          DPROF_ADD(1, tp1, tp2);
          // (3) the far tiles, through their summaries: lane t holds tile r = base + t
          if (!brk && more && fl < far_hi) {
            const int nr = jt_top - (fl >> 5) + 1;
            const int slot_top0 = (jt_top * 32) % DENSE_CAP;
            for (int base = 0; base < nr && !brk; base += 32) {
              DPROF_INC(5, 1);
              const int rr = base + lane;
              const bool have = rr < nr;
              const u32 tV = have ? SB.V[c][rr] : 0u;
              const u32 tMa = have ? SB.MK[c][2 + rr] : 0u;
              const int tmx = have ? SB.TM[c][rr] : NEG_INF;
              const u32 tM = __brev(tMa);                        // visiting order
              // cells of the tile inside the window: all 32, except in the window's lowest tile
              const int jt = jt_top - rr;
              u32 tA = have ? 0xFFFFFFFFu : 0u;
              if (have && jt * 32 < fl) tA = low_mask(jt * 32 + 32 - fl);
              const int nround = min(32, nr - base);
              int curt = 0;
              while (curt < nround) {
                // tiles [curt, tf) hold no score above max_f, i.e. no record (lchain.rs:84): n_skip only grows there
                const u32 cand = __ballot_sync(0xFFFFFFFFu, lane >= curt && lane < nround && tmx > max_f);
                const int tf = cand ? (__ffs(cand) - 1) : nround;
                if (tf > curt) {
                  const bool mine = lane >= curt && lane < tf;
                  const int cnt = mine ? __popc(tM & tV) : 0;
                  const int total = __reduce_add_sync(0xFFFFFFFFu, cnt);
                  const int need = max(max_skip + 1 - n_skip, 1);
                  if (total >= need) {                            // lchain.rs:85: the need-th marked slot breaks the loop
                    if (count_cells) {
                      int incl = cnt;
#pragma unroll
                      for (int d = 1; d < 32; d <<= 1) {
                        const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                        if (lane >= d) incl += o;
                      }
                      const u32 hit_m = __ballot_sync(0xFFFFFFFFu, incl >= need);
                      const int tb = __ffs(hit_m) - 1;
                      const int before = __reduce_add_sync(0xFFFFFFFFu, (mine && lane < tb) ? __popc(tA) : 0);
                      const int excl_tb = __shfl_sync(0xFFFFFFFFu, incl - cnt, tb);
                      const u32 mtb = __shfl_sync(0xFFFFFFFFu, tM & tV, tb);
                      cells += (unsigned)(before + nth_set_bit(mtb, need - excl_tb) + 1);
                    }
                    brk = true;
                    break;
                  }
                  n_skip += total;
                  if (count_cells) cells += (unsigned)__reduce_add_sync(0xFFFFFFFFu, mine ? __popc(tA) : 0);
                }
                if (tf >= nround) break;
                // a record tile: its cells one by one (recomputed: the predecessors' state is in shared memory)
                DPROF_INC(6, 1);
                {
                  const int jtf = jt_top - (base + tf);
                  const int j = jtf * 32 + 31 - lane;            // visiting order
                  int slot = slot_top0 - 32 * (base + tf); if (slot < 0) slot += DENSE_CAP;
                  slot += 31 - lane;
                  const int2 xq = sxq[slot], fp = sfp[slot];
                  const bool act = j >= fl;
                  int s1;
                  const bool v2 = chain_sc_flat(act, ri, qi, xq.x, xq.y, (int)ss[slot], mdx, mdy, bw, p.chn_pen_gap, p.chn_pen_skip, G.half_log, s1);
                  const int sc2 = v2 ? wadd(s1, fp.x) : NEG_INF;
                  const u32 Vt = __shfl_sync(0xFFFFFFFFu, tV, tf), Mt = __shfl_sync(0xFFFFFFFFu, tM, tf), At = __shfl_sync(0xFFFFFFFFu, tA, tf);
                  int rec_last;
                  const bool b2 = chain_tile_walk(lane, Vt, Mt, At, sc2, max_skip, max_f, n_skip, rec_last, cells);
                  if (rec_last >= 0) max_j = jtf * 32 + 31 - rec_last;
                  if (b2) { brk = true; break; }
                }
                curt = tf + 1;
              }
            }
          }
          DPROF_T(tp3);
          DPROF_ADD(2, tp2, tp3);
        }
        if (lane == c) {                                         // anchor i takes over its ring slot
          rj = i; rx = cx; rq = cq; rsp = csp; rhi = chi;
          rf = max_f; rpp = max_j;                               // lchain.rs:89
        }
      }
      if (lane >= done && lane < tile_n) {
        rj = i0 + lane; rx = cx; rq = cq; rsp = csp; rhi = chi;
        rf = csp; rpp = -1;
      }
      if (lane < tile_n) {                                       // every lane now holds its own anchor of this tile
        *reinterpret_cast<int2*>(&A[i0 + lane]) = make_int2(rf, rpp);   // {v, chain length} and B follow from dense_reduce_tile
        const int slot = (i0 + lane) % DENSE_CAP;                // ... and they enter the shared window ring
        sh->sxq[slot] = make_int2(rx, rq); sh->sfp[slot] = make_int2(rf, rpp); sh->ss[slot] = (u8)rsp;
      }
      if (bnd) blk_start = i0 + 31 - __clz(bnd);
      if (lane == 0) {
        // what phase A of tile T + 2 needs to know about the last committed anchor (i0 + tile_n - 1)
        sh->cm_blk[T & 1] = blk_start;
      }
      if (lane == tile_n - 1) sh->cm_hi[T & 1] = chi;
      {  // lchain.rs:163: the LAST maximum of f
        const int fl2 = lane < tile_n ? rf : NEG_INF * 4;
        const int m = __reduce_max_sync(0xFFFFFFFFu, fl2);
        if (m >= bf) {
          const u32 eq = __ballot_sync(0xFFFFFFFFu, fl2 == m);
          bf = m;
          bi = i0 + 31 - __clz(eq);
        }
      }
      __threadfence_block();
      DPROF_T(tb0);
      dense_bar_all<NW>();                                       // tile T is committed; the summaries of tile T + 1 are complete
      DPROF_T(tb1);
      DPROF_ADD(3, tb0, tb1);
      DPROF_INC(10, 1);
    }
    dense_bar_all<NW>();                                         // dense_reduce_tile has finished the last tile
    best = bi;
    bestA = A[best]; bestB = B[best];
    bool rescue = false;
    if (pass == 0 && G.do_rescue) {
      // lchain.rs:321-330 rescue_long_join on the single (fallback) chain
      const ulonglong2 ab = an[best];
      const int qe = wadd((int)(u32)ab.y, 1);
      const int qs = max(bestB.x, 0);
      const int best_cov = max(wsub(qe, qs), 0);
      const int uncovered = max(wsub(qlen, best_cov), 0);
      rescue = uncovered > p.rmq_rescue_size || (float)best_cov < __fmul_rn((float)qlen, __fsub_rn(1.0f, p.rmq_rescue_ratio));
    }
    if (lane == 0) sh->rescue_flag = rescue ? 1 : 0;
    __threadfence_block();
    dense_bar_all<NW>();                                         // the helpers read the decision
    if (!rescue) break;
    hit.flags |= 1u;
    dense_bar_all<NW>();                                         // ... before the flag is rewritten
  }
  if (wid == 0) chain_finish(G, r, lane, an, A, a0, qlen, m0, m1, best, bestA, bestB, hit, cells);
#ifdef MM2_DENSE_PROF
  if (wid == 0 && lane == 0) {
    dprof[7] = (unsigned long long)(clock64() - dprof_t0);
    for (int x = 0; x < 12; ++x) if (x != 8 && x != 9) atomicAdd(&g_dense_prof[x], dprof[x]);
  }
#endif
}

template <bool COUNT, bool FAST>
__global__ void __launch_bounds__(CH_WARPS * 32, MM2_CH_OCC) chain_ring_kernel(ChainArgs G) {
  const u32 r = blockIdx.x * CH_WARPS + (threadIdx.x >> 5);
  if (r >= G.nreads) return;
  if (chain_is_dense(G, r)) return;                               // chain_dense_kernel's
  chain_read<COUNT, FAST>(G, r, threadIdx.x & 31);
}

// NW warps per dense read, one resident CTA per SM (the window ring and the two summary buffers take ~210 KB of shared memory).
template <int NW>
__global__ void __launch_bounds__(NW * 32, 1) chain_dense_kernel(ChainArgs G) {
  __shared__ DenseSh sh;
  extern __shared__ __align__(16) unsigned char dense_dyn[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const u32 ndense = G.dense[0];
  if (threadIdx.x == 0) {
    sh.sxq = reinterpret_cast<int2*>(dense_dyn);
    sh.sfp = sh.sxq + DENSE_CAP;
    sh.buf = reinterpret_cast<DenseBuf*>(dense_dyn + DENSE_CAP * 16);
    sh.pen16 = reinterpret_cast<u16*>(dense_dyn + DENSE_CAP * 16 + 2 * sizeof(DenseBuf));
    sh.ss = reinterpret_cast<u8*>(dense_dyn + DENSE_CAP * 16 + 2 * sizeof(DenseBuf) + DENSE_LUT * 2);
  }
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) sh.next = atomicAdd(&G.dense[1], 1u);
    __syncthreads();
    const u32 k = sh.next;
    if (k >= ndense) return;
    chain_read_dense<NW>(G, dense_read_of_ticket(G, k), lane, wid, &sh);
  }
}


}  // namespace

// Opt-in to more than 48 KB of dynamic shared memory.  The attribute is per device, so mm2_ctx_create calls this for
// every context (after cudaSetDevice) instead of once per process.
int lchain_init_device() {
  CUDA_TRY(cudaFuncSetAttribute(chain_dense_kernel<DENSE_NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, DENSE_DYN));
  return MM2_OK;
}

// host-built table of 0.5 * mg_log2(dd + 1) (lchain.rs:15,30-31): glibc logf, division by f32 LN_2, exact halving
static std::vector<float> build_half_log(int n) {
  std::vector<float> t((size_t)n);
  for (int dd = 0; dd < n; ++dd) {
    float log_pen = 0.0f;
    if (dd >= 1) {
      const int x = dd + 1;
      volatile float lx = logf((float)x);
      volatile float q = lx / 0.6931472f;
      log_pen = x <= 1 ? 0.0f : q;
    }
    volatile float h = 0.5f * log_pen;
    t[(size_t)dd] = h;
  }
  return t;
}

int chain_batch(mm2_ctx* ctx, const ulonglong2* d_anchors, const u64* d_read_aoff, const u64* d_read_off, const u64* d_mini_off,
                const u64* d_mval, const u32* d_sum_span, u32 nreads, const mm2_chain_params_t& p, int do_rescue, int4* d_A,
                int4* d_B, int* d_T, int* d_W, int* d_chain, ReadHit* d_hits, unsigned long long* d_cells) {
  if (!nreads) return MM2_OK;
  const int max_bw = std::max(p.bw, do_rescue ? p.bw_long : p.bw);
  if (max_bw < 0 || max_bw > (1 << 26)) { mm2_set_error("chain: bandwidth out of range"); return MM2_E_ARG; }
  const int nl = max_bw + 2;
  if (ctx->lut_n < nl || ctx->lut_gap != p.chn_pen_gap) {
    // the tables depend on dd (and the second one on chn_pen_gap) only: upload once per context and keep them (no per-batch
    // H2D).  [0, n2): 0.5 * mg_log2(dd + 1); [n2, 2 n2): the integer penalty int(chn_pen_gap * dd + 0.5 * mg_log2(dd + 1)) of
    // lchain.rs:28-32 for chn_pen_skip == 0, with the reference's float operations in the reference's order
    const int n2 = std::max(nl, ctx->lut_n);
    std::vector<float> hl = build_half_log(n2);
    std::vector<int> pi((size_t)n2);
    for (int dd = 0; dd < n2; ++dd) {
      volatile float lin = p.chn_pen_gap * (float)dd;
      volatile float sum = lin + hl[(size_t)dd];
      pi[(size_t)dd] = (int)sum;
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));   // earlier launches of this context may still read the old tables
    MM2_TRY(ctx->lut.ensure((size_t)n2 * 8));
    MM2_TRY(ctx->pin_small.ensure((size_t)n2 * 8));
    memcpy(ctx->pin_small.p, hl.data(), (size_t)n2 * 4);
    memcpy((u8*)ctx->pin_small.p + (size_t)n2 * 4, pi.data(), (size_t)n2 * 4);
    CUDA_TRY(cudaMemcpyAsync(ctx->lut.p, ctx->pin_small.p, (size_t)n2 * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->lut_n = n2; ctx->lut_gap = p.chn_pen_gap;
  }
  ChainArgs G;
  G.anchors = d_anchors; G.read_aoff = d_read_aoff; G.read_off = d_read_off; G.mini_off = d_mini_off; G.mval = d_mval;
  G.sum_span = d_sum_span; G.nreads = nreads; G.p = p; G.do_rescue = do_rescue; G.half_log = ctx->lut.as<float>();
  G.pen_int = reinterpret_cast<const int*>(ctx->lut.as<float>() + ctx->lut_n);
  const bool lut = p.chn_pen_skip == 0.0f && p.max_chain_iter >= 32 && p.bw >= 0 && p.bw_long >= 0;   // chain_read<.., FAST>
  G.A = d_A; G.B = d_B; G.T = d_T; G.W = d_W; G.chain = d_chain; G.hits = d_hits; G.cells = d_cells;
  G.dense = nullptr; G.dense_min = 0x7fffffff; G.dense_ratio5 = 0;
  const int grid = (int)((nreads + CH_WARPS - 1) / CH_WARPS);
  {
    MM2_TRY(ctx->read_class.ensure(((size_t)nreads * DENSE_BINS + 16) * 4));
    G.dense = ctx->read_class.as<u32>();
    // the CTA-per-read kernel keeps the window in shared memory: larger max_chain_iter values stay with the warp-per-read kernel
    G.dense_min = p.max_chain_iter <= DENSE_CAP - 64 ? ctx->chain_dense_min : 0x7fffffff;
    G.dense_ratio5 = ctx->chain_dense_ratio5;
    CUDA_TRY(cudaMemsetAsync(G.dense, 0, 64, ctx->stream));
    MM2_LAUNCH(ctx, chain_classify_kernel, (int)((nreads + 255) / 256), 256, 0, G);
    if (d_cells) { if (lut) MM2_LAUNCH(ctx, (chain_ring_kernel<true, true>), grid, CH_WARPS * 32, 0, G); else MM2_LAUNCH(ctx, (chain_ring_kernel<true, false>), grid, CH_WARPS * 32, 0, G); }
    else { if (lut) MM2_LAUNCH(ctx, (chain_ring_kernel<false, true>), grid, CH_WARPS * 32, 0, G); else MM2_LAUNCH(ctx, (chain_ring_kernel<false, false>), grid, CH_WARPS * 32, 0, G); }
    // persistent CTAs (one per SM), one dense read at a time each; with no dense read they exit at once
    if (G.dense_min != 0x7fffffff) MM2_LAUNCH(ctx, chain_dense_kernel<DENSE_NW>, (int)std::min<u64>(nreads, (u64)ctx->n_sm), DENSE_NW * 32, DENSE_DYN, G);
  }
#ifdef MM2_DENSE_PROF
  {
    cudaStreamSynchronize(ctx->stream);
    unsigned long long h[12];
    cudaMemcpyFromSymbol(h, g_dense_prof, sizeof h);
    if (h[4]) fprintf(stderr, "[dense prof] anchors-with-window %llu tiles %llu far-rounds %llu record-tiles %llu far-pairs %llu | warp 0 cycles per anchor: ring %.0f late %.0f far %.0f barrier-wait %.0f total %.0f | helper warp: phase A %.0f wait %.0f cycles per tile\n",
                      h[4], h[10], h[5], h[6], h[11], (double)h[0] / h[4], (double)h[1] / h[4], (double)h[2] / h[4], (double)h[3] / h[4], (double)h[7] / h[4],
                      (double)h[8] / std::max(1ull, h[10]), (double)h[9] / std::max(1ull, h[10]));
    memset(h, 0, sizeof h);
    cudaMemcpyToSymbol(g_dense_prof, h, sizeof h);
  }
#endif
  CUDA_TRY(cudaGetLastError());
  return MM2_OK;
}
