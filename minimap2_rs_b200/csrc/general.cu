// general.cu — the non-default tail of `mm2rs align` (main.rs:209-218) for `-n < 2`, where chain_dp_all really returns
// many chains: backtrack extraction (lchain.rs:93-160), rescue_long_join (:321-330),
// merge_adjacent_chains_with_gap (:288-314), select_and_filter_chains (:220-260) and one PAF record per kept chain
// (paf.rs:130-222, 238-248).  The forward DP — the only heavy part — still runs on the device (chain_kernel, once with
// bw and, if any read asks for the rescue, once more with bw_long); this file is the per-read glue over a handful of
// tiny chains, executed by host threads on the downloaded DP state.
#include "mm2_internal.cuh"

#include <algorithm>
#include <cmath>
#include <thread>

#include "stages.cuh"

namespace {

struct Chain { std::vector<u32> idx; i32 score; };

inline i32 aq(const mm2_anchor_t& a) { return (i32)(u32)a.y; }
inline i32 asp(const mm2_anchor_t& a) { return (i32)((a.y >> 32) & 0xff); }
inline i32 ar(const mm2_anchor_t& a) { return (i32)(u32)a.x; }
inline bool arev(const mm2_anchor_t& a) { return (a.x >> 63) != 0; }
inline i32 arid(const mm2_anchor_t& a) { return (i32)((a.x >> 32) & 0x7fffffffULL); }
inline i32 wsub(i32 a, i32 b) { return (i32)((u32)a - (u32)b); }
inline i32 wadd(i32 a, i32 b) { return (i32)((u32)a + (u32)b); }

// lchain.rs:179-200: (start clamped at 0, end) over the chain's anchors on the query / target axis
void qrange(const mm2_anchor_t* a, const std::vector<u32>& c, i32* s, i32* e) {
  i32 lo = INT32_MAX, hi = -1;
  for (u32 i : c) { lo = std::min(lo, wsub(aq(a[i]), asp(a[i]) - 1)); hi = std::max(hi, wadd(aq(a[i]), 1)); }
  *s = std::max(lo, 0); *e = hi;
}
void trange(const mm2_anchor_t* a, const std::vector<u32>& c, i32* s, i32* e) {
  i32 lo = INT32_MAX, hi = -1;
  for (u32 i : c) { lo = std::min(lo, wsub(ar(a[i]), asp(a[i]) - 1)); hi = std::max(hi, wadd(ar(a[i]), 1)); }
  *s = std::max(lo, 0); *e = hi;
}

// lchain.rs:202-218: stable order by (score desc, query start asc, target start asc)
void sort_chains(const mm2_anchor_t* a, std::vector<Chain>& ch) {
  struct Key { i32 sc, qs, ts; size_t i; };
  std::vector<Key> k(ch.size());
  for (size_t i = 0; i < ch.size(); ++i) { i32 e; k[i].sc = ch[i].score; qrange(a, ch[i].idx, &k[i].qs, &e); trange(a, ch[i].idx, &k[i].ts, &e); k[i].i = i; }
  std::stable_sort(k.begin(), k.end(), [](const Key& x, const Key& y) {
    if (x.sc != y.sc) return y.sc < x.sc;
    if (x.qs != y.qs) return x.qs < y.qs;
    return x.ts < y.ts;
  });
  std::vector<Chain> o;
  o.reserve(ch.size());
  for (auto& kk : k) o.push_back(std::move(ch[kk.i]));
  ch.swap(o);
}

// lchain.rs:93-176 given the forward DP (f, pprev, v).  The walk-back helper never takes a second step (F3), so every
// candidate is {i0} with score f[i0] - f[pprev[i0]] (or empty with score 0), visited by (f desc, index desc).
void chains_from_dp(const mm2_anchor_t* a, const int4* A, size_t n, const mm2_chain_params_t& p, std::vector<Chain>& out) {
  out.clear();
  if (n == 0) return;
  std::vector<u32> order;
  for (size_t i = 0; i < n; ++i) if (A[i].x > 0) order.push_back((u32)i);
  if (order.empty()) return;
  std::stable_sort(order.begin(), order.end(), [&](u32 x, u32 y) { return A[x].x < A[y].x; });
  for (size_t q = order.size(); q-- > 0;) {
    const u32 i0 = order[q];
    const i32 prev = A[i0].y;
    const i32 s = prev < 0 ? A[i0].x : wsub(A[i0].x, A[prev].x);
    Chain c;
    if (s > 0) { c.idx.push_back(i0); c.score = s; } else c.score = 0;
    if (c.score >= p.min_chain_score && (i32)c.idx.size() >= p.min_cnt) out.push_back(std::move(c));
  }
  if (out.empty()) {  // lchain.rs:162-173
    size_t best = 0;
    for (size_t i = 1; i < n; ++i) if (A[i].x >= A[best].x) best = i;
    Chain c;
    for (i32 i = (i32)best; i >= 0; i = A[i].y) c.idx.push_back((u32)i);
    std::reverse(c.idx.begin(), c.idx.end());
    c.score = A[best].z;
    out.push_back(std::move(c));
  }
  sort_chains(a, out);
}

// lchain.rs:288-314
std::vector<std::vector<u32>> merge_with_gap(const mm2_anchor_t* a, const std::vector<Chain>& ch, i32 gq, i32 gt) {
  std::vector<std::pair<i32, size_t>> items;
  for (size_t i = 0; i < ch.size(); ++i) { i32 s, e; qrange(a, ch[i].idx, &s, &e); items.emplace_back(s, i); }
  std::stable_sort(items.begin(), items.end(), [](const std::pair<i32, size_t>& x, const std::pair<i32, size_t>& y) { return x.first < y.first; });  // F10
  std::vector<std::vector<u32>> m;
  for (auto& it : items) {
    const std::vector<u32>& c = ch[it.second].idx;
    if (m.empty()) { m.push_back(c); continue; }
    std::vector<u32>& last = m.back();
    // an empty chain has no first/last anchor: the reference would panic on `.unwrap()`; it cannot be produced with
    // min_cnt >= 1, and with min_cnt <= 0 the caller rejects the options
    const mm2_anchor_t& al = a[last.back()];
    const mm2_anchor_t& af = a[c.front()];
    const bool same = arid(al) == arid(af) && arev(al) == arev(af);
    i32 lqs, lqe, cqs, cqe, lts, lte, cts, cte;
    qrange(a, last, &lqs, &lqe); qrange(a, c, &cqs, &cqe); trange(a, last, &lts, &lte); trange(a, c, &cts, &cte);
    const i32 qg = wsub(cqs, lqe), tg = wsub(cts, lte);
    if (same && qg >= 0 && tg >= 0 && qg <= gq && tg <= gt) last.insert(last.end(), c.begin(), c.end());
    else m.push_back(c);
  }
  return m;
}

struct Kept { std::vector<u32> idx; bool primary; };

// lchain.rs:220-260 (main.rs:216-217 passes the MERGED chains with the UNMERGED scores: scores[i] for i < n_merged)
void select_filter(const mm2_anchor_t* a, std::vector<std::vector<u32>>& merged, const std::vector<i32>& scores_in, float mask_level,
                   float pri_ratio, size_t best_n, std::vector<Kept>& out, i32* s1, i32* s2) {
  out.clear(); *s1 = 0; *s2 = 0;
  if (merged.empty()) return;
  std::vector<Chain> ch(merged.size());
  for (size_t i = 0; i < merged.size(); ++i) { ch[i].idx = std::move(merged[i]); ch[i].score = scores_in[i]; }
  sort_chains(a, ch);
  std::vector<std::pair<i32, i32>> pri;
  std::vector<char> is_pri(ch.size(), 1);
  for (size_t ci = 0; ci < ch.size(); ++ci) {
    i32 qs, qe; qrange(a, ch[ci].idx, &qs, &qe);
    bool ov = false;
    for (auto& pr : pri) {
      const float o = (float)std::max(wsub(std::min(qe, pr.second), std::max(qs, pr.first)), 0);
      const float len = (float)std::max(wsub(qe, qs), 1);
      if (o / len >= mask_level) { ov = true; break; }
    }
    if (ov) is_pri[ci] = 0; else pri.emplace_back(qs, qe);
  }
  *s1 = ch[0].score;
  size_t sec = 0;
  for (size_t i = 0; i < ch.size(); ++i) {
    if (i == 0) { out.push_back(Kept{ch[i].idx, true}); continue; }
    if (!is_pri[i]) continue;
    if ((float)ch[i].score >= pri_ratio * (float)*s1 && sec < best_n) { out.push_back(Kept{ch[i].idx, false}); sec += 1; }
    if (*s2 == 0) *s2 = ch[i].score;
  }
}

// paf.rs:130-222 for one chain; mini_pos / sum_k / n_mini describe the query sketched with the index's w,k (paf.rs:156)
// returns false when the reference would panic (idx.seq[rid0] out of bounds)
bool paf_record(const mm2_index* idx, const mm2_anchor_t* a, const std::vector<u32>& chain, i32 qlen, const u64* mval, size_t n_mini,
                u64 sum_k, bool primary, mm2_paf_rec_t& rec) {
  const bool rev = arev(a[chain[0]]);
  i32 qs, qe, ts, te;
  qrange(a, chain, &qs, &qe); trange(a, chain, &ts, &te);
  const u32 rid0 = (u32)((a[chain[0]].x >> 32) & 0x7fffffffULL);
  if (rid0 >= idx->lens.size()) return false;
  rec.rid = rid0; rec.qlen = (u32)qlen; rec.qstart = (u32)qs; rec.qend = (u32)qe; rec.tlen = idx->lens[rid0];
  rec.tstart = (u32)ts; rec.tend = (u32)te; rec.nm = (u32)std::max(wsub(qe, qs), 0); rec.blen = (u32)std::max(wsub(te, ts), 0);
  rec.cm = (u32)chain.size(); rec.s1 = rec.s2 = 0; rec.rl = 0; rec.strand = rev ? '-' : '+'; rec.mapq = 60; rec.tp = primary ? 'P' : 'S';
  rec.flags = 0;
  const float avg_k = n_mini ? (float)sum_k / (float)n_mini : (float)idx->k;
  auto fwd = [&](const mm2_anchor_t& x) { return arev(x) ? wsub(wsub(qlen, 1), wsub(wadd(aq(x), 1), asp(x))) : aq(x); };
  std::vector<i32> cq;
  cq.reserve(chain.size());
  if (rev) for (size_t t = chain.size(); t-- > 0;) cq.push_back(fwd(a[chain[t]]));
  else for (u32 i : chain) cq.push_back(fwd(a[i]));
  auto mp = [&](size_t j) { return (i32)(u32)((mval[j] >> 1) & 0xffffffffULL); };
  float dv = 0.0f;
  if (n_mini && !cq.empty()) {
    const i32 first = cq[0];
    size_t lo = 0, hi = n_mini, st = 0;
    bool found = false;
    while (lo < hi) {  // slice::binary_search, then rewind to the first equal element (paf.rs:178-180)
      const size_t mid = lo + (hi - lo) / 2;
      if (mp(mid) == first) { found = true; st = mid; break; }
      if (mp(mid) < first) lo = mid + 1; else hi = mid;
    }
    if (found) {
      while (st > 0 && mp(st - 1) == first) st -= 1;
      size_t j = st, kq = 1, en = st;
      i32 n_match = 1;
      while (j + 1 < n_mini && kq < cq.size()) { j += 1; if (mp(j) == cq[kq]) { n_match += 1; en = j; kq += 1; } }
      i32 n_tot = (i32)(en - st + 1);
      const i32 r_qs = rev ? qlen - qe : qs, r_qe = rev ? qlen - qs : qe;
      const i32 ak = avg_k != avg_k ? 0 : (avg_k >= 2147483648.0f ? INT32_MAX : (avg_k <= -2147483648.0f ? INT32_MIN : (i32)avg_k));
      if (r_qs > ak && ts > ak) n_tot += 1;
      if ((qlen - r_qe) > ak && ((i32)rec.tlen - te) > ak) n_tot += 1;
      const float frac = (float)n_match / (float)n_tot;
      dv = frac >= 1.0f ? 0.0f : 1.0f - powf(frac, 1.0f / std::max(avg_k, 1.0f));
    }
  }
  rec.dv = dv;
  return true;
}

}  // namespace

// Called by map_device_impl once the sorted anchors of the batch are resident (ctx->anchors / ctx->read_aoff).
int map_general_finish(mm2_ctx* ctx, const mm2_index* idx, const u64* d_read_off, const u64* h_off, size_t nreads,
                       const mm2_map_opts_t* o, const mm2_chain_params_t& p, const SketchOut& so, const u32* d_sum_span, u64 nm, u64 na,
                       mm2_map_result_t* out) {
  cudaStream_t st = ctx->stream;
  if (p.min_cnt < 1) {
    mm2_set_error("mm2_map_batch: -n < 1 lets empty chains through, on which the reference panics (lchain.rs:300 unwrap)");
    return MM2_E_REF_PANIC;
  }
  std::vector<mm2_anchor_t> anchors((size_t)na);
  std::vector<int4> A1((size_t)na), A2;
  std::vector<u64> aoff(nreads + 1, 0), moff(nreads + 1, 0), mval((size_t)nm);
  std::vector<u32> sum_span(nreads, 0);
  auto run_dp = [&](const mm2_chain_params_t& pp, std::vector<int4>& dst) -> int {
    MM2_TRY(chain_batch(ctx, ctx->anchors.as<ulonglong2>(), ctx->read_aoff.as<u64>(), d_read_off, so.seq_off, so.val, d_sum_span, (u32)nreads,
                        pp, 0, ctx->dpA.as<int4>(), ctx->dpB.as<int4>(), ctx->dpT.as<int>(), ctx->dpW.as<int>(), nullptr,
                        ctx->hits.as<ReadHit>(), nullptr));
    if (na) CUDA_TRY(cudaMemcpyAsync(dst.data(), ctx->dpA.p, na * 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return MM2_OK;
  };
  ctx->timer.mark(st, "chain");
  MM2_TRY(run_dp(p, A1));
  if (na) CUDA_TRY(cudaMemcpyAsync(anchors.data(), ctx->anchors.p, na * 16, cudaMemcpyDeviceToHost, st));
  if (nm) CUDA_TRY(cudaMemcpyAsync(mval.data(), so.val, nm * 8, cudaMemcpyDeviceToHost, st));
  if (nreads) {
    CUDA_TRY(cudaMemcpyAsync(aoff.data(), ctx->read_aoff.p, (nreads + 1) * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(moff.data(), so.seq_off, (nreads + 1) * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(sum_span.data(), d_sum_span, nreads * 4, cudaMemcpyDeviceToHost, st));
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  // first extraction + rescue decision (lchain.rs:321-326)
  std::vector<std::vector<Chain>> chains(nreads);
  std::vector<char> rescue(nreads, 0);
  const int nth = nreads >= 2048 ? 8 : 1;
  auto par = [&](auto fn) {
    std::vector<std::thread> th;
    for (int t = 1; t < nth; ++t) th.emplace_back([&, t]() { for (size_t r = (size_t)t; r < nreads; r += (size_t)nth) fn(r); });
    for (size_t r = 0; r < nreads; r += (size_t)nth) fn(r);
    for (auto& t : th) t.join();
  };
  par([&](size_t r) {
    const size_t n = (size_t)(aoff[r + 1] - aoff[r]);
    if (!n) return;
    const mm2_anchor_t* a = anchors.data() + aoff[r];
    chains_from_dp(a, A1.data() + aoff[r], n, p, chains[r]);
    if (chains[r].empty()) return;
    const i32 qlen = (i32)(h_off[r + 1] - h_off[r]);
    i32 qs, qe; qrange(a, chains[r][0].idx, &qs, &qe);
    const i32 cov = std::max(wsub(qe, qs), 0), unc = std::max(wsub(qlen, cov), 0);
    rescue[r] = unc > p.rmq_rescue_size || (float)cov < (float)qlen * (1.0f - p.rmq_rescue_ratio);
  });
  bool any_rescue = false;
  for (char c : rescue) any_rescue |= c != 0;
  if (any_rescue) {  // lchain.rs:327-329: rerun with bw = bw_long and REPLACE chains + scores
    mm2_chain_params_t p2 = p;
    p2.bw = p.bw_long;
    A2.resize((size_t)na);
    MM2_TRY(run_dp(p2, A2));
    par([&](size_t r) {
      if (!rescue[r]) return;
      const size_t n = (size_t)(aoff[r + 1] - aoff[r]);
      chains_from_dp(anchors.data() + aoff[r], A2.data() + aoff[r], n, p2, chains[r]);
    });
  }
  ctx->timer.mark(st, "d2h");
  ctx->timer.mark(st, "end");
  CUDA_TRY(cudaStreamSynchronize(st));
  ctx->timer.finish();
  // merge, select, records
  std::vector<std::vector<mm2_paf_rec_t>> recs(nreads);
  std::vector<char> panic(nreads, 0);
  par([&](size_t r) {
    if (chains[r].empty()) return;
    const mm2_anchor_t* a = anchors.data() + aoff[r];
    const i32 qlen = (i32)(h_off[r + 1] - h_off[r]);
    std::vector<i32> scores;
    for (auto& c : chains[r]) scores.push_back(c.score);
    std::vector<std::vector<u32>> merged = merge_with_gap(a, chains[r], p.max_dist_y, p.max_dist_y);
    std::vector<Kept> kept;
    i32 s1, s2;
    select_filter(a, merged, scores, o->mask_level, o->pri_ratio, (size_t)std::max(o->best_n, 0), kept, &s1, &s2);
    for (size_t ci = 0; ci < kept.size(); ++ci) {
      mm2_paf_rec_t rec;
      if (!paf_record(idx, a, kept[ci].idx, qlen, mval.data() + moff[r], (size_t)(moff[r + 1] - moff[r]), sum_span[r], ci == 0, rec)) {
        panic[r] = 1; recs[r].clear(); break;
      }
      rec.read_id = (u32)r; rec.s1 = (u32)std::max(s1, 0); rec.s2 = (u32)std::max(s2, 0); rec.flags = rescue[r] ? 1 : 0;
      recs[r].push_back(rec);
    }
  });
  memset(out, 0, sizeof *out);
  out->n_reads = nreads; out->n_bases = nreads ? h_off[nreads] - h_off[0] : 0; out->n_minimizers = nm; out->n_anchors = na;
  size_t tot = 0, npan = 0;
  for (size_t r = 0; r < nreads; ++r) { tot += recs[r].size(); npan += panic[r]; out->n_rescued += rescue[r]; }
  out->recs = (mm2_paf_rec_t*)malloc(std::max<size_t>(1, tot) * sizeof(mm2_paf_rec_t));
  out->panic_reads = (u32*)malloc(std::max<size_t>(1, npan) * 4);
  for (size_t r = 0; r < nreads; ++r) {
    for (auto& rec : recs[r]) out->recs[out->n_recs++] = rec;
    if (panic[r]) out->panic_reads[out->n_panic++] = (u32)r;
  }
  return MM2_OK;
}
