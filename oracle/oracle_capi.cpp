// oracle_capi.cpp — flat C entry points over the CPU restatement, for ctypes.
// TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs.  Never linked into libmm2b200.so.
#include "mm2_oracle.hpp"

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>

using namespace orc;

extern "C" {

void orc_free(void* p) { free(p); }

int orc_sketch(const uint8_t* seq, size_t len, int w, int k, uint32_t rid, int is_hpc, Minimizer** out, size_t* n) {
  std::vector<Minimizer> v;
  sketch_sequence(seq, len, (size_t)w, (size_t)k, rid, is_hpc != 0, v);
  *n = v.size();
  *out = (Minimizer*)malloc(std::max<size_t>(1, v.size()) * sizeof(Minimizer));
  if (!v.empty()) memcpy(*out, v.data(), v.size() * sizeof(Minimizer));
  return 0;
}

// sequences are given as one concatenated byte buffer + (nseq+1) offsets; names likewise
void* orc_index_build(const uint8_t* cat, const uint64_t* offs, const char* names_cat, const uint64_t* name_offs,
                      size_t nseq, int w, int k, int b, int flag, int n_threads) {
  std::vector<FastaRecord> recs(nseq);
  for (size_t i = 0; i < nseq; ++i) {
    recs[i].seq.assign(cat + offs[i], cat + offs[i + 1]);
    if (names_cat) recs[i].name.assign(names_cat + name_offs[i], names_cat + name_offs[i + 1]);
  }
  return build_index_from_records(recs, w, k, b, flag, n_threads);
}
void* orc_index_build_fasta(const char* path, int w, int k, int b, int flag, int n_threads) {
  std::string err;
  return build_index_from_fasta(path, w, k, b, flag, n_threads, &err);
}
void orc_index_free(void* h) { delete (Index*)h; }
int orc_index_save_mmi(void* h, const char* path) { std::string e; return ((Index*)h)->save_to_mmi(path, &e) ? 0 : -1; }
void* orc_index_load_mmi(const char* path) { std::string e; return Index::load_from_mmi(path, &e); }
int orc_index_save_native(void* h, const char* path) { std::string e; return ((Index*)h)->save_to_file(path, &e) ? 0 : -1; }
void* orc_index_load_native(const char* path) { std::string e; return Index::load_from_file(path, &e); }
void orc_index_stats(void* h, uint64_t* n_keys, double* avg_occ, double* avg_spacing, uint64_t* total_len) {
  ((Index*)h)->stats(n_keys, avg_occ, avg_spacing, total_len);
}
int32_t orc_index_calc_mid_occ(void* h, float frac) { return ((Index*)h)->calc_mid_occ(frac); }
void orc_index_params(void* h, int32_t* wkbf, uint32_t* n_seq) {
  Index* i = (Index*)h; wkbf[0] = i->w; wkbf[1] = i->k; wkbf[2] = i->b; wkbf[3] = i->flag; *n_seq = i->n_seq;
}
// 0 = none, 1 = single (occ[0]), 2 = multi (copied into occ up to cap; *n = full count)
int orc_index_get(void* h, uint64_t minier, uint64_t* occ, size_t cap, size_t* n) {
  u64 single = 0; const u64* multi = nullptr; size_t cnt = 0;
  int r = ((Index*)h)->get(minier, &single, &multi, &cnt);
  if (r == 1) { if (cap) occ[0] = single; *n = 1; }
  else if (r == 2) { for (size_t i = 0; i < cnt && i < cap; ++i) occ[i] = multi[i]; *n = cnt; }
  else *n = 0;
  return r;
}

size_t orc_filter_query_minimizers(Minimizer* mv, size_t n, int32_t q_occ_max, float q_occ_frac) {
  std::vector<Minimizer> v(mv, mv + n);
  filter_query_minimizers(v, q_occ_max, q_occ_frac);
  if (!v.empty()) memcpy(mv, v.data(), v.size() * sizeof(Minimizer));
  return v.size();
}

int orc_build_anchors_filtered(void* h, const Minimizer* mv, size_t n, int32_t qlen, int32_t mid_occ, Anchor** out, size_t* n_out) {
  std::vector<Minimizer> v(mv, mv + n);
  std::vector<Anchor> a = build_anchors_filtered(*(Index*)h, v, qlen, mid_occ);
  *n_out = a.size();
  *out = (Anchor*)malloc(std::max<size_t>(1, a.size()) * sizeof(Anchor));
  if (!a.empty()) memcpy(*out, a.data(), a.size() * sizeof(Anchor));
  return 0;
}

// chain_dp_all with the forward-DP trace.  f/v/pprev (n each, caller-allocated, may be NULL);
// chains flattened: *chain_offs (n_chains+1), *chain_idx, *scores (malloc'd).
int orc_chain_dp_all(const Anchor* a, size_t n, const ChainParams* p, int32_t* f, int32_t* v, int64_t* pprev, uint64_t* cells,
                     size_t* n_chains, uint64_t** chain_offs, uint64_t** chain_idx, int32_t** scores) {
  std::vector<Anchor> av(a, a + n);
  Chains chains; std::vector<i32> sc; DpTrace tr;
  chain_dp_all(av, *p, chains, sc, &tr);
  if (f && !tr.f.empty()) memcpy(f, tr.f.data(), n * 4);
  if (v && !tr.v.empty()) memcpy(v, tr.v.data(), n * 4);
  if (pprev && !tr.pprev.empty()) memcpy(pprev, tr.pprev.data(), n * 8);
  if (cells) *cells = tr.cells;
  *n_chains = chains.size();
  size_t tot = 0;
  for (auto& c : chains) tot += c.size();
  *chain_offs = (uint64_t*)malloc((chains.size() + 1) * 8);
  *chain_idx = (uint64_t*)malloc(std::max<size_t>(1, tot) * 8);
  *scores = (int32_t*)malloc(std::max<size_t>(1, chains.size()) * 4);
  size_t o = 0;
  for (size_t i = 0; i < chains.size(); ++i) {
    (*chain_offs)[i] = o;
    for (size_t x : chains[i]) (*chain_idx)[o++] = x;
    (*scores)[i] = sc[i];
  }
  (*chain_offs)[chains.size()] = o;
  return 0;
}

void orc_default_chain_params(int32_t k, ChainParams* p) { *p = default_chain_params(k); }

// penalty LUT of comput_sc for chn_pen_skip == 0 (lchain.rs:28-32), dd = 0..n-1; used to cross-check the
// table the product library computes for itself.
void orc_chain_pen_lut(float chn_pen_gap, int32_t n, int32_t* lut) {
  for (int32_t dd = 0; dd < n; ++dd) {
    float lin_pen = chn_pen_gap * (float)dd + 0.0f * 0.0f;
    float log_pen = dd >= 1 ? (dd + 1 <= 1 ? 0.0f : logf((float)(dd + 1)) / 0.6931472f) : 0.0f;
    float s = lin_pen + 0.5f * log_pen;
    lut[dd] = (i32)s;
  }
}

struct orc_align_opts {
  int32_t w, k;
  float frac_top_repetitive;
  int32_t max_gap;
  int32_t bw, bw_long;  // <0: not given
  int32_t min_cnt, min_chain_score;
  float mask_level, pri_ratio;
  int32_t best_n;
};
struct orc_align_stats {
  uint64_t n_reads, n_bases, n_minimizers, n_minimizers_kept, n_anchors, cells, n_rescued, n_lines, n_panic;
  double seconds;
};

// Maps reads [0,nreads) with n_threads host threads (1 = what `mm2rs align` does, main.rs:193-219,
// looped over reads).  PAF lines come back as one '\n'-joined malloc'd buffer in read order.
int orc_align_batch(void* h, const uint8_t* cat, const uint64_t* offs, const char* names_cat, const uint64_t* name_offs,
                    size_t nreads, const orc_align_opts* o, int n_threads, char** paf_out, size_t* paf_len,
                    orc_align_stats* stats) {
  Index* idx = (Index*)h;
  AlignOpts ao;
  ao.w = o->w; ao.k = o->k; ao.frac_top_repetitive = o->frac_top_repetitive; ao.max_gap = o->max_gap;
  if (o->bw >= 0) {
    ao.has_r = true;
    ao.r = std::to_string(o->bw);
    if (o->bw_long >= 0) ao.r += "," + std::to_string(o->bw_long);
  }
  ao.min_cnt = o->min_cnt; ao.min_chain_score = o->min_chain_score;
  ao.mask_level = o->mask_level; ao.pri_ratio = o->pri_ratio; ao.best_n = (size_t)o->best_n;
  i32 mid_occ = idx->calc_mid_occ(ao.frac_top_repetitive);  // main.rs:196-197
  if (mid_occ < 10) mid_occ = 10;
  std::vector<std::vector<std::string>> lines(nreads);
  if (n_threads < 1) n_threads = 1;
  std::vector<AlignStats> tst((size_t)n_threads);
  std::vector<uint64_t> t_resc((size_t)n_threads, 0), t_panic((size_t)n_threads, 0);
  std::atomic<size_t> next(0);
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&](int tid) {
    for (;;) {
      size_t s = next.fetch_add(8);
      if (s >= nreads) break;
      size_t e = std::min(nreads, s + 8);
      for (size_t r = s; r < e; ++r) {
        std::string qname = names_cat ? std::string(names_cat + name_offs[r], names_cat + name_offs[r + 1]) : std::string("*");
        AlignStats one;
        lines[r] = align_read(*idx, mid_occ, ao, qname, cat + offs[r], (size_t)(offs[r + 1] - offs[r]), &one);
        AlignStats& a = tst[(size_t)tid];
        a.n_minimizers += one.n_minimizers; a.n_minimizers_kept += one.n_minimizers_kept;
        a.n_anchors += one.n_anchors; a.cells += one.cells;
        if (one.rescued) t_resc[(size_t)tid] += 1;
        if (one.panic) t_panic[(size_t)tid] += 1;
      }
    }
  };
  if (n_threads == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  std::string all;
  size_t n_lines = 0;
  for (auto& v : lines) for (auto& l : v) { all += l; all += '\n'; n_lines += 1; }
  if (paf_out) {
    *paf_out = (char*)malloc(all.size() + 1);
    memcpy(*paf_out, all.data(), all.size());
    (*paf_out)[all.size()] = 0;
    *paf_len = all.size();
  }
  if (stats) {
    memset(stats, 0, sizeof *stats);
    stats->n_reads = nreads; stats->n_bases = offs[nreads] - offs[0];
    for (int t = 0; t < n_threads; ++t) {
      stats->n_minimizers += tst[(size_t)t].n_minimizers; stats->n_minimizers_kept += tst[(size_t)t].n_minimizers_kept;
      stats->n_anchors += tst[(size_t)t].n_anchors; stats->cells += tst[(size_t)t].cells;
      stats->n_rescued += t_resc[(size_t)t]; stats->n_panic += t_panic[(size_t)t];
    }
    stats->n_lines = n_lines;
    stats->seconds = std::chrono::duration<double>(t1 - t0).count();
  }
  return 0;
}

}  // extern "C"
