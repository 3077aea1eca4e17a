// mm2_oracle.hpp — CPU restatement of the mm2rs mapping hot path.
//
// TEST INFRASTRUCTURE ONLY.  This is the parity checker for the B200 path: it is
// linked/loaded only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs.  Nothing under minimap2_rs_b200/ may include or call it.
//
// PARITY UNPINNED: the reference (xuzhougeng/minimap2_rs) ships no tests, no golden
// vectors and no fixtures and cannot be compiled here (no rustc/cargo,
// un-vendored noodles-fasta git dependency).  This restatement follows the Rust text
// line by line; every function cites the file:line it follows.  The only known-answer
// material in the reference (README.md:24-27) is for input files that are not in the
// repository.
//
// Semantics copied on purpose: wrapping u64/i32 arithmetic of a
// Rust release build, `as i32` truncation, `i32 as u64` sign extension, saturating
// float->int casts, f32 arithmetic without FMA contraction (compile with
// -ffp-contract=off), last-maximum `max_by_key`, stable `sort_by`.
#pragma once
#include <cstdint>
#include <cstddef>
#include <string>
#include <vector>

namespace orc {

typedef uint64_t u64;
typedef uint32_t u32;
typedef int32_t i32;
typedef int64_t i64;

// src/sketch.rs:15-19
struct Minimizer {
  u64 key_span;        // hash<<8 | span
  u64 rid_pos_strand;  // rid<<32 | last_pos<<1 | strand
};
// src/seeds.rs:4-5
struct Anchor {
  u64 x, y;
};
// src/lchain.rs:36-52
struct ChainParams {
  i32 max_dist_x, max_dist_y, bw, max_chain_iter, min_chain_score, min_cnt;
  float chn_pen_gap, chn_pen_skip;
  i32 max_chain_skip, max_drop, bw_long, rmq_rescue_size;
  float rmq_rescue_ratio;
};

uint8_t nt4(uint8_t b);                                    // src/nt4.rs:2-10
u64 hash64(u64 key, u64 mask);                             // src/sketch.rs:4-13
void sketch_sequence(const uint8_t* seq, size_t len, size_t w, size_t k, u32 rid, bool is_hpc,
                     std::vector<Minimizer>& out);         // src/sketch.rs:29-100

// ---- index (src/index.rs) -------------------------------------------------------------
struct IndexSeq {
  bool has_name;
  std::string name;
  u64 offset;
  u32 len;
  bool is_alt;
};

// Open-addressing u64->u64 map standing in for std::collections::HashMap (index.rs:31).
// Iteration order of the Rust map is randomly seeded; every consumer here
// either is order-independent or sorts by key first.
struct HashTab {
  std::vector<u64> keys, vals;
  size_t n = 0, cap_mask = 0;
  void reserve(size_t n_items);
  void insert(u64 k, u64 v);
  bool get(u64 k, u64* v) const;
  template <class F> void for_each(F f) const {
    for (size_t i = 0; i < keys.size(); ++i)
      if (keys[i] != ~0ULL) f(keys[i], vals[i]);
  }
};

struct Bucket {
  std::vector<Minimizer> a;
  std::vector<u64> p;
  bool has_h = false;
  HashTab h;
};

struct Index {
  i32 w, k, b, flag;
  u32 n_seq;
  std::vector<IndexSeq> seq;
  std::vector<u32> S;
  std::vector<Bucket> B;
  Index(i32 w, i32 k, i32 b, i32 flag);                    // index.rs:47-51
  void add_minimizers(const std::vector<Minimizer>& v);    // index.rs:69-72
  void post_process(int n_threads);                        // index.rs:74-109
  void stats(u64* n_keys, double* avg_occ, double* avg_spacing, u64* total_len) const;  // :111-122
  i32 calc_mid_occ(float frac) const;                      // index.rs:124-141
  // index.rs:143-154.  returns 0 = None, 1 = Single (val in *single), 2 = Multi (ptr,n)
  int get(u64 minier, u64* single, const u64** multi, size_t* n) const;
  std::vector<uint8_t> get_ref_subseq(size_t rid, i32 st, i32 en) const;  // index.rs:53-67
  bool save_to_mmi(const std::string& path, std::string* err) const;      // index.rs:233-307
  static Index* load_from_mmi(const std::string& path, std::string* err); // index.rs:361-424
  bool save_to_file(const std::string& path, std::string* err) const;     // index.rs:156-230
  static Index* load_from_file(const std::string& path, std::string* err);// index.rs:309-358
};

struct FastaRecord {
  std::string name;
  std::vector<uint8_t> seq;
};
// Stand-in for noodles_fasta::io::Reader::records() (index.rs:431-437, main.rs:94-98).
bool read_fasta(const std::string& path, std::vector<FastaRecord>& out, bool first_only,
                std::string* err);
// index.rs:427-475 (sketch parallel over sequences, pack serial, post_process parallel over buckets)
Index* build_index_from_records(const std::vector<FastaRecord>& recs, i32 w, i32 k, i32 b, i32 flag,
                                int n_threads);
Index* build_index_from_fasta(const std::string& path, i32 w, i32 k, i32 b, i32 flag, int n_threads,
                              std::string* err);

// ---- seeds (src/seeds.rs) ---------------------------------------------------------------
std::vector<Minimizer> collect_query_minimizers(const uint8_t* seq, size_t len, size_t w, size_t k);  // :7-11
void filter_query_minimizers(std::vector<Minimizer>& mv, i32 q_occ_max, float q_occ_frac);            // :13-36
std::vector<Anchor> build_anchors_filtered(const Index& idx, const std::vector<Minimizer>& mv, i32 qlen,
                                           i32 mid_occ);                                              // :42-60

// ---- lchain (src/lchain.rs) -------------------------------------------------------------
struct DpTrace {  // optional capture of the forward DP for stage-level parity tests
  std::vector<i32> f, v;
  std::vector<i64> pprev;
  u64 cells = 0;  // inner-loop iterations of lchain.rs:80
};
typedef std::vector<std::vector<size_t>> Chains;
void chain_dp_all(const std::vector<Anchor>& a, const ChainParams& p, Chains& chains, std::vector<i32>& scores,
                  DpTrace* trace = nullptr);                                                          // :59-176
std::vector<size_t> chain_dp(const std::vector<Anchor>& a, const ChainParams& p);                     // :54-57
void sort_chains_stable(const std::vector<Anchor>& a, Chains& chains, std::vector<i32>& scores);      // :202-218
std::vector<bool> select_primary_secondary(const std::vector<Anchor>& a, const Chains& chains,
                                           const std::vector<i32>& scores, float mask_level);         // :220-235
void select_and_filter_chains(const std::vector<Anchor>& a, const Chains& chains, const std::vector<i32>& scores,
                              float mask_level, float pri_ratio, size_t best_n, Chains& out_chains,
                              std::vector<i32>& out_scores, std::vector<bool>& out_is_primary, i32* s1,
                              i32* s2);                                                               // :237-260
Chains merge_adjacent_chains(const std::vector<Anchor>& a, const Chains& chains);                     // :262-286
Chains merge_adjacent_chains_with_gap(const std::vector<Anchor>& a, const Chains& chains, i32 max_gap_q,
                                      i32 max_gap_t);                                                 // :288-314
i32 chain_query_coverage(const std::vector<Anchor>& a, const std::vector<size_t>& chain);             // :316-319
void rescue_long_join(const std::vector<Anchor>& a, const Chains& chains, const std::vector<i32>& scores,
                      const ChainParams& p, i32 qlen, Chains& out_chains, std::vector<i32>& out_scores,
                      u64* cells = nullptr, bool* reran = nullptr);                                   // :321-330

// ---- paf (src/paf.rs) ---------------------------------------------------------------------
struct PafRecord {
  std::string qname;
  u32 qlen, qstart, qend;
  char strand;
  std::string tname;
  u32 tlen, tstart, tend, nm, blen;
  uint8_t mapq;
  char tp;
  u32 cm, s1, s2;
  float dv;
  u32 rl;
};
// returns false for None.  If the reference would panic (idx.seq[rid0] out of bounds, F5)
// *panic is set and false is returned.
bool paf_from_chain_with_primary(const Index& idx, const std::vector<Anchor>& a, const std::vector<size_t>& chain,
                                 const std::string& qname, const uint8_t* qseq, size_t qlen, bool is_primary,
                                 PafRecord& rec, bool* panic);                                        // :130-222
std::string write_paf(const PafRecord& rec);                                                          // :224-236
std::vector<std::string> write_paf_many_with_scores(const Index& idx, const std::vector<Anchor>& a,
                                                    const Chains& chains, i32 top_s1, i32 top_s2,
                                                    const std::string& qname, const uint8_t* qseq, size_t qlen,
                                                    bool* panic);                                     // :238-248

// ---- main.rs orchestration -----------------------------------------------------------------
ChainParams default_chain_params(i32 k);                                                              // main.rs:105-123
void apply_preset(const std::string& preset, i32* w, i32* k);                                         // main.rs:125-133
struct AlignOpts {  // the `align` flag surface, main.rs:55-89
  i32 w = 10, k = 15;
  float frac_top_repetitive = 2e-4f;
  i32 max_gap = 5000;
  bool has_r = false;
  std::string r;
  i32 min_cnt = 3, min_chain_score = 40;
  float mask_level = 0.5f, pri_ratio = 0.8f;
  size_t best_n = 5;
};
struct AlignStats {
  u64 n_minimizers = 0, n_minimizers_kept = 0, n_anchors = 0, cells = 0;
  bool rescued = false, panic = false;
};
// main.rs:189-219 for one read (mid_occ is hoisted: it depends on the index only).
std::vector<std::string> align_read(const Index& idx, i32 mid_occ, const AlignOpts& o, const std::string& qname,
                                    const uint8_t* q, size_t qlen, AlignStats* st = nullptr);
ChainParams align_chain_params(const AlignOpts& o);                                                   // main.rs:199-208

}  // namespace orc
