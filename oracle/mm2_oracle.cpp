// mm2_oracle.cpp — CPU restatement of the mm2rs mapping hot path (see mm2_oracle.hpp).
//
// TEST INFRASTRUCTURE ONLY (parity checker + CPU baseline).  PARITY UNPINNED: the reference
// ships no tests/golden vectors and cannot be compiled here (no rustc); this file follows
// the Rust text function by function, citing file:line of /root/reference for each.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fPIC (no -ffast-math).
#include "mm2_oracle.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <thread>
#include <atomic>
#include <climits>

namespace orc {

static const u64 U64MAX = ~0ULL;

// ---- nt4.rs:2-10 --------------------------------------------------------------------------
uint8_t nt4(uint8_t b) {
  switch (b) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
  }
}

// ---- sketch.rs:4-13 -----------------------------------------------------------------------
u64 hash64(u64 key, u64 mask) {
  key = (~key + (key << 21)) & mask;
  key ^= key >> 24;
  key = (key + (key << 3) + (key << 8)) & mask;
  key ^= key >> 14;
  key = (key + (key << 2) + (key << 4)) & mask;
  key ^= key >> 28;
  key = (key + (key << 31)) & mask;
  return key;
}

// ---- sketch.rs:21-27 (TinyQueue) ----------------------------------------------------------
namespace {
struct TinyQueue {
  size_t front = 0, count = 0;
  i32 a[32] = {0};
  void clear() { front = 0; count = 0; }
  void push(i32 x) { a[(count + front) & 0x1f] = x; count += 1; }
  i32 shift() {
    if (count == 0) return -1;
    i32 x = a[front];
    front = (front + 1) & 0x1f;
    count -= 1;
    return x;
  }
};
struct Info { u64 k, v; };
}  // namespace

// ---- sketch.rs:29-100 ---------------------------------------------------------------------
void sketch_sequence(const uint8_t* seq, size_t len, size_t w, size_t k, u32 rid, bool is_hpc,
                     std::vector<Minimizer>& out) {
  // asserts of sketch.rs:30-32 are checked by callers (a Rust panic has no C++ analogue here)
  if (len == 0 || !(w > 0 && w < 256) || !(k > 0 && k <= 28)) return;
  const u64 shift1 = 2 * ((u64)k - 1);
  const u64 mask = (1ULL << (2 * k)) - 1;
  u64 kmer[2] = {0, 0};
  i32 l = 0;
  size_t buf_pos = 0, min_pos = 0;
  i32 kmer_span = 0;
  std::vector<Info> buf(w, Info{U64MAX, U64MAX});
  Info min = {U64MAX, U64MAX};
  TinyQueue tq;
  const i32 wi = (i32)w, ki = (i32)k;
  for (size_t i = 0; i < len; ++i) {
    i32 c = (i32)nt4(seq[i]);
    Info info = {U64MAX, U64MAX};
    if (c < 4) {
      if (is_hpc) {  // sketch.rs:51-61: every base pushes its remaining run length (F12)
        size_t skip_len = 1;
        if (i + 1 < len && (i32)nt4(seq[i + 1]) == c) {
          size_t t = i + 2;
          while (t < len && (i32)nt4(seq[t]) == c) t += 1;
          skip_len = t - i;
        }
        tq.push((i32)skip_len);
        kmer_span += (i32)skip_len;
        if ((i32)tq.count > ki) kmer_span -= tq.shift();
      } else {
        kmer_span = (l + 1 < ki) ? l + 1 : ki;
      }
      kmer[0] = ((kmer[0] << 2) | (u64)c) & mask;
      kmer[1] = (kmer[1] >> 2) | ((u64)(3 ^ c) << shift1);
      if (kmer[0] != kmer[1]) {
        int z = kmer[0] < kmer[1] ? 0 : 1;
        l += 1;
        if (l >= ki && kmer_span < 256) {
          info.k = (hash64(kmer[z], mask) << 8) | (u64)kmer_span;
          info.v = ((u64)rid << 32) | ((u64)i << 1) | (u64)z;
        }
      }
    } else {
      l = 0; tq.clear(); kmer_span = 0;
    }
    buf[buf_pos] = info;
    if (l == wi + ki - 1 && min.k != U64MAX) {
      for (size_t j = buf_pos + 1; j < w; ++j)
        if (min.k == buf[j].k && buf[j].v != min.v) out.push_back(Minimizer{buf[j].k, buf[j].v});
      for (size_t j = 0; j < buf_pos; ++j)
        if (min.k == buf[j].k && buf[j].v != min.v) out.push_back(Minimizer{buf[j].k, buf[j].v});
    }
    if (info.k <= min.k) {
      if (l >= wi + ki && min.k != U64MAX) out.push_back(Minimizer{min.k, min.v});
      min = info; min_pos = buf_pos;
    } else if (buf_pos == min_pos) {
      if (l >= wi + ki - 1 && min.k != U64MAX) out.push_back(Minimizer{min.k, min.v});
      min.k = U64MAX;
      for (size_t j = buf_pos + 1; j < w; ++j) if (min.k >= buf[j].k) { min = buf[j]; min_pos = j; }
      for (size_t j = 0; j <= buf_pos; ++j) if (min.k >= buf[j].k) { min = buf[j]; min_pos = j; }
      if (l >= wi + ki - 1 && min.k != U64MAX) {
        for (size_t j = buf_pos + 1; j < w; ++j)
          if (min.k == buf[j].k && min.v != buf[j].v) out.push_back(Minimizer{buf[j].k, buf[j].v});
        for (size_t j = 0; j <= buf_pos; ++j)
          if (min.k == buf[j].k && min.v != buf[j].v) out.push_back(Minimizer{buf[j].k, buf[j].v});
      }
    }
    buf_pos += 1; if (buf_pos == w) buf_pos = 0;
  }
  if (min.k != U64MAX) out.push_back(Minimizer{min.k, min.v});
}

// ---- HashTab (stand-in for std HashMap, index.rs:31) --------------------------------------
static inline u64 mix64(u64 x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
void HashTab::reserve(size_t n_items) {
  size_t cap = 16;
  while (cap < n_items * 2) cap <<= 1;
  keys.assign(cap, U64MAX);
  vals.assign(cap, 0);
  cap_mask = cap - 1;
  n = 0;
}
void HashTab::insert(u64 k, u64 v) {
  if (keys.empty() || (n + 1) * 2 > keys.size()) {  // grow
    std::vector<u64> ok; std::vector<u64> ov;
    ok.swap(keys); ov.swap(vals);
    size_t want = ok.empty() ? 8 : ok.size();
    reserve(want);  // reserve doubles relative to n_items
    for (size_t i = 0; i < ok.size(); ++i) if (ok[i] != U64MAX) insert(ok[i], ov[i]);
  }
  size_t i = mix64(k) & cap_mask;
  while (keys[i] != U64MAX && keys[i] != k) i = (i + 1) & cap_mask;
  if (keys[i] == U64MAX) { keys[i] = k; n += 1; }
  vals[i] = v;
}
bool HashTab::get(u64 k, u64* v) const {
  if (keys.empty()) return false;
  size_t i = mix64(k) & cap_mask;
  while (keys[i] != U64MAX) {
    if (keys[i] == k) { *v = vals[i]; return true; }
    i = (i + 1) & cap_mask;
  }
  return false;
}

// ---- index.rs:10-26 -----------------------------------------------------------------------
static inline size_t kroundup64(size_t x) {
  x -= 1; x |= x >> 1; x |= x >> 2; x |= x >> 4; x |= x >> 8; x |= x >> 16; x |= x >> 32;
  return x + 1;
}
static inline void mm_seq4_set(std::vector<u32>& S, u64 o, uint8_t c) {
  size_t i = (size_t)(o >> 3);
  unsigned shift = (unsigned)((o & 7) << 2);
  u32 v = S[i];
  S[i] = (v & ~(0xFu << shift)) | (((u32)c & 0xF) << shift);
}
static inline uint8_t mm_seq4_get(const std::vector<u32>& S, u64 o) {
  return (uint8_t)((S[(size_t)(o >> 3)] >> ((o & 7) << 2)) & 0xF);
}

// ---- index.rs:47-51 -----------------------------------------------------------------------
Index::Index(i32 w_, i32 k_, i32 b_, i32 flag_) : w(w_), k(k_), b(b_), flag(flag_), n_seq(0) {
  B.resize((size_t)1 << b_);
}

// ---- index.rs:53-67 -----------------------------------------------------------------------
std::vector<uint8_t> Index::get_ref_subseq(size_t rid, i32 st, i32 en) const {
  std::vector<uint8_t> out;
  if (rid >= seq.size()) return out;
  const IndexSeq& s = seq[rid];
  u64 st0 = (u64)std::max(st, 0);
  u64 en0 = (u64)std::max(std::min(en, (i32)s.len), 0);
  if (st0 >= en0) return out;
  st0 += s.offset; en0 += s.offset;
  for (u64 o = st0; o < en0; ++o) {
    uint8_t c = mm_seq4_get(S, o);
    out.push_back(c == 0 ? 'A' : c == 1 ? 'C' : c == 2 ? 'G' : c == 3 ? 'T' : 'N');
  }
  return out;
}

// ---- index.rs:69-72 -----------------------------------------------------------------------
void Index::add_minimizers(const std::vector<Minimizer>& v) {
  u64 mask = (1ULL << b) - 1;
  for (const Minimizer& m : v) B[(size_t)((m.key_span >> 8) & mask)].a.push_back(m);
}

// ---- index.rs:74-109 ----------------------------------------------------------------------
static void post_process_bucket(Bucket& bk, i32 b_bits) {
  if (bk.a.empty()) return;
  std::stable_sort(bk.a.begin(), bk.a.end(),
                   [](const Minimizer& x, const Minimizer& y) { return (x.key_span >> 8) < (y.key_span >> 8); });
  i32 n = 1, n_keys = 0; size_t total_p = 0;
  for (size_t j = 1; j <= bk.a.size(); ++j) {
    if (j == bk.a.size() || (bk.a[j].key_span >> 8) != (bk.a[j - 1].key_span >> 8)) {
      n_keys += 1; if (n > 1) total_p += (size_t)n; n = 1;
    } else n += 1;
  }
  bk.p.assign(total_p, 0);
  bk.h.reserve((size_t)n_keys);
  n = 1; size_t start_a = 0, start_p = 0;
  for (size_t j = 1; j <= bk.a.size(); ++j) {
    if (j == bk.a.size() || (bk.a[j].key_span >> 8) != (bk.a[j - 1].key_span >> 8)) {
      const Minimizer& p = bk.a[j - 1];
      u64 key_top = ((p.key_span >> 8) >> b_bits) << 1;
      if (n == 1) {
        bk.h.insert(key_top | 1, p.rid_pos_strand);
      } else {
        for (i32 q = 0; q < n; ++q) bk.p[start_p + (size_t)q] = bk.a[start_a + (size_t)q].rid_pos_strand;
        std::sort(bk.p.begin() + start_p, bk.p.begin() + start_p + n);
        bk.h.insert(key_top, ((u64)start_p << 32) | (u64)n);
        start_p += (size_t)n;
      }
      start_a = j; n = 1;
    } else n += 1;
  }
  bk.has_h = true;
  std::vector<Minimizer>().swap(bk.a);
}

template <class F>
static void parallel_for(size_t n, int n_threads, F f) {
  if (n_threads <= 1 || n <= 1) { for (size_t i = 0; i < n; ++i) f(i); return; }
  std::atomic<size_t> next(0);
  std::vector<std::thread> th;
  const size_t grain = std::max<size_t>(1, n / ((size_t)n_threads * 16));
  for (int t = 0; t < n_threads; ++t)
    th.emplace_back([&]() {
      for (;;) {
        size_t s = next.fetch_add(grain);
        if (s >= n) break;
        size_t e = std::min(n, s + grain);
        for (size_t i = s; i < e; ++i) f(i);
      }
    });
  for (auto& t : th) t.join();
}

void Index::post_process(int n_threads) {
  const i32 b_bits = b;
  parallel_for(B.size(), n_threads, [&](size_t i) { post_process_bucket(B[i], b_bits); });
}

// ---- index.rs:111-122 ---------------------------------------------------------------------
void Index::stats(u64* n_keys_o, double* avg_occ, double* avg_spacing, u64* total_len_o) const {
  u64 n_keys = 0, sum_occ = 0;
  for (const Bucket& bk : B)
    if (bk.has_h)
      bk.h.for_each([&](u64 k, u64 v) {
        if ((k & 1) == 1) { n_keys += 1; sum_occ += 1; } else { n_keys += 1; sum_occ += (v & 0xffffffffULL); }
      });
  u64 total_len = 0;
  for (const IndexSeq& s : seq) total_len += (u64)s.len;
  *n_keys_o = n_keys;
  *avg_occ = n_keys > 0 ? (double)sum_occ / (double)n_keys : 0.0;
  *avg_spacing = sum_occ > 0 ? (double)total_len / (double)sum_occ : 0.0;
  *total_len_o = total_len;
}

// ---- index.rs:124-141 ---------------------------------------------------------------------
i32 Index::calc_mid_occ(float frac) const {
  std::vector<u32> counts;
  for (const Bucket& bk : B)
    if (bk.has_h)
      bk.h.for_each([&](u64 k, u64 v) { counts.push_back((k & 1) == 1 ? 1u : (u32)(v & 0xffffffffULL)); });
  if (counts.empty()) return INT32_MAX;
  std::sort(counts.begin(), counts.end());
  size_t n = counts.size();
  double x = (1.0 - (double)frac) * (double)n;
  size_t idx = x <= 0.0 ? 0 : (size_t)x;  // Rust `as usize` saturates at 0 for negatives
  idx = std::min(idx, n - 1);
  return (i32)counts[idx] + 1;
}

// ---- index.rs:143-154 ---------------------------------------------------------------------
int Index::get(u64 minier, u64* single, const u64** multi, size_t* n) const {
  u64 mask = (1ULL << b) - 1;
  const Bucket& bk = B[(size_t)(minier & mask)];
  if (!bk.has_h) return 0;
  u64 key = (minier >> b) << 1;
  u64 val;
  if (bk.h.get(key | 1, &val)) { *single = val; return 1; }
  if (bk.h.get(key, &val)) {
    size_t off = (size_t)(val >> 32), cnt = (size_t)(val & 0xffffffffULL);
    *multi = bk.p.data() + off; *n = cnt;
    return 2;
  }
  return 0;
}

// ---- little-endian IO helpers -------------------------------------------------------------
namespace {
struct Writer {
  FILE* f; bool ok = true;
  explicit Writer(FILE* f_) : f(f_) {}
  void bytes(const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) ok = false; }
  void u8(uint8_t v) { bytes(&v, 1); }
  void u32_(u32 v) { bytes(&v, 4); }
  void i32_(i32 v) { bytes(&v, 4); }
  void u64_(u64 v) { bytes(&v, 8); }
};
struct Reader {
  FILE* f; bool ok = true;
  explicit Reader(FILE* f_) : f(f_) {}
  void bytes(void* p, size_t n) { if (n && fread(p, 1, n, f) != n) ok = false; }
  uint8_t u8() { uint8_t v = 0; bytes(&v, 1); return v; }
  u32 u32_() { u32 v = 0; bytes(&v, 4); return v; }
  i32 i32_() { i32 v = 0; bytes(&v, 4); return v; }
  u64 u64_() { u64 v = 0; bytes(&v, 8); return v; }
};
// canonical (ascending key) listing of a bucket's (key,val) pairs: SURVEY.md F4
std::vector<std::pair<u64, u64>> sorted_entries(const Bucket& bk) {
  std::vector<std::pair<u64, u64>> e;
  if (bk.has_h) bk.h.for_each([&](u64 k, u64 v) { e.emplace_back(k, v); });
  std::sort(e.begin(), e.end());
  return e;
}
}  // namespace

// ---- index.rs:233-307 ---------------------------------------------------------------------
bool Index::save_to_mmi(const std::string& path, std::string* err) const {
  FILE* fp = fopen(path.c_str(), "wb");
  if (!fp) { if (err) *err = "cannot create " + path; return false; }
  static char iobuf[1 << 20];
  setvbuf(fp, iobuf, _IOFBF, sizeof iobuf);
  Writer wr(fp);
  wr.bytes("MMI\2", 4);
  wr.u32_((u32)w); wr.u32_((u32)k); wr.u32_((u32)b); wr.u32_((u32)seq.size()); wr.u32_((u32)flag);
  u64 sum_len = 0;
  for (const IndexSeq& s : seq) {
    if (s.has_name) {
      uint8_t l = (uint8_t)std::min<size_t>(s.name.size(), 255);
      wr.u8(l); wr.bytes(s.name.data(), l);
    } else wr.u8(0);
    wr.u32_(s.len);
    sum_len += (u64)s.len;
  }
  size_t nb = (size_t)1 << b;
  for (size_t i = 0; i < nb; ++i) {
    const Bucket& bk = B[i];
    wr.u32_((u32)bk.p.size());
    wr.bytes(bk.p.data(), bk.p.size() * 8);
    u32 size = bk.has_h ? (u32)bk.h.n : 0u;
    wr.u32_(size);
    if (bk.has_h) for (auto& kv : sorted_entries(bk)) { wr.u64_(kv.first); wr.u64_(kv.second); }
  }
  size_t words = (size_t)((sum_len + 7) / 8);
  wr.bytes(S.data(), words * 4);
  bool ok = wr.ok;
  if (fclose(fp) != 0) ok = false;
  if (!ok && err) *err = "write error on " + path;
  return ok;
}

// ---- index.rs:361-424 ---------------------------------------------------------------------
Index* Index::load_from_mmi(const std::string& path, std::string* err) {
  FILE* fp = fopen(path.c_str(), "rb");
  if (!fp) { if (err) *err = "cannot open " + path; return nullptr; }
  Reader rd(fp);
  char magic[4]; rd.bytes(magic, 4);
  if (!rd.ok || memcmp(magic, "MMI\2", 4) != 0) { fclose(fp); if (err) *err = "invalid MMI magic"; return nullptr; }
  i32 w = (i32)rd.u32_(), k = (i32)rd.u32_(), b = (i32)rd.u32_();
  u32 n_seq = rd.u32_();
  i32 flag = (i32)rd.u32_();
  if (!rd.ok || b < 0 || b > 30) { fclose(fp); if (err) *err = "truncated MMI header"; return nullptr; }
  Index* idx = new Index(w, k, b, flag);
  idx->n_seq = n_seq;
  u64 sum_len = 0;
  for (u32 i = 0; i < n_seq && rd.ok; ++i) {
    IndexSeq s; size_t nl = rd.u8();
    s.has_name = nl > 0;
    if (nl > 0) { s.name.resize(nl); rd.bytes(&s.name[0], nl); }
    s.len = rd.u32_(); s.offset = sum_len; s.is_alt = false;
    sum_len += (u64)s.len;
    idx->seq.push_back(s);
  }
  for (size_t i = 0; i < idx->B.size() && rd.ok; ++i) {
    Bucket& bk = idx->B[i];
    size_t n = rd.u32_();
    bk.p.resize(n); rd.bytes(bk.p.data(), n * 8);
    size_t size = rd.u32_();
    if (size > 0) {
      bk.has_h = true; bk.h.reserve(size);
      for (size_t j = 0; j < size && rd.ok; ++j) { u64 kk = rd.u64_(), vv = rd.u64_(); bk.h.insert(kk, vv); }
    }
  }
  size_t words = (size_t)((sum_len + 7) / 8);
  idx->S.assign(words, 0);
  rd.bytes(idx->S.data(), words * 4);
  bool ok = rd.ok;
  fclose(fp);
  if (!ok) { delete idx; if (err) *err = "truncated MMI file"; return nullptr; }
  return idx;
}

// ---- index.rs:156-230 ---------------------------------------------------------------------
bool Index::save_to_file(const std::string& path, std::string* err) const {
  FILE* fp = fopen(path.c_str(), "wb");
  if (!fp) { if (err) *err = "cannot create " + path; return false; }
  Writer wr(fp);
  wr.bytes("MM2RSIDX\0", 9);
  wr.u32_(1);
  wr.i32_(w); wr.i32_(k); wr.i32_(b); wr.i32_(flag); wr.u32_(n_seq);
  wr.u32_((u32)seq.size());
  for (const IndexSeq& s : seq) {
    wr.u8(s.has_name ? 1 : 0);
    if (s.has_name) { wr.u32_((u32)s.name.size()); wr.bytes(s.name.data(), s.name.size()); }
    wr.u64_(s.offset); wr.u32_(s.len); wr.u8(s.is_alt ? 1 : 0);
  }
  wr.u64_((u64)S.size());
  wr.bytes(S.data(), S.size() * 4);
  wr.u32_((u32)B.size());
  for (const Bucket& bk : B) {
    wr.u64_((u64)bk.p.size());
    wr.bytes(bk.p.data(), bk.p.size() * 8);
    wr.u8(bk.has_h ? 1 : 0);
    if (bk.has_h) {
      wr.u64_((u64)bk.h.n);
      for (auto& kv : sorted_entries(bk)) { wr.u64_(kv.first); wr.u64_(kv.second); }
    }
  }
  bool ok = wr.ok;
  if (fclose(fp) != 0) ok = false;
  if (!ok && err) *err = "write error on " + path;
  return ok;
}

// ---- index.rs:309-358 ---------------------------------------------------------------------
Index* Index::load_from_file(const std::string& path, std::string* err) {
  FILE* fp = fopen(path.c_str(), "rb");
  if (!fp) { if (err) *err = "cannot open " + path; return nullptr; }
  Reader rd(fp);
  char magic[9]; rd.bytes(magic, 9);
  if (!rd.ok || memcmp(magic, "MM2RSIDX\0", 9) != 0) { fclose(fp); if (err) *err = "invalid index file magic"; return nullptr; }
  (void)rd.u32_();
  i32 w = rd.i32_(), k = rd.i32_(), b = rd.i32_(), flag = rd.i32_();
  u32 n_seq_decl = rd.u32_();
  if (!rd.ok || b < 0 || b > 30) { fclose(fp); if (err) *err = "truncated index header"; return nullptr; }
  Index* idx = new Index(w, k, b, flag);
  idx->n_seq = n_seq_decl;
  size_t n_seq = rd.u32_();
  for (size_t i = 0; i < n_seq && rd.ok; ++i) {
    IndexSeq s;
    s.has_name = rd.u8() != 0;
    if (s.has_name) { size_t l = rd.u32_(); s.name.resize(l); rd.bytes(&s.name[0], l); }
    s.offset = rd.u64_(); s.len = rd.u32_(); s.is_alt = rd.u8() != 0;
    idx->seq.push_back(s);
  }
  size_t s_words = (size_t)rd.u64_();
  if (rd.ok) { idx->S.resize(s_words); rd.bytes(idx->S.data(), s_words * 4); }
  size_t nb = rd.u32_();
  idx->B.clear(); idx->B.resize(nb);
  for (size_t i = 0; i < nb && rd.ok; ++i) {
    Bucket& bk = idx->B[i];
    size_t p_len = (size_t)rd.u64_();
    bk.p.resize(p_len); rd.bytes(bk.p.data(), p_len * 8);
    bk.has_h = rd.u8() != 0;
    if (bk.has_h) {
      size_t h_len = (size_t)rd.u64_();
      bk.h.reserve(h_len);
      for (size_t j = 0; j < h_len && rd.ok; ++j) { u64 kk = rd.u64_(), vv = rd.u64_(); bk.h.insert(kk, vv); }
    }
  }
  bool ok = rd.ok;
  fclose(fp);
  if (!ok) { delete idx; if (err) *err = "truncated index file"; return nullptr; }
  return idx;
}

// ---- FASTA (stand-in for noodles_fasta::io::Reader, index.rs:431-437; parity unpinned) ----
bool read_fasta(const std::string& path, std::vector<FastaRecord>& out, bool first_only, std::string* err) {
  FILE* fp = fopen(path.c_str(), "rb");
  if (!fp) { if (err) *err = "cannot open " + path; return false; }
  std::vector<char> buf(1 << 20);
  bool in_header = false, at_line_start = true, have = false, name_done = false, stop = false;
  FastaRecord cur;
  size_t n;
  while (!stop && (n = fread(buf.data(), 1, buf.size(), fp)) > 0) {
    for (size_t i = 0; i < n; ++i) {
      char c = buf[i];
      if (in_header) {
        if (c == '\n') { in_header = false; at_line_start = true; }
        else if (!name_done) { if (c == ' ' || c == '\t' || c == '\r') name_done = true; else cur.name.push_back(c); }
        continue;
      }
      if (at_line_start && c == '>') {
        if (have) { out.push_back(std::move(cur)); cur = FastaRecord(); if (first_only) { stop = true; break; } }
        have = true; in_header = true; name_done = false; at_line_start = false;
        continue;
      }
      if (c == '\n') { at_line_start = true; continue; }
      at_line_start = false;
      if (c == '\r') continue;
      if (have) cur.seq.push_back((uint8_t)c);
    }
  }
  if (have && !stop) out.push_back(std::move(cur));
  fclose(fp);
  return true;
}

// ---- index.rs:427-475 ---------------------------------------------------------------------
Index* build_index_from_records(const std::vector<FastaRecord>& recs, i32 w, i32 k, i32 b, i32 flag, int n_threads) {
  Index* idx = new Index(w, k, b, flag);
  idx->n_seq = (u32)recs.size();
  const bool is_hpc = (flag & 1) != 0;
  std::vector<std::vector<Minimizer>> minis(recs.size());
  parallel_for(recs.size(), n_threads, [&](size_t rid) {  // rayon par_iter over sequences, index.rs:442-452
    if (!recs[rid].seq.empty())
      sketch_sequence(recs[rid].seq.data(), recs[rid].seq.size(), (size_t)w, (size_t)k, (u32)rid, is_hpc, minis[rid]);
  });
  u64 total_len = 0;
  for (auto& r : recs) total_len += (u64)r.seq.size();
  size_t words = total_len == 0 ? 0 : kroundup64((size_t)((total_len + 7) / 8));
  idx->S.assign(words, 0);
  u64 sum_len = 0;
  for (size_t rid = 0; rid < recs.size(); ++rid) {
    const auto& seq = recs[rid].seq;
    for (size_t j = 0; j < seq.size(); ++j) mm_seq4_set(idx->S, sum_len + (u64)j, nt4(seq[j]));
    IndexSeq s; s.has_name = true; s.name = recs[rid].name; s.offset = sum_len; s.len = (u32)seq.size(); s.is_alt = false;
    idx->seq.push_back(s);
    idx->add_minimizers(minis[rid]);
    std::vector<Minimizer>().swap(minis[rid]);
    sum_len += (u64)seq.size();
  }
  idx->post_process(n_threads);
  return idx;
}

Index* build_index_from_fasta(const std::string& path, i32 w, i32 k, i32 b, i32 flag, int n_threads, std::string* err) {
  std::vector<FastaRecord> recs;
  if (!read_fasta(path, recs, false, err)) return nullptr;
  return build_index_from_records(recs, w, k, b, flag, n_threads);
}

// ---- seeds.rs:7-11 ------------------------------------------------------------------------
std::vector<Minimizer> collect_query_minimizers(const uint8_t* seq, size_t len, size_t w, size_t k) {
  std::vector<Minimizer> v;
  sketch_sequence(seq, len, w, k, 0, false, v);
  return v;
}

// ---- seeds.rs:13-36 -----------------------------------------------------------------------
void filter_query_minimizers(std::vector<Minimizer>& mv, i32 q_occ_max, float q_occ_frac) {
  if (mv.empty() || q_occ_frac <= 0.0f || q_occ_max <= 0) return;
  if ((i32)mv.size() <= q_occ_max) return;
  std::vector<std::pair<u64, size_t>> keys;
  keys.reserve(mv.size());
  for (size_t i = 0; i < mv.size(); ++i) keys.emplace_back(mv[i].key_span >> 8, i);
  std::sort(keys.begin(), keys.end());  // tie order is irrelevant: only group sizes/membership are used
  std::vector<char> keep(mv.size(), 1);
  size_t st = 0, n = keys.size();
  float cf = (float)mv.size() * q_occ_frac;
  size_t cutoff = cf <= 0.0f ? 0 : (size_t)cf;
  for (size_t i = 1; i <= n; ++i) {
    if (i == n || keys[i].first != keys[st].first) {
      size_t cnt = i - st;
      if ((i32)cnt > q_occ_max && cnt > cutoff)
        for (size_t j = st; j < i; ++j) keep[keys[j].second] = 0;
      st = i;
    }
  }
  size_t j = 0;
  for (size_t i = 0; i < mv.size(); ++i) if (keep[i]) { mv[j] = mv[i]; j += 1; }
  mv.resize(j);
}

// ---- seeds.rs:62-79 -----------------------------------------------------------------------
static inline void push_anchor(std::vector<Anchor>& out, u64 r, const Minimizer& m, i32 qlen) {
  u64 rid = (r >> 32) & 0xffffffffULL;
  i32 rpos = (i32)(u32)((r >> 1) & 0xffffffffULL);
  i32 rstrand = (i32)(r & 1);
  i32 qpos = (i32)(u32)((m.rid_pos_strand >> 1) & 0xffffffffULL);
  i32 qstrand = (i32)(m.rid_pos_strand & 1);
  i32 qspan = (i32)(m.key_span & 0xff);
  bool forward = rstrand == qstrand;
  u64 rpos64 = (u64)(i64)rpos;  // Rust `i32 as u64` sign-extends
  u64 x = forward ? ((rid << 32) | rpos64) : ((1ULL << 63) | (rid << 32) | rpos64);
  u64 y;
  if (forward) {
    y = ((u64)(i64)qspan << 32) | (u64)(i64)qpos;
  } else {
    i32 qp32 = (i32)((u32)qlen - ((u32)qpos + 1u - (u32)qspan) - 1u);  // wrapping i32 arithmetic
    y = ((u64)(i64)qspan << 32) | (u64)(i64)qp32;
  }
  out.push_back(Anchor{x, y});
}

// ---- seeds.rs:42-60 -----------------------------------------------------------------------
std::vector<Anchor> build_anchors_filtered(const Index& idx, const std::vector<Minimizer>& mv, i32 qlen, i32 mid_occ) {
  std::vector<Anchor> a;
  for (const Minimizer& m : mv) {
    u64 minier = m.key_span >> 8;
    u64 single = 0; const u64* multi = nullptr; size_t n = 0;
    int r = idx.get(minier, &single, &multi, &n);
    if (r == 1) push_anchor(a, single, m, qlen);
    else if (r == 2) {
      if ((i32)n > mid_occ) continue;
      for (size_t i = 0; i < n; ++i) push_anchor(a, multi[i], m, qlen);
    }
  }
  std::stable_sort(a.begin(), a.end(), [](const Anchor& p, const Anchor& q) { return p.x == q.x ? p.y < q.y : p.x < q.x; });
  return a;
}

// ---- lchain.rs:3-15 -----------------------------------------------------------------------
static inline i32 a_qpos(const Anchor& a) { return (i32)(u32)(a.y & 0xffffffffULL); }
static inline i32 a_qspan(const Anchor& a) { return (i32)((a.y >> 32) & 0xff); }
static inline i32 a_rpos(const Anchor& a) { return (i32)(u32)(a.x & 0xffffffffULL); }
static inline bool a_rev(const Anchor& a) { return (a.x >> 63) != 0; }
static inline i32 a_rid(const Anchor& a) { return (i32)((a.x >> 32) & 0x7fffffffULL); }
static inline float mg_log2(i32 x) { return x <= 1 ? 0.0f : logf((float)x) / 0.6931472f /* f32::consts::LN_2 */; }

static inline i32 f32_to_i32(float v) {  // Rust `as i32`: truncate toward zero, saturate, NaN -> 0
  if (v != v) return 0;
  if (v >= 2147483648.0f) return INT32_MAX;
  if (v <= -2147483648.0f) return INT32_MIN;
  return (i32)v;
}

// ---- lchain.rs:17-34 ----------------------------------------------------------------------
static inline bool comput_sc(const Anchor& ai, const Anchor& aj, i32 max_dist_x, i32 max_dist_y, i32 bw,
                             float chn_pen_gap, float chn_pen_skip, i32* sc_out) {
  i32 dq = a_qpos(ai) - a_qpos(aj);
  if (dq <= 0 || dq > max_dist_x) return false;
  i32 dr = a_rpos(ai) - a_rpos(aj);
  if (dr == 0 || dq > max_dist_y) return false;
  i32 dd = dr - dq; if (dd < 0) dd = -dd;
  if (dd > bw) return false;
  i32 dg = std::min(dr, dq);
  i32 q_span = a_qspan(aj);
  i32 sc = std::min(q_span, dg);
  if (dd != 0 || dg > q_span) {
    float lin_pen = chn_pen_gap * (float)dd + chn_pen_skip * (float)dg;
    float log_pen = dd >= 1 ? mg_log2(dd + 1) : 0.0f;
    sc -= f32_to_i32(lin_pen + 0.5f * log_pen);
  }
  *sc_out = sc;
  return true;
}

// ---- lchain.rs:179-200 --------------------------------------------------------------------
static void chain_qrange(const std::vector<Anchor>& a, const std::vector<size_t>& chain, i32* qs_o, i32* qe_o) {
  i32 qs = INT32_MAX, qe = -1;
  for (size_t i : chain) {
    i32 s = a_qpos(a[i]) - (a_qspan(a[i]) - 1), e = a_qpos(a[i]) + 1;
    if (s < qs) qs = s;
    if (e > qe) qe = e;
  }
  *qs_o = std::max(qs, 0); *qe_o = qe;
}
static void chain_trange(const std::vector<Anchor>& a, const std::vector<size_t>& chain, i32* ts_o, i32* te_o) {
  i32 ts = INT32_MAX, te = -1;
  for (size_t i : chain) {
    i32 s = a_rpos(a[i]) - (a_qspan(a[i]) - 1), e = a_rpos(a[i]) + 1;
    if (s < ts) ts = s;
    if (e > te) te = e;
  }
  *ts_o = std::max(ts, 0); *te_o = te;
}

// ---- lchain.rs:202-218 --------------------------------------------------------------------
void sort_chains_stable(const std::vector<Anchor>& a, Chains& chains, std::vector<i32>& scores) {
  std::vector<size_t> idxs(chains.size());
  for (size_t i = 0; i < idxs.size(); ++i) idxs[i] = i;
  std::stable_sort(idxs.begin(), idxs.end(), [&](size_t i, size_t j) {
    i32 si = scores[i], sj = scores[j];
    if (si != sj) return sj < si;
    i32 qi, qj, ti, tj, e;
    chain_qrange(a, chains[i], &qi, &e); chain_qrange(a, chains[j], &qj, &e);
    if (qi != qj) return qi < qj;
    chain_trange(a, chains[i], &ti, &e); chain_trange(a, chains[j], &tj, &e);
    return ti < tj;
  });
  Chains c2; std::vector<i32> s2;
  for (size_t i : idxs) { c2.push_back(chains[i]); s2.push_back(scores[i]); }
  chains.swap(c2); scores.swap(s2);
}

// mg_chain_bk_end-like helper shared by both backtrack passes (lchain.rs:104-121 == :133-150).
// NB: t[i] is set to 2 and then tested ==0 on the same i, so the loop body runs once.
static inline i64 bk_end(i64 i0, i32 zf, const std::vector<i32>& f, const std::vector<i64>& pprev, std::vector<i32>& t,
                         i32 max_drop) {
  i64 i = i0, end_i = -1, max_i = i;
  i32 max_s = 0;
  if (i >= 0 && t[(size_t)i] == 0) {
    for (;;) {
      t[(size_t)i] = 2;
      end_i = pprev[(size_t)i];
      i32 s = end_i < 0 ? zf : zf - f[(size_t)end_i];
      if (s > max_s) { max_s = s; max_i = end_i; } else if (max_s - s > max_drop) break;
      if (!(i >= 0 && t[(size_t)i] == 0 && end_i >= 0)) break;
      i = end_i;
    }
    i64 ii = i0;
    while (ii >= 0 && ii != end_i) { t[(size_t)ii] = 0; ii = pprev[(size_t)ii]; }
  }
  return max_i;
}

// ---- lchain.rs:59-176 ---------------------------------------------------------------------
void chain_dp_all(const std::vector<Anchor>& a, const ChainParams& p, Chains& chains, std::vector<i32>& scores,
                  DpTrace* trace) {
  chains.clear(); scores.clear();
  const size_t n = a.size();
  if (trace) { trace->f.clear(); trace->v.clear(); trace->pprev.clear(); }
  if (n == 0) return;
  i32 max_dist_x = p.max_dist_x, max_dist_y = p.max_dist_y;
  if (max_dist_x < p.bw) max_dist_x = p.bw;
  if (max_dist_y < p.bw) max_dist_y = p.bw;
  std::vector<i32> f(n, 0), v(n, 0), t(n, 0);
  std::vector<i64> pprev(n, -1);
  size_t st = 0;
  u64 cells = 0;
  for (size_t i = 0; i < n; ++i) {
    while (st < i && (a_rid(a[st]) != a_rid(a[i]) || a_rev(a[st]) != a_rev(a[i]) ||
                      a_rpos(a[i]) > a_rpos(a[st]) + max_dist_x)) st += 1;
    i64 max_j = -1;
    i32 max_f = a_qspan(a[i]);
    size_t start_j = ((i32)i - p.max_chain_iter > (i32)st) ? (size_t)((i32)i - p.max_chain_iter) : st;
    i32 n_skip = 0;
    for (size_t j = i; j-- > start_j;) {
      cells += 1;
      if (a_rid(a[j]) != a_rid(a[i]) || a_rev(a[j]) != a_rev(a[i])) continue;
      i32 sc0;
      if (comput_sc(a[i], a[j], max_dist_x, max_dist_y, p.bw, p.chn_pen_gap, p.chn_pen_skip, &sc0)) {
        i32 sc = sc0 + f[j];
        if (sc > max_f) { max_f = sc; max_j = (i64)j; if (n_skip > 0) n_skip -= 1; }
        else if (t[j] == (i32)i) { n_skip += 1; if (n_skip > p.max_chain_skip) break; }
        if (pprev[j] >= 0) t[(size_t)pprev[j]] = (i32)i;
      }
    }
    f[i] = max_f; pprev[i] = max_j;
    v[i] = (max_j >= 0 && v[(size_t)max_j] > max_f) ? v[(size_t)max_j] : max_f;
  }
  if (trace) { trace->f = f; trace->v = v; trace->pprev = pprev; trace->cells += cells; }
  // z: (f[i], i) with f[i] > 0, ascending f; the Rust sort is unstable (F10): ties -> ascending index
  std::vector<std::pair<i32, size_t>> z;
  for (size_t i = 0; i < n; ++i) if (f[i] > 0) z.emplace_back(f[i], i);
  if (z.empty()) return;
  std::stable_sort(z.begin(), z.end(), [](const std::pair<i32, size_t>& x, const std::pair<i32, size_t>& y) { return x.first < y.first; });
  std::fill(t.begin(), t.end(), 0);
  size_t n_v = 0, n_u = 0;
  for (size_t kk = z.size(); kk-- > 0;) {  // first pass: sizes only (lchain.rs:99-126)
    size_t i0 = z[kk].second;
    if (t[i0] != 0) continue;
    i64 end_i = bk_end((i64)i0, z[kk].first, f, pprev, t, p.max_drop);
    size_t len0 = n_v;
    i64 i = (i64)i0;
    while (i >= 0 && i != end_i) { n_v += 1; t[(size_t)i] = 1; i = pprev[(size_t)i]; }
    i32 sc = i < 0 ? z[kk].first : z[kk].first - f[(size_t)i];
    if (sc >= p.min_chain_score && n_v > len0 && (i32)(n_v - len0) >= p.min_cnt) n_u += 1; else n_v = len0;
  }
  chains.reserve(n_u); scores.reserve(n_u);
  std::fill(t.begin(), t.end(), 0);
  for (size_t kk = z.size(); kk-- > 0;) {  // second pass (lchain.rs:127-160)
    size_t i0 = z[kk].second;
    if (t[i0] != 0) continue;
    i64 end_i = bk_end((i64)i0, z[kk].first, f, pprev, t, p.max_drop);
    std::vector<size_t> v_idxs;
    i64 i = (i64)i0;
    while (i >= 0 && i != end_i) { v_idxs.push_back((size_t)i); t[(size_t)i] = 1; i = pprev[(size_t)i]; }
    i32 sc = i < 0 ? z[kk].first : z[kk].first - f[(size_t)i];
    if (sc >= p.min_chain_score && (i32)v_idxs.size() >= p.min_cnt) {
      std::reverse(v_idxs.begin(), v_idxs.end());
      scores.push_back(sc);
      chains.push_back(std::move(v_idxs));
    }
  }
  if (chains.empty()) {  // fallback (lchain.rs:162-173): LAST maximum of f (Iterator::max_by_key)
    size_t best_i = 0;
    for (size_t i = 1; i < n; ++i) if (f[i] >= f[best_i]) best_i = i;
    std::vector<size_t> v_idxs;
    i64 i = (i64)best_i;
    while (i >= 0) { v_idxs.push_back((size_t)i); i = pprev[(size_t)i]; }
    std::reverse(v_idxs.begin(), v_idxs.end());
    if (!v_idxs.empty()) { chains.push_back(std::move(v_idxs)); scores.push_back(v[best_i]); }
  }
  sort_chains_stable(a, chains, scores);
}

// ---- lchain.rs:54-57 ----------------------------------------------------------------------
std::vector<size_t> chain_dp(const std::vector<Anchor>& a, const ChainParams& p) {
  Chains c; std::vector<i32> s;
  chain_dp_all(a, p, c, s);
  return c.empty() ? std::vector<size_t>() : c[0];
}

// ---- lchain.rs:220-235 --------------------------------------------------------------------
std::vector<bool> select_primary_secondary(const std::vector<Anchor>& a, const Chains& chains, const std::vector<i32>& scores,
                                           float mask_level) {
  (void)scores;
  std::vector<std::pair<i32, i32>> primaries;
  std::vector<bool> is_primary(chains.size(), true);
  for (size_t ci = 0; ci < chains.size(); ++ci) {
    i32 qs, qe; chain_qrange(a, chains[ci], &qs, &qe);
    bool overlapped = false;
    for (auto& pr : primaries) {
      float ov = (float)std::max(std::min(qe, pr.second) - std::max(qs, pr.first), 0);
      float len = (float)std::max(qe - qs, 1);
      if (ov / len >= mask_level) { overlapped = true; break; }
    }
    if (overlapped) is_primary[ci] = false; else primaries.emplace_back(qs, qe);
  }
  return is_primary;
}

// ---- lchain.rs:237-260 --------------------------------------------------------------------
void select_and_filter_chains(const std::vector<Anchor>& a, const Chains& chains_in, const std::vector<i32>& scores_in,
                              float mask_level, float pri_ratio, size_t best_n, Chains& out_chains,
                              std::vector<i32>& out_scores, std::vector<bool>& out_is_primary, i32* s1_o, i32* s2_o) {
  out_chains.clear(); out_scores.clear(); out_is_primary.clear();
  *s1_o = 0; *s2_o = 0;
  if (chains_in.empty()) return;
  Chains chains = chains_in;
  std::vector<i32> scores = scores_in;
  // NB: main.rs:216-217 passes merged chains with the UNMERGED scores; sort_chains_stable indexes
  // scores[i] for i < chains.len() (always in range because merging never adds chains) and its map
  // only picks scores for the surviving indices, so scores2.len() == chains.len().
  {
    std::vector<i32> sc_trunc(scores.begin(), scores.begin() + std::min(scores.size(), chains.size()));
    sort_chains_stable(a, chains, sc_trunc);
    scores.swap(sc_trunc);
  }
  std::vector<bool> is_primary = select_primary_secondary(a, chains, scores, mask_level);
  i32 s1 = scores[0], s2 = 0;
  size_t sec_kept = 0;
  for (size_t i = 0; i < chains.size(); ++i) {
    if (i == 0) { out_chains.push_back(chains[i]); out_scores.push_back(scores[i]); out_is_primary.push_back(true); }
    else {
      if (!is_primary[i]) continue;
      if ((float)scores[i] >= pri_ratio * (float)s1) {
        if (sec_kept < best_n) { out_chains.push_back(chains[i]); out_scores.push_back(scores[i]); out_is_primary.push_back(false); sec_kept += 1; }
      }
      if (s2 == 0) s2 = scores[i];
    }
  }
  *s1_o = s1; *s2_o = s2;
}

// ---- lchain.rs:262-286 / :288-314 ---------------------------------------------------------
static Chains merge_impl(const std::vector<Anchor>& a, const Chains& chains, bool with_gap, i32 max_gap_q, i32 max_gap_t) {
  std::vector<std::pair<i32, size_t>> items;
  for (size_t i = 0; i < chains.size(); ++i) { i32 qs, qe; chain_qrange(a, chains[i], &qs, &qe); items.emplace_back(qs, i); }
  // Rust: sort_unstable_by_key(qs) (F10): ties -> ascending original index
  std::stable_sort(items.begin(), items.end(), [](const std::pair<i32, size_t>& x, const std::pair<i32, size_t>& y) { return x.first < y.first; });
  Chains merged;
  for (auto& it : items) {
    const std::vector<size_t>& ch = chains[it.second];
    if (merged.empty()) { merged.push_back(ch); continue; }
    std::vector<size_t>& last = merged.back();
    const Anchor& a_last = a[last.back()];
    const Anchor& a_first = a[ch.front()];
    bool same = a_rid(a_last) == a_rid(a_first) && a_rev(a_last) == a_rev(a_first);
    i32 last_qs, last_qe, ch_qs, ch_qe;
    chain_qrange(a, last, &last_qs, &last_qe);
    chain_qrange(a, ch, &ch_qs, &ch_qe);
    bool do_merge;
    if (!with_gap) do_merge = same && ch_qs <= last_qe;
    else {
      i32 last_ts, last_te, ch_ts, ch_te;
      chain_trange(a, last, &last_ts, &last_te);
      chain_trange(a, ch, &ch_ts, &ch_te);
      i32 q_gap = ch_qs - last_qe, t_gap = ch_ts - last_te;
      do_merge = same && q_gap >= 0 && t_gap >= 0 && q_gap <= max_gap_q && t_gap <= max_gap_t;
    }
    if (do_merge) last.insert(last.end(), ch.begin(), ch.end()); else merged.push_back(ch);
  }
  return merged;
}
Chains merge_adjacent_chains(const std::vector<Anchor>& a, const Chains& chains) { return merge_impl(a, chains, false, 0, 0); }
Chains merge_adjacent_chains_with_gap(const std::vector<Anchor>& a, const Chains& chains, i32 max_gap_q, i32 max_gap_t) {
  return merge_impl(a, chains, true, max_gap_q, max_gap_t);
}

// ---- lchain.rs:316-319 --------------------------------------------------------------------
i32 chain_query_coverage(const std::vector<Anchor>& a, const std::vector<size_t>& chain) {
  i32 qs, qe; chain_qrange(a, chain, &qs, &qe);
  return std::max(qe - qs, 0);
}

// ---- lchain.rs:321-330 --------------------------------------------------------------------
void rescue_long_join(const std::vector<Anchor>& a, const Chains& chains, const std::vector<i32>& scores, const ChainParams& p,
                      i32 qlen, Chains& out_chains, std::vector<i32>& out_scores, u64* cells, bool* reran) {
  if (reran) *reran = false;
  if (chains.empty()) { out_chains = chains; out_scores = scores; return; }
  i32 best_cov = chain_query_coverage(a, chains[0]);
  i32 uncovered = std::max(qlen - best_cov, 0);
  bool rescue = uncovered > p.rmq_rescue_size || (float)best_cov < (float)qlen * (1.0f - p.rmq_rescue_ratio);
  if (!rescue) { out_chains = chains; out_scores = scores; return; }
  if (reran) *reran = true;
  ChainParams p2 = p;
  p2.bw = p.bw_long;
  DpTrace tr;
  chain_dp_all(a, p2, out_chains, out_scores, cells ? &tr : nullptr);
  if (cells) *cells += tr.cells;
}

// ---- paf.rs:130-222 -----------------------------------------------------------------------
bool paf_from_chain_with_primary(const Index& idx, const std::vector<Anchor>& a, const std::vector<size_t>& chain,
                                 const std::string& qname, const uint8_t* qseq, size_t qlen_sz, bool is_primary,
                                 PafRecord& rec, bool* panic) {
  if (panic) *panic = false;
  if (chain.empty()) return false;
  const i32 qlen = (i32)qlen_sz;
  char strand = a_rev(a[chain[0]]) ? '-' : '+';
  i32 qs = INT32_MAX, qe = -1, ts = INT32_MAX, te = -1;
  u32 cm = 0;
  for (size_t i : chain) {
    const Anchor& an = a[i];
    cm += 1;
    i32 s = a_qpos(an) - (a_qspan(an) - 1), e = a_qpos(an) + 1;
    if (s < qs) qs = s;
    if (e > qe) qe = e;
    i32 rs = a_rpos(an) - (a_qspan(an) - 1), re = a_rpos(an) + 1;
    if (rs < ts) ts = rs;
    if (re > te) te = re;
  }
  if (qs < 0) qs = 0;
  if (ts < 0) ts = 0;
  size_t rid0 = (size_t)((a[chain[0]].x >> 32) & 0x7fffffffULL);
  if (rid0 >= idx.seq.size()) { if (panic) *panic = true; return false; }  // Rust: index out of bounds panic (F5)
  const IndexSeq& tseq = idx.seq[rid0];
  std::string tname = tseq.has_name ? tseq.name : std::string("*");
  u32 tlen = tseq.len;
  i32 qs2 = qs, qe2 = qe, ts2 = ts, te2 = te;
  u32 mlen = (u32)std::max(qe2 - qs2, 0), blen = (u32)std::max(te2 - ts2, 0);
  std::vector<Minimizer> mv = collect_query_minimizers(qseq, qlen_sz, (size_t)idx.w, (size_t)idx.k);
  std::vector<i32> mini_pos; mini_pos.reserve(mv.size());
  u64 sum_k = 0;
  for (const Minimizer& m : mv) { mini_pos.push_back((i32)(u32)((m.rid_pos_strand >> 1) & 0xffffffffULL)); sum_k += (m.key_span & 0xff); }
  float avg_k = !mv.empty() ? (float)sum_k / (float)mv.size() : (float)idx.k;
  auto qpos_fwd = [&](const Anchor& an) -> i32 {
    i32 qp = a_qpos(an), qsp = a_qspan(an);
    return a_rev(an) ? qlen - 1 - (qp + 1 - qsp) : qp;
  };
  std::vector<i32> cq; cq.reserve(chain.size());
  if (strand == '-') for (size_t t = chain.size(); t-- > 0;) cq.push_back(qpos_fwd(a[chain[t]]));
  else for (size_t i : chain) cq.push_back(qpos_fwd(a[i]));
  float dv = 0.0f;
  if (!mini_pos.empty() && !cq.empty()) {
    i32 first = cq[0];
    // slice::binary_search (any equal element) followed by the rewind to the first equal one
    size_t lo = 0, hi = mini_pos.size(); bool found = false; size_t st = 0;
    while (lo < hi) {
      size_t mid = lo + (hi - lo) / 2;
      if (mini_pos[mid] == first) { found = true; st = mid; break; }
      if (mini_pos[mid] < first) lo = mid + 1; else hi = mid;
    }
    if (found) {
      while (st > 0 && mini_pos[st - 1] == first) st -= 1;
      size_t j = st, kq = 1, en = st;
      i32 n_match = 1;
      while (j + 1 < mini_pos.size() && kq < cq.size()) {
        j += 1;
        if (mini_pos[j] == cq[kq]) { n_match += 1; en = j; kq += 1; }
      }
      i32 n_tot = (i32)((en - st) + 1);
      i32 r_qs_final = strand == '-' ? qlen - qe2 : qs2;
      i32 r_qe_final = strand == '-' ? qlen - qs2 : qe2;
      i32 r_rs = ts2, r_re = te2;
      i32 avg_k_i = f32_to_i32(avg_k);
      if (r_qs_final > avg_k_i && r_rs > avg_k_i) n_tot += 1;
      if ((qlen - r_qe_final) > avg_k_i && ((i32)tlen - r_re) > avg_k_i) n_tot += 1;
      float frac = (float)n_match / (float)n_tot;
      if (frac >= 1.0f) dv = 0.0f;
      else dv = 1.0f - powf(frac, 1.0f / std::max(avg_k, 1.0f));
    }
  }
  rec.qname = qname; rec.qlen = (u32)qlen_sz; rec.qstart = (u32)qs2; rec.qend = (u32)qe2; rec.strand = strand;
  rec.tname = tname; rec.tlen = tlen; rec.tstart = (u32)ts2; rec.tend = (u32)te2; rec.nm = mlen; rec.blen = blen;
  rec.mapq = 60; rec.tp = is_primary ? 'P' : 'S'; rec.cm = cm; rec.s1 = 0; rec.s2 = 0; rec.dv = dv; rec.rl = 0;
  return true;
}

// ---- paf.rs:224-236 -----------------------------------------------------------------------
std::string write_paf(const PafRecord& rec) {
  u32 qs, qe;
  if (rec.strand == '-') { qs = rec.qlen - rec.qend; qe = rec.qlen - rec.qstart; } else { qs = rec.qstart; qe = rec.qend; }
  char nums[256];
  snprintf(nums, sizeof nums, "\t%u\t%u\t%u\t%c\t", rec.qlen, qs, qe, rec.strand);
  std::string s = rec.qname + nums + rec.tname;
  snprintf(nums, sizeof nums, "\t%u\t%u\t%u\t%u\t%u\t%u\ttp:A:%c\tcm:i:%u\ts1:i:%u\ts2:i:%u\tdv:f:%.4f\trl:i:%u", rec.tlen,
           rec.tstart, rec.tend, rec.nm, rec.blen, (unsigned)rec.mapq, rec.tp, rec.cm, rec.s1, rec.s2, (double)rec.dv, rec.rl);
  s += nums;
  return s;
}

// ---- paf.rs:238-248 -----------------------------------------------------------------------
std::vector<std::string> write_paf_many_with_scores(const Index& idx, const std::vector<Anchor>& a, const Chains& chains,
                                                    i32 top_s1, i32 top_s2, const std::string& qname, const uint8_t* qseq,
                                                    size_t qlen, bool* panic) {
  std::vector<std::string> out;
  if (panic) *panic = false;
  for (size_t ci = 0; ci < chains.size(); ++ci) {
    PafRecord rec; bool pn = false;
    if (paf_from_chain_with_primary(idx, a, chains[ci], qname, qseq, qlen, ci == 0, rec, &pn)) {
      rec.s1 = (u32)std::max(top_s1, 0);
      rec.s2 = (u32)std::max(top_s2, 0);
      out.push_back(write_paf(rec));
    }
    if (pn) { if (panic) *panic = true; break; }
  }
  return out;
}

// ---- main.rs:105-123 ----------------------------------------------------------------------
ChainParams default_chain_params(i32 k) {
  float chain_gap_scale = 0.8f;
  float chn_pen_gap = 0.01f * chain_gap_scale * (float)k;
  ChainParams p;
  p.max_dist_x = 5000; p.max_dist_y = 5000; p.bw = 500; p.max_chain_iter = 5000; p.min_chain_score = 40; p.min_cnt = 3;
  p.chn_pen_gap = chn_pen_gap; p.chn_pen_skip = 0.0f; p.max_chain_skip = 25; p.max_drop = 500; p.bw_long = 20000;
  p.rmq_rescue_size = 1000; p.rmq_rescue_ratio = 0.1f;
  return p;
}

// ---- main.rs:125-133 ----------------------------------------------------------------------
void apply_preset(const std::string& preset, i32* w, i32* k) {
  if (preset == "map-ont") { *k = 15; *w = 10; }
  else if (preset == "map-hifi" || preset == "lr:hq") { *k = 19; *w = 10; }
  else if (preset == "sr") { *k = 21; *w = 11; }
}

static bool parse_i32(const std::string& s, i32* v) {  // str::parse::<i32>: optional sign, digits only, no overflow
  if (s.empty()) return false;
  size_t i = 0; bool neg = false;
  if (s[0] == '+' || s[0] == '-') { neg = s[0] == '-'; i = 1; if (s.size() == 1) return false; }
  long long acc = 0;
  for (; i < s.size(); ++i) {
    if (s[i] < '0' || s[i] > '9') return false;
    acc = acc * 10 + (s[i] - '0');
    if (acc > 2147483648LL) return false;
  }
  if (neg) acc = -acc;
  if (acc > INT32_MAX || acc < INT32_MIN) return false;
  *v = (i32)acc;
  return true;
}

// ---- main.rs:199-208 ----------------------------------------------------------------------
ChainParams align_chain_params(const AlignOpts& o) {
  ChainParams p = default_chain_params(o.k);
  p.max_dist_x = o.max_gap; p.max_dist_y = o.max_gap;
  p.min_cnt = o.min_cnt; p.min_chain_score = o.min_chain_score;
  if (o.has_r && !o.r.empty()) {
    size_t c = o.r.find(',');
    std::string p0 = o.r.substr(0, c);
    i32 v;
    if (parse_i32(p0, &v)) p.bw = v;
    if (c != std::string::npos) {
      std::string rest = o.r.substr(c + 1);
      size_t c2 = rest.find(',');
      std::string p1 = rest.substr(0, c2);
      if (parse_i32(p1, &v)) p.bw_long = v;
    }
  }
  return p;
}

// ---- main.rs:193-219 for one read ---------------------------------------------------------
std::vector<std::string> align_read(const Index& idx, i32 mid_occ, const AlignOpts& o, const std::string& qname, const uint8_t* q,
                                    size_t qlen, AlignStats* st) {
  std::vector<std::string> lines;
  std::vector<Minimizer> mv = collect_query_minimizers(q, qlen, (size_t)o.w, (size_t)o.k);
  if (st) st->n_minimizers += mv.size();
  filter_query_minimizers(mv, 10, 0.01f);
  if (st) st->n_minimizers_kept += mv.size();
  std::vector<Anchor> anchors = build_anchors_filtered(idx, mv, (i32)qlen, mid_occ);
  if (st) st->n_anchors += anchors.size();
  ChainParams p = align_chain_params(o);
  Chains chains_all; std::vector<i32> scores_all;
  DpTrace tr;
  chain_dp_all(anchors, p, chains_all, scores_all, st ? &tr : nullptr);
  if (st) st->cells += tr.cells;
  bool panic = false;
  if (chains_all.empty()) {
    // main.rs:211-213: chain_dp again (same empty result) -> paf_from_chain returns None on an empty chain
    std::vector<size_t> chain = chain_dp(anchors, p);
    PafRecord rec;
    if (paf_from_chain_with_primary(idx, anchors, chain, qname, q, qlen, true, rec, &panic)) lines.push_back(write_paf(rec));
  } else {
    Chains chains_rescued; std::vector<i32> scores_rescued;
    u64 cells2 = 0;
    bool reran = false;
    rescue_long_join(anchors, chains_all, scores_all, p, (i32)qlen, chains_rescued, scores_rescued, st ? &cells2 : nullptr, &reran);
    if (st) { st->cells += cells2; st->rescued = reran; }
    Chains chains_merged = merge_adjacent_chains_with_gap(anchors, chains_rescued, p.max_dist_y, p.max_dist_y);
    Chains chains; std::vector<i32> scores; std::vector<bool> is_pri; i32 s1, s2;
    select_and_filter_chains(anchors, chains_merged, scores_rescued, o.mask_level, o.pri_ratio, o.best_n, chains, scores, is_pri, &s1, &s2);
    lines = write_paf_many_with_scores(idx, anchors, chains, s1, s2, qname, q, qlen, &panic);
  }
  if (st && panic) st->panic = true;
  return lines;
}

}  // namespace orc
