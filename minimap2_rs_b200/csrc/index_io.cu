// index_io.cu — host side of the index object: FASTA ingest, .mmi / native persistence, stats, calc_mid_occ.
// Replaces index.rs:111-141 (stats, calc_mid_occ), :156-424 (save/load in both formats), :427-438 (FASTA ingest) and
// main.rs:135-145 (load_index_auto).  The numeric content is produced on the device (index.cu); this file only
// (de)serialises the flat arrays.  Hash entries of a bucket are written in ascending key order (the reference's order
// is the random iteration order of a std HashMap).
#include "mm2_internal.cuh"

#include <algorithm>
#include <new>
#include <stdexcept>

#include "stages.cuh"

namespace {

// minimal FASTA reader with noodles' record semantics: name = bytes after '>' up to the first whitespace, sequence =
// concatenation of the following lines (index.rs:431-437)
int read_fasta_all(const char* path, std::vector<std::string>& names, std::vector<u8>& cat, std::vector<u64>& offs, bool first_only) {
  FILE* fp = fopen(path, "rb");
  if (!fp) { mm2_set_error("cannot open %s", path); return MM2_E_IO; }
  std::vector<char> buf(1 << 22);
  bool in_header = false, line_start = true, have = false, name_done = false, stop = false;
  size_t n;
  offs.clear(); names.clear(); cat.clear();
  while (!stop && (n = fread(buf.data(), 1, buf.size(), fp)) > 0) {
    for (size_t i = 0; i < n; ++i) {
      const char c = buf[i];
      if (in_header) {
        if (c == '\n') { in_header = false; line_start = true; }
        else if (!name_done) { if (c == ' ' || c == '\t' || c == '\r') name_done = true; else names.back().push_back(c); }
        continue;
      }
      if (line_start && c == '>') {
        if (have && first_only) { stop = true; break; }
        have = true; in_header = true; name_done = false; line_start = false;
        names.emplace_back(); offs.push_back(cat.size());
        continue;
      }
      if (c == '\n') { line_start = true; continue; }
      line_start = false;
      if (c == '\r') continue;
      if (have) cat.push_back((u8)c);
    }
  }
  const bool err = ferror(fp) != 0;
  fclose(fp);
  if (err) { mm2_set_error("read error on %s", path); return MM2_E_IO; }
  offs.push_back(cat.size());
  return MM2_OK;
}

struct FileW {
  FILE* f; bool ok = true;
  void bytes(const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) ok = false; }
  void u8_(u8 v) { bytes(&v, 1); }
  void u32_(u32 v) { bytes(&v, 4); }
  void i32_(i32 v) { bytes(&v, 4); }
  void u64_(u64 v) { bytes(&v, 8); }
};
struct MemR {  // bounds-checked little-endian reader over a whole file in memory
  const u8* p; size_t n, o = 0; bool ok = true;
  bool need(size_t k) { if (k > n - o) { ok = false; return false; } return true; }   // o <= n always; no wrap-around
  // k records of `rec` bytes each, with the multiplication checked against what is left in the file
  const u8* take_recs(size_t k, size_t rec) { if (k > (n - o) / rec) { ok = false; return nullptr; } return take(k * rec); }
  u8 u8_() { if (!need(1)) return 0; return p[o++]; }
  u32 u32_() { if (!need(4)) return 0; u32 v; memcpy(&v, p + o, 4); o += 4; return v; }
  i32 i32_() { return (i32)u32_(); }
  u64 u64_() { if (!need(8)) return 0; u64 v; memcpy(&v, p + o, 8); o += 8; return v; }
  const u8* take(size_t k) { if (!need(k)) return nullptr; const u8* q = p + o; o += k; return q; }
};

int slurp(const char* path, std::vector<u8>& data) {
  FILE* fp = fopen(path, "rb");
  if (!fp) { mm2_set_error("cannot open %s", path); return MM2_E_IO; }
  fseek(fp, 0, SEEK_END);
  const long sz = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  if (sz < 0) { fclose(fp); mm2_set_error("cannot stat %s", path); return MM2_E_IO; }
  data.resize((size_t)sz);
  const size_t got = sz ? fread(data.data(), 1, (size_t)sz, fp) : 0;
  fclose(fp);
  if (got != (size_t)sz) { mm2_set_error("short read on %s", path); return MM2_E_IO; }
  return MM2_OK;
}

// host copies of the flat device arrays
struct HostIndex {
  std::vector<u64> hkeys, hvals, koff, poff, p;
  std::vector<u32> S;
};
int download(const mm2_index* idx, HostIndex& h, bool want_S) {
  CUDA_TRY(cudaSetDevice(idx->device));
  const size_t nb = (size_t)1 << idx->b;
  h.hkeys.resize(idx->n_keys); h.hvals.resize(idx->n_keys); h.koff.resize(nb + 1); h.poff.resize(nb + 1); h.p.resize(idx->n_p);
  if (idx->n_keys) {   // device layout: interleaved {key, value} records
    std::vector<u64> kv((size_t)idx->n_keys * 2);
    CUDA_TRY(cudaMemcpy(kv.data(), idx->kv.p, idx->n_keys * 16, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < idx->n_keys; ++i) { h.hkeys[i] = kv[2 * i]; h.hvals[i] = kv[2 * i + 1]; }
  }
  CUDA_TRY(cudaMemcpy(h.koff.data(), idx->bkt_koff.p, (nb + 1) * 8, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(h.poff.data(), idx->bkt_poff.p, (nb + 1) * 8, cudaMemcpyDeviceToHost));
  if (idx->n_p) CUDA_TRY(cudaMemcpy(h.p.data(), idx->p.p, idx->n_p * 8, cudaMemcpyDeviceToHost));
  if (want_S) {
    h.S.resize(idx->S_words_alloc);
    if (idx->S_words_alloc) CUDA_TRY(cudaMemcpy(h.S.data(), idx->S.p, idx->S_words_alloc * 4, cudaMemcpyDeviceToHost));
  }
  return MM2_OK;
}

// upload flat host arrays into a fresh index object and build the lookup table + occurrence histogram
int upload(mm2_ctx* ctx, mm2_index* idx, HostIndex& h) {
  const size_t nb = (size_t)1 << idx->b;
  idx->n_keys = h.hkeys.size(); idx->n_p = h.p.size();
  MM2_TRY(idx->kv.ensure(std::max<size_t>(1, idx->n_keys) * 16));
  MM2_TRY(idx->p.ensure(std::max<size_t>(1, idx->n_p) * 8));
  MM2_TRY(idx->bkt_koff.ensure((nb + 1) * 8));
  MM2_TRY(idx->bkt_poff.ensure((nb + 1) * 8));
  MM2_TRY(idx->S.ensure(std::max<size_t>(1, h.S.size()) * 4));
  MM2_TRY(idx->seq_len.ensure(std::max<size_t>(1, idx->lens.size()) * 4));
  if (idx->n_keys) {
    std::vector<u64> kv((size_t)idx->n_keys * 2);
    for (size_t i = 0; i < idx->n_keys; ++i) { kv[2 * i] = h.hkeys[i]; kv[2 * i + 1] = h.hvals[i]; }
    CUDA_TRY(cudaMemcpy(idx->kv.p, kv.data(), idx->n_keys * 16, cudaMemcpyHostToDevice));
  }
  if (idx->n_p) CUDA_TRY(cudaMemcpy(idx->p.p, h.p.data(), idx->n_p * 8, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(idx->bkt_koff.p, h.koff.data(), (nb + 1) * 8, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(idx->bkt_poff.p, h.poff.data(), (nb + 1) * 8, cudaMemcpyHostToDevice));
  if (!h.S.empty()) CUDA_TRY(cudaMemcpy(idx->S.p, h.S.data(), h.S.size() * 4, cudaMemcpyHostToDevice));
  if (!idx->lens.empty()) CUDA_TRY(cudaMemcpy(idx->seq_len.p, idx->lens.data(), idx->lens.size() * 4, cudaMemcpyHostToDevice));
  idx->S_words_alloc = h.S.size();
  idx->occ_hist.assign(65536, 0);
  idx->occ_big.clear();
  u64 sum_occ = 0;
  for (size_t i = 0; i < idx->n_keys; ++i) {
    const u64 c = (h.hkeys[i] & 1) ? 1 : (h.hvals[i] & 0xffffffffULL);
    sum_occ += c;
    if (c < 65536) idx->occ_hist[c] += 1; else idx->occ_big.push_back((u32)c);
  }
  std::sort(idx->occ_big.begin(), idx->occ_big.end());
  idx->n_minimizers = sum_occ;
  MM2_TRY(index_build_lookup(ctx, idx));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return MM2_OK;
}

// Slot order of C minimap2's per-bucket hash table (klib khash 0.2.8 as instantiated in minimap2's index.c:
// KHASH_INIT(idx, uint64_t, uint64_t, 1, idx_hash, idx_eq) with idx_hash(a) = a >> 1, truncated to khint_t), filled the way
// mm_idx_post's worker does: kh_resize(h, n_keys), then one kh_put per key in ascending key order.  mm_idx_dump walks the
// slots 0 .. kh_end and writes the occupied ones, so this permutation is the entry order of an .mmi written by C minimap2.
// `ent` comes in ascending key order and leaves in slot order.  (Restated from the published klib algorithm; no minimap2
// binary exists in this image to compare with, so the order is "as specified", and any order loads.)
struct KhashSim {
  std::vector<u64> keys, vals;
  std::vector<u8> st;   // 0 empty, 1 occupied, 2 deleted
  u32 n_buckets = 0, size = 0, n_occupied = 0, upper_bound = 0;
  static u32 roundup32(u32 x) { --x; x |= x >> 1; x |= x >> 2; x |= x >> 4; x |= x >> 8; x |= x >> 16; return ++x; }
  static u32 hash(u64 key) { return (u32)(key >> 1); }
  void resize(u32 want) {
    u32 nn = roundup32(want);
    if (nn < 4) nn = 4;
    if (size >= (u32)(nn * 0.77 + 0.5)) return;   // requested size is too small
    std::vector<u8> nst(nn, 0);
    if (n_buckets < nn) { keys.resize(nn, 0); vals.resize(nn, 0); }
    for (u32 j = 0; j != n_buckets; ++j) {
      if (st[j] != 1) continue;
      u64 key = keys[j], val = vals[j];
      const u32 new_mask = nn - 1;
      st[j] = 2;
      for (;;) {   // kick-out process
        u32 i = hash(key) & new_mask, step = 0;
        while (nst[i] != 0) i = (i + (++step)) & new_mask;
        nst[i] = 1;
        if (i < n_buckets && st[i] == 1) { std::swap(keys[i], key); std::swap(vals[i], val); st[i] = 2; }
        else { keys[i] = key; vals[i] = val; break; }
      }
    }
    if (n_buckets > nn) { keys.resize(nn); vals.resize(nn); }
    st.swap(nst);
    n_buckets = nn; n_occupied = size; upper_bound = (u32)(n_buckets * 0.77 + 0.5);
  }
  void put(u64 key, u64 val) {   // the key is absent (distinct keys); its low bit does not take part in hash / equality
    if (n_occupied >= upper_bound) { if (n_buckets > (size << 1)) resize(n_buckets - 1); else resize(n_buckets + 1); }
    const u32 mask = n_buckets - 1;
    u32 i = hash(key) & mask, step = 0, site = n_buckets, x = n_buckets;
    if (st[i] == 0) x = i;
    else {
      const u32 last = i;
      while (st[i] != 0 && (st[i] == 2 || (keys[i] >> 1) != (key >> 1))) {
        if (st[i] == 2) site = i;
        i = (i + (++step)) & mask;
        if (i == last) { x = site; break; }
      }
      if (x == n_buckets) x = (st[i] == 0 && site != n_buckets) ? site : i;
    }
    if (st[x] == 0) { ++size; ++n_occupied; } else if (st[x] == 2) ++size;
    keys[x] = key; vals[x] = val; st[x] = 1;
  }
};
void khash_slot_order(u64* kv, size_t n) {   // kv: n (key, value) pairs in ascending key order -> khash slot order
  if (n < 2) return;
  KhashSim h;
  h.resize((u32)n);
  for (size_t i = 0; i < n; ++i) h.put(kv[2 * i], kv[2 * i + 1]);
  size_t o = 0;
  for (u32 sidx = 0; sidx < h.n_buckets; ++sidx)
    if (h.st[sidx] == 1) { kv[2 * o] = h.keys[sidx]; kv[2 * o + 1] = h.vals[sidx]; ++o; }
}

bool ends_with(const char* s, const char* suf) {
  const size_t a = strlen(s), b = strlen(suf);
  return a >= b && memcmp(s + a - b, suf, b) == 0;
}

}  // namespace

extern "C" void mm2_index_free(mm2_index_t* idx) {
  if (!idx) return;
  cudaSetDevice(idx->device);
  idx->S.release(); idx->kv.release(); idx->bkt_koff.release(); idx->bkt_poff.release();
  idx->p.release(); idx->seq_len.release(); idx->tab.release(); idx->bloom.release();
  delete idx;
}

extern "C" int mm2_index_build_fasta(mm2_ctx_t* ctx, const char* path, int w, int k, int b, int flag, mm2_index_t** out) {
  if (!ctx || !path || !out) { mm2_set_error("mm2_index_build_fasta: NULL argument"); return MM2_E_ARG; }
  std::vector<std::string> names; std::vector<u8> cat; std::vector<u64> offs;
  MM2_TRY(read_fasta_all(path, names, cat, offs, false));
  std::vector<const char*> np;
  for (auto& s : names) np.push_back(s.c_str());
  const size_t nseq = names.size();
  return mm2_index_build_seqs(ctx, cat.data(), offs.data(), np.data(), nseq, w, k, b, flag, out);
}

// ---- index.rs:233-307 ---------------------------------------------------------------------------------------------------
static int save_mmi_impl(const mm2_index_t* idx, const char* path, bool khash_order);
extern "C" int mm2_index_save_mmi(const mm2_index_t* idx, const char* path) {
  if (!idx || !path) { mm2_set_error("mm2_index_save_mmi: NULL argument"); return MM2_E_ARG; }
  try { return save_mmi_impl(idx, path, false); }
  catch (const std::bad_alloc&) { mm2_set_error("out of host memory"); return MM2_E_OOM; }
}
extern "C" int mm2_index_save_mmi_khash(const mm2_index_t* idx, const char* path) {
  if (!idx || !path) { mm2_set_error("mm2_index_save_mmi_khash: NULL argument"); return MM2_E_ARG; }
  try { return save_mmi_impl(idx, path, true); }
  catch (const std::bad_alloc&) { mm2_set_error("out of host memory"); return MM2_E_OOM; }
}
static int save_mmi_impl(const mm2_index_t* idx, const char* path, bool khash_order) {
  HostIndex h;
  MM2_TRY(download(idx, h, true));
  FILE* fp = fopen(path, "wb");
  if (!fp) { mm2_set_error("cannot create %s", path); return MM2_E_IO; }
  std::vector<char> iobuf(1 << 22);
  setvbuf(fp, iobuf.data(), _IOFBF, iobuf.size());
  FileW wr{fp};
  wr.bytes("MMI\2", 4);
  wr.u32_((u32)idx->w); wr.u32_((u32)idx->k); wr.u32_((u32)idx->b); wr.u32_((u32)idx->lens.size()); wr.u32_((u32)idx->flag);
  u64 sum_len = 0;
  for (size_t i = 0; i < idx->lens.size(); ++i) {
    if (idx->has_name[i]) {
      const u8 l = (u8)std::min<size_t>(idx->names[i].size(), 255);
      wr.u8_(l); wr.bytes(idx->names[i].data(), l);
    } else wr.u8_(0);
    wr.u32_(idx->lens[i]);
    sum_len += idx->lens[i];
  }
  const size_t nb = (size_t)1 << idx->b;
  std::vector<u64> kv;
  for (size_t bi = 0; bi < nb; ++bi) {
    const u64 p0 = h.poff[bi], p1 = h.poff[bi + 1], k0 = h.koff[bi], k1 = h.koff[bi + 1];
    wr.u32_((u32)(p1 - p0));
    wr.bytes(h.p.data() + p0, (size_t)(p1 - p0) * 8);
    wr.u32_((u32)(k1 - k0));
    kv.resize((size_t)(k1 - k0) * 2);
    for (u64 q = k0; q < k1; ++q) { kv[(size_t)(q - k0) * 2] = h.hkeys[q]; kv[(size_t)(q - k0) * 2 + 1] = h.hvals[q]; }
    if (khash_order) khash_slot_order(kv.data(), (size_t)(k1 - k0));
    wr.bytes(kv.data(), kv.size() * 8);
  }
  // index.rs:291 writes the packed sequence unconditionally ("C checks NO_SEQ bit itself"); an index that was LOADED from a
  // file flagged MM_I_NO_SEQ (2) has none, and is written back the way C minimap2 dumps such an index: without it
  const size_t words = (size_t)((sum_len + 7) / 8);
  if (words > h.S.size()) {
    if (!(idx->flag & 2)) { fclose(fp); mm2_set_error("index has no sequence array to write"); return MM2_E_FORMAT; }
  } else wr.bytes(h.S.data(), words * 4);
  bool ok = wr.ok;
  if (fclose(fp) != 0) ok = false;
  if (!ok) { mm2_set_error("write error on %s", path); return MM2_E_IO; }
  return MM2_OK;
}

// ---- index.rs:361-424 ---------------------------------------------------------------------------------------------------
static int load_mmi_impl(mm2_ctx_t* ctx, const char* path, mm2_index_t** out);
extern "C" int mm2_index_load_mmi(mm2_ctx_t* ctx, const char* path, mm2_index_t** out) {
  if (!ctx || !path || !out) { mm2_set_error("mm2_index_load_mmi: NULL argument"); return MM2_E_ARG; }
  try { return load_mmi_impl(ctx, path, out); }   // nothing may throw across the C boundary
  catch (const std::bad_alloc&) { mm2_set_error("out of host memory while loading %s", path); return MM2_E_OOM; }
  catch (const std::exception& e) { mm2_set_error("invalid index file %s (%s)", path, e.what()); return MM2_E_FORMAT; }
}
static int load_mmi_impl(mm2_ctx_t* ctx, const char* path, mm2_index_t** out) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  std::vector<u8> data;
  MM2_TRY(slurp(path, data));
  MemR rd{data.data(), data.size()};
  const u8* magic = rd.take(4);
  if (!magic || memcmp(magic, "MMI\2", 4) != 0) { mm2_set_error("invalid MMI magic"); return MM2_E_FORMAT; }
  mm2_index* idx = new mm2_index();
  idx->device = ctx->device;
  idx->w = (i32)rd.u32_(); idx->k = (i32)rd.u32_(); idx->b = (i32)rd.u32_(); idx->n_seq = rd.u32_(); idx->flag = (i32)rd.u32_();
  auto fail = [&](int rc, const char* msg) { mm2_set_error("%s", msg); mm2_index_free(idx); return rc; };
  if (!rd.ok || idx->b < 0 || idx->b > 30) return fail(MM2_E_FORMAT, "truncated or invalid MMI header");
  u64 sum_len = 0;
  for (u32 i = 0; i < idx->n_seq && rd.ok; ++i) {
    const size_t nl = rd.u8_();
    const u8* nm = rd.take(nl);
    idx->has_name.push_back(nl > 0);
    idx->names.push_back(nm && nl ? std::string((const char*)nm, nl) : std::string());
    const u32 len = rd.u32_();
    idx->lens.push_back(len); idx->seq_offset.push_back(sum_len); idx->is_alt.push_back(0);
    sum_len += len;
  }
  if (!rd.ok) return fail(MM2_E_FORMAT, "truncated MMI sequence table");
  idx->total_len = sum_len;
  const size_t nb = (size_t)1 << idx->b;
  HostIndex h;
  h.koff.assign(nb + 1, 0); h.poff.assign(nb + 1, 0);
  std::vector<std::pair<u64, u64>> ent;
  for (size_t bi = 0; bi < nb && rd.ok; ++bi) {
    const size_t n = rd.u32_();
    const u8* pp = rd.take_recs(n, 8);
    if (!rd.ok) break;
    h.poff[bi] = h.p.size();
    h.p.resize(h.p.size() + n);
    if (n) memcpy(h.p.data() + h.poff[bi], pp, n * 8);
    const size_t size = rd.u32_();
    const u8* ee = rd.take_recs(size, 16);
    if (!rd.ok) break;
    h.koff[bi] = h.hkeys.size();
    ent.resize(size);
    for (size_t q = 0; q < size; ++q) { memcpy(&ent[q].first, ee + q * 16, 8); memcpy(&ent[q].second, ee + q * 16 + 8, 8); }
    std::sort(ent.begin(), ent.end());  // any writer's order (C minimap2 dumps khash order) -> ascending key
    for (auto& kv : ent) { h.hkeys.push_back(kv.first); h.hvals.push_back(kv.second); }
  }
  if (!rd.ok) return fail(MM2_E_FORMAT, "truncated MMI bucket table");
  h.koff[nb] = h.hkeys.size(); h.poff[nb] = h.p.size();
  // MM_I_NO_SEQ (flag bit 1, value 2): C minimap2 dumps such an index without the packed sequence.  The reference's loader
  // (index.rs:417-419) reads the array unconditionally and fails on these files; honouring the flag is what makes indexes
  // written by `minimap2 -d` with --idx-no-seq usable.  Mapping never touches S.
  if (!(idx->flag & 2)) {
    const size_t words = (size_t)((sum_len + 7) / 8);
    const u8* sp = rd.take_recs(words, 4);
    if (!rd.ok) return fail(MM2_E_FORMAT, "truncated MMI sequence array");
    h.S.resize(words);
    if (words) memcpy(h.S.data(), sp, words * 4);
  }
  const int rc = upload(ctx, idx, h);
  if (rc != MM2_OK) { mm2_index_free(idx); return rc; }
  *out = idx;
  return MM2_OK;
}

// ---- index.rs:156-230 ---------------------------------------------------------------------------------------------------
extern "C" int mm2_index_save_native(const mm2_index_t* idx, const char* path) {
  if (!idx || !path) { mm2_set_error("mm2_index_save_native: NULL argument"); return MM2_E_ARG; }
  HostIndex h;
  MM2_TRY(download(idx, h, true));
  FILE* fp = fopen(path, "wb");
  if (!fp) { mm2_set_error("cannot create %s", path); return MM2_E_IO; }
  std::vector<char> iobuf(1 << 22);
  setvbuf(fp, iobuf.data(), _IOFBF, iobuf.size());
  FileW wr{fp};
  wr.bytes("MM2RSIDX\0", 9);
  wr.u32_(1);
  wr.i32_(idx->w); wr.i32_(idx->k); wr.i32_(idx->b); wr.i32_(idx->flag); wr.u32_(idx->n_seq);
  wr.u32_((u32)idx->lens.size());
  for (size_t i = 0; i < idx->lens.size(); ++i) {
    wr.u8_(idx->has_name[i] ? 1 : 0);
    if (idx->has_name[i]) { wr.u32_((u32)idx->names[i].size()); wr.bytes(idx->names[i].data(), idx->names[i].size()); }
    wr.u64_(idx->seq_offset[i]); wr.u32_(idx->lens[i]); wr.u8_(idx->is_alt[i] ? 1 : 0);
  }
  wr.u64_((u64)h.S.size());
  wr.bytes(h.S.data(), h.S.size() * 4);
  const size_t nb = (size_t)1 << idx->b;
  wr.u32_((u32)nb);
  std::vector<u64> kv;
  for (size_t bi = 0; bi < nb; ++bi) {
    const u64 p0 = h.poff[bi], p1 = h.poff[bi + 1], k0 = h.koff[bi], k1 = h.koff[bi + 1];
    wr.u64_(p1 - p0);
    wr.bytes(h.p.data() + p0, (size_t)(p1 - p0) * 8);
    const bool has_h = k1 > k0;  // index.rs:78: an empty bucket keeps h = None
    wr.u8_(has_h ? 1 : 0);
    if (has_h) {
      wr.u64_(k1 - k0);
      kv.resize((size_t)(k1 - k0) * 2);
      for (u64 q = k0; q < k1; ++q) { kv[(size_t)(q - k0) * 2] = h.hkeys[q]; kv[(size_t)(q - k0) * 2 + 1] = h.hvals[q]; }
      wr.bytes(kv.data(), kv.size() * 8);
    }
  }
  bool ok = wr.ok;
  if (fclose(fp) != 0) ok = false;
  if (!ok) { mm2_set_error("write error on %s", path); return MM2_E_IO; }
  return MM2_OK;
}

// ---- index.rs:309-358 ---------------------------------------------------------------------------------------------------
static int load_native_impl(mm2_ctx_t* ctx, const char* path, mm2_index_t** out);
extern "C" int mm2_index_load_native(mm2_ctx_t* ctx, const char* path, mm2_index_t** out) {
  if (!ctx || !path || !out) { mm2_set_error("mm2_index_load_native: NULL argument"); return MM2_E_ARG; }
  try { return load_native_impl(ctx, path, out); }
  catch (const std::bad_alloc&) { mm2_set_error("out of host memory while loading %s", path); return MM2_E_OOM; }
  catch (const std::exception& e) { mm2_set_error("invalid index file %s (%s)", path, e.what()); return MM2_E_FORMAT; }
}
static int load_native_impl(mm2_ctx_t* ctx, const char* path, mm2_index_t** out) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  std::vector<u8> data;
  MM2_TRY(slurp(path, data));
  MemR rd{data.data(), data.size()};
  const u8* magic = rd.take(9);
  if (!magic || memcmp(magic, "MM2RSIDX\0", 9) != 0) { mm2_set_error("invalid index file magic"); return MM2_E_FORMAT; }
  (void)rd.u32_();
  mm2_index* idx = new mm2_index();
  idx->device = ctx->device;
  idx->w = rd.i32_(); idx->k = rd.i32_(); idx->b = rd.i32_(); idx->flag = rd.i32_(); idx->n_seq = rd.u32_();
  auto fail = [&](int rc, const char* msg) { mm2_set_error("%s", msg); mm2_index_free(idx); return rc; };
  if (!rd.ok || idx->b < 0 || idx->b > 30) return fail(MM2_E_FORMAT, "truncated or invalid index header");
  const size_t n_seq = rd.u32_();
  u64 total = 0;
  for (size_t i = 0; i < n_seq && rd.ok; ++i) {
    const bool hn = rd.u8_() != 0;
    std::string name;
    if (hn) { const size_t l = rd.u32_(); const u8* nm = rd.take(l); if (nm) name.assign((const char*)nm, l); }
    idx->has_name.push_back(hn); idx->names.push_back(name);
    idx->seq_offset.push_back(rd.u64_()); idx->lens.push_back(rd.u32_()); idx->is_alt.push_back(rd.u8_() != 0);
    total += idx->lens.back();
  }
  if (!rd.ok) return fail(MM2_E_FORMAT, "truncated index sequence table");
  idx->total_len = total;
  HostIndex h;
  const size_t s_words = (size_t)rd.u64_();
  const u8* sp = rd.take_recs(s_words, 4);
  if (!rd.ok) return fail(MM2_E_FORMAT, "truncated index sequence array");
  h.S.resize(s_words);
  if (s_words) memcpy(h.S.data(), sp, s_words * 4);
  const size_t nb = rd.u32_();
  if (!rd.ok || nb != ((size_t)1 << idx->b)) return fail(MM2_E_FORMAT, "bucket count does not match b");
  h.koff.assign(nb + 1, 0); h.poff.assign(nb + 1, 0);
  std::vector<std::pair<u64, u64>> ent;
  for (size_t bi = 0; bi < nb && rd.ok; ++bi) {
    const size_t n = (size_t)rd.u64_();
    const u8* pp = rd.take_recs(n, 8);
    if (!rd.ok) break;
    h.poff[bi] = h.p.size();
    h.p.resize(h.p.size() + n);
    if (n) memcpy(h.p.data() + h.poff[bi], pp, n * 8);
    h.koff[bi] = h.hkeys.size();
    if (rd.u8_() != 0) {
      const size_t size = (size_t)rd.u64_();
      const u8* ee = rd.take_recs(size, 16);
      if (!rd.ok) break;
      ent.resize(size);
      for (size_t q = 0; q < size; ++q) { memcpy(&ent[q].first, ee + q * 16, 8); memcpy(&ent[q].second, ee + q * 16 + 8, 8); }
      std::sort(ent.begin(), ent.end());
      for (auto& kv : ent) { h.hkeys.push_back(kv.first); h.hvals.push_back(kv.second); }
    }
  }
  if (!rd.ok) return fail(MM2_E_FORMAT, "truncated index bucket table");
  h.koff[nb] = h.hkeys.size(); h.poff[nb] = h.p.size();
  const int rc = upload(ctx, idx, h);
  if (rc != MM2_OK) { mm2_index_free(idx); return rc; }
  *out = idx;
  return MM2_OK;
}

// ---- main.rs:135-145 ----------------------------------------------------------------------------------------------------
extern "C" int mm2_index_load_auto(mm2_ctx_t* ctx, const char* path, int w, int k, int b, int flag, mm2_index_t** out) {
  if (!ctx || !path || !out) { mm2_set_error("mm2_index_load_auto: NULL argument"); return MM2_E_ARG; }
  if (ends_with(path, ".mmi")) return mm2_index_load_mmi(ctx, path, out);
  if (mm2_index_load_native(ctx, path, out) == MM2_OK) return MM2_OK;
  return mm2_index_build_fasta(ctx, path, w, k, b, flag, out);
}

// ---- index.rs:111-122 ---------------------------------------------------------------------------------------------------
extern "C" int mm2_index_stats(const mm2_index_t* idx, uint64_t* n_keys, double* avg_occ, double* avg_spacing, uint64_t* total_len) {
  if (!idx || !n_keys || !avg_occ || !avg_spacing || !total_len) { mm2_set_error("mm2_index_stats: NULL argument"); return MM2_E_ARG; }
  u64 nk = 0, so = 0;
  for (size_t c = 0; c < idx->occ_hist.size(); ++c) { nk += idx->occ_hist[c]; so += idx->occ_hist[c] * (u64)c; }
  for (u32 c : idx->occ_big) { nk += 1; so += c; }
  u64 tl = 0;
  for (u32 l : idx->lens) tl += l;
  *n_keys = nk;
  *avg_occ = nk > 0 ? (double)so / (double)nk : 0.0;
  *avg_spacing = so > 0 ? (double)tl / (double)so : 0.0;
  *total_len = tl;
  return MM2_OK;
}

// ---- index.rs:124-141: the ((1-frac)*n)-th smallest occurrence count, + 1 ------------------------------------------------
extern "C" int mm2_index_calc_mid_occ(const mm2_index_t* idx, float frac, int32_t* out) {
  if (!idx || !out) { mm2_set_error("mm2_index_calc_mid_occ: NULL argument"); return MM2_E_ARG; }
  u64 n = 0;
  for (u64 c : idx->occ_hist) n += c;
  n += idx->occ_big.size();
  if (n == 0) { *out = INT32_MAX; return MM2_OK; }
  const double x = (1.0 - (double)frac) * (double)n;  // frac is rounded to f32 first, then widened (index.rs:138)
  u64 pos = x <= 0.0 ? 0 : (x >= 18446744073709551615.0 ? ~0ULL : (u64)x);
  pos = std::min(pos, n - 1);
  u64 seen = 0;
  for (size_t c = 0; c < idx->occ_hist.size(); ++c) {
    seen += idx->occ_hist[c];
    if (pos < seen) { *out = (i32)c + 1; return MM2_OK; }
  }
  const u32 c = idx->occ_big[(size_t)(pos - seen)];
  *out = (i32)c + 1;
  return MM2_OK;
}

extern "C" int mm2_index_params(const mm2_index_t* idx, int32_t* w, int32_t* k, int32_t* b, int32_t* flag, uint32_t* n_seq) {
  if (!idx) { mm2_set_error("NULL index"); return MM2_E_ARG; }
  if (w) *w = idx->w;
  if (k) *k = idx->k;
  if (b) *b = idx->b;
  if (flag) *flag = idx->flag;
  if (n_seq) *n_seq = idx->n_seq;
  return MM2_OK;
}

extern "C" int mm2_index_seq(const mm2_index_t* idx, uint32_t rid, const char** name, uint32_t* len) {
  if (!idx) { mm2_set_error("NULL index"); return MM2_E_ARG; }
  if (rid >= idx->lens.size()) { mm2_set_error("rid %u out of range (%zu sequences)", rid, idx->lens.size()); return MM2_E_REF_PANIC; }
  if (name) *name = idx->has_name[rid] ? idx->names[rid].c_str() : "*";
  if (len) *len = idx->lens[rid];
  return MM2_OK;
}

// ---- index.rs:53-67 -----------------------------------------------------------------------------------------------------
extern "C" int mm2_index_get_ref_subseq(const mm2_index_t* idx, uint32_t rid, int32_t st, int32_t en, uint8_t** out, size_t* n) {
  if (!idx || !out || !n) { mm2_set_error("mm2_index_get_ref_subseq: NULL argument"); return MM2_E_ARG; }
  *out = (u8*)malloc(1); *n = 0;
  if (rid >= idx->lens.size()) return MM2_OK;
  u64 st0 = (u64)std::max(st, 0);
  u64 en0 = (u64)std::max(std::min(en, (i32)idx->lens[rid]), 0);
  if (st0 >= en0) return MM2_OK;
  st0 += idx->seq_offset[rid]; en0 += idx->seq_offset[rid];
  const u64 w0 = st0 >> 3, w1 = (en0 + 7) >> 3;
  if (w1 > idx->S_words_alloc) { mm2_set_error("index has no sequence array"); return MM2_E_FORMAT; }
  std::vector<u32> words((size_t)(w1 - w0));
  CUDA_TRY(cudaSetDevice(idx->device));
  CUDA_TRY(cudaMemcpy(words.data(), idx->S.as<u32>() + w0, words.size() * 4, cudaMemcpyDeviceToHost));
  free(*out);
  *out = (u8*)malloc((size_t)(en0 - st0));
  if (!*out) { mm2_set_error("out of host memory"); return MM2_E_OOM; }
  for (u64 o = st0; o < en0; ++o) {
    const u32 c = (words[(size_t)((o >> 3) - w0)] >> ((o & 7) << 2)) & 0xF;
    (*out)[o - st0] = c == 0 ? 'A' : c == 1 ? 'C' : c == 2 ? 'G' : c == 3 ? 'T' : 'N';
  }
  *n = (size_t)(en0 - st0);
  return MM2_OK;
}

extern "C" int mm2_index_build_timings(const mm2_index_t* idx, float* ms5, uint64_t* n_bases, uint64_t* n_minimizers) {
  if (!idx) { mm2_set_error("NULL index"); return MM2_E_ARG; }
  if (ms5) memcpy(ms5, idx->build_ms, sizeof idx->build_ms);
  if (n_bases) *n_bases = idx->total_len;
  if (n_minimizers) *n_minimizers = idx->n_minimizers;
  return MM2_OK;
}
