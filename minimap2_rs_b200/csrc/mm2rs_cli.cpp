// mm2rs_cli.cpp — the `mm2rs` command line over libmm2b200.so: same subcommands, flags and stdout as the reference's
// src/main.rs (index | anchors | chain | align), with every computation done on the GPU through the C ABI.
// Extension (off by default): `align --all-reads` maps every record of the query FASTA instead of only
// the first one (main.rs:92-103).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "mm2b200.h"

namespace {

[[noreturn]] void die(const std::string& msg, int code = 1) {
  fprintf(stderr, "Error: %s\n", msg.c_str());  // anyhow's `Error: ...` (main.rs:147)
  exit(code);
}
void check(int rc) {
  if (rc == MM2_OK) return;
  // inputs the reference asserts on exit with a Rust panic (101); I/O and format errors with 1
  die(mm2_last_error(), (rc == MM2_E_ARG || rc == MM2_E_REF_PANIC) ? 101 : 1);
}

struct Fasta {
  std::vector<std::string> names;
  std::vector<uint8_t> cat;
  std::vector<uint64_t> offs;
};

// One batch of query records: the bases live in a page-locked buffer (mm2_host_alloc) so that mm2_map_batch uploads them
// at full PCIe speed straight from where the parser wrote them.
struct Batch {
  std::vector<std::string> names;
  std::vector<uint64_t> offs;   // nreads + 1
  uint8_t* cat = nullptr;
  size_t len = 0, cap = 0;
  void clear() { names.clear(); offs.clear(); len = 0; }
  void push(uint8_t c) {
    if (len == cap) {
      const size_t ncap = cap ? cap * 2 : (size_t)1 << 20;
      uint8_t* nb = (uint8_t*)mm2_host_alloc(ncap);
      if (!nb) die("out of page-locked host memory");
      if (len) memcpy(nb, cat, len);
      if (cat) mm2_host_free(cat);
      cat = nb; cap = ncap;
    }
    cat[len++] = c;
  }
  ~Batch() { if (cat) mm2_host_free(cat); }
};

// Streaming FASTA / FASTQ reader.  noodles-like record semantics (name = up to the first whitespace;
// sequence lines concatenated).  FASTQ is an extension (the reference reads FASTA only): a file whose first byte is '@' --
// '@name ...', sequence lines up to a line starting with '+', then as many quality characters as there were bases.
// fill() appends whole records to a batch until it holds max_bases bases (or max_reads records) and stops at the next record
// boundary, so a file of any size is mapped in bounded memory.
struct ReadStream {
  FILE* fp = nullptr;
  std::vector<char> buf;
  size_t pos = 0, n = 0;
  bool fastq = false, eof = false;
  enum { FA_SEQ, FA_HEADER, FQ_HEADER, FQ_SEQ, FQ_PLUS, FQ_QUAL } st = FA_SEQ;
  bool line_start = true, name_done = false, have = false;   // have: inside a record
  size_t need_q = 0, rec_start = 0;
  bool open(const std::string& path) {
    fp = fopen(path.c_str(), "rb");
    if (!fp) return false;
    buf.resize(4 << 20);
    const int c0 = fgetc(fp);
    if (c0 != EOF) ungetc(c0, fp);
    fastq = c0 == '@';
    st = fastq ? FQ_QUAL : FA_SEQ;
    return true;
  }
  ~ReadStream() { if (fp) fclose(fp); }
  // true: the batch holds at least one record
  bool fill(Batch& b, size_t max_bases, size_t max_reads) {
    b.clear();
    if (have) {   // the previous batch stopped in front of this record's first character
      have = false;
    }
    for (;;) {
      if (pos == n) {
        if (eof) break;
        n = fread(buf.data(), 1, buf.size(), fp);
        pos = 0;
        if (n == 0) { eof = true; break; }
      }
      const char c = buf[pos];
      const bool starts_record = fastq ? (st == FQ_QUAL && need_q == 0 && c == '@') : (st == FA_SEQ && line_start && c == '>');
      if (starts_record) {
        if (!b.names.empty() && (b.len >= max_bases || b.names.size() >= max_reads)) {   // leave the character for the next batch
          b.offs.push_back(b.len);
          return true;
        }
        ++pos;
        b.names.emplace_back(); b.offs.push_back(b.len);
        name_done = false; line_start = false;
        st = fastq ? FQ_HEADER : FA_HEADER;
        continue;
      }
      ++pos;
      switch (st) {
        case FA_HEADER:
        case FQ_HEADER:
          if (c == '\n') { st = fastq ? FQ_SEQ : FA_SEQ; line_start = true; }
          else if (!name_done) { if (c == ' ' || c == '\t' || c == '\r') name_done = true; else b.names.back().push_back(c); }
          break;
        case FA_SEQ:
          if (c == '\n') { line_start = true; break; }
          line_start = false;
          if (c == '\r') break;
          if (!b.names.empty()) b.push((uint8_t)c);
          break;
        case FQ_SEQ:
          if (c == '\r') break;
          if (c == '\n') { line_start = true; break; }
          if (line_start && c == '+') { st = FQ_PLUS; need_q = b.len - (size_t)b.offs.back(); break; }
          line_start = false;
          b.push((uint8_t)c);
          break;
        case FQ_PLUS:
          if (c == '\n') { st = FQ_QUAL; line_start = true; }
          break;
        case FQ_QUAL:   // quality strings may contain '@', but only inside their need_q characters
          if (c == '\r') break;
          if (need_q) { if (c != '\n') --need_q; break; }
          break;
      }
    }
    if (b.names.empty()) return false;
    b.offs.push_back(b.len);
    return true;
  }
};

// the first record only (main.rs:92-103), or everything, into one in-memory Fasta (anchors / chain subcommands)
bool read_fasta(const std::string& path, bool first_only, Fasta& fa) {
  ReadStream rs;
  if (!rs.open(path)) return false;
  Batch b;
  if (rs.fill(b, first_only ? 0 : ~(size_t)0, first_only ? 1 : ~(size_t)0)) {
    fa.names = b.names; fa.offs = b.offs; fa.cat.assign(b.cat, b.cat + b.len);
  } else fa.offs.push_back(0);
  return true;
}

struct Args {
  std::vector<std::string> pos;
  int w = 10, k = 15, b = 14, bw = 5000;
  bool hpc = false, all_reads = false, has_r = false, khash = false;
  int gpus = 1, batch_mb = 1024;
  std::string dump, r, preset, output;
  bool has_dump = false, has_output = false, has_preset = false;
  float f = 2e-4f, mask_level = 0.5f, pri_ratio = 0.8f;
  int g = 5000, n = 3, m = 40, best_n = 5;
};

bool parse_i32(const std::string& s, int* v) {  // str::parse::<i32>
  if (s.empty()) return false;
  char* e = nullptr;
  long long x = strtoll(s.c_str(), &e, 10);
  if (*e || x > 2147483647LL || x < -2147483648LL) return false;
  for (size_t i = (s[0] == '+' || s[0] == '-') ? 1 : 0; i < s.size(); ++i) if (s[i] < '0' || s[i] > '9') return false;
  *v = (int)x;
  return true;
}
int need_i32(const std::string& flag, const std::string& s) {
  int v;
  if (!parse_i32(s, &v)) { fprintf(stderr, "error: invalid value '%s' for '%s'\n", s.c_str(), flag.c_str()); exit(2); }
  return v;
}
float need_f32(const std::string& flag, const std::string& s) {
  char* e = nullptr;
  float v = strtof(s.c_str(), &e);
  if (s.empty() || *e) { fprintf(stderr, "error: invalid value '%s' for '%s'\n", s.c_str(), flag.c_str()); exit(2); }
  return v;
}

void usage() {
  fprintf(stderr,
          "mm2rs (B200): Rust rewrite of minimap2 (WIP) — GPU hot path\n\nUsage: mm2rs <COMMAND>\n\nCommands:\n"
          "  index    <fasta> [-w 10] [-k 15] [-b 14] [-H|--hpc] [-d|--dump FILE] [--gpus N] [--khash-order]\n"
          "  anchors  <ref> <qry> [-w 10] [-k 15] [-H]\n"
          "  chain    <ref> <qry> [-w 10] [-k 15] [-r 5000] [-H]\n"
          "  align    <ref> <qry> [-w 10] [-k 15] [-H] [-f 2e-4] [-g 5000] [-r NUM[,NUM]] [-n 3] [-m 40] [-M 0.5] [-p 0.8]\n"
          "           [-N 5] [-x PRESET] [-a] [-o FILE] [--all-reads] [--batch-mb 1024]\n"
          "Extensions over the reference CLI: index --gpus N (bucket-sharded build on N GPUs), --khash-order (.mmi entries in C\n"
          "minimap2's hash-slot order); align --all-reads (every record of the query, streamed in batches of --batch-mb), FASTQ queries.\n");
}

// clap-style parsing of one subcommand's arguments: `-w 10`, `-w10`, `-w=10`, `--long v`, `--long=v`
Args parse(const std::string& cmd, int argc, char** argv) {
  Args a;
  if (cmd == "chain") a.bw = 5000;
  auto takes_value = [&](const std::string& f) {
    static const char* v[] = {"-w", "-k", "-b", "-d", "--dump", "-r", "-f", "-g", "-n", "-m", "-M", "--mask-level", "-p", "--pri-ratio",
                              "-N", "--best-n", "-x", "-o", "--gpus", "--batch-mb"};
    for (const char* x : v) if (f == x) return true;
    return false;
  };
  for (int i = 0; i < argc; ++i) {
    std::string t = argv[i];
    if (t.size() < 2 || t[0] != '-' || (t[1] >= '0' && t[1] <= '9')) { a.pos.push_back(t); continue; }
    std::string flag = t, val;
    bool has_val = false;
    if (t[1] != '-') {
      flag = t.substr(0, 2);
      if (t.size() > 2) { val = t.substr(t[2] == '=' ? 3 : 2); has_val = true; }
    } else {
      const size_t eq = t.find('=');
      if (eq != std::string::npos) { flag = t.substr(0, eq); val = t.substr(eq + 1); has_val = true; }
    }
    if (takes_value(flag)) {
      if (!has_val) {
        if (i + 1 >= argc) { fprintf(stderr, "error: a value is required for '%s'\n", flag.c_str()); exit(2); }
        val = argv[++i];
      }
      if (flag == "-w") a.w = need_i32(flag, val);
      else if (flag == "-k") a.k = need_i32(flag, val);
      else if (flag == "-b") a.b = need_i32(flag, val);
      else if (flag == "-d" || flag == "--dump") { a.dump = val; a.has_dump = true; }
      else if (flag == "-r") { if (cmd == "chain") a.bw = need_i32(flag, val); else { a.r = val; a.has_r = true; } }
      else if (flag == "-f") a.f = need_f32(flag, val);
      else if (flag == "-g") a.g = need_i32(flag, val);
      else if (flag == "-n") a.n = need_i32(flag, val);
      else if (flag == "-m") a.m = need_i32(flag, val);
      else if (flag == "-M" || flag == "--mask-level") a.mask_level = need_f32(flag, val);
      else if (flag == "-p" || flag == "--pri-ratio") a.pri_ratio = need_f32(flag, val);
      else if (flag == "-N" || flag == "--best-n") a.best_n = need_i32(flag, val);
      else if (flag == "-x") { a.preset = val; a.has_preset = true; }
      else if (flag == "-o") { a.output = val; a.has_output = true; }
      else if (flag == "--gpus") a.gpus = need_i32(flag, val);
      else if (flag == "--batch-mb") a.batch_mb = need_i32(flag, val);
    } else if (flag == "-H" || flag == "--hpc") a.hpc = true;
    else if (flag == "-a") { /* parsed and ignored (main.rs:85-86) */ }
    else if (flag == "--all-reads") a.all_reads = true;
    else if (flag == "--khash-order") a.khash = true;
    else { fprintf(stderr, "error: unexpected argument '%s' found\n", t.c_str()); exit(2); }
  }
  return a;
}

void apply_preset(const std::string& p, int* w, int* k) {  // main.rs:125-133
  if (p == "map-ont") { *k = 15; *w = 10; }
  else if (p == "map-hifi" || p == "lr:hq") { *k = 19; *w = 10; }
  else if (p == "sr") { *k = 21; *w = 11; }
}

struct Session {
  mm2_ctx_t* ctx = nullptr;
  mm2_index_t* idx = nullptr;
  ~Session() { if (idx) mm2_index_free(idx); if (ctx) mm2_ctx_destroy(ctx); }
};

// anchors of the first query record exactly as main.rs:160-168 / :172-181 compute them
void first_read_anchors(Session& s, const Args& a, const Fasta& q, mm2_anchor_t** anchors, size_t* n_anchors) {
  const size_t qlen = q.offs.size() > 1 ? (size_t)(q.offs[1] - q.offs[0]) : 0;
  mm2_mini_t* mv = nullptr;
  size_t nm = 0;
  check(mm2_sketch(s.ctx, q.cat.data(), qlen, a.w, a.k, 0, 0, &mv, &nm));   // seeds.rs:7-11 (asserts on an empty read)
  check(mm2_filter_query_minimizers(s.ctx, mv, &nm, 10, 0.01f));
  int32_t mid_occ = 0;
  check(mm2_index_calc_mid_occ(s.idx, 2e-4f, &mid_occ));
  if (mid_occ < 10) mid_occ = 10;
  check(mm2_build_anchors_filtered(s.ctx, s.idx, mv, nm, (int32_t)qlen, mid_occ, anchors, n_anchors));
  mm2_free(mv);
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2 || !strcmp(argv[1], "-h") || !strcmp(argv[1], "--help")) { usage(); return argc < 2 ? 2 : 0; }
  const std::string cmd = argv[1];
  if (cmd != "index" && cmd != "anchors" && cmd != "chain" && cmd != "align") { usage(); return 2; }
  Args a = parse(cmd, argc - 2, argv + 2);
  const size_t need = cmd == "index" ? 1 : 2;
  if (a.pos.size() != need) { fprintf(stderr, "error: expected %zu positional argument(s)\n", need); usage(); return 2; }
  Session s;
  const char* dev = getenv("MM2RS_DEVICE");
  check(mm2_ctx_create(dev ? atoi(dev) : 0, &s.ctx));

  if (cmd == "index") {  // main.rs:150-159
    const int flag = a.hpc ? 1 : 0;
    std::vector<mm2_ctx_t*> extra_ctx;
    std::vector<mm2_index_t*> replicas;
    if (a.gpus > 1) {
      // extension: bucket-sharded build on N GPUs of this box (mm2_index_build_multi); the replica on the first GPU is dumped
      Fasta ref;
      if (!read_fasta(a.pos[0], false, ref)) die("cannot open " + a.pos[0]);
      std::vector<mm2_ctx_t*> ctxs{s.ctx};
      for (int d = 1; d < a.gpus; ++d) { mm2_ctx_t* c = nullptr; check(mm2_ctx_create((dev ? atoi(dev) : 0) + d, &c)); ctxs.push_back(c); extra_ctx.push_back(c); }
      std::vector<const char*> np;
      for (auto& nm : ref.names) np.push_back(nm.c_str());
      replicas.assign(ctxs.size(), nullptr);
      check(mm2_index_build_multi(ctxs.data(), (int)ctxs.size(), ref.cat.data(), ref.offs.data(), np.data(), ref.names.size(), a.w, a.k, a.b, flag,
                                  replicas.data()));
      s.idx = replicas[0];
    } else check(mm2_index_build_fasta(s.ctx, a.pos[0].c_str(), a.w, a.k, a.b, flag, &s.idx));
    uint64_t n_keys = 0, total_len = 0;
    double avg_occ = 0, avg_spacing = 0;
    uint32_t n_seq = 0;
    check(mm2_index_stats(s.idx, &n_keys, &avg_occ, &avg_spacing, &total_len));
    check(mm2_index_params(s.idx, nullptr, nullptr, nullptr, nullptr, &n_seq));
    printf("kmer size: %d; skip: %d; is_hpc: %d; #seq: %u\n", a.k, a.w, a.hpc ? 1 : 0, n_seq);
    printf("distinct minimizers: %llu (avg occ %.2f) avg spacing %.3f total length %llu\n", (unsigned long long)n_keys, avg_occ,
           avg_spacing, (unsigned long long)total_len);
    if (a.has_dump) {
      const std::string& p = a.dump;
      const bool mmi = p.size() >= 4 && p.compare(p.size() - 4, 4, ".mmi") == 0;
      check(mmi ? (a.khash ? mm2_index_save_mmi_khash(s.idx, p.c_str()) : mm2_index_save_mmi(s.idx, p.c_str())) : mm2_index_save_native(s.idx, p.c_str()));
    }
    for (size_t d = 1; d < replicas.size(); ++d) mm2_index_free(replicas[d]);
    for (mm2_ctx_t* c : extra_ctx) mm2_ctx_destroy(c);
    return 0;
  }

  if (cmd == "align" && a.has_preset) apply_preset(a.preset, &a.w, &a.k);  // main.rs:190
  const int flag = a.hpc ? 1 : 0;
  check(mm2_index_load_auto(s.ctx, a.pos[0].c_str(), a.w, a.k, 14, flag, &s.idx));  // main.rs:135-145, b = 14
  Fasta q;
  if (cmd != "align") {
    if (!read_fasta(a.pos[1], true, q)) die("cannot open " + a.pos[1]);
    if (q.names.empty()) { q.names.push_back("*"); q.offs.assign(2, 0); }  // main.rs:101: ("*", empty)
  }

  if (cmd == "anchors") {  // main.rs:160-171
    mm2_anchor_t* an = nullptr; size_t n = 0;
    first_read_anchors(s, a, q, &an, &n);
    printf("anchors: %zu\n", n);
    for (size_t i = 0; i < n && i < 10; ++i) printf("x=0x%016llx y=0x%016llx\n", (unsigned long long)an[i].x, (unsigned long long)an[i].y);
    mm2_free(an);
    return 0;
  }
  if (cmd == "chain") {  // main.rs:172-188
    mm2_anchor_t* an = nullptr; size_t n = 0;
    first_read_anchors(s, a, q, &an, &n);
    mm2_chain_params_t p;
    mm2_default_chain_params(a.k, &p);
    p.bw = a.bw;
    mm2_chains_t ch;
    check(mm2_chain_dp_all(s.ctx, an, n, &p, &ch));
    const size_t len = ch.n_chains ? (size_t)(ch.chain_offs[1] - ch.chain_offs[0]) : 0;
    printf("best_chain_len: %zu\n", len);
    if (len) {
      const mm2_anchor_t& st = an[ch.chain_idx[0]];
      const mm2_anchor_t& en = an[ch.chain_idx[len - 1]];
      printf("start: x=0x%016llx y=0x%016llx\n", (unsigned long long)st.x, (unsigned long long)st.y);
      printf("end:   x=0x%016llx y=0x%016llx\n", (unsigned long long)en.x, (unsigned long long)en.y);
    }
    mm2_chains_free(&ch);
    mm2_free(an);
    return 0;
  }

  // align: main.rs:189-230
  mm2_map_opts_t o;
  mm2_default_map_opts(&o);
  o.w = a.w; o.k = a.k; o.frac_top_repetitive = a.f; o.max_gap = a.g; o.min_cnt = a.n; o.min_chain_score = a.m;
  o.mask_level = a.mask_level; o.pri_ratio = a.pri_ratio; o.best_n = a.best_n;
  if (a.has_r && !a.r.empty()) {  // main.rs:202-208: unparsable parts are silently ignored
    const size_t c = a.r.find(',');
    int v;
    if (parse_i32(a.r.substr(0, c), &v)) o.bw = v;
    if (c != std::string::npos) {
      const std::string rest = a.r.substr(c + 1);
      if (parse_i32(rest.substr(0, rest.find(',')), &v)) o.bw_long = v;
    }
  }
  // Streaming ingest: a reader thread parses batch i + 1 into page-locked memory while the GPU maps batch i; the PAF lines are
  // written per batch, i.e. in input order, and memory stays bounded by two batches whatever the size of the file.
  ReadStream rs;
  if (!rs.open(a.pos[1])) die("cannot open " + a.pos[1]);
  const size_t max_bases = a.all_reads ? (size_t)std::max(1, a.batch_mb) << 20 : 0, max_reads = a.all_reads ? ~(size_t)0 : 1;
  Batch B[2];
  bool have = rs.fill(B[0], max_bases, max_reads);
  if (!have) { B[0].names.push_back("*"); B[0].offs.assign(2, 0); have = true; }  // main.rs:101: ("*", empty) -> the sketch asserts
  FILE* out = stdout;
  if (a.has_output && a.output != "-") {
    out = fopen(a.output.c_str(), "wb");
    if (!out) die("cannot create " + a.output);
  }
  size_t read_base = 0;
  for (int cur = 0; have; cur ^= 1) {
    Batch& b = B[cur];
    bool next_have = false;
    std::thread reader;
    if (a.all_reads) reader = std::thread([&]() { next_have = rs.fill(B[cur ^ 1], max_bases, max_reads); });
    const size_t nreads = b.names.size();
    int rc_exit = 0;
    std::string err;
    for (size_t i = 0; i < nreads && !rc_exit; ++i)
      if (b.offs[i + 1] == b.offs[i]) { rc_exit = 101; err = "empty query sequence (sketch.rs:30 asserts !seq.is_empty())"; }
    mm2_map_result_t res;
    memset(&res, 0, sizeof res);
    if (!rc_exit) {
      const int rc = mm2_map_batch(s.ctx, s.idx, b.cat, b.offs.data(), nreads, &o, &res);
      if (rc != MM2_OK) { rc_exit = (rc == MM2_E_ARG || rc == MM2_E_REF_PANIC) ? 101 : 1; err = mm2_last_error(); }
    }
    if (!rc_exit && res.n_panic) {
      char msg[256];
      snprintf(msg, sizeof msg, "thread 'main' panicked: index out of bounds (anchor rid beyond the sequence table; read %zu) — "
                                "the reference mis-encodes odd target ids (seeds.rs:64-71)", read_base + res.panic_reads[0]);
      err = msg; rc_exit = 101;
    }
    if (!rc_exit) {
      std::vector<const char*> qn;
      for (auto& nm : b.names) qn.push_back(nm.c_str());
      char* txt = nullptr; size_t tl = 0;
      const int rc = mm2_paf_format_batch(s.idx, &res, qn.data(), &txt, &tl);
      if (rc != MM2_OK) { rc_exit = 1; err = mm2_last_error(); }
      else { fwrite(txt, 1, tl, out); mm2_free(txt); }
    }
    mm2_map_result_free(&res);
    if (reader.joinable()) reader.join();
    if (rc_exit) {
      if (out != stdout) fclose(out);
      if (rc_exit == 101 && err.rfind("thread", 0) == 0) { fprintf(stderr, "%s\n", err.c_str()); return 101; }
      die(err, rc_exit);
    }
    read_base += nreads;
    have = next_have;
  }
  if (out != stdout) fclose(out);
  return 0;
}
