// gen.cpp — deterministic synthetic genomes and reads.  Shared by tests and bench so the
// oracle and the GPU path see identical bytes.  xoshiro256** seeded through splitmix64.
// Build: g++ -O2 -std=c++17 -fPIC -shared -pthread -o libmm2gen.so gen.cpp
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <thread>
#include <vector>
#include <algorithm>

namespace {
struct Rng {
  uint64_t s[4];
  static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
  }
  explicit Rng(uint64_t seed) { for (int i = 0; i < 4; ++i) s[i] = splitmix(seed); }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  uint64_t below(uint64_t n) { return (uint64_t)(((__uint128_t)next() * n) >> 64); }
};
const char ACGT[4] = {'A', 'C', 'G', 'T'};
inline char comp(char c) {
  switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; }
}
}  // namespace

extern "C" {

// iid ACGT; n_run_rate = probability per base of starting an N run (geometric length, mean n_run_mean)
void mm2gen_genome(uint64_t seed, uint8_t* out, size_t len, double n_run_rate, double n_run_mean) {
  const size_t CH = 1 << 22;
  size_t nch = (len + CH - 1) / CH;
  unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t)
    th.emplace_back([=]() {
      for (size_t c = t; c < nch; c += nt) {
        Rng r(seed * 0x100000001b3ULL + c);
        size_t s = c * CH, e = std::min(len, s + CH);
        for (size_t i = s; i < e;) {
          uint64_t x = r.next();
          for (int j = 0; j < 32 && i < e; ++j, ++i) out[i] = ACGT[(x >> (2 * j)) & 3];
        }
        if (n_run_rate > 0) {
          size_t i = s;
          while (i < e) {
            // distance to next run start ~ geometric(n_run_rate)
            double u = r.uni();
            size_t gap = (size_t)(-std::max(1.0, 1.0 / n_run_rate) * __builtin_log(1.0 - u));
            i += gap;
            if (i >= e) break;
            size_t rl = 1 + (size_t)(-n_run_mean * __builtin_log(1.0 - r.uni()));
            for (size_t j = 0; j < rl && i < e; ++j, ++i) out[i] = 'N';
          }
        }
      }
    });
  for (auto& x : th) x.join();
}

// Repeat-rich genome (C5): start from iid ACGT, then overwrite with tandem arrays and dispersed repeat families.
void mm2gen_repeat_genome(uint64_t seed, uint8_t* out, size_t len, double tandem_frac, double dispersed_frac) {
  mm2gen_genome(seed, out, len, 0.0, 0.0);
  Rng r(seed ^ 0xabcdef12345ULL);
  // tandem arrays: unit 150-500 bp, 200-2000 copies, 1-3 % per-copy divergence
  size_t tandem_target = (size_t)(tandem_frac * (double)len), done = 0;
  while (done < tandem_target) {
    size_t unit = 150 + r.below(351), copies = 200 + r.below(1801);
    size_t tot = unit * copies;
    if (tot + 1 >= len) { copies = std::max<size_t>(2, len / (4 * unit)); tot = unit * copies; }
    size_t pos = r.below(len - tot);
    double div = 0.01 + 0.02 * r.uni();
    std::vector<uint8_t> u(unit);
    for (auto& c : u) c = ACGT[r.below(4)];
    for (size_t c = 0; c < copies; ++c)
      for (size_t j = 0; j < unit; ++j) out[pos + c * unit + j] = r.uni() < div ? ACGT[r.below(4)] : u[j];
    done += tot;
  }
  // dispersed: 64 families of 6 kb, copies at 5 % divergence
  size_t disp_target = (size_t)(dispersed_frac * (double)len);
  const size_t FL = 6000, NF = 64;
  if (len > 4 * FL && disp_target > 0) {
    size_t copies_each = std::max<size_t>(1, disp_target / (FL * NF));
    for (size_t f = 0; f < NF; ++f) {
      std::vector<uint8_t> fam(FL);
      for (auto& c : fam) c = ACGT[r.below(4)];
      for (size_t c = 0; c < copies_each; ++c) {
        size_t pos = r.below(len - FL);
        for (size_t j = 0; j < FL; ++j) out[pos + j] = r.uni() < 0.05 ? ACGT[r.below(4)] : fam[j];
      }
    }
  }
}

// Reads of exactly read_len bases drawn from sequences [seq_offs[i], seq_offs[i+1]) of `genome` (only sequences with
// length >= 4*read_len are used), 50 % reverse-complemented, iid substitution/insertion/deletion errors.
// out must hold nreads*read_len bytes.  src_seq/src_pos/src_rev (nreads each, may be NULL) record the truth.
// `first`: index of the first read of this call within the read set (read i depends only on (seed, i)), so that ranks can
// generate disjoint shards of one read set.
void mm2gen_reads_from(uint64_t seed, const uint8_t* genome, const uint64_t* seq_offs, size_t nseq, size_t first, size_t nreads, size_t read_len,
                  double p_sub, double p_ins, double p_del, uint8_t* out, uint32_t* src_seq, uint64_t* src_pos, uint8_t* src_rev) {
  std::vector<size_t> ok;
  std::vector<uint64_t> cum;
  uint64_t tot = 0;
  for (size_t i = 0; i < nseq; ++i) {
    uint64_t l = seq_offs[i + 1] - seq_offs[i];
    if (l >= 4 * read_len) { ok.push_back(i); tot += l - 3 * read_len; cum.push_back(tot); }
  }
  if (ok.empty()) { memset(out, 'A', nreads * read_len); return; }
  unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t)
    th.emplace_back([&, t]() {
      std::vector<uint8_t> tmp(read_len);
      for (size_t rd = t; rd < nreads; rd += nt) {
        Rng r(seed * 0x9e3779b97f4a7c15ULL + (first + rd) * 2 + 1);
        uint64_t x = r.below(tot);
        size_t si = std::upper_bound(cum.begin(), cum.end(), x) - cum.begin();
        uint64_t base = si ? cum[si - 1] : 0;
        size_t s = ok[si];
        uint64_t pos = x - base;  // start within sequence s; 3*read_len of slack to the end
        const uint8_t* src = genome + seq_offs[s] + pos;
        size_t o = 0, i = 0;
        while (o < read_len) {
          double u = r.uni();
          if (u < p_del) { i++; continue; }
          if (u < p_del + p_ins) { tmp[o++] = ACGT[r.below(4)]; continue; }
          uint8_t c = src[i++];
          if (u < p_del + p_ins + p_sub) { uint8_t d; do { d = ACGT[r.below(4)]; } while (d == c); c = d; }
          tmp[o++] = c;
        }
        bool rev = (r.next() & 1) != 0;
        uint8_t* dst = out + rd * read_len;
        if (rev) for (size_t j = 0; j < read_len; ++j) dst[j] = comp((char)tmp[read_len - 1 - j]);
        else memcpy(dst, tmp.data(), read_len);
        if (src_seq) src_seq[rd] = (uint32_t)s;
        if (src_pos) src_pos[rd] = pos;
        if (src_rev) src_rev[rd] = rev ? 1 : 0;
      }
    });
  for (auto& x : th) x.join();
}

void mm2gen_reads(uint64_t seed, const uint8_t* genome, const uint64_t* seq_offs, size_t nseq, size_t nreads, size_t read_len,
                  double p_sub, double p_ins, double p_del, uint8_t* out, uint32_t* src_seq, uint64_t* src_pos, uint8_t* src_rev) {
  mm2gen_reads_from(seed, genome, seq_offs, nseq, 0, nreads, read_len, p_sub, p_ins, p_del, out, src_seq, src_pos, src_rev);
}

}  // extern "C"
