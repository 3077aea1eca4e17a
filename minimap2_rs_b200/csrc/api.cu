// api.cu — C ABI of libmm2b200.so: context, sketch / seeds / chain stage entry points, and the batched mapping
// pipeline that `mm2rs align` drives (main.rs:193-219 for every read of a batch).  Host code here only moves data,
// sequences kernels and finishes the handful of scalar float operations that must match glibc (powf for dv).
#include "mm2_internal.cuh"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <mutex>
#include <atomic>
#include <chrono>
#include <malloc.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <cctype>
#include <thread>

#include "stages.cuh"

// ---- errors ----------------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void mm2_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
extern "C" const char* mm2_last_error(void) { return g_err; }
extern "C" void mm2_free(void* p) { free(p); }
extern "C" void* mm2_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { mm2_set_error("cudaMallocHost(%zu) failed", bytes); return nullptr; }
  return p;
}
extern "C" void mm2_host_free(void* p) { if (p) cudaFreeHost(p); }

// Page-locked memory on the NUMA node the GPU hangs off.  With one process per GPU on a two-socket box, half of the ranks
// otherwise stream their reads across the socket interconnect (round 1: 21 GB/s of upload per GPU with eight ranks copying
// at once, against 55 GB/s for one).  The node comes from sysfs (PCI bus id of the device); the allocation runs under a
// temporary MPOL_BIND policy (raw syscall: no libnuma in the image).  Any failure falls back to the default placement.
static int gpu_numa_node(int device) {
  char bus[32] = "";
  if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return -1; }
  for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
  char path[128];
  snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE* f = fopen(path, "r");
  if (!f) return -1;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  return node;
}
extern "C" void* mm2_host_alloc_on(int device, size_t bytes) {
  const int node = gpu_numa_node(device);
  bool bound = false;
  if (node >= 0 && node < 1024 && !getenv("MM2_NO_NUMA")) {
    unsigned long mask[16] = {0};
    mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
    bound = syscall(SYS_set_mempolicy, 2 /* MPOL_BIND */, mask, (unsigned long)(sizeof mask * 8)) == 0;
  }
  void* p = mm2_host_alloc(bytes);
  if (bound) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0ul);
  return p;
}

// ---- context ---------------------------------------------------------------------------------------------------------
extern "C" int mm2_ctx_create(int device, mm2_ctx_t** out) {
  if (!out) { mm2_set_error("mm2_ctx_create: NULL out"); return MM2_E_ARG; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    mm2_set_error("no CUDA device available (%s); libmm2b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count=0");
    return MM2_E_CUDA;
  }
  if (device < 0 || device >= ndev) { mm2_set_error("device %d out of range (%d devices)", device, ndev); return MM2_E_ARG; }
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) { mm2_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return MM2_E_CUDA; }
  {  // keep freed index pages in the default memory pool (see DevBuf::pooled)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long thr = ~0ULL;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  MM2_TRY(lchain_init_device()); MM2_TRY(radix_init_device()); MM2_TRY(seeds_init_device());
  mm2_ctx* c = new mm2_ctx();
  c->device = device;
  e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete c; mm2_set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); return MM2_E_CUDA; }
  c->own_stream = true;
  { const char* e = getenv("MM2_PIPELINE"); if (e && atoi(e) == 0) c->pipeline = false; }
  // The three waits of a mapping call spin (lowest latency).  MM2_SYNC=block makes them sleep on a blocking event; measured on
  // 8 ranks x 5 threads / 32 cores it does not help the end-to-end time (bound by the host-to-GPU links) and costs the
  // device-resident path 60 % (wake-up latency between dependent launches), so it stays opt-in.
  { const char* e = getenv("MM2_SYNC"); if (e && !strcmp(e, "block")) c->block_sync = true; }
  { const char* e = getenv("MM2_WORKERS"); if (e && atoi(e) >= 2 && atoi(e) <= 4) c->n_workers = atoi(e); }
  { const char* e = getenv("MM2_SUBBATCH_MB"); if (e && atoi(e) > 0) c->subbatch_bytes = (u64)atoi(e) << 20; }
  // test hook: MM2_CHAIN_DENSE_MIN=n sends every read with >= n anchors to the CTA-per-read chaining kernel
  { const char* e = getenv("MM2_CHAIN_DENSE_MIN"); if (e && atoi(e) > 0) { c->chain_dense_min = atoi(e); c->chain_dense_ratio5 = 0; } }
  c->n_sm = prop.multiProcessorCount;
  *out = c;
  return MM2_OK;
}

extern "C" void mm2_ctx_destroy(mm2_ctx_t* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  DevBuf* bufs[] = {&c->seq, &c->seq_off, &c->tile_seq, &c->tile_first, &c->tile_status, &c->misc, &c->mkey, &c->mval,
                    &c->mini_off, &c->keep, &c->occ_cnt, &c->occ_loc, &c->anchor_off_m, &c->scan_status, &c->anchors,
                    &c->read_aoff, &c->read_class, &c->read_flag, &c->read_nhit, &c->read_na, &c->flag_list, &c->dpA, &c->dpB, &c->dpT, &c->dpW, &c->hits, &c->chain_idx, &c->lut, &c->sort_tmp,
                    &c->sort_tmp2, &c->sort_keys2, &c->sort_vals2, &c->runidx, &c->run_start, &c->run_gp, &c->rs_counts, &c->rs_offs, &c->diag, &c->mg_recv_k, &c->mg_recv_v, &c->mg_stage, &c->fine_tmp, &c->packed, &c->packed_n};
  for (DevBuf* b : bufs) b->release();
  c->pin_in.release(); c->pin_out.release(); c->pin_small.release(); c->pin_scalar.release();
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  for (int w = 0; w < 4; ++w) if (c->worker[w]) mm2_ctx_destroy(c->worker[w]);
  if (c->sync_event) { cudaEventDestroy(c->sync_event); c->sync_event = nullptr; }
  for (cudaEvent_t e : c->copy_events) cudaEventDestroy(e);
  c->copy_events.clear();
  if (c->copy_stream) { cudaStreamDestroy(c->copy_stream); c->copy_stream = nullptr; }
  delete c;
}

extern "C" int mm2_ctx_set_stream(mm2_ctx_t* c, void* cuda_stream) {
  if (!c) { mm2_set_error("NULL ctx"); return MM2_E_ARG; }
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (cuda_stream) {
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)cuda_stream; c->own_stream = false;
  } else if (!c->own_stream) {
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  return MM2_OK;
}
extern "C" int mm2_ctx_synchronize(mm2_ctx_t* c) {
  if (!c) { mm2_set_error("NULL ctx"); return MM2_E_ARG; }
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return MM2_OK;
}
extern "C" uint64_t mm2_ctx_launch_count(const mm2_ctx_t* c) { return c ? c->launches : 0; }
extern "C" int mm2_ctx_count_cells(mm2_ctx_t* c, int on) {
  if (!c) { mm2_set_error("NULL ctx"); return MM2_E_ARG; }
  c->count_cells = on != 0;
  for (int w = 0; w < 4; ++w) if (c->worker[w]) c->worker[w]->count_cells = c->count_cells;
  return MM2_OK;
}
extern "C" uint64_t mm2_ctx_last_cells(const mm2_ctx_t* c) { return c ? c->last_cells : 0; }
extern "C" int mm2_ctx_last_timings(const mm2_ctx_t* c, const char** names, const float** ms, int* n) {
  if (!c || !names || !ms || !n) { mm2_set_error("NULL argument"); return MM2_E_ARG; }
  *names = c->timer.names_blob.c_str(); *ms = c->timer.ms.data(); *n = (int)c->timer.ms.size();
  return MM2_OK;
}

// ---- small kernels of the API layer --------------------------------------------------------------------------------------
namespace {
__global__ void interleave_kernel(const u64* __restrict__ a, const u64* __restrict__ b, ulonglong2* __restrict__ out, u64 n) {
  for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) out[i] = make_ulonglong2(a[i], b[i]);
}
__global__ void deinterleave_kernel(const ulonglong2* __restrict__ in, u64* __restrict__ a, u64* __restrict__ b, u64 n) {
  for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) { const ulonglong2 v = in[i]; a[i] = v.x; b[i] = v.y; }
}
__global__ void dp_unpack_kernel(const int4* __restrict__ A, int* __restrict__ f, int* __restrict__ v, int* __restrict__ pprev, u64 n) {
  for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) { const int4 a = A[i]; f[i] = a.x; pprev[i] = a.y; v[i] = a.z; }
}
inline int grid_for(u64 n, int block = 256) { return (int)std::max<u64>(1, std::min<u64>((n + block - 1) / block, 148ull * 32)); }

template <class T>
T* xmalloc(size_t n) { return (T*)malloc(std::max<size_t>(1, n) * sizeof(T)); }

// The record array of a large batch is megabytes of freshly mapped pages on every call (a page fault per 4 KB, ~2 ms per
// 100k reads).  mm2_map_result_free parks up to four such buffers here and the next batch takes one back.
struct RecCache {
  std::mutex mu;
  struct Slot { void* p; size_t cap; } slot[4] = {{nullptr, 0}, {nullptr, 0}, {nullptr, 0}, {nullptr, 0}};
  static constexpr size_t MIN_BYTES = 1u << 20;
  void* take(size_t bytes) {
    if (bytes >= MIN_BYTES) {
      std::lock_guard<std::mutex> lk(mu);
      for (auto& s : slot)
        if (s.p && s.cap >= bytes && s.cap <= 2 * bytes) { void* p = s.p; s.p = nullptr; s.cap = 0; return p; }
    }
    return malloc(std::max<size_t>(1, bytes));
  }
  void give(void* p) {
    if (!p) return;
    const size_t cap = malloc_usable_size(p);
    if (cap >= MIN_BYTES) {
      std::lock_guard<std::mutex> lk(mu);
      for (auto& s : slot)
        if (!s.p) { s.p = p; s.cap = cap; return; }
    }
    free(p);
  }
};
RecCache g_rec_cache;
}  // namespace

// ---- sketch ------------------------------------------------------------------------------------------------------------
extern "C" int mm2_sketch_batch(mm2_ctx_t* ctx, const uint8_t* cat, const uint64_t* offs, size_t nseq, int w, int k, uint32_t rid_base,
                                uint32_t rid_step, int is_hpc, mm2_mini_t** out, uint64_t** out_offs) {
  if (!ctx || !offs || !out || !out_offs || (!cat && nseq && offs[nseq] > offs[0])) { mm2_set_error("mm2_sketch_batch: NULL argument"); return MM2_E_ARG; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  *out = nullptr; *out_offs = nullptr;
  const u64 base = nseq ? offs[0] : 0, total = nseq ? offs[nseq] - base : 0;
  std::vector<u64> off0(nseq + 1, 0);
  for (size_t i = 0; i <= nseq && nseq; ++i) off0[i] = offs[i] - base;
  MM2_TRY(ctx->seq.ensure(total + 64));
  MM2_TRY(ctx->seq_off.ensure((nseq + 1) * 8));
  if (total) CUDA_TRY(cudaMemcpyAsync(ctx->seq.p, cat + base, total, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(ctx->seq_off.p, off0.data(), (nseq + 1) * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  SketchOut so;
  MM2_TRY(sketch_device(ctx, ctx->seq.as<u8>(), ctx->seq_off.as<u64>(), off0.data(), nseq, w, k, rid_base, rid_step, is_hpc, &so));
  mm2_mini_t* h = xmalloc<mm2_mini_t>(so.total);
  u64* ho = xmalloc<u64>(nseq + 1);
  if (!h || !ho) { free(h); free(ho); mm2_set_error("out of host memory"); return MM2_E_OOM; }
  ho[0] = 0;
  if (nseq) {
    MM2_TRY(ctx->anchors.ensure(std::max<u64>(1, so.total) * 16));
    if (so.total) MM2_LAUNCH(ctx, interleave_kernel, grid_for(so.total), 256, 0, so.key, so.val, ctx->anchors.as<ulonglong2>(), so.total);
    if (so.total) CUDA_TRY(cudaMemcpyAsync(h, ctx->anchors.p, so.total * 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ho, so.seq_off, (nseq + 1) * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  *out = h; *out_offs = ho;
  return MM2_OK;
}

extern "C" int mm2_sketch(mm2_ctx_t* ctx, const uint8_t* seq, size_t len, int w, int k, uint32_t rid, int is_hpc, mm2_mini_t** out, size_t* n) {
  if (!n || !out) { mm2_set_error("mm2_sketch: NULL argument"); return MM2_E_ARG; }
  if (len == 0) { mm2_set_error("mm2_sketch: empty sequence (sketch.rs:30 asserts !seq.is_empty())"); return MM2_E_ARG; }
  const u64 offs[2] = {0, (u64)len};
  u64* oo = nullptr;
  MM2_TRY(mm2_sketch_batch(ctx, seq, offs, 1, w, k, rid, 0, is_hpc, out, &oo));
  *n = (size_t)oo[1];
  free(oo);
  return MM2_OK;
}

// ---- index build entry points (device work in index.cu; FASTA parsing in index_io.cu) ----------------------------------
extern "C" int mm2_index_build_seqs(mm2_ctx_t* ctx, const uint8_t* cat, const uint64_t* offs, const char* const* names, size_t nseq,
                                    int w, int k, int b, int flag, mm2_index_t** out) {
  if (!ctx || !out || (nseq && (!cat || !offs))) { mm2_set_error("mm2_index_build_seqs: NULL argument"); return MM2_E_ARG; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  static const u64 zero_off[1] = {0};
  return index_build_device(ctx, cat, nseq ? offs : zero_off, names, nseq, w, k, b, flag, out);
}

// ---- seeds stage entry points ---------------------------------------------------------------------------------------------
static int upload_minis(mm2_ctx* ctx, const mm2_mini_t* mv, size_t n) {
  MM2_TRY(ctx->mkey.ensure(std::max<size_t>(1, n) * 8));
  MM2_TRY(ctx->mval.ensure(std::max<size_t>(1, n) * 8));
  MM2_TRY(ctx->anchors.ensure(std::max<size_t>(1, n) * 16));
  if (n) {
    CUDA_TRY(cudaMemcpyAsync(ctx->anchors.p, mv, n * 16, cudaMemcpyHostToDevice, ctx->stream));
    MM2_LAUNCH(ctx, deinterleave_kernel, grid_for(n), 256, 0, ctx->anchors.as<ulonglong2>(), ctx->mkey.as<u64>(), ctx->mval.as<u64>(), (u64)n);
  }
  return MM2_OK;
}

extern "C" int mm2_filter_query_minimizers(mm2_ctx_t* ctx, mm2_mini_t* mv, size_t* n, int32_t q_occ_max, float q_occ_frac) {
  if (!ctx || !n || (*n && !mv)) { mm2_set_error("mm2_filter_query_minimizers: NULL argument"); return MM2_E_ARG; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t cnt = *n;
  if (cnt == 0) return MM2_OK;
  MM2_TRY(upload_minis(ctx, mv, cnt));
  MM2_TRY(ctx->mini_off.ensure(16));
  MM2_TRY(ctx->keep.ensure(cnt + 16));
  MM2_TRY(ctx->misc.ensure(64));
  const u64 mo[2] = {0, (u64)cnt};
  CUDA_TRY(cudaMemcpyAsync(ctx->mini_off.p, mo, 16, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  MM2_TRY(seeds_filter(ctx, ctx->mkey.as<u64>(), ctx->mini_off.as<u64>(), 1, cnt, q_occ_max, q_occ_frac, ctx->keep.as<u8>(), ctx->misc.as<u32>()));
  std::vector<u8> keep(cnt);
  CUDA_TRY(cudaMemcpyAsync(keep.data(), ctx->keep.p, cnt, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  size_t j = 0;
  for (size_t i = 0; i < cnt; ++i) if (keep[i]) mv[j++] = mv[i];  // seeds.rs:33-35: survivors keep their order
  *n = j;
  return MM2_OK;
}

extern "C" int mm2_build_anchors_filtered(mm2_ctx_t* ctx, const mm2_index_t* idx, const mm2_mini_t* mv, size_t n, int32_t qlen,
                                          int32_t mid_occ, mm2_anchor_t** out, size_t* n_out) {
  if (!ctx || !idx || !out || !n_out || (n && !mv)) { mm2_set_error("mm2_build_anchors_filtered: NULL argument"); return MM2_E_ARG; }
  if (idx->device != ctx->device) { mm2_set_error("index lives on device %d, context on %d", idx->device, ctx->device); return MM2_E_ARG; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  *out = nullptr; *n_out = 0;
  MM2_TRY(upload_minis(ctx, mv, n));
  MM2_TRY(ctx->mini_off.ensure(16));
  MM2_TRY(ctx->seq_off.ensure(16));
  MM2_TRY(ctx->misc.ensure(64));
  const u64 mo[2] = {0, (u64)n}, ro[2] = {0, (u64)(u32)qlen};
  CUDA_TRY(cudaMemcpyAsync(ctx->mini_off.p, mo, 16, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(ctx->seq_off.p, ro, 16, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const IndexView V = idx->view();
  u64 na = 0;
  // the minimizers are already filtered by the caller (seeds.rs:42-45): q_occ_frac = 0 switches the filter off
  MM2_TRY(seeds_hits(ctx, V, ctx->mkey.as<u64>(), ctx->mval.as<u64>(), ctx->mini_off.as<u64>(), 1, n, 0, 0.0f, mid_occ, false,
                     ctx->misc.as<u32>(), &na));
  MM2_TRY(ctx->anchors.ensure(std::max<u64>(1, na) * 16));
  // the deinterleave above used ctx->anchors as staging; it has completed (stream order), safe to reuse/grow
  MM2_TRY(seeds_fill_and_sort(ctx, V, ctx->mini_off.as<u64>(), ctx->seq_off.as<u64>(), 1, ctx->anchors.as<ulonglong2>()));
  mm2_anchor_t* h = xmalloc<mm2_anchor_t>(na);
  if (!h) { mm2_set_error("out of host memory"); return MM2_E_OOM; }
  if (na) CUDA_TRY(cudaMemcpyAsync(h, ctx->anchors.p, na * 16, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  *out = h; *n_out = (size_t)na;
  return MM2_OK;
}

// ---- chaining stage entry point ---------------------------------------------------------------------------------------------
extern "C" void mm2_default_chain_params(int32_t k, mm2_chain_params_t* p) {  // main.rs:105-123
  const float chain_gap_scale = 0.8f;
  volatile float t = 0.01f * chain_gap_scale;
  volatile float g = t * (float)k;
  p->max_dist_x = 5000; p->max_dist_y = 5000; p->bw = 500; p->max_chain_iter = 5000; p->min_chain_score = 40; p->min_cnt = 3;
  p->chn_pen_gap = g; p->chn_pen_skip = 0.0f; p->max_chain_skip = 25; p->max_drop = 500; p->bw_long = 20000;
  p->rmq_rescue_size = 1000; p->rmq_rescue_ratio = 0.1f;
}

namespace {
inline i32 a_qpos(const mm2_anchor_t& a) { return (i32)(u32)a.y; }
inline i32 a_qspan(const mm2_anchor_t& a) { return (i32)((a.y >> 32) & 0xff); }
inline i32 a_rpos(const mm2_anchor_t& a) { return (i32)(u32)a.x; }
struct HostChain { std::vector<u64> idx; i32 score; i32 qstart, tstart; };
// start coordinates used as tie-breakers by sort_chains_stable (lchain.rs:179-200; an empty chain gives i32::MAX)
void chain_starts(const mm2_anchor_t* a, HostChain& c) {
  i32 qs = INT32_MAX, ts = INT32_MAX;
  for (u64 i : c.idx) {
    qs = std::min(qs, (i32)((u32)a_qpos(a[i]) - (u32)(a_qspan(a[i]) - 1)));
    ts = std::min(ts, (i32)((u32)a_rpos(a[i]) - (u32)(a_qspan(a[i]) - 1)));
  }
  c.qstart = std::max(qs, 0); c.tstart = std::max(ts, 0);
}
}  // namespace

extern "C" int mm2_chain_dp_all(mm2_ctx_t* ctx, const mm2_anchor_t* a, size_t n, const mm2_chain_params_t* p, mm2_chains_t* out) {
  if (!ctx || !p || !out || (n && !a)) { mm2_set_error("mm2_chain_dp_all: NULL argument"); return MM2_E_ARG; }
  memset(out, 0, sizeof *out);
  out->chain_offs = xmalloc<u64>(1); out->chain_offs[0] = 0;
  if (n == 0) return MM2_OK;  // lchain.rs:61
  if (n > 0x7fffffffull) { mm2_set_error("too many anchors"); return MM2_E_ARG; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  MM2_TRY(ctx->anchors.ensure(n * 16));
  MM2_TRY(ctx->dpA.ensure(n * 16)); MM2_TRY(ctx->dpB.ensure(n * 16)); MM2_TRY(ctx->dpT.ensure(n * 4)); MM2_TRY(ctx->dpW.ensure(n * 4));
  MM2_TRY(ctx->chain_idx.ensure(n * 4 * 4));
  MM2_TRY(ctx->read_aoff.ensure(16)); MM2_TRY(ctx->seq_off.ensure(16)); MM2_TRY(ctx->mini_off.ensure(16));
  MM2_TRY(ctx->hits.ensure(sizeof(ReadHit) + 16)); MM2_TRY(ctx->misc.ensure(64)); MM2_TRY(ctx->mval.ensure(16));
  const u64 ao[2] = {0, (u64)n}, zo[2] = {0, 0};
  CUDA_TRY(cudaMemcpyAsync(ctx->anchors.p, a, n * 16, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(ctx->read_aoff.p, ao, 16, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(ctx->seq_off.p, zo, 16, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(ctx->mini_off.p, zo, 16, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(ctx->misc.p, 0, 64, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  MM2_TRY(chain_batch(ctx, ctx->anchors.as<ulonglong2>(), ctx->read_aoff.as<u64>(), ctx->seq_off.as<u64>(), ctx->mini_off.as<u64>(),
                      ctx->mval.as<u64>(), ctx->misc.as<u32>(), 1, *p, 0, ctx->dpA.as<int4>(), ctx->dpB.as<int4>(), ctx->dpT.as<int>(),
                      ctx->dpW.as<int>(), nullptr, ctx->hits.as<ReadHit>(), nullptr));
  int* d_f = ctx->chain_idx.as<int>();
  MM2_LAUNCH(ctx, dp_unpack_kernel, grid_for(n), 256, 0, ctx->dpA.as<int4>(), d_f, d_f + n, d_f + 2 * n, (u64)n);
  std::vector<i32> f(n), v(n), pp(n);
  CUDA_TRY(cudaMemcpyAsync(f.data(), d_f, n * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(v.data(), d_f + n, n * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(pp.data(), d_f + 2 * n, n * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  out->f = xmalloc<i32>(n); out->v = xmalloc<i32>(n); out->pprev = xmalloc<i64>(n);
  for (size_t i = 0; i < n; ++i) { out->f[i] = f[i]; out->v[i] = v[i]; out->pprev[i] = pp[i]; }
  // Backtracking (lchain.rs:93-160).  The helper that walks back from a chain end sets t[i] = 2 and then tests
  // t[i] == 0 on the same i, so it never takes a second step: every candidate chain is {i0} with score
  // f[i0] - f[pprev[i0]] (or is empty when that difference is <= 0), visited in order of (f desc, index desc).
  std::vector<u64> order;
  for (size_t i = 0; i < n; ++i) if (f[i] > 0) order.push_back(i);
  std::vector<HostChain> chains;
  if (!order.empty()) {
    std::stable_sort(order.begin(), order.end(), [&](u64 x, u64 y) { return f[x] < f[y]; });  // F10: ties keep index order
    for (size_t q = order.size(); q-- > 0;) {
      const u64 i0 = order[q];
      const i32 prev = pp[i0];
      const i32 s = prev < 0 ? f[i0] : (i32)((u32)f[i0] - (u32)f[prev]);
      HostChain c;
      if (s > 0) { c.idx.push_back(i0); c.score = s; } else c.score = 0;
      if (c.score >= p->min_chain_score && (i32)c.idx.size() >= p->min_cnt) chains.push_back(std::move(c));
    }
    if (chains.empty()) {  // lchain.rs:162-173 fallback: last maximum of f, full walk, score v[best]
      size_t best = 0;
      for (size_t i = 1; i < n; ++i) if (f[i] >= f[best]) best = i;
      HostChain c;
      for (i64 i = (i64)best; i >= 0; i = pp[(size_t)i]) c.idx.push_back((u64)i);
      std::reverse(c.idx.begin(), c.idx.end());
      c.score = v[best];
      chains.push_back(std::move(c));
    }
    for (auto& c : chains) chain_starts(a, c);
    std::stable_sort(chains.begin(), chains.end(), [](const HostChain& x, const HostChain& y) {  // lchain.rs:202-218
      if (x.score != y.score) return y.score < x.score;
      if (x.qstart != y.qstart) return x.qstart < y.qstart;
      return x.tstart < y.tstart;
    });
  }
  size_t tot = 0;
  for (auto& c : chains) tot += c.idx.size();
  free(out->chain_offs);
  out->n_chains = chains.size();
  out->chain_offs = xmalloc<u64>(chains.size() + 1);
  out->chain_idx = xmalloc<u64>(tot);
  out->scores = xmalloc<i32>(chains.size());
  size_t o = 0;
  for (size_t i = 0; i < chains.size(); ++i) {
    out->chain_offs[i] = o;
    for (u64 x : chains[i].idx) out->chain_idx[o++] = x;
    out->scores[i] = chains[i].score;
  }
  out->chain_offs[chains.size()] = o;
  return MM2_OK;
}

extern "C" void mm2_chains_free(mm2_chains_t* c) {
  if (!c) return;
  free(c->chain_offs); free(c->chain_idx); free(c->scores); free(c->f); free(c->v); free(c->pprev);
  memset(c, 0, sizeof *c);
}

// ---- batched mapping ------------------------------------------------------------------------------------------------------
extern "C" void mm2_default_map_opts(mm2_map_opts_t* o) {  // main.rs:55-89 defaults
  o->w = 10; o->k = 15; o->frac_top_repetitive = 2e-4f; o->max_gap = 5000; o->bw = -1; o->bw_long = -1; o->min_cnt = 3;
  o->min_chain_score = 40; o->mask_level = 0.5f; o->pri_ratio = 0.8f; o->best_n = 5; o->q_occ_max = 10; o->q_occ_frac = 0.01f;
  o->mid_occ_floor = 10; o->want_stage_dump = 0;
}

// concatenate the results of consecutive read ranges (first[i] = index of the first read of part i) in read order
static void merge_map_results(mm2_map_result_t* parts, const size_t* first, size_t nparts, mm2_map_result_t* out) {
  memset(out, 0, sizeof *out);
  size_t nrec = 0, npan = 0;
  for (size_t i = 0; i < nparts; ++i) { nrec += parts[i].n_recs; npan += parts[i].n_panic; }
  out->recs = (mm2_paf_rec_t*)g_rec_cache.take(std::max<size_t>(1, nrec) * sizeof(mm2_paf_rec_t));
  out->panic_reads = xmalloc<u32>(npan);
  for (size_t i = 0; i < nparts; ++i) {
    mm2_map_result_t& p = parts[i];
    for (size_t j = 0; j < p.n_recs; ++j) { mm2_paf_rec_t r = p.recs[j]; r.read_id += (u32)first[i]; out->recs[out->n_recs++] = r; }
    for (size_t j = 0; j < p.n_panic; ++j) out->panic_reads[out->n_panic++] = p.panic_reads[j] + (u32)first[i];
    out->n_reads += p.n_reads; out->n_bases += p.n_bases; out->n_minimizers += p.n_minimizers;
    out->n_minimizers_kept += p.n_minimizers_kept; out->n_anchors += p.n_anchors; out->n_rescued += p.n_rescued;
    mm2_map_result_free(&p);
  }
}

// Same, for parts whose records were written straight into one destination array (part i at base + first[i], read ids and
// panic ids already global): close the gaps left by reads without a record; nothing is copied when every read has one.
static void merge_in_place(mm2_map_result_t* parts, const size_t* first, size_t nparts, mm2_paf_rec_t* base, mm2_map_result_t* out) {
  memset(out, 0, sizeof *out);
  out->recs = base;
  size_t npan = 0, w = 0;
  for (size_t i = 0; i < nparts; ++i) npan += parts[i].n_panic;
  out->panic_reads = xmalloc<u32>(npan);
  for (size_t i = 0; i < nparts; ++i) {
    mm2_map_result_t& p = parts[i];
    if (p.n_recs && base + first[i] != base + w) memmove(base + w, base + first[i], p.n_recs * sizeof(mm2_paf_rec_t));
    w += p.n_recs;
    for (size_t j = 0; j < p.n_panic; ++j) out->panic_reads[out->n_panic++] = p.panic_reads[j];
    out->n_reads += p.n_reads; out->n_bases += p.n_bases; out->n_minimizers += p.n_minimizers;
    out->n_minimizers_kept += p.n_minimizers_kept; out->n_anchors += p.n_anchors; out->n_rescued += p.n_rescued;
    p.recs = nullptr;   // not owned by the part
    mm2_map_result_free(&p);
  }
  out->n_recs = w;
}

// paf.rs:130-222 for the single reported chain of a read: 0 = no record (no anchors, main.rs:211-213), 1 = record,
// 2 = the reference panics on this read (idx.seq[rid0] out of bounds after the odd-rid sign extension, F5)
static int build_record(const mm2_index* idx, const ReadHit& h, u32 r, i32 qlen, mm2_paf_rec_t& rec) {
  if (h.n_anchors == 0) return 0;
  const u32 rid0 = h.rid_rev & 0x7fffffffu;
  if (rid0 >= idx->lens.size()) return 2;   // idx.seq[rid0] (paf.rs:149): bounded by the table actually held
  const bool rev = (h.rid_rev >> 31) != 0;
  rec.read_id = r; rec.rid = rid0; rec.qlen = (u32)qlen; rec.qstart = (u32)h.qs; rec.qend = (u32)h.qe;
  rec.tlen = idx->lens[rid0]; rec.tstart = (u32)h.ts; rec.tend = (u32)h.te;
  rec.nm = (u32)std::max((i32)((u32)h.qe - (u32)h.qs), 0); rec.blen = (u32)std::max((i32)((u32)h.te - (u32)h.ts), 0);
  rec.cm = h.cm; rec.s1 = (u32)std::max(h.score, 0); rec.s2 = 0; rec.rl = 0;
  rec.strand = rev ? '-' : '+'; rec.mapq = 60; rec.tp = 'P'; rec.flags = (u8)(h.flags & 1u);
  // dv (paf.rs:156-200).  With the index's w/k equal to the query's, every chain position is one of the read's
  // minimizer positions, so n_match = cm and n_tot spans the ranks of the chain's two end positions.
  float dv = 0.0f;
  if (h.n_mini > 0 && h.st_rank >= 0) {
    const float avg_k = (float)h.sum_span / (float)h.n_mini;
    i32 n_match = (i32)h.cm;
    const i32 en = h.en_rank >= 0 ? h.en_rank : h.st_rank;
    if (h.en_rank < 0) n_match = 1;
    i32 n_tot = en - h.st_rank + 1;
    const i32 r_qs = rev ? qlen - h.qe : h.qs, r_qe = rev ? qlen - h.qs : h.qe;
    const i32 avg_k_i = avg_k != avg_k ? 0 : (avg_k >= 2147483648.0f ? INT32_MAX : (avg_k <= -2147483648.0f ? INT32_MIN : (i32)avg_k));
    if (r_qs > avg_k_i && h.ts > avg_k_i) n_tot += 1;
    if ((qlen - r_qe) > avg_k_i && ((i32)rec.tlen - h.te) > avg_k_i) n_tot += 1;
    const float frac = (float)n_match / (float)n_tot;
    if (frac >= 1.0f) dv = 0.0f;
    else dv = 1.0f - powf(frac, 1.0f / std::max(avg_k, 1.0f));
  }
  rec.dv = dv;
  return 1;
}

// MM2_TRACE=1: host-clock timeline of the mapping calls on stderr (one line per event: ms since the first event, context,
// event), to see how the sub-batch pipeline overlaps copies, kernels and record assembly
static void mm2_trace(const mm2_ctx* ctx, const char* what) {
  static const bool on = [] { const char* e = getenv("MM2_TRACE"); return e && *e == '1'; }();
  if (!on) return;
  static const auto t0 = std::chrono::steady_clock::now();
  static std::mutex mu;
  const float t = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  std::lock_guard<std::mutex> lk(mu);
  fprintf(stderr, "[mm2 trace] %9.3f ctx=%p %s\n", t, (const void*)ctx, what);
}

// rec_dst != NULL: the records (at most one per read on this path) go to rec_dst[0 .. nreads) -- a slice of the caller's
// array, not owned by *out -- and read / panic ids are offset by read_id_base.  Ignored by the general multi-chain tail.
static int map_device_impl(mm2_ctx* ctx, const mm2_index* idx, const u8* d_cat, const u64* d_off, const u64* h_off, size_t nreads,
                           const mm2_map_opts_t* o, mm2_map_result_t* out, bool timer_started, mm2_paf_rec_t* rec_dst = nullptr,
                           u32 read_id_base = 0) {
  cudaStream_t st = ctx->stream;
  const auto wall0 = std::chrono::steady_clock::now();
  if (nreads > 0xFFFFFFF0ull) { mm2_set_error("too many reads in one batch"); return MM2_E_ARG; }
  // paf.rs:156 re-sketches the query with the INDEX's w/k for dv while everything else uses the CLI's
  const bool wk_mismatch = (o->w != idx->w || o->k != idx->k);
  i32 mid_occ = 0;
  MM2_TRY(mm2_index_calc_mid_occ(idx, o->frac_top_repetitive, &mid_occ));  // main.rs:196-197
  if (mid_occ < o->mid_occ_floor) mid_occ = o->mid_occ_floor;
  mm2_chain_params_t p;
  mm2_default_chain_params(o->k, &p);  // main.rs:199-208
  p.max_dist_x = o->max_gap; p.max_dist_y = o->max_gap; p.min_cnt = o->min_cnt; p.min_chain_score = o->min_chain_score;
  if (o->bw >= 0) p.bw = o->bw;
  if (o->bw_long >= 0) p.bw_long = o->bw_long;

  if (!timer_started) ctx->timer.reset();
  ctx->timer.mark(st, "sketch");
  SketchOut so;
  mm2_trace(ctx, "sketch issue");
  MM2_TRY(sketch_device(ctx, d_cat, d_off, h_off, nreads, o->w, o->k, 0, 0, 0, &so));  // seeds.rs:7-11: rid 0, no HPC
  mm2_trace(ctx, "sketch done");
  const u64 nm = so.total;
  ctx->timer.mark(st, "lookup");
  MM2_TRY(ctx->misc.ensure((nreads + 16) * 4 + 64));
  u32* d_sum_span = ctx->misc.as<u32>() + 16;
  CUDA_TRY(cudaMemsetAsync(ctx->misc.p, 0, 64, st));
  const IndexView V = idx->view();
  u64 na = 0, n_dropped = 0;
  MM2_TRY(seeds_hits(ctx, V, so.key, so.val, so.seq_off, (u32)nreads, nm, o->q_occ_max, o->q_occ_frac, mid_occ, o->want_stage_dump != 0,
                     d_sum_span, &na, &n_dropped));
  mm2_trace(ctx, "hits done");
  {
    // Anchors and DP state take 56 B per anchor.  If a batch of repeat-rich reads needs more than what is free, map its
    // two halves one after the other (reads are independent; the cheap stages above are simply redone per half).
    auto grow = [](const DevBuf& b, u64 bytes) -> u64 { return bytes > b.cap ? bytes + bytes / 8 + 256 : 0; };
    const u64 need = grow(ctx->anchors, na * 16) + grow(ctx->dpA, na * 16) + grow(ctx->dpB, na * 16) + grow(ctx->dpT, na * 4) +
                     grow(ctx->dpW, na * 4);
    size_t free_b = 0, total_b = 0;
    bool have_info = need && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess;   // asked only when an arena must grow: the call takes milliseconds
    if (const char* e = getenv("MM2_ANCHOR_BUDGET_MB")) { free_b = (size_t)atoll(e) << 20; have_info = true; }  // test hook
    if (need && have_info && need > (u64)(free_b * 0.9) && nreads > 1 && !o->want_stage_dump) {
      const size_t mid = nreads / 2;
      mm2_map_result_t part[2];
      memset(part, 0, sizeof part);
      const size_t first[2] = {0, mid};
      if (rec_dst) {
        int rc = map_device_impl(ctx, idx, d_cat, d_off, h_off, mid, o, &part[0], true, rec_dst, read_id_base);
        if (rc == MM2_OK) rc = map_device_impl(ctx, idx, d_cat, d_off + mid, h_off + mid, nreads - mid, o, &part[1], true, rec_dst + mid, read_id_base + (u32)mid);
        if (rc != MM2_OK) { part[0].recs = part[1].recs = nullptr; mm2_map_result_free(&part[0]); mm2_map_result_free(&part[1]); return rc; }
        merge_in_place(part, first, 2, rec_dst, out);
        return MM2_OK;
      }
      int rc = map_device_impl(ctx, idx, d_cat, d_off, h_off, mid, o, &part[0], true);
      if (rc == MM2_OK) rc = map_device_impl(ctx, idx, d_cat, d_off + mid, h_off + mid, nreads - mid, o, &part[1], true);
      if (rc != MM2_OK) { mm2_map_result_free(&part[0]); mm2_map_result_free(&part[1]); return rc; }
      merge_map_results(part, first, 2, out);
      return MM2_OK;
    }
  }
  ctx->timer.mark(st, "anchor_fill");
  MM2_TRY(ctx->anchors.ensure(std::max<u64>(1, na) * 16));
  MM2_TRY(seeds_fill_and_sort(ctx, V, so.seq_off, d_off, (u32)nreads, ctx->anchors.as<ulonglong2>()));
  MM2_TRY(ctx->dpA.ensure(std::max<u64>(1, na) * 16));
  MM2_TRY(ctx->dpB.ensure(std::max<u64>(1, na) * 16));
  MM2_TRY(ctx->dpT.ensure(std::max<u64>(1, na) * 4));
  MM2_TRY(ctx->dpW.ensure(std::max<u64>(1, na) * 4));
  MM2_TRY(ctx->hits.ensure((nreads + 1) * sizeof(ReadHit)));
  if (o->min_cnt < 2 || wk_mismatch) {
    // general tail of main.rs:209-218: chain_dp_all may return many (single-anchor) chains, and/or dv needs the
    // minimizers of the query under the index's own w/k
    SketchOut dvs = so;
    u64 nm_dv = nm;
    if (wk_mismatch) {
      MM2_TRY(sketch_device(ctx, d_cat, d_off, h_off, nreads, idx->w, idx->k, 0, 0, 0, &dvs));
      nm_dv = dvs.total;
      MM2_TRY(ctx->keep.ensure(nm_dv + 16));
      MM2_TRY(seeds_filter(ctx, dvs.key, dvs.seq_off, (u32)nreads, nm_dv, 0, 0.0f, ctx->keep.as<u8>(), d_sum_span));  // sum of spans only
    }
    const int rc = map_general_finish(ctx, idx, d_off, h_off, nreads, o, p, dvs, d_sum_span, nm_dv, na, out);
    if (rc == MM2_OK) { out->n_minimizers = nm; out->n_minimizers_kept = nm - n_dropped; }
    return rc;
  }
  ctx->timer.mark(st, "chain");
  // The 64-byte per-read results are written by the chain kernels straight into page-locked host memory (zero-copy:
  // cudaMallocHost memory is device-addressable under unified addressing): one posted PCIe write per read while the kernel
  // runs, instead of a device buffer plus a separate D2H copy at the end (measured 0.8 ms per 100k reads).
  MM2_TRY(ctx->pin_out.ensure((nreads + 1) * sizeof(ReadHit) + 64));
  ReadHit* hits = ctx->pin_out.as<ReadHit>();
  unsigned long long* d_cells = nullptr;
  if (ctx->count_cells) {
    MM2_TRY(ctx->diag.ensure(64));
    d_cells = ctx->diag.as<unsigned long long>();
    CUDA_TRY(cudaMemsetAsync(d_cells, 0, 8, st));
  }
  MM2_TRY(chain_batch(ctx, ctx->anchors.as<ulonglong2>(), ctx->read_aoff.as<u64>(), d_off, so.seq_off, so.val, d_sum_span, (u32)nreads,
                      p, 1, ctx->dpA.as<int4>(), ctx->dpB.as<int4>(), ctx->dpT.as<int>(), ctx->dpW.as<int>(), nullptr,
                      hits, d_cells));   // cell counter: a diagnostic (mm2b200_diag.h), off by default
  ctx->timer.mark(st, "d2h");
  ctx->timer.mark(st, "end");
  mm2_trace(ctx, "chain issued");
  CUDA_TRY(mm2_stream_wait(ctx));
  mm2_trace(ctx, "chain+d2h done");
  ctx->timer.finish();
  if (d_cells) { u64 c = 0; MM2_TRY(read_scalar_u64(ctx, (const u64*)d_cells, &c)); ctx->last_cells = c; }
  const auto wall1 = std::chrono::steady_clock::now();

  // ---- records (paf.rs:130-222) ---------------------------------------------------------------------------------------------
  memset(out, 0, sizeof *out);
  out->n_reads = nreads; out->n_bases = nreads ? h_off[nreads] - h_off[0] : 0; out->n_minimizers = nm; out->n_anchors = na;
  out->recs = rec_dst ? rec_dst : (mm2_paf_rec_t*)g_rec_cache.take(std::max<size_t>(1, nreads) * sizeof(mm2_paf_rec_t));
  std::vector<u32> panics;
  size_t nr = 0;
  {
    // host threads over contiguous read ranges: count the records of each range, then write them in place in read order
    const int nth = nreads >= 16384 ? (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency())) : 1;
    std::vector<size_t> cnt((size_t)nth + 1, 0);
    std::vector<std::vector<u32>> ppan((size_t)nth);
    std::vector<u64> presc((size_t)nth, 0);
    auto range = [&](int t, size_t& lo, size_t& hi) { lo = nreads * (size_t)t / (size_t)nth; hi = nreads * (size_t)(t + 1) / (size_t)nth; };
    auto count = [&](int t) {
      size_t lo, hi, c = 0; range(t, lo, hi);
      for (size_t r = lo; r < hi; ++r) c += (hits[r].n_anchors != 0 && (hits[r].rid_rev & 0x7fffffffu) < idx->lens.size()) ? 1 : 0;
      cnt[(size_t)t + 1] = c;
    };
    auto fill = [&](int t) {
      size_t lo, hi; range(t, lo, hi);
      mm2_paf_rec_t* dst = out->recs + cnt[(size_t)t];
      for (size_t r = lo; r < hi; ++r) {
        const ReadHit& h = hits[r];
        if (h.flags & 1u) presc[(size_t)t] += 1;
        const int kind = build_record(idx, h, (u32)r + read_id_base, (i32)(h_off[r + 1] - h_off[r]), *dst);
        if (kind == 1) ++dst;
        else if (kind == 2) ppan[(size_t)t].push_back((u32)r + read_id_base);
      }
    };
    auto run = [&](auto fn) {
      std::vector<std::thread> th;
      for (int t = 1; t < nth; ++t) th.emplace_back(fn, t);
      fn(0);
      for (auto& t : th) t.join();
    };
    if (nth == 1) { count(0); fill(0); }
    else {
      // one set of threads for both passes: a barrier on an atomic counter after the counts
      std::atomic<int> arrived{0};
      auto both = [&](int t) {
        count(t);
        if (arrived.fetch_add(1) + 1 == nth) { for (int x = 0; x < nth; ++x) cnt[(size_t)x + 1] += cnt[(size_t)x]; arrived.store(nth + 1); }
        while (arrived.load() != nth + 1) std::this_thread::yield();
        fill(t);
      };
      run(both);
    }
    if (nth == 1) cnt[1] += cnt[0];
    nr = cnt[(size_t)nth];
    for (int t = 0; t < nth; ++t) {
      panics.insert(panics.end(), ppan[(size_t)t].begin(), ppan[(size_t)t].end());
      out->n_rescued += presc[(size_t)t];
    }
  }
  out->n_recs = nr;
  out->n_panic = panics.size();
  out->panic_reads = xmalloc<u32>(panics.size());
  if (!panics.empty()) memcpy(out->panic_reads, panics.data(), panics.size() * 4);
  out->n_minimizers_kept = nm - n_dropped;   // seeds.rs:13-36 survivors (counted by the exact-filter passes)
  {  // host wall clock of this call next to the device stage times: whole call up to here, and the record assembly alone
    const auto wall2 = std::chrono::steady_clock::now();
    ctx->timer.add_host("host_records", std::chrono::duration<float, std::milli>(wall2 - wall1).count());
    ctx->timer.add_host("host_call", std::chrono::duration<float, std::milli>(wall2 - wall0).count());
  }

  mm2_trace(ctx, "records done");
  if (o->want_stage_dump) {
    out->mini_offs = xmalloc<u64>(nreads + 1); out->minis = xmalloc<mm2_mini_t>(nm); out->mini_keep = xmalloc<u8>(nm);
    out->anchor_offs = xmalloc<u64>(nreads + 1); out->anchors = xmalloc<mm2_anchor_t>(na);
    out->f = xmalloc<i32>(na); out->v = xmalloc<i32>(na); out->pprev = xmalloc<i32>(na);
    out->mini_offs[0] = 0; out->anchor_offs[0] = 0;
    if (nreads) {
      CUDA_TRY(cudaMemcpyAsync(out->mini_offs, so.seq_off, (nreads + 1) * 8, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaMemcpyAsync(out->anchor_offs, ctx->read_aoff.p, (nreads + 1) * 8, cudaMemcpyDeviceToHost, st));
    }
    if (nm) {
      CUDA_TRY(cudaMemcpyAsync(out->mini_keep, ctx->keep.p, nm, cudaMemcpyDeviceToHost, st));
      MM2_TRY(ctx->sort_tmp.ensure(nm * 16));
      MM2_LAUNCH(ctx, interleave_kernel, grid_for(nm), 256, 0, so.key, so.val, ctx->sort_tmp.as<ulonglong2>(), nm);
      CUDA_TRY(cudaMemcpyAsync(out->minis, ctx->sort_tmp.p, nm * 16, cudaMemcpyDeviceToHost, st));
    }
    if (na) {
      CUDA_TRY(cudaMemcpyAsync(out->anchors, ctx->anchors.p, na * 16, cudaMemcpyDeviceToHost, st));
      MM2_TRY(ctx->chain_idx.ensure(na * 12));
      int* d_f = ctx->chain_idx.as<int>();
      MM2_LAUNCH(ctx, dp_unpack_kernel, grid_for(na), 256, 0, ctx->dpA.as<int4>(), d_f, d_f + na, d_f + 2 * na, na);
      CUDA_TRY(cudaMemcpyAsync(out->f, d_f, na * 4, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaMemcpyAsync(out->v, d_f + na, na * 4, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaMemcpyAsync(out->pprev, d_f + 2 * na, na * 4, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    if (nm) { u64 kept = 0; for (u64 i = 0; i < nm; ++i) kept += out->mini_keep[i]; out->n_minimizers_kept = kept; }
  }
  return MM2_OK;
}

extern "C" int mm2_map_batch_device(mm2_ctx_t* ctx, const mm2_index_t* idx, const void* d_cat, const void* d_offs, const uint64_t* h_offs,
                                    size_t nreads, const mm2_map_opts_t* opts, mm2_map_result_t* out) {
  if (!ctx || !idx || !h_offs || !opts || !out || (nreads && (!d_cat || !d_offs))) { mm2_set_error("mm2_map_batch_device: NULL argument"); return MM2_E_ARG; }
  if (idx->device != ctx->device) { mm2_set_error("index lives on device %d, context on %d", idx->device, ctx->device); return MM2_E_ARG; }
  if (nreads && h_offs[0] != 0) { mm2_set_error("mm2_map_batch_device: offsets must start at 0"); return MM2_E_ARG; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  return map_device_impl(ctx, idx, (const u8*)d_cat, (const u64*)d_offs, h_offs, nreads, opts, out, false);
}

// one sub-batch from host memory on `ctx`: H2D of the reads, the device pipeline, D2H of the hits
// ---- reads as 2-bit codes (mm2_map_batch_packed): a quarter of the H2D bytes; unpacked to the ASCII layout on the device ----
namespace {
struct PackedSrc { const u8* packed; const u64* n_pos; size_t n_n; u64* d_npos; };   // d_npos: the positions, resident
// 16 bases per thread: one packed word -> 16 letters (a 16-byte store).  first: 16-base group of the first thread.
// first_valid: bases below it (in the first group) belong to an earlier upload whose N stamps must survive.
__global__ void unpack_reads_kernel(const u32* __restrict__ packed, u64 group_lo, u64 ngroups, u64 group_base, u64 first_valid, uint4* __restrict__ out) {
  const u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngroups) return;
  const u32 wd = packed[group_lo - group_base + g];
  u32 o[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    u32 x = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const u32 c = (wd >> (8 * q + 2 * j)) & 3u;
      x |= ((0x54474341u >> (8 * c)) & 0xFFu) << (8 * j);     // "ACGT"[c]
    }
    o[q] = x;
  }
  const u64 b16 = 16 * (group_lo + g);
  if (b16 >= first_valid) out[group_lo - group_base + g] = make_uint4(o[0], o[1], o[2], o[3]);
  else {
    u8* ob = reinterpret_cast<u8*>(out + (group_lo - group_base + g));
    for (int j = 0; j < 16; ++j) if (b16 + j >= first_valid) ob[j] = (u8)(o[j >> 2] >> (8 * (j & 3)));
  }
}
__global__ void stamp_n_kernel(const u64* __restrict__ pos, u64 n, u64 base, u8* __restrict__ seq) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) seq[pos[i] - base] = (u8)'N';
}
}  // namespace
// bases [b0, b1) of the batch (relative to `base`) -> ctx->seq, on stream st: a plain copy, or packed copy + unpack + N stamps
static cudaError_t upload_reads(mm2_ctx* ctx, cudaStream_t st, const u8* cat, const PackedSrc* pk, u64 base, u64 b0, u64 b1) {
  if (b1 <= b0) return cudaSuccess;
  if (!pk) return cudaMemcpyAsync(ctx->seq.as<u8>() + b0, cat + base + b0, b1 - b0, cudaMemcpyHostToDevice, st);
  const u64 gb = base / 16, g0 = (base + b0) / 16, g1 = (base + b1 + 15) / 16;   // base is a multiple of 16 (checked by the caller)
  cudaError_t e = cudaMemcpyAsync(ctx->packed.as<u8>() + 4 * (g0 - gb), pk->packed + 4 * g0, 4 * (g1 - g0), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  unpack_reads_kernel<<<(unsigned)((g1 - g0 + 255) / 256), 256, 0, st>>>(ctx->packed.as<u32>(), g0, g1 - g0, gb, base + b0, ctx->seq.as<uint4>());
  ctx->launches += 1;
  if (pk->n_n) {
    const u64* lo = std::lower_bound(pk->n_pos, pk->n_pos + pk->n_n, base + b0);
    const u64* hi = std::lower_bound(pk->n_pos, pk->n_pos + pk->n_n, base + b1);
    if (hi > lo) {
      stamp_n_kernel<<<(unsigned)((hi - lo + 255) / 256), 256, 0, st>>>(pk->d_npos + (lo - pk->n_pos), (u64)(hi - lo), base, ctx->seq.as<u8>());
      ctx->launches += 1;
    }
  }
  return cudaGetLastError();
}
// device buffers of a packed batch (+ the N positions, uploaded once, on stream st)
static int packed_prepare(mm2_ctx* ctx, cudaStream_t st, PackedSrc* pk, u64 total) {
  MM2_TRY(ctx->packed.ensure(total / 4 + 64));
  pk->d_npos = nullptr;
  if (pk->n_n) {
    MM2_TRY(ctx->packed_n.ensure(pk->n_n * 8));
    pk->d_npos = ctx->packed_n.as<u64>();
    CUDA_TRY(cudaMemcpyAsync(pk->d_npos, pk->n_pos, pk->n_n * 8, cudaMemcpyHostToDevice, st));
  }
  return MM2_OK;
}

static int map_host_single(mm2_ctx* ctx, const mm2_index* idx, const u8* cat, const u64* offs, size_t nreads, const mm2_map_opts_t* opts,
                           mm2_map_result_t* out, PackedSrc* pk = nullptr) {
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const u64 base = nreads ? offs[0] : 0, total = nreads ? offs[nreads] - base : 0;
  std::vector<u64> off0(nreads + 1, 0);
  for (size_t i = 0; i <= nreads && nreads; ++i) off0[i] = offs[i] - base;
  ctx->timer.reset();
  ctx->timer.mark(st, "h2d");
  MM2_TRY(ctx->seq.ensure(total + 64));
  MM2_TRY(ctx->seq_off.ensure((nreads + 1) * 8));
  MM2_TRY(ctx->pin_in.ensure((nreads + 1) * 8));
  {
    // one H2D at a time: concurrent copies from several workers would share the link and all finish together, which
    // keeps the workers in lock step (copy, copy, copy, then compute, compute, compute) instead of overlapping
    static std::mutex h2d_mutex;
    std::lock_guard<std::mutex> lk(h2d_mutex);
    mm2_trace(ctx, "h2d start");
    if (pk) MM2_TRY(packed_prepare(ctx, st, pk, total));
    CUDA_TRY(upload_reads(ctx, st, cat, pk, base, 0, total));
    memcpy(ctx->pin_in.p, off0.data(), (nreads + 1) * 8);  // pinned bounce: pageable sources serialise the streams
    CUDA_TRY(cudaMemcpyAsync(ctx->seq_off.p, ctx->pin_in.p, (nreads + 1) * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    mm2_trace(ctx, "h2d done");
  }
  return map_device_impl(ctx, idx, ctx->seq.as<u8>(), ctx->seq_off.as<u64>(), off0.data(), nreads, opts, out, true);
}

// Large host batches are cut into sub-batches that rotate over a few worker contexts (own stream, own arenas, own
// host thread), so the H2D copy and the host-side record assembly of one sub-batch overlap the kernels of the other.
// Reads are independent (main.rs:193-219), so the records are simply concatenated in input order.
static int map_host_pipelined(mm2_ctx* ctx, const mm2_index* idx, const u8* cat, const u64* offs, size_t nreads,
                              const mm2_map_opts_t* opts, mm2_map_result_t* out, size_t nsub, PackedSrc* pk = nullptr) {
  const int NW = ctx->n_workers;
  for (int w = 0; w < NW; ++w)
    if (!ctx->worker[w]) MM2_TRY(mm2_ctx_create(ctx->device, &ctx->worker[w]));
  for (int w = 0; w < NW; ++w) { ctx->worker[w]->count_cells = ctx->count_cells; ctx->worker[w]->last_cells = 0; }
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (!ctx->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  while (ctx->copy_events.size() < nsub + 1) {   // + 1: the first sub-batch may be split below
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->copy_events.push_back(e);
  }
  // The whole batch gets one device buffer; a dedicated copy stream uploads it sub-batch by sub-batch, back to back, with an
  // event after each, so the link never waits for a worker (before: a worker uploaded its own next sub-batch only after it
  // had finished the previous one, and the copy engine idled ~35 % of the call).  Offsets stay absolute (rebased to the
  // batch start): every stage only uses differences, and the sketch kernel's halo loads are bounded by the sub-batch end.
  const u64 base = offs[0], total = offs[nreads] - base;
  MM2_TRY(ctx->seq.ensure(total + 64));
  MM2_TRY(ctx->seq_off.ensure((nreads + 1) * 8));
  MM2_TRY(ctx->pin_in.ensure((nreads + 1) * 8));
  if (pk) MM2_TRY(packed_prepare(ctx, ctx->copy_stream, pk, total));
  u64* h_off0 = ctx->pin_in.as<u64>();
  for (size_t i = 0; i <= nreads; ++i) h_off0[i] = offs[i] - base;
  // sub-batch boundaries balanced by bases; the first sub-batch is split 1/4 + 3/4 so that the kernels start after a quarter
  // of a sub-batch has been uploaded instead of a whole one
  const size_t nunit = nsub;
  if (nsub >= 4) nsub += 1;
  std::vector<size_t> cut(nsub + 1, nreads);
  cut[0] = 0;
  for (size_t sidx = 1; sidx < nsub; ++sidx) {
    const u64 target = nsub == nunit ? total * sidx / nunit : (sidx == 1 ? total / (4 * nunit) : total * (sidx - 1) / nunit);
    cut[sidx] = (size_t)(std::lower_bound(h_off0, h_off0 + nreads + 1, target) - h_off0);
    if (cut[sidx] < cut[sidx - 1]) cut[sidx] = cut[sidx - 1];
  }
  std::vector<mm2_map_result_t> part(nsub);
  for (auto& p : part) memset(&p, 0, sizeof p);
  int rc[4] = {MM2_OK, MM2_OK, MM2_OK, MM2_OK};
  std::string err[4];
  std::vector<std::vector<float>> ms_sum(4);
  u64 cells_sum[4] = {0, 0, 0, 0};
  // one record array for the whole batch, written in place by the workers (the default path reports at most one chain per
  // read); the general multi-chain tail keeps per-part arrays that are concatenated afterwards
  const bool direct = !(opts->min_cnt < 2 || opts->w != idx->w || opts->k != idx->k);
  mm2_paf_rec_t* final_recs = direct ? (mm2_paf_rec_t*)g_rec_cache.take(std::max<size_t>(1, nreads) * sizeof(mm2_paf_rec_t)) : nullptr;
  std::atomic<size_t> n_issued{0};   // sub-batches whose upload (and event) has been enqueued on the copy stream
  std::atomic<int> copy_rc{MM2_OK};
  auto work = [&](int w) {
    mm2_ctx* c = ctx->worker[w];
    cudaSetDevice(c->device);
    for (size_t sidx = (size_t)w; sidx < nsub; sidx += (size_t)NW) {
      const size_t lo = cut[sidx], hi = cut[sidx + 1];
      while (n_issued.load(std::memory_order_acquire) <= sidx && copy_rc.load() == MM2_OK) std::this_thread::yield();
      if (copy_rc.load() != MM2_OK) { rc[w] = copy_rc.load(); err[w] = "host to device copy failed"; return; }
      if (cudaStreamWaitEvent(c->stream, ctx->copy_events[sidx], 0) != cudaSuccess) { rc[w] = MM2_E_CUDA; err[w] = "cudaStreamWaitEvent failed"; return; }
      const int r = map_device_impl(c, idx, ctx->seq.as<u8>(), ctx->seq_off.as<u64>() + lo, h_off0 + lo, hi - lo, opts, &part[sidx], false,
                                    final_recs ? final_recs + lo : nullptr, final_recs ? (u32)lo : 0u);
      if (r != MM2_OK) { rc[w] = r; err[w] = mm2_last_error(); return; }
      cells_sum[w] += c->last_cells;
      if (ms_sum[w].size() < c->timer.ms.size()) ms_sum[w].resize(c->timer.ms.size(), 0.f);
      for (size_t i = 0; i < c->timer.ms.size(); ++i) ms_sum[w][i] += c->timer.ms[i];
    }
  };
  u64 l0 = 0, l1 = 0;
  for (int w = 0; w < NW; ++w) l0 += ctx->worker[w]->launches;
  std::vector<std::thread> th;
  for (int w = 0; w < NW; ++w) th.emplace_back(work, w);
  {  // this thread feeds the copy stream (a pageable source makes cudaMemcpyAsync block, so the workers are already running)
    cudaStream_t cs = ctx->copy_stream;
    cudaError_t e = cudaMemcpyAsync(ctx->seq_off.p, h_off0, (nreads + 1) * 8, cudaMemcpyHostToDevice, cs);
    for (size_t sidx = 0; sidx < nsub && e == cudaSuccess; ++sidx) {
      const u64 b0 = h_off0[cut[sidx]], b1 = h_off0[cut[sidx + 1]];
      e = upload_reads(ctx, cs, cat, pk, base, b0, b1);
      if (e == cudaSuccess) e = cudaEventRecord(ctx->copy_events[sidx], cs);
      if (e == cudaSuccess) n_issued.store(sidx + 1, std::memory_order_release);
    }
    if (e != cudaSuccess) { copy_rc.store(MM2_E_CUDA); cudaGetLastError(); }
  }
  for (auto& t : th) t.join();
  cudaStreamSynchronize(ctx->copy_stream);
  for (int w = 0; w < NW; ++w) l1 += ctx->worker[w]->launches;
  ctx->launches += l1 - l0;
  ctx->last_cells = cells_sum[0] + cells_sum[1] + cells_sum[2] + cells_sum[3];
  for (int w = 0; w < NW; ++w)
    if (rc[w] != MM2_OK) {
      for (auto& p : part) { if (final_recs) p.recs = nullptr; mm2_map_result_free(&p); }
      g_rec_cache.give(final_recs);
      mm2_set_error("%s", err[w].c_str());
      return rc[w];
    }
  // stage times: sum over both workers (device time spent per stage, not wall time)
  ctx->timer.names = ctx->worker[0]->timer.names;
  ctx->timer.ms.assign(ms_sum[0].size(), 0.f);
  for (int w = 1; w < NW; ++w) if (ms_sum[w].size() > ms_sum[0].size()) ms_sum[0].resize(ms_sum[w].size(), 0.f);
  ctx->timer.ms.assign(ms_sum[0].size(), 0.f);
  for (int w = 0; w < NW; ++w) for (size_t i = 0; i < ms_sum[w].size() && i < ctx->timer.ms.size(); ++i) ctx->timer.ms[i] += ms_sum[w][i];
  ctx->timer.names_blob.clear();
  for (size_t i = 0; i < ctx->timer.ms.size() && i < ctx->timer.names.size(); ++i) { ctx->timer.names_blob += ctx->timer.names[i]; ctx->timer.names_blob.push_back('\0'); }
  ctx->timer.names_blob.push_back('\0');
  if (final_recs) merge_in_place(part.data(), cut.data(), nsub, final_recs, out);
  else merge_map_results(part.data(), cut.data(), nsub, out);
  return MM2_OK;
}

static int map_host_any(mm2_ctx_t* ctx, const mm2_index_t* idx, const uint8_t* cat, PackedSrc* pk, const uint64_t* offs, size_t nreads,
                        const mm2_map_opts_t* opts, mm2_map_result_t* out) {
  if (idx->device != ctx->device) { mm2_set_error("index lives on device %d, context on %d", idx->device, ctx->device); return MM2_E_ARG; }
  const u64 total = nreads ? offs[nreads] - offs[0] : 0;
  // sub-batches of ~64 Mbase (MM2_SUBBATCH_MB) over 4 worker contexts (MM2_WORKERS): measured best on configs[1]; small batches and stage dumps take the single-context path
  // (packed reads: the upload of a sub-batch is four times shorter, twice the sub-batch size was measured best: 18.9 vs 19.8 ms)
  const u64 sub_bytes = ctx->subbatch_bytes * (pk ? 2 : 1);
  size_t nsub = (size_t)std::min<u64>(64, total / sub_bytes);
  if (nsub > nreads) nsub = nreads;
  if (nsub >= 2 && !opts->want_stage_dump && ctx->pipeline) {
    // the pipelined path keeps the whole batch resident: batches above MM2_RESIDENT_MB (default 4096) go through it in pieces
    const char* e_res = getenv("MM2_RESIDENT_MB");
    const u64 resident = (u64)(e_res && atoll(e_res) > 0 ? atoll(e_res) : 4096) << 20;
    if (total <= resident) return map_host_pipelined(ctx, idx, cat, offs, nreads, opts, out, nsub, pk);
    std::vector<size_t> first;
    std::vector<mm2_map_result_t> part;
    size_t lo = 0;
    while (lo < nreads) {
      size_t hi = (size_t)(std::upper_bound(offs + lo, offs + nreads + 1, offs[lo] + resident) - offs) - 1;
      if (hi <= lo) hi = lo + 1;
      if (hi > nreads) hi = nreads;
      if (pk) while (hi < nreads && (offs[hi] & 15)) ++hi;   // packed pieces start on a 16-base boundary
      const u64 piece = offs[hi] - offs[lo];
      size_t ns = (size_t)std::min<u64>(64, piece / sub_bytes);
      if (ns > hi - lo) ns = hi - lo;
      mm2_map_result_t r;
      memset(&r, 0, sizeof r);
      const int rc = ns >= 2 ? map_host_pipelined(ctx, idx, cat, offs + lo, hi - lo, opts, &r, ns, pk) : map_host_single(ctx, idx, cat, offs + lo, hi - lo, opts, &r, pk);
      if (rc != MM2_OK) { for (auto& p : part) mm2_map_result_free(&p); return rc; }
      part.push_back(r); first.push_back(lo);
      lo = hi;
    }
    first.push_back(nreads);
    merge_map_results(part.data(), first.data(), part.size(), out);
    return MM2_OK;
  }
  return map_host_single(ctx, idx, cat, offs, nreads, opts, out, pk);
}

extern "C" int mm2_map_batch(mm2_ctx_t* ctx, const mm2_index_t* idx, const uint8_t* cat, const uint64_t* offs, size_t nreads,
                             const mm2_map_opts_t* opts, mm2_map_result_t* out) {
  if (!ctx || !idx || !offs || !opts || !out || (nreads && !cat)) { mm2_set_error("mm2_map_batch: NULL argument"); return MM2_E_ARG; }
  return map_host_any(ctx, idx, cat, nullptr, offs, nreads, opts, out);
}

extern "C" int mm2_map_batch_packed(mm2_ctx_t* ctx, const mm2_index_t* idx, const uint8_t* packed, const uint64_t* n_pos, size_t n_n,
                                    const uint64_t* offs, size_t nreads, const mm2_map_opts_t* opts, mm2_map_result_t* out) {
  if (!ctx || !idx || !offs || !opts || !out || (nreads && !packed) || (n_n && !n_pos)) { mm2_set_error("mm2_map_batch_packed: NULL argument"); return MM2_E_ARG; }
  if (nreads && (offs[0] & 15)) { mm2_set_error("mm2_map_batch_packed: offs[0] must be a multiple of 16 bases"); return MM2_E_ARG; }
  PackedSrc pk{packed, n_pos, n_n, nullptr};
  return map_host_any(ctx, idx, nullptr, &pk, offs, nreads, opts, out);
}

// nt4.rs:2-10 on the host: ASCII -> 2-bit codes + the positions of everything that is not ACGTacgt
extern "C" int mm2_pack_reads(const uint8_t* cat, uint64_t n_bases, uint8_t* packed, uint64_t* n_pos, size_t n_cap, size_t* n_n) {
  if ((n_bases && (!cat || !packed)) || !n_n) { mm2_set_error("mm2_pack_reads: NULL argument"); return MM2_E_ARG; }
  static const struct Tab { u8 t[256]; Tab() { memset(t, 4, 256); t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3; } } tab;
  size_t nn = 0;
  for (u64 i = 0; i < n_bases; i += 4) {
    u8 b = 0;
    for (u64 j = i; j < i + 4 && j < n_bases; ++j) {
      const u8 c = tab.t[cat[j]];
      if (c > 3) { if (nn < n_cap && n_pos) n_pos[nn] = j; ++nn; }
      b |= (u8)((c & 3) << (2 * (j - i)));
    }
    packed[i / 4] = b;
  }
  *n_n = nn;   // > n_cap: call again with a larger n_pos
  return MM2_OK;
}

extern "C" void mm2_map_result_free(mm2_map_result_t* r) {
  if (!r) return;
  g_rec_cache.give(r->recs); free(r->panic_reads); free(r->mini_offs); free(r->minis); free(r->mini_keep); free(r->anchor_offs); free(r->anchors);
  free(r->f); free(r->v); free(r->pprev);
  memset(r, 0, sizeof *r);
}

// ---- PAF text (paf.rs:224-236) ---------------------------------------------------------------------------------------------
namespace {
inline char* put_u32(char* p, u32 v) {
  char tmp[10]; int n = 0;
  do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}
inline char* put_str(char* p, const char* s, size_t n) { memcpy(p, s, n); return p + n; }
// upper bound of one formatted line excluding the two names
constexpr size_t PAF_FIXED = 200;
size_t paf_line(const mm2_paf_rec_t& rec, const char* qname, size_t qn, const char* tname, size_t tn, char* p0) {
  char* p = p0;
  u32 qs, qe;
  if (rec.strand == '-') { qs = rec.qlen - rec.qend; qe = rec.qlen - rec.qstart; } else { qs = rec.qstart; qe = rec.qend; }
  p = put_str(p, qname, qn); *p++ = '\t';
  p = put_u32(p, rec.qlen); *p++ = '\t'; p = put_u32(p, qs); *p++ = '\t'; p = put_u32(p, qe); *p++ = '\t';
  *p++ = (char)rec.strand; *p++ = '\t';
  p = put_str(p, tname, tn); *p++ = '\t';
  p = put_u32(p, rec.tlen); *p++ = '\t'; p = put_u32(p, rec.tstart); *p++ = '\t'; p = put_u32(p, rec.tend); *p++ = '\t';
  p = put_u32(p, rec.nm); *p++ = '\t'; p = put_u32(p, rec.blen); *p++ = '\t'; p = put_u32(p, rec.mapq);
  p = put_str(p, "\ttp:A:", 6); *p++ = (char)rec.tp;
  p = put_str(p, "\tcm:i:", 6); p = put_u32(p, rec.cm);
  p = put_str(p, "\ts1:i:", 6); p = put_u32(p, rec.s1);
  p = put_str(p, "\ts2:i:", 6); p = put_u32(p, rec.s2);
  p = put_str(p, "\tdv:f:", 6);
  p += snprintf(p, 48, "%.4f", (double)rec.dv);  // `{:.4}` of an f32: correctly rounded decimal of the exact value
  p = put_str(p, "\trl:i:", 6); p = put_u32(p, rec.rl);
  return (size_t)(p - p0);
}
}  // namespace

extern "C" int mm2_paf_format(const mm2_paf_rec_t* rec, const char* qname, const char* tname, char* buf, size_t cap) {
  if (!rec || !buf) { mm2_set_error("mm2_paf_format: NULL argument"); return MM2_E_ARG; }
  if (!qname) qname = "*";
  if (!tname) tname = "*";
  const size_t qn = strlen(qname), tn = strlen(tname);
  if (cap < qn + tn + PAF_FIXED) { mm2_set_error("mm2_paf_format: buffer too small"); return MM2_E_ARG; }
  const size_t n = paf_line(*rec, qname, qn, tname, tn, buf);
  buf[n] = 0;
  return (int)n;
}

extern "C" int mm2_paf_format_batch(const mm2_index_t* idx, const mm2_map_result_t* res, const char* const* qnames, char** out, size_t* out_len) {
  if (!idx || !res || !out || !out_len) { mm2_set_error("mm2_paf_format_batch: NULL argument"); return MM2_E_ARG; }
  size_t cap = 1;
  for (size_t i = 0; i < res->n_recs; ++i) {
    const mm2_paf_rec_t& r = res->recs[i];
    const char* qn = qnames && qnames[r.read_id] ? qnames[r.read_id] : "*";
    cap += strlen(qn) + (idx->has_name[r.rid] ? idx->names[r.rid].size() : 1) + PAF_FIXED;
  }
  char* buf = (char*)malloc(cap);
  if (!buf) { mm2_set_error("out of host memory"); return MM2_E_OOM; }
  char* p = buf;
  for (size_t i = 0; i < res->n_recs; ++i) {
    const mm2_paf_rec_t& r = res->recs[i];
    const char* qn = qnames && qnames[r.read_id] ? qnames[r.read_id] : "*";
    const bool hn = idx->has_name[r.rid] != 0;
    const char* tn = hn ? idx->names[r.rid].c_str() : "*";
    p += paf_line(r, qn, strlen(qn), tn, hn ? idx->names[r.rid].size() : 1, p);
    *p++ = '\n';
  }
  *p = 0;
  *out = buf; *out_len = (size_t)(p - buf);
  return MM2_OK;
}
