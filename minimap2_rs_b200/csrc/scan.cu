// scan.cu — single-pass exclusive prefix sum (u32 counts -> u64 offsets) with decoupled look-back.
// Used wherever the reference does a serial `push` into a Vec: anchor offsets (seeds.rs:42-57), run/bucket offsets
// of the index build (index.rs:80-107).
#include "mm2_internal.cuh"

namespace {
constexpr int SC_NT = 256, SC_PER = 8, SC_TILE = SC_NT * SC_PER;

// ACC = u32: a tile of 2048 counts is known to sum below 2^32 (occurrence counts); ACC = u64: no such bound (per-read
// anchor counts).
template <class ACC>
__global__ void __launch_bounds__(SC_NT) scan_kernel(const u32* __restrict__ in, u64* __restrict__ out, u64 n,
                                                     u64* status, u32* ticket) {
  __shared__ ACC s_wsum[SC_NT / 32];
  __shared__ u32 s_tile;
  __shared__ u64 s_base;
  const int tid = threadIdx.x;
  const u64 ntiles = (n + SC_TILE - 1) / SC_TILE;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const u64 tile = s_tile;
    if (tile >= ntiles) break;
    const u64 i0 = tile * SC_TILE + (u64)tid * SC_PER;
    u32 c[SC_PER];
    if (i0 + SC_PER <= n) {
      const uint4 a = *reinterpret_cast<const uint4*>(in + i0), b = *reinterpret_cast<const uint4*>(in + i0 + 4);
      c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < SC_PER; ++j) c[j] = (i0 + j < n) ? in[i0 + j] : 0u;
    }
    ACC sum = 0;
#pragma unroll
    for (int j = 0; j < SC_PER; ++j) sum += c[j];
    ACC inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const ACC t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
      if ((tid & 31) >= d) inc += t;
    }
    if ((tid & 31) == 31) s_wsum[tid >> 5] = inc;
    __syncthreads();
    ACC wbase = 0, tot = 0;
#pragma unroll
    for (int x = 0; x < SC_NT / 32; ++x) {
      const ACC ws = s_wsum[x];
      if (x < (tid >> 5)) wbase += ws;
      tot += ws;
    }
    if (tid < 32) {
      volatile u64* st = status;
      u64 excl = 0;
      if (tile == 0) {
        if (tid == 0) st[0] = (2ULL << 62) | (u64)tot;
      } else {
        if (tid == 0) st[tile] = (1ULL << 62) | (u64)tot;
        i64 look = (i64)tile - 1;
        for (;;) {
          const i64 idx = look - tid;
          u64 v;
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
        if (tid == 0) st[tile] = (2ULL << 62) | (excl + (u64)tot);
      }
      if (tid == 0) {
        s_base = excl;
        if (tile == ntiles - 1) out[n] = excl + (u64)tot;
      }
    }
    __syncthreads();
    u64 run = s_base + wbase + inc - sum;
#pragma unroll
    for (int j = 0; j < SC_PER; ++j) {
      if (i0 + j < n) out[i0 + j] = run;
      run += c[j];
    }
  }
}

__global__ void scan_zero_total(u64* out) { out[0] = 0; }
}  // namespace

int scan_u32_to_u64(mm2_ctx* ctx, const u32* d_in, u64* d_out, size_t n, bool wide) {
  if (n == 0) { MM2_LAUNCH(ctx, scan_zero_total, 1, 1, 0, d_out); return MM2_OK; }
  const size_t ntiles = (n + SC_TILE - 1) / SC_TILE;
  MM2_TRY(ctx->scan_status.ensure(ntiles * 8 + 16));
  CUDA_TRY(cudaMemsetAsync(ctx->scan_status.p, 0, ntiles * 8 + 16, ctx->stream));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  const int grid = (int)std::min<size_t>(ntiles, (size_t)sms * 8);
  if (wide)
    MM2_LAUNCH(ctx, scan_kernel<u64>, grid, SC_NT, 0, d_in, d_out, (u64)n, ctx->scan_status.as<u64>(), (u32*)((u8*)ctx->scan_status.p + ntiles * 8));
  else
    MM2_LAUNCH(ctx, scan_kernel<u32>, grid, SC_NT, 0, d_in, d_out, (u64)n, ctx->scan_status.as<u64>(), (u32*)((u8*)ctx->scan_status.p + ntiles * 8));
  CUDA_TRY(cudaGetLastError());
  return MM2_OK;
}

int read_scalar_u64(mm2_ctx* ctx, const u64* d_src, u64* out) {
  MM2_TRY(ctx->pin_scalar.ensure(64));
  CUDA_TRY(cudaMemcpyAsync(ctx->pin_scalar.p, d_src, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(mm2_stream_wait(ctx));
  *out = *ctx->pin_scalar.as<u64>();
  return MM2_OK;
}
