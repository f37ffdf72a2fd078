// Micro-benchmark: how fast can one persistent CTA per SM stream a (N x 257) fp32 matrix with different access
// patterns?  (development aid for fbank_tc.cu; not part of the library)
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ float ld_stream1(const float *p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_stream4(const float *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// pattern 0: column chunks, rows warp+16j (what fbank_tc v1 does); 1: rows 8*warp+j; DEPTH items in flight
template <int DEPTH, int MODE, int NW>
__global__ void __launch_bounds__(NW * 32, 1) chunk_kernel(const float *x, float *out, int N, int F, int rows_per_cta) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_begin = blockIdx.x * rows_per_cta, row_end = min(N, row_begin + rows_per_cta);
  const int ntiles = (row_end - row_begin + 127) / 128, nch = 8;
  const int nitems = ntiles * nch;
  constexpr int RPW = 128 / NW;   // rows per warp per tile
  float v[DEPTH][RPW];
  float acc = 0.f;
  int li = 0;
  auto load = [&](float (&g)[RPW]) {
    const int t = li / nch, c = li % nch;
    const int row0 = row_begin + t * 128;
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
      const int r = MODE == 0 ? warp + NW * j : RPW * warp + j;
      const int row = row0 + r;
      g[j] = row < row_end ? ld_stream1(x + (size_t)row * F + c * 32 + lane) : 0.f;
    }
    ++li;
  };
#pragma unroll
  for (int d = 0; d < DEPTH; ++d) if (d < nitems) load(v[d]);
  for (int q0 = 0; q0 < nitems; q0 += DEPTH) {
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      if (q0 + d < nitems) {
#pragma unroll
        for (int j = 0; j < RPW; ++j) acc += v[d][j] * v[d][j];
        if (li < nitems) load(v[d]);
      }
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

// flat float4 streaming of the CTA's row range, U float4 in flight per thread
template <int U, int NW>
__global__ void __launch_bounds__(NW * 32, 1) flat_kernel(const float *x, float *out, int N, int F, int rows_per_cta) {
  const int row_begin = blockIdx.x * rows_per_cta, row_end = min(N, row_begin + rows_per_cta);
  const size_t b4 = ((size_t)row_begin * F) / 4, e4 = ((size_t)row_end * F) / 4;
  float acc = 0.f;
  for (size_t i = b4 + threadIdx.x; i < e4; i += (size_t)NW * 32 * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t k = i + (size_t)u * NW * 32;
      v[u] = k < e4 ? ld_stream4(x + 4 * k) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x * v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 123.456f) out[0] = acc;
}


// pattern 2: like pattern 0 (rows warp+16j, one 32-float item per row and step), but every warp load is a 128 B
// ALIGNED window of flat memory (one cache line, 4 sectors) instead of the row's k-chunk (4 B aligned: 2 lines,
// 5 sectors).  A 257-float row spans 9 such windows; lanes whose element belongs to a neighbouring row are wasted.
template <int DEPTH, int NW>
__global__ void __launch_bounds__(NW * 32, 1) window_kernel(const float *x, float *out, int N, int F, int rows_per_cta) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_begin = blockIdx.x * rows_per_cta, row_end = min(N, row_begin + rows_per_cta);
  const int ntiles = (row_end - row_begin + 127) / 128, nch = 9;
  const int nitems = ntiles * nch;
  constexpr int RPW = 128 / NW;
  float v[DEPTH][RPW];
  float acc = 0.f;
  int li = 0;
  const size_t total = (size_t)N * F;
  auto load = [&](float (&g)[RPW]) {
    const int t = li / nch, c = li % nch;
    const int row0 = row_begin + t * 128;
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
      const int row = row0 + warp + NW * j;
      const size_t f = (((size_t)row * F) & ~(size_t)31) + 32 * c + lane;
      g[j] = (row < row_end && f < total) ? ld_stream1(x + f) : 0.f;
    }
    ++li;
  };
#pragma unroll
  for (int d = 0; d < DEPTH; ++d) if (d < nitems) load(v[d]);
  for (int q0 = 0; q0 < nitems; q0 += DEPTH) {
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      if (q0 + d < nitems) {
#pragma unroll
        for (int j = 0; j < RPW; ++j) acc += v[d][j] * v[d][j];
        if (li < nitems) load(v[d]);
      }
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

// pattern 3: the CTA's row range staged through shared memory by 1-D bulk async copies (cp.async.bulk, the TMA unit):
// blocks of RB rows (RB % 4 == 0 keeps every block 16 B aligned although a row is 1028 B), NST blocks in flight;
// NW consumer warps read each landed block back from shared memory (stand-in for the converters).
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int RB, int NST, int NW>
__global__ void __launch_bounds__(NW * 32, 1) bulk_kernel(const float *x, float *out, int N, int F, int rows_per_cta) {
  extern __shared__ __align__(128) unsigned char smraw[];
  float *stage = reinterpret_cast<float *>(smraw);
  __shared__ unsigned long long full[NST], empty[NST];
  const int tid = threadIdx.x;
  const int row_begin = blockIdx.x * rows_per_cta, row_end = min(N, row_begin + rows_per_cta);
  const int nblk = (row_end - row_begin + RB - 1) / RB;
  const int stage_floats = RB * F;
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(NW));
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  auto issue = [&](int q) {
    const int st = q % NST;
    const int r0 = row_begin + q * RB;
    const unsigned bytes = (unsigned)(min(RB, row_end - r0) * F * 4);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[st])), "r"(bytes));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(stage + (size_t)st * stage_floats)),
                 "l"(x + (size_t)r0 * F), "r"(bytes), "r"(smem_u32(&full[st]))
                 : "memory");
  };
  auto wait = [&](unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(bar)),
        "r"(parity));
  };
  if (tid == 0)
    for (int q = 0; q < NST && q < nblk; ++q) issue(q);
  float acc = 0.f;
  for (int q = 0; q < nblk; ++q) {
    const int st = q % NST;
    const unsigned ph = (unsigned)((q / NST) & 1);
    wait(&full[st], ph);
    const int n = min(RB, row_end - (row_begin + q * RB)) * F;
    const float *sp = stage + (size_t)st * stage_floats;
    for (int i = tid; i < n; i += NW * 32) acc += sp[i] * sp[i];
    __syncwarp();
    if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[st])));
    if (tid == 0 && q + NST < nblk) {
      wait(&empty[st], ph);
      issue(q + NST);
    }
  }
  if (acc == 123.456f) out[0] = acc;
}


// pattern 4: rows warp+16j, items of 64 bins, 8 B per lane (256 B per warp load).  Row r starts at byte 1028 r: 8 B
// aligned for even r; odd rows load the pairs (2L+1, 2L+2) instead of (2L, 2L+1), which are 8 B aligned again
// (a converter would fetch the missing even element from the neighbouring lane with one shuffle).
__device__ __forceinline__ float2 ld_stream2(const float *p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
template <int DEPTH, int NW>
__global__ void __launch_bounds__(NW * 32, 1) pair_kernel(const float *x, float *out, int N, int F, int rows_per_cta) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_begin = blockIdx.x * rows_per_cta, row_end = min(N, row_begin + rows_per_cta);
  const int ntiles = (row_end - row_begin + 127) / 128, nch = 4;
  const int nitems = ntiles * nch;
  constexpr int RPW = 128 / NW;
  float2 v[DEPTH][RPW];
  float acc = 0.f;
  int li = 0;
  auto load = [&](float2 (&g)[RPW]) {
    const int t = li / nch, c = li % nch;
    const int row0 = row_begin + t * 128;
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
      const int row = row0 + warp + NW * j;
      g[j] = row < row_end ? ld_stream2(x + (size_t)row * F + 64 * c + 2 * lane + (row & 1)) : make_float2(0.f, 0.f);
    }
    ++li;
  };
#pragma unroll
  for (int d = 0; d < DEPTH; ++d) if (d < nitems) load(v[d]);
  for (int q0 = 0; q0 < nitems; q0 += DEPTH) {
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      if (q0 + d < nitems) {
#pragma unroll
        for (int j = 0; j < RPW; ++j) acc += v[d][j].x * v[d][j].x + v[d][j].y;
        if (li < nitems) load(v[d]);
      }
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

template <typename K>
void run(const char *name, K launch, size_t bytes) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) launch();
  cudaDeviceSynchronize();
  float best = 1e9f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  printf("%-44s %8.2f us  %7.1f GB/s  (%s)\n", name, best * 1e3f, bytes / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int N = 25600, F = 257, sms = 148;
  const int rpc = (N + sms - 1) / sms;
  float *x, *x2, *out, *flush;
  const size_t nb = (size_t)N * F * 4;
  cudaMalloc(&x, nb + 64); cudaMalloc(&x2, nb + 64); cudaMalloc(&out, 64); cudaMalloc(&flush, 256u << 20);
  cudaMemset(x, 0, nb); cudaMemset(x2, 0, nb);
  // note: inputs (26 MB) fit the 126 MB L2; flush between launches for DRAM numbers
  auto fl = [&]() { cudaMemsetAsync(flush, 1, 256u << 20); };
  run("chunk rows w+16j  DEPTH4 16w (flush)", [&]() { fl(); chunk_kernel<4, 0, 16><<<sms, 512>>>(x, out, N, F, rpc); }, nb);
#define R(NAME, ...) run(NAME, [&]() { __VA_ARGS__; }, nb)
  R("chunk rows w+16j  DEPTH4 16w (L2 warm)", chunk_kernel<4, 0, 16><<<sms, 512>>>(x, out, N, F, rpc));
  R("chunk rows w+16j  DEPTH8 16w (L2 warm)", chunk_kernel<8, 0, 16><<<sms, 512>>>(x, out, N, F, rpc));
  R("chunk rows 8w+j   DEPTH4 16w (L2 warm)", chunk_kernel<4, 1, 16><<<sms, 512>>>(x, out, N, F, rpc));
  R("chunk rows w+32j  DEPTH4 32w (L2 warm)", chunk_kernel<4, 0, 32><<<sms, 1024>>>(x, out, N, F, rpc));
  R("chunk rows w+32j  DEPTH8 32w (L2 warm)", chunk_kernel<8, 0, 32><<<sms, 1024>>>(x, out, N, F, rpc));
  R("flat float4 U4 16w (L2 warm)", flat_kernel<4, 16><<<sms, 512>>>(x, out, N, F, rpc));
  R("flat float4 U8 16w (L2 warm)", flat_kernel<8, 16><<<sms, 512>>>(x, out, N, F, rpc));
  R("flat float4 U8 32w (L2 warm)", flat_kernel<8, 32><<<sms, 1024>>>(x, out, N, F, rpc));
  run("flat float4 U8 32w (flush)", [&]() { fl(); flat_kernel<8, 32><<<sms, 1024>>>(x, out, N, F, rpc); }, nb);
  R("aligned windows w+16j DEPTH4 16w (L2 warm)", window_kernel<4, 16><<<sms, 512>>>(x, out, N, F, rpc));
  R("aligned windows w+16j DEPTH7 16w (L2 warm)", window_kernel<7, 16><<<sms, 512>>>(x, out, N, F, rpc));
  run("aligned windows w+16j DEPTH4 16w (flush)", [&]() { fl(); window_kernel<4, 16><<<sms, 512>>>(x, out, N, F, rpc); }, nb);
  R("pair rows w+16j float2 DEPTH2 16w (L2 warm)", pair_kernel<2, 16><<<sms, 512>>>(x, out, N, F, rpc));
  R("pair rows w+16j float2 DEPTH4 16w (L2 warm)", pair_kernel<4, 16><<<sms, 512>>>(x, out, N, F, rpc));
  run("pair rows w+16j float2 DEPTH2 16w (flush)", [&]() { fl(); pair_kernel<2, 16><<<sms, 512>>>(x, out, N, F, rpc); }, nb);
  run("pair rows w+16j float2 DEPTH4 16w (flush)", [&]() { fl(); pair_kernel<4, 16><<<sms, 512>>>(x, out, N, F, rpc); }, nb);
  run("chunk rows w+16j  DEPTH8 16w (flush)", [&]() { fl(); chunk_kernel<8, 0, 16><<<sms, 512>>>(x, out, N, F, rpc); }, nb);
  {
    const int rpc4 = (rpc + 3) / 4 * 4;   // 16 B aligned row blocks
    auto k1 = bulk_kernel<16, 8, 16>;
    auto k2 = bulk_kernel<8, 16, 16>;
    auto k3 = bulk_kernel<44, 4, 16>;
    cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 8 * 257 * 4);
    cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16 * 257 * 4);
    cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, 44 * 4 * 257 * 4);
    R("bulk copy 16 rows x 8 stages (L2 warm)", k1<<<sms, 512, 16 * 8 * 257 * 4>>>(x, out, N, F, rpc4));
    R("bulk copy 8 rows x 16 stages (L2 warm)", k2<<<sms, 512, 8 * 16 * 257 * 4>>>(x, out, N, F, rpc4));
    R("bulk copy 44 rows x 4 stages = whole range (L2 warm)", k3<<<sms, 512, 44 * 4 * 257 * 4>>>(x, out, N, F, rpc4));
    run("bulk copy 16 rows x 8 stages (flush)", [&]() { fl(); k1<<<sms, 512, 16 * 8 * 257 * 4>>>(x, out, N, F, rpc4); }, nb);
    run("bulk copy 44 rows x 4 stages (flush)", [&]() { fl(); k3<<<sms, 512, 44 * 4 * 257 * 4>>>(x, out, N, F, rpc4); }, nb);
  }
  run("memset 256MB alone (for flush cost)", [&]() { fl(); }, 256u << 20);
  return 0;
}
