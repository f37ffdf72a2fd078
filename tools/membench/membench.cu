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
  run("memset 256MB alone (for flush cost)", [&]() { fl(); }, 256u << 20);
  return 0;
}
