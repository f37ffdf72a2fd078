// LSTMCell step of the attention decoder (model/e2e_decoder.py:128 `self.decoder[0](ey, (z_list[0], c_list[0]))`,
// torch.nn.LSTMCell arithmetic, gate order i, f, g, o) as pieces that fit the decoder loop:
//   * the embedding half of the input product, W_ih[:, :Z] . embed(y_s) + b_ih + b_hh, does not depend on the recurrence:
//     ALL positions at once on the tensor-core GEMM before the loop (the caller);
//   * per step only the two batch-sized products that do depend on it -- context . W_ih[:, Z:]^T and h . W_hh^T --
//     (re2e_skinny_nt, accumulating into one gate buffer) and ONE fused pointwise kernel (this file);
//   * backward per step: this file's pointwise kernel (gate gradients, d c_prev) + two re2e_skinny_nn products
//     (d context, d h_prev); the weight gradients of all steps are two dense GEMMs after the loop.
#include "common.cuh"

namespace re2e {
namespace {

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }

// gates (B,4Z): in = the two recurrent products; out = the gate ACTIVATIONS (i, f, g, o), kept for the backward
__global__ void __launch_bounds__(256) lstm_pointwise_fwd_kernel(float *__restrict__ gates, const float *__restrict__ egate,
                                                                  const float *__restrict__ c_prev,
                                                                  float *__restrict__ c_out, float *__restrict__ h_out,
                                                                  int B, int Z) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * Z) return;
  const int m = idx / Z, u = idx - m * Z;
  float *g = gates + (size_t)m * 4 * Z + u;
  const float *e = egate ? egate + (size_t)m * 4 * Z + u : nullptr;
  const float pi = g[0] + (e ? e[0] : 0.0f), pf = g[Z] + (e ? e[Z] : 0.0f);
  const float pg = g[2 * Z] + (e ? e[2 * Z] : 0.0f), po = g[3 * Z] + (e ? e[3 * Z] : 0.0f);
  const float i = sigmoid_f(pi), f = sigmoid_f(pf), gg = tanhf(pg), o = sigmoid_f(po);
  const float c = f * (c_prev ? c_prev[idx] : 0.0f) + i * gg;
  g[0] = i; g[Z] = f; g[2 * Z] = gg; g[3 * Z] = o;
  c_out[idx] = c;
  h_out[idx] = o * tanhf(c);
}

// act (B,4Z) gate activations, c_new the cell state this step produced; dh / dc: gradients arriving at (h', c') (either may
// be NULL = zero).  dgates (B,4Z): gradient w.r.t. the PRE-activation gates; dc_prev (B,Z).
__global__ void __launch_bounds__(256) lstm_pointwise_bwd_kernel(const float *__restrict__ act, const float *__restrict__ c_prev,
                                                                  const float *__restrict__ c_new, const float *__restrict__ dh,
                                                                  const float *__restrict__ dc, float *__restrict__ dgates,
                                                                  float *__restrict__ dc_prev, int B, int Z) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * Z) return;
  const int m = idx / Z, u = idx - m * Z;
  const float *a = act + (size_t)m * 4 * Z + u;
  const float i = a[0], f = a[Z], g = a[2 * Z], o = a[3 * Z];
  const float tc = tanhf(c_new[idx]);
  const float dhv = dh ? dh[idx] : 0.0f;
  const float dct = (dc ? dc[idx] : 0.0f) + dhv * o * (1.0f - tc * tc);
  float *d = dgates + (size_t)m * 4 * Z + u;
  d[0] = dct * g * i * (1.0f - i);
  d[Z] = dct * (c_prev ? c_prev[idx] : 0.0f) * f * (1.0f - f);
  d[2 * Z] = dct * i * (1.0f - g * g);
  d[3 * Z] = dhv * tc * o * (1.0f - o);
  dc_prev[idx] = dct * f;
}

// out[m,n] (+)= sum_k X[m,k] * W[n,k]   for a batch-sized M (the decoder's per-position products: M = utterances or beam
// rows, N up to 4Z = 1200, K up to 4Z).  One CTA per 8 output columns (150 CTAs for N = 1200: the 1.5 MB weight matrix is
// streamed once, by all SMs), warp <-> column, lane <-> row; the reduction runs in chunks of kBK columns staged in shared
// memory with 128-bit accesses on both sides: X rows at a pitch of kBK + 4 floats (conflict-free 128-bit reads with
// lane <-> row), the warp's W row as broadcast reads.  Four independent FMA chains per thread.
constexpr int kBMN = 8;        // columns (warps) per CTA
constexpr int kBK = 320;       // reduction chunk
constexpr int kBM = 32;        // rows per pass (lane <-> row)

__global__ void __launch_bounds__(kBMN * 32) batch_nt_kernel(const float *__restrict__ X, const float *__restrict__ W,
                                                            float *__restrict__ out, int M, int N, int K, int accumulate) {
  extern __shared__ __align__(16) float batch_smem[];
  float *x_s = batch_smem, *w_s = batch_smem + kBM * (kBK + 4);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.x * kBMN + warp;
  for (int m0 = 0; m0 < M; m0 += kBM) {
    const int rows = min(kBM, M - m0);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int k0 = 0; k0 < K; k0 += kBK) {
      const int kc = min(kBK, K - k0), kq = kc >> 2;          // K % 4 == 0
      __syncthreads();                                         // previous chunk consumed
      for (int i = tid; i < rows * kq; i += kBMN * 32) {
        const int r = i / kq, q = i - r * kq;
        *reinterpret_cast<float4 *>(x_s + r * (kBK + 4) + 4 * q) =
            __ldg(reinterpret_cast<const float4 *>(X + (size_t)(m0 + r) * K + k0) + q);
      }
      for (int i = tid; i < kBMN * kq; i += kBMN * 32) {
        const int r = i / kq, q = i - r * kq;
        const int nn = blockIdx.x * kBMN + r;
        *reinterpret_cast<float4 *>(w_s + r * kBK + 4 * q) =
            nn < N ? __ldg(reinterpret_cast<const float4 *>(W + (size_t)nn * K + k0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();
      if (lane < rows) {
        const float4 *xr = reinterpret_cast<const float4 *>(x_s + lane * (kBK + 4));
        const float4 *wr = reinterpret_cast<const float4 *>(w_s + warp * kBK);
#pragma unroll 4
        for (int q = 0; q < kq; ++q) {
          const float4 x = xr[q], w = wr[q];
          a0 = fmaf(x.x, w.x, a0); a1 = fmaf(x.y, w.y, a1); a2 = fmaf(x.z, w.z, a2); a3 = fmaf(x.w, w.w, a3);
        }
      }
    }
    if (lane < rows && n < N) {
      float *o = out + (size_t)(m0 + lane) * N + n;
      const float v = (a0 + a1) + (a2 + a3);
      *o = accumulate ? *o + v : v;
    }
  }
}

}  // namespace
}  // namespace re2e

using namespace re2e;

extern "C" int re2e_batch_nt(const float *X, const float *W, float *out, int M, int N, int K, int accumulate,
                             void *stream) {
  RE2E_CHECK_ARG(X && W && out && M > 0 && N > 0 && K > 0);
  if ((K & 3) || !aligned16(X) || !aligned16(W)) return RE2E_E_UNSUPPORTED;
  const size_t smem = sizeof(float) * (kBM * (kBK + 4) + kBMN * kBK);
  int rc = ensure_smem(reinterpret_cast<const void *>(batch_nt_kernel), smem);
  if (rc != RE2E_OK) return rc;
  batch_nt_kernel<<<(N + kBMN - 1) / kBMN, kBMN * 32, smem, static_cast<cudaStream_t>(stream)>>>(X, W, out, M, N, K,
                                                                                               accumulate);
  count_launch();
  return launch_status();
}

extern "C" int re2e_lstm_pointwise_fwd(float *gates, const float *egate, const float *c_prev, float *c_out, float *h_out,
                                       int B, int Z, void *stream) {
  RE2E_CHECK_ARG(gates && c_out && h_out && B > 0 && Z > 0);
  const int n = B * Z;
  lstm_pointwise_fwd_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(gates, egate, c_prev, c_out,
                                                                                          h_out, B, Z);
  count_launch();
  return launch_status();
}

extern "C" int re2e_lstm_pointwise_bwd(const float *act, const float *c_prev, const float *c_new, const float *dh,
                                       const float *dc, float *dgates, float *dc_prev, int B, int Z, void *stream) {
  RE2E_CHECK_ARG(act && c_new && dgates && dc_prev && B > 0 && Z > 0);
  const int n = B * Z;
  lstm_pointwise_bwd_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(act, c_prev, c_new, dh, dc,
                                                                                          dgates, dc_prev, B, Z);
  count_launch();
  return launch_status();
}
