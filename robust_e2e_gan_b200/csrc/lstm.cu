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

}  // namespace
}  // namespace re2e

using namespace re2e;

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
