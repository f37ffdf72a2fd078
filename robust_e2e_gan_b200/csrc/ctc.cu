// CTC loss (alpha/beta recursion) and its gradient, log-softmax, best path and batched prefix scoring
// for sm_100a.  Replaces the warp_ctc.CTCLoss call of model/e2e_ctc.py:30,63 (third-party
// warpctc_pytorch; algorithm restated from Graves et al. 2006), F.log_softmax at :75 and the
// per-hypothesis numpy loop of CTCPrefixScore.__call__ at :109-155.
//
// Three kernels per loss evaluation, all HBM-bound on the (B,Th,V) logits:
//   ctc_lse_kernel   : one warp per frame; 128-bit streaming loads, online log-sum-exp per lane,
//                      warp-shuffle merge; gathers the S = 2U+1 label log-probs lp[b,t,s] (compact).
//   ctc_ab_kernel    : one CTA per utterance; alpha (forward in t) and beta (backward in t) run
//                      concurrently in the two halves of the CTA on separate named barriers; the
//                      lattice column lives in shared memory, lp is register-prefetched 8 frames
//                      ahead.  The column is renormalised every 8 frames and the running offset kept
//                      in fp64, so |alpha| stays O(100) and fp32 rounding does not grow with T
//                      (plain fp32 log-domain CTC loses ~1e-4 at |alpha| ~ 1e3).
//   ctc_grad_kernel  : one warp per frame; grad = gscale * (softmax - occupancy), written once.
// Algorithmic bytes: 4*V per valid frame (lse) + 2*4*V per frame (grad read + write).
#include <math_constants.h>

#include "common.cuh"

namespace re2e {
namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kRenorm = 8;

struct CtcWs {
  unsigned int *counter;
  float *lse;      // (B,Th)
  float *lpmax;    // (B,Th)  max_s lp[b,t,s]: per-frame normaliser of the lattice column
  double *offA;    // (B,Th)
  double *offB;    // (B,Th)
  double *nll_d;   // (B) negative log-likelihood in fp64 (float nll loses 1e-4 at |nll| ~ 1e3)
  float *lpc;      // (B,Th,Smax)
  float *alpha;    // (B,Th,Smax)
  float *beta;     // (B,Th,Smax)
  size_t bytes;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline CtcWs carve(void *ws, int B, int Th, int Smax) {
  CtcWs w;
  char *p = static_cast<char *>(ws);
  size_t off = 0;
  w.counter = reinterpret_cast<unsigned int *>(p + off); off += 256;
  w.offA = reinterpret_cast<double *>(p + off); off += align_up(sizeof(double) * (size_t)B * Th, 256);
  w.offB = reinterpret_cast<double *>(p + off); off += align_up(sizeof(double) * (size_t)B * Th, 256);
  w.nll_d = reinterpret_cast<double *>(p + off); off += align_up(sizeof(double) * (size_t)B, 256);
  w.lse = reinterpret_cast<float *>(p + off); off += align_up(sizeof(float) * (size_t)B * Th, 256);
  w.lpmax = reinterpret_cast<float *>(p + off); off += align_up(sizeof(float) * (size_t)B * Th, 256);
  size_t lat = align_up(sizeof(float) * (size_t)B * Th * Smax, 256);
  w.lpc = reinterpret_cast<float *>(p + off); off += lat;
  w.alpha = reinterpret_cast<float *>(p + off); off += lat;
  w.beta = reinterpret_cast<float *>(p + off); off += lat;
  w.bytes = off;
  return w;
}

// merge two (max, sum-of-exp) pairs
__device__ __forceinline__ void lse_merge(float &m, float &s, float m2, float s2) {
  float mn = fmaxf(m, m2);
  if (mn == -CUDART_INF_F) { m = mn; s = 0.f; return; }
  s = s * __expf(m - mn) + s2 * __expf(m2 - mn);
  m = mn;
}
__device__ __forceinline__ void lse_push(float &m, float &s, float x) {
  if (x > m) { s = s * __expf(m - x) + 1.0f; m = x; }
  else s += __expf(x - m);
}
__device__ __forceinline__ void lse_push4(float &m, float &s, float4 v) {
  float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
  if (mx > m) { s *= __expf(m - mx); m = mx; }
  s += __expf(v.x - m) + __expf(v.y - m) + __expf(v.z - m) + __expf(v.w - m);
}

// warp-cooperative (max, sumexp) over one row of V floats; row only 4 B aligned in general
__device__ __forceinline__ void warp_row_lse(const float *row, int V, int lane, float &m, float &s) {
  m = -CUDART_INF_F; s = 0.f;
  int head = (int)(((16u - (unsigned)(reinterpret_cast<uintptr_t>(row) & 15u)) & 15u) >> 2);
  if (head > V) head = V;
  if (lane < head) lse_push(m, s, ld_stream1(row + lane));
  const float *body = row + head;
  const int nvec = (V - head) >> 2;
  int i4 = lane;
  for (; i4 + 96 < nvec; i4 += 128) {
    float4 a = ld_stream4(body + 4 * i4), b = ld_stream4(body + 4 * (i4 + 32)),
           c = ld_stream4(body + 4 * (i4 + 64)), d = ld_stream4(body + 4 * (i4 + 96));
    lse_push4(m, s, a); lse_push4(m, s, b); lse_push4(m, s, c); lse_push4(m, s, d);
  }
  for (; i4 < nvec; i4 += 32) lse_push4(m, s, ld_stream4(body + 4 * i4));
  for (int i = head + 4 * nvec + lane; i < V; i += 32) lse_push(m, s, ld_stream1(row + i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    lse_merge(m, s, m2, s2);
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerCta * 32)
ctc_lse_kernel(const float *__restrict__ logits, long long stride_b, long long stride_t,
               const int32_t *__restrict__ labels, const int32_t *__restrict__ label_offs,
               const int32_t *__restrict__ label_lens, const int32_t *__restrict__ input_lens, int blank,
               float *__restrict__ lse, float *__restrict__ lpmax, float *__restrict__ lpc, int B, int Th,
               int V, int Smax) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (row >= (long long)B * Th) return;
  const int b = (int)(row / Th), t = (int)(row - (long long)b * Th);
  if (t >= min(Th, __ldg(input_lens + b))) return;  // padded frame: never read
  const float *x = logits + b * stride_b + t * stride_t;
  float m, s;
  warp_row_lse(x, V, lane, m, s);
  const float l = m + logf(s);
  if (lane == 0) lse[row] = l;
  const int U = __ldg(label_lens + b), S = 2 * U + 1;
  const int32_t *lab = labels + __ldg(label_offs + b);
  float *out = lpc + row * Smax;
  float mx = -CUDART_INF_F;
  for (int i = lane; i < S; i += 32) {
    int v = (i & 1) ? __ldg(lab + (i >> 1)) : blank;
    const float lpv = __ldg(x + v) - l;   // row was just streamed: L2 hit
    out[i] = lpv;
    mx = fmaxf(mx, lpv);
  }
  mx = warp_max(mx);
  if (lane == 0) lpmax[row] = mx;
}

// log(e^a + e^b + e^c) with -inf handling
__device__ __forceinline__ float lse3(float a, float b, float c) {
  float m = fmaxf(a, fmaxf(b, c));
  if (m == -CUDART_INF_F) return m;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// one CTA per utterance; blockDim = 2*Sp (Sp = Smax rounded to 32): first half alpha, second beta
__global__ void __launch_bounds__(1024)
ctc_ab_kernel(const float *__restrict__ lpc, const float *__restrict__ lpmax,
              const int32_t *__restrict__ labels,
              const int32_t *__restrict__ label_offs, const int32_t *__restrict__ label_lens,
              const int32_t *__restrict__ input_lens, int blank, float *__restrict__ alpha,
              float *__restrict__ beta, double *__restrict__ offA, double *__restrict__ offB,
              float *__restrict__ nll, double *__restrict__ nll_d, float *__restrict__ loss,
              unsigned int *counter, int B, int Th, int Smax, int Sp) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.x;
  const int half = threadIdx.x / Sp;     // warp-uniform (Sp % 32 == 0)
  const int s = threadIdx.x - half * Sp;
  const int pitch = Sp + 4;
  float *col = smem + half * 2 * pitch;  // [2][pitch], element s at index s+2
  float *wred = smem + 4 * pitch + half * 32;
  const int T = min(Th, max(0, __ldg(input_lens + b)));
  const int U = __ldg(label_lens + b), S = 2 * U + 1;
  const int32_t *lab = labels + __ldg(label_offs + b);
  const bool live = s < S;
  // skip transition allowed?  alpha: s-2 -> s ; beta: s+2 -> s
  bool skip = false;
  if (live && (s & 1)) {
    if (half == 0) skip = s >= 3 && __ldg(lab + (s >> 1)) != __ldg(lab + (s >> 1) - 1);
    else skip = s + 2 < S && __ldg(lab + (s >> 1)) != __ldg(lab + (s >> 1) + 1);
  }
  const float NEG = -CUDART_INF_F;
  if (s < 2) { col[s] = NEG; col[pitch + s] = NEG; col[Sp + 2 + s] = NEG; col[pitch + Sp + 2 + s] = NEG; }
  const float *lp_b = lpc + (size_t)b * Th * Smax;
  float *out_b = (half == 0 ? alpha : beta) + (size_t)b * Th * Smax;
  double *off_b = (half == 0 ? offA : offB) + (size_t)b * Th;
  const int nwarps = Sp >> 5;
  const int barid = 1 + half;

  const float *lpm_b = lpmax + (size_t)b * Th;
  float pf[kRenorm], pm[kRenorm];
#pragma unroll
  for (int j = 0; j < kRenorm; ++j) {
    int t = half ? T - 1 - j : j;
    pf[j] = (live && j < T) ? __ldg(lp_b + (size_t)t * Smax + s) : NEG;
    pm[j] = j < T ? __ldg(lpm_b + t) : 0.0f;
  }
  double off = 0.0;
  int cur = 0;
  for (int i0 = 0; i0 < T; i0 += kRenorm) {
#pragma unroll
    for (int j = 0; j < kRenorm; ++j) {
      const int i = i0 + j;
      if (i >= T) break;
      const int t = half ? T - 1 - i : i;
      // every frame is shifted by its own max label log-prob (kept in `off`), so a column only drifts
      // by the gap to the best label between the exact renormalisations below
      const float lpm = pm[j];
      const float lpv = pf[j] - lpm;
      off += (double)lpm;
      {  // prefetch frame i + kRenorm
        const int i2 = i + kRenorm;
        const int t2 = half ? T - 1 - i2 : i2;
        pf[j] = (live && i2 < T) ? __ldg(lp_b + (size_t)t2 * Smax + s) : NEG;
        pm[j] = i2 < T ? __ldg(lpm_b + t2) : 0.0f;
      }
      float v;
      if (i == 0) {
        const bool start = half == 0 ? (s < 2) : (s >= S - 2);
        v = (live && start) ? lpv : NEG;
      } else {
        const float *c = col + cur * pitch + 2;
        float a0 = c[s];
        float a1 = half == 0 ? c[s - 1] : c[s + 1];
        float a2 = skip ? (half == 0 ? c[s - 2] : c[s + 2]) : NEG;
        v = live ? lse3(a0, a1, a2) + lpv : NEG;
      }
      float *n = col + (cur ^ 1) * pitch + 2;
      if (j == kRenorm - 1) {
        // renormalise: subtract the column max, remember it in fp64
        float wm = warp_max(v);
        if ((threadIdx.x & 31) == 0) wred[s >> 5] = wm;
        named_bar(barid, Sp);
        float gm = NEG;
        for (int w = 0; w < nwarps; ++w) gm = fmaxf(gm, wred[w]);
        if (gm != NEG) { v -= gm; off += (double)gm; }
      }
      n[s] = v;
      if (live) out_b[(size_t)t * Smax + s] = v;
      if (s == 0) off_b[t] = off;
      cur ^= 1;
      named_bar(barid, Sp);
    }
  }
  if (half == 0 && s == 0) {
    double r;
    if (T == 0) r = (U == 0) ? 0.0 : (double)CUDART_INF_F;
    else {
      const float *c = col + cur * pitch + 2;
      float a = c[S - 1], bb = S > 1 ? c[S - 2] : NEG;
      float m = fmaxf(a, bb);
      if (m == NEG) r = (double)CUDART_INF_F;
      else r = -((double)m + (double)logf(expf(a - m) + expf(bb - m)) + off);
    }
    nll[b] = (float)r;
    nll_d[b] = r;
    __threadfence();
    unsigned int done = atomicAdd(counter, 1u);
    if (done == (unsigned)B - 1) {      // last utterance to finish: deterministic ordered sum
      __threadfence();
      double acc = 0.0;
      for (int k = 0; k < B; ++k) acc += *(volatile double *)(nll_d + k);
      loss[0] = (float)(acc / (double)B);
      *counter = 0u;
    }
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerCta * 32)
ctc_grad_kernel(const float *__restrict__ logits, long long stride_b, long long stride_t,
                const int32_t *__restrict__ labels, const int32_t *__restrict__ label_offs,
                const int32_t *__restrict__ label_lens, const int32_t *__restrict__ input_lens, int blank,
                const double *__restrict__ nll_d, const float *__restrict__ grad_out,
                const float *__restrict__ lse, const float *__restrict__ lpc,
                const float *__restrict__ alpha, const float *__restrict__ beta,
                const double *__restrict__ offA, const double *__restrict__ offB, float *__restrict__ grad,
                int B, int Th, int V, int Smax, int vec_ok) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (row >= (long long)B * Th) return;
  const int b = (int)(row / Th), t = (int)(row - (long long)b * Th);
  const float *x = logits + b * stride_b + t * stride_t;
  float *g = grad + b * stride_b + t * stride_t;
  const bool valid = t < min(Th, __ldg(input_lens + b));
  const float gs = (grad_out ? __ldg(grad_out) : 1.0f) / (float)B;
  const float l = valid ? lse[row] : 0.f;
  int head = (int)(((16u - (unsigned)(reinterpret_cast<uintptr_t>(g) & 15u)) & 15u) >> 2);
  if (!vec_ok || head > V) head = V;
  for (int i = lane; i < head; i += 32) g[i] = valid ? gs * __expf(ld_stream1(x + i) - l) : 0.f;
  const int nvec = (V - head) >> 2;
  const float *xb = x + head;
  float *gb = g + head;
  if (valid) {
    int i4 = lane;
    for (; i4 + 96 < nvec; i4 += 128) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ld_stream4(xb + 4 * (i4 + 32 * u));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float4 o = make_float4(gs * __expf(v[u].x - l), gs * __expf(v[u].y - l), gs * __expf(v[u].z - l),
                               gs * __expf(v[u].w - l));
        *reinterpret_cast<float4 *>(gb + 4 * (i4 + 32 * u)) = o;
      }
    }
    for (; i4 < nvec; i4 += 32) {
      float4 v = ld_stream4(xb + 4 * i4);
      *reinterpret_cast<float4 *>(gb + 4 * i4) = make_float4(
          gs * __expf(v.x - l), gs * __expf(v.y - l), gs * __expf(v.z - l), gs * __expf(v.w - l));
    }
  } else {
    for (int i4 = lane; i4 < nvec; i4 += 32) *reinterpret_cast<float4 *>(gb + 4 * i4) = make_float4(0, 0, 0, 0);
  }
  for (int i = head + 4 * nvec + lane; i < V; i += 32) g[i] = valid ? gs * __expf(ld_stream1(x + i) - l) : 0.f;
  if (!valid) return;
  const double nl = nll_d[b];
  if (!isfinite(nl)) return;  // infeasible labelling: leave softmax only
  __syncwarp();               // row stores above are ordered before the corrections below
  const int U = __ldg(label_lens + b), S = 2 * U + 1;
  const int32_t *lab = labels + __ldg(label_offs + b);
  const double base = offA[row] + offB[row] + nl;
  const size_t lo = (size_t)row * Smax;
  for (int i = lane; i < S; i += 32) {
    float a = alpha[lo + i], bt = beta[lo + i];
    if (a == -CUDART_INF_F || bt == -CUDART_INF_F) continue;
    float occ = expf((float)((double)a + (double)bt - (double)lpc[lo + i] + base));
    int v = (i & 1) ? __ldg(lab + (i >> 1)) : blank;
    atomicAdd(g + v, -gs * occ);
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerCta * 32)
log_softmax_kernel(const float *__restrict__ x, float *__restrict__ out, int32_t *__restrict__ best,
                   long long rows, int V) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float *xr = x + row * V;
  float *o = out + row * V;
  float m = -CUDART_INF_F, s = 0.f;
  int am = 0x7fffffff;
  float av = -CUDART_INF_F;
  for (int i = lane; i < V; i += 32) {
    float v = xr[i];
    lse_push(m, s, v);
    if (v > av) { av = v; am = i; }
  }
#pragma unroll
  for (int of = 16; of > 0; of >>= 1) {
    float m2 = __shfl_xor_sync(0xffffffffu, m, of), s2 = __shfl_xor_sync(0xffffffffu, s, of);
    lse_merge(m, s, m2, s2);
    float av2 = __shfl_xor_sync(0xffffffffu, av, of);
    int am2 = __shfl_xor_sync(0xffffffffu, am, of);
    if (av2 > av || (av2 == av && am2 < am)) { av = av2; am = am2; }
  }
  const float l = m + logf(s);
  for (int i = lane; i < V; i += 32) o[i] = xr[i] - l;
  if (best && lane == 0) best[row] = am;
}

// ---------------------------------------------------------------------------------------------
// Batched prefix scoring: one warp per (hypothesis h, candidate j); sequential over t.
__device__ __forceinline__ float logaddexpf_(float a, float b) {
  float m = fmaxf(a, b), d = fminf(a, b) - m;
  return m + log1pf(expf(d));
}
__global__ void __launch_bounds__(128)
ctc_prefix_kernel(const float *__restrict__ lpz, const float *__restrict__ r_prev,
                  const int32_t *__restrict__ cs, const int32_t *__restrict__ last,
                  const int32_t *__restrict__ out_len, float *__restrict__ log_psi,
                  float *__restrict__ r_new, int T, int V, int H, int Cc, int blank, int eos) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * Cc) return;
  const int h = idx / Cc;
  const int c = __ldg(cs + idx);
  const float LZ = -10000000000.0f;
  const int ol = __ldg(out_len + h);
  const float *rp = r_prev + (size_t)h * T * 2;
  float *rn = r_new + (size_t)idx * T * 2;
  const bool same = ol > 0 && c == __ldg(last + h);
  const int start = ol > 1 ? ol : 1;
  for (int t = 0; t < start - 1; ++t) { rn[2 * t] = LZ; rn[2 * t + 1] = LZ; }
  float rn_prev, rb_prev;
  if (ol == 0) { rn_prev = __ldg(lpz + c); rb_prev = LZ; }
  else { rn_prev = LZ; rb_prev = LZ; }
  rn[2 * (start - 1)] = rn_prev;
  rn[2 * (start - 1) + 1] = rb_prev;
  float psi = rn_prev;
  for (int t = start; t < T; ++t) {
    const float pn = rp[2 * (t - 1)], pb = rp[2 * (t - 1) + 1];
    const float phi = same ? pb : logaddexpf_(pn, pb);
    const float xc = __ldg(lpz + (size_t)t * V + c), xb = __ldg(lpz + (size_t)t * V + blank);
    const float nn = logaddexpf_(rn_prev, phi) + xc;
    const float nb = logaddexpf_(rn_prev, rb_prev) + xb;
    psi = logaddexpf_(psi, phi + xc);
    rn[2 * t] = nn; rn[2 * t + 1] = nb;
    rn_prev = nn; rb_prev = nb;
  }
  if (c == eos) psi = logaddexpf_(rp[2 * (T - 1)], rp[2 * (T - 1) + 1]);
  log_psi[idx] = psi;
}

}  // namespace
}  // namespace re2e

using namespace re2e;

extern "C" size_t re2e_ctc_ws_bytes(int B, int Th, int V, int Umax) {
  (void)V;
  if (B <= 0 || Th <= 0 || Umax < 0) return 0;
  return carve(nullptr, B, Th, 2 * Umax + 1).bytes;
}

extern "C" int re2e_ctc_loss_fwd(const float *logits, long long stride_b, long long stride_t,
                                 const int32_t *labels, const int32_t *label_offs,
                                 const int32_t *label_lens, const int32_t *input_lens, int blank,
                                 float *nll, float *loss, void *ws, size_t ws_bytes, int B, int Th, int V,
                                 int Umax, void *stream) {
  RE2E_CHECK_ARG(logits && labels && label_offs && label_lens && input_lens && nll && loss && ws);
  RE2E_CHECK_ARG(B > 0 && Th > 0 && V > 0 && Umax >= 0 && blank >= 0 && blank < V);
  const int Smax = 2 * Umax + 1;
  const int Sp = (Smax + 31) & ~31;
  if (2 * Sp > 1024) return RE2E_E_UNSUPPORTED;
  CtcWs w = carve(ws, B, Th, Smax);
  if (ws_bytes < w.bytes) return RE2E_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RE2E_CUDA(cudaMemsetAsync(w.counter, 0, 256, st));
  const long long rows = (long long)B * Th;
  const int grid = (int)((rows + kWarpsPerCta - 1) / kWarpsPerCta);
  ctc_lse_kernel<<<grid, kWarpsPerCta * 32, 0, st>>>(logits, stride_b, stride_t, labels, label_offs,
                                                     label_lens, input_lens, blank, w.lse, w.lpmax, w.lpc, B,
                                                     Th, V, Smax);
  count_launch();
  int rc = launch_status();
  if (rc != RE2E_OK) return rc;
  const size_t smem = sizeof(float) * (4 * (size_t)(Sp + 4) + 64);
  ctc_ab_kernel<<<B, 2 * Sp, smem, st>>>(w.lpc, w.lpmax, labels, label_offs, label_lens, input_lens, blank, w.alpha,
                                         w.beta, w.offA, w.offB, nll, w.nll_d, loss, w.counter, B, Th, Smax, Sp);
  count_launch();
  return launch_status();
}

extern "C" int re2e_ctc_loss_bwd(const float *logits, long long stride_b, long long stride_t,
                                 const int32_t *labels, const int32_t *label_offs,
                                 const int32_t *label_lens, const int32_t *input_lens, int blank,
                                 const float *nll, const float *grad_out, const void *ws, size_t ws_bytes,
                                 float *grad, int B, int Th, int V, int Umax, void *stream) {
  RE2E_CHECK_ARG(logits && labels && label_offs && label_lens && input_lens && nll && ws && grad);
  RE2E_CHECK_ARG(B > 0 && Th > 0 && V > 0 && Umax >= 0);
  const int Smax = 2 * Umax + 1;
  CtcWs w = carve(const_cast<void *>(ws), B, Th, Smax);
  if (ws_bytes < w.bytes) return RE2E_E_WORKSPACE;
  const long long rows = (long long)B * Th;
  const int grid = (int)((rows + kWarpsPerCta - 1) / kWarpsPerCta);
  const int vec_ok = ((reinterpret_cast<uintptr_t>(logits) ^ reinterpret_cast<uintptr_t>(grad)) & 15u) == 0;
  ctc_grad_kernel<<<grid, kWarpsPerCta * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, stride_b, stride_t, labels, label_offs, label_lens, input_lens, blank, w.nll_d, grad_out, w.lse,
      w.lpc, w.alpha, w.beta, w.offA, w.offB, grad, B, Th, V, Smax, vec_ok);
  count_launch();
  return launch_status();
}

extern "C" int re2e_log_softmax(const float *logits, float *out, int32_t *best, long long rows, int V,
                                void *stream) {
  RE2E_CHECK_ARG(logits && out && rows > 0 && V > 0);
  const int grid = (int)((rows + kWarpsPerCta - 1) / kWarpsPerCta);
  log_softmax_kernel<<<grid, kWarpsPerCta * 32, 0, static_cast<cudaStream_t>(stream)>>>(logits, out, best,
                                                                                       rows, V);
  count_launch();
  return launch_status();
}

extern "C" int re2e_ctc_prefix_score(const float *lpz, const float *r_prev, const int32_t *cs,
                                     const int32_t *last, const int32_t *out_len, float *log_psi,
                                     float *r_new, int T, int V, int H, int Ccand, int blank, int eos,
                                     void *stream) {
  RE2E_CHECK_ARG(lpz && r_prev && cs && last && out_len && log_psi && r_new);
  RE2E_CHECK_ARG(T > 0 && V > 0 && H > 0 && Ccand > 0);
  const int n = H * Ccand;
  ctc_prefix_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      lpz, r_prev, cs, last, out_len, log_psi, r_new, T, V, H, Ccand, blank, eos);
  count_launch();
  return launch_status();
}
