// Location-aware attention step (AttLoc.forward, model/e2e_attention.py:258-299) for sm_100a,
// forward and backward, one thread-block-cluster kernel each.
//
// Mapping.  One CLUSTER of CL CTAs per utterance b; CTA `rank` owns the encoder frames
// [t0,t1) = rank*ceil(Th/CL) ....  The two big operands of a step, pre[b,t0:t1,:] and
// enc_h[b,t0:t1,:], are contiguous in HBM, so they are fetched with 1-D bulk async copies (TMA,
// cp.async.bulk -> SASS UBLKCP) in chunks of 8 frames into a shared-memory ring, each chunk
// signalled on its own mbarrier; the loads are all in flight while the CTA computes the location
// convolution.  Softmax statistics and the context vector are combined across the CTAs of the
// cluster through distributed shared memory (ld.shared::cluster) -- no global round trip and no
// second kernel.  In the backward the gradient w.r.t. pre is formed in place in the ring and
// accumulated into HBM by the TMA unit itself (cp.reduce.async.bulk .add.f32): the SM never reads
// d_pre.  Algorithmic bytes per step (fp32): fwd 4*B*Th*(A+D); bwd 4*B*Th*(A+D) + 4*B*Th*A.
//
// Thread mapping inside a CTA (256 threads = 8 warps): for the energy part a row (frame) is
// handled by a PAIR of warps, each owning half of the A attention channels, lane <-> channel
// (a = half*A/2 + lane + 32j): W_att (A x C) then lives in registers (APL*CP floats per lane).
#include <math_constants.h>

#include "common.cuh"

namespace re2e {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kChunkRows = 8;
constexpr int kDplMax = 16;   // D <= 512
constexpr int kMaxStages = 30;

__host__ __device__ inline int round4(int v) { return (v + 3) & ~3; }

struct AttGeom {
  int tloc_max, nch, ns, stage_floats, Thp, CKp;
};

struct AttFwdParams {
  const float *pre, *enc, *att_prev, *dec_proj, *W_att, *W_conv, *gvec, *gvec_b;
  float scaling;
  float *c, *w, *conv;
  int B, Th, D, A, C, K;
  AttGeom g;
};

struct AttBwdParams {
  const float *dc, *dw, *pre, *enc, *att_prev, *w, *dec_proj, *conv, *W_att, *W_conv, *gvec;
  float scaling;
  float *d_pre, *d_decproj, *d_att_prev, *dW_att, *dW_conv, *dgvec, *dgvec_b;
  int accumulate_pre;
  int B, Th, D, A, C, K;
  AttGeom g;
};

// issue chunk `q` of the [first | second] operand sequence into ring stage q % ns
__device__ __forceinline__ void issue_chunk(int q, const AttGeom &g, const float *first, int wfirst,
                                            const float *second, int wsecond, int b, int Th, int t0,
                                            int t1, float *stages, uint64_t *full) {
  const int st = q % g.ns;
  const bool is_first = q < g.nch;
  const int qq = is_first ? q : q - g.nch;
  const int r0 = t0 + kChunkRows * qq;
  const int rows = min(kChunkRows, t1 - r0);
  const int width = is_first ? wfirst : wsecond;
  const float *src = (is_first ? first : second) + ((size_t)b * Th + r0) * width;
  const uint32_t bytes = (uint32_t)rows * width * 4u;
  mbar_expect_tx(&full[st], bytes);
  bulk_g2s(stages + (size_t)st * g.stage_floats, src, bytes, &full[st]);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int APL, int CP>
__global__ void __launch_bounds__(kThreads, 1) attloc_fwd_kernel(const AttFwdParams p) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const AttGeom g = p.g;
  const int Th = p.Th, D = p.D, A = p.A, C = p.C, K = p.K, filts = (p.K - 1) / 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, half = warp & 1, rw = warp >> 1;
  const int CL = (int)cluster_nctarank(), rank = (int)cluster_ctarank();
  const int b = blockIdx.x / CL;
  const int t0 = min(Th, rank * g.tloc_max), t1 = min(Th, t0 + g.tloc_max), tloc = t1 - t0;
  const int nch = (tloc + kChunkRows - 1) / kChunkRows;  // chunks actually used by this CTA
  AttGeom gl = g;
  gl.nch = nch;
  const int total = 2 * nch;

  uint64_t *full = reinterpret_cast<uint64_t *>(smraw);
  float *stages = reinterpret_cast<float *>(smraw + 256);
  float *ap_s = stages + (size_t)g.ns * g.stage_floats;
  float *wc_s = ap_s + g.Thp;
  float *conv_s = wc_s + g.CKp;
  float *e_s = conv_s + g.tloc_max * CP;
  float *p_s = e_s + 2 * g.tloc_max;
  float *cred = p_s + g.tloc_max;
  float *cpart = cred + kWarps * D;
  float *xch = cpart + round4(D);

  if (tid == 0) {
    for (int i = 0; i < g.ns; ++i) mbar_init(&full[i], 1);
    mbar_fence_init();
    const int first = total < g.ns ? total : g.ns;
    for (int q = 0; q < first; ++q) issue_chunk(q, gl, p.pre, A, p.enc, D, b, Th, t0, t1, stages, full);
  }
  // stage the small operands
  for (int i = tid; i < Th; i += kThreads) ap_s[i] = __ldg(p.att_prev + (size_t)b * Th + i);
  for (int i = tid; i < C * K; i += kThreads) wc_s[i] = __ldg(p.W_conv + i);
  float Watt[APL][CP], dp[APL], gv[APL];
#pragma unroll
  for (int j = 0; j < APL; ++j) {
    const int a = half * (A / 2) + lane + 32 * j;
    dp[j] = __ldg(p.dec_proj + (size_t)b * A + a);
    gv[j] = __ldg(p.gvec + a);
#pragma unroll
    for (int c = 0; c < CP; ++c) Watt[j][c] = c < C ? __ldg(p.W_att + (size_t)a * C + c) : 0.0f;
  }
  __syncthreads();

  // ---- location convolution: conv[t,c] = sum_k Wc[c,k] * att_prev[t + k - filts]  (zero padded)
  for (int item = tid; item < tloc * CP; item += kThreads) {
    const int tl = item / CP, c = item - tl * CP;
    float acc = 0.0f;
    if (c < C) {
      const int t = t0 + tl;
      const int klo = max(0, filts - t), khi = min(K, Th + filts - t);
      const float *wr = wc_s + c * K;
      const float *ar = ap_s + (t - filts);
      int k = klo;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (; k + 3 < khi; k += 4) {
        a0 = fmaf(wr[k], ar[k], a0);
        a1 = fmaf(wr[k + 1], ar[k + 1], a1);
        a2 = fmaf(wr[k + 2], ar[k + 2], a2);
        a3 = fmaf(wr[k + 3], ar[k + 3], a3);
      }
      for (; k < khi; ++k) a0 = fmaf(wr[k], ar[k], a0);
      acc = (a0 + a1) + (a2 + a3);
      if (p.conv) p.conv[((size_t)b * Th + t) * C + c] = acc;
    }
    conv_s[item] = acc;
  }
  __syncthreads();

  // ---- energies: e[t] = g . tanh(W_att conv[t] + pre[t] + dec_proj) (+ g_b), two warps per frame
  for (int q = 0; q < nch; ++q) {
    const int st = q % g.ns;
    mbar_wait(&full[st], (uint32_t)((q / g.ns) & 1));
    const float *tile = stages + (size_t)st * g.stage_floats;
    const int rows = min(kChunkRows, tloc - kChunkRows * q);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = rw + 4 * rr;
      if (r < rows) {
        const int tl = kChunkRows * q + r;
        float cv[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) cv[c] = conv_s[tl * CP + c];
        const float *row = tile + r * A + half * (A / 2) + lane;
        float part = 0.0f;
#pragma unroll
        for (int j = 0; j < APL; ++j) {
          float u = dp[j] + row[32 * j];
#pragma unroll
          for (int c = 0; c < CP; ++c) u = fmaf(Watt[j][c], cv[c], u);
          part = fmaf(gv[j], tanh_fast(u), part);
        }
        part = warp_sum(part);
        if (lane == 0) e_s[half * g.tloc_max + tl] = part;
      }
    }
    if (q + g.ns < total) {  // ring wrap: everyone is done with this stage, refill it
      __syncthreads();
      if (tid == 0) issue_chunk(q + g.ns, gl, p.pre, A, p.enc, D, b, Th, t0, t1, stages, full);
    }
  }
  __syncthreads();

  // ---- local softmax statistics over this CTA's frames (the softmax spans ALL Th frames,
  //      padding included: e2e_attention.py:282-288 applies no length mask)
  if (warp == 0) {
    const float gb = __ldg(p.gvec_b);
    float m = -CUDART_INF_F;
    for (int tl = lane; tl < tloc; tl += 32) {
      float ev = p.scaling * (e_s[tl] + e_s[g.tloc_max + tl] + gb);
      e_s[tl] = ev;
      m = fmaxf(m, ev);
    }
    m = warp_max(m);
    float s = 0.0f;
    for (int tl = lane; tl < tloc; tl += 32) {
      float pv = expf(e_s[tl] - m);
      p_s[tl] = pv;
      s += pv;
    }
    s = warp_sum(s);
    if (lane == 0) { xch[0] = m; xch[1] = s; }
  }
  __syncthreads();

  // ---- un-normalised context over this CTA's frames: one warp per frame, lane <-> d
  float acc[kDplMax];
#pragma unroll
  for (int j = 0; j < kDplMax; ++j) acc[j] = 0.0f;
  for (int q = nch; q < total; ++q) {
    const int st = q % g.ns;
    mbar_wait(&full[st], (uint32_t)((q / g.ns) & 1));
    const float *tile = stages + (size_t)st * g.stage_floats;
    const int qq = q - nch;
    const int rows = min(kChunkRows, tloc - kChunkRows * qq);
    if (warp < rows) {
      const float pw = p_s[kChunkRows * qq + warp];
      const float *row = tile + warp * D + lane;
#pragma unroll
      for (int j = 0; j < kDplMax; ++j)
        if (lane + 32 * j < D) acc[j] = fmaf(pw, row[32 * j], acc[j]);
    }
    if (q + g.ns < total) {
      __syncthreads();
      if (tid == 0) issue_chunk(q + g.ns, gl, p.pre, A, p.enc, D, b, Th, t0, t1, stages, full);
    }
  }
#pragma unroll
  for (int j = 0; j < kDplMax; ++j)
    if (lane + 32 * j < D) cred[warp * D + lane + 32 * j] = acc[j];
  __syncthreads();
  for (int d = tid; d < D; d += kThreads) {
    float s = 0.0f;
#pragma unroll
    for (int w8 = 0; w8 < kWarps; ++w8) s += cred[w8 * D + d];
    cpart[d] = s;
  }
  // ---- combine across the cluster through distributed shared memory
  cluster_sync_all();
  float M = -CUDART_INF_F;
  for (int r = 0; r < CL; ++r) M = fmaxf(M, dsmem_ld(dsmem_addr(xch, r)));
  float S = 0.0f;
  for (int r = 0; r < CL; ++r) {
    const float mr = dsmem_ld(dsmem_addr(xch, r)), sr = dsmem_ld(dsmem_addr(xch + 1, r));
    S += sr * expf(mr - M);
  }
  const float inv = 1.0f / S;
  const float mine = expf(xch[0] - M) * inv;
  for (int tl = tid; tl < tloc; tl += kThreads) p.w[(size_t)b * Th + t0 + tl] = p_s[tl] * mine;
  const int dper = (D + CL - 1) / CL;
  for (int d = rank * dper + tid; d < min(D, (rank + 1) * dper); d += kThreads) {
    float s = 0.0f;
    for (int r = 0; r < CL; ++r) {
      const float mr = dsmem_ld(dsmem_addr(xch, r));
      s += dsmem_ld(dsmem_addr(cpart + d, r)) * expf(mr - M);
    }
    p.c[(size_t)b * D + d] = s * inv;
  }
  cluster_sync_all();  // nobody leaves while a peer may still read its shared memory
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <int APL, int CP>
__global__ void __launch_bounds__(kThreads, 1) attloc_bwd_kernel(const AttBwdParams p) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const AttGeom g = p.g;
  const int Th = p.Th, D = p.D, A = p.A, C = p.C, K = p.K, filts = (p.K - 1) / 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, half = warp & 1, rw = warp >> 1;
  const int CL = (int)cluster_nctarank(), rank = (int)cluster_ctarank();
  const int b = blockIdx.x / CL;
  const int t0 = min(Th, rank * g.tloc_max), t1 = min(Th, t0 + g.tloc_max), tloc = t1 - t0;
  const int nch = (tloc + kChunkRows - 1) / kChunkRows;
  AttGeom gl = g;
  gl.nch = nch;
  const int total = 2 * nch;

  uint64_t *full = reinterpret_cast<uint64_t *>(smraw);
  float *stages = reinterpret_cast<float *>(smraw + 256);
  float *ap_s = stages + (size_t)g.ns * g.stage_floats;   // Thp
  float *wc_s = ap_s + g.Thp;                             // CKp
  float *conv_s = wc_s + g.CKp;                           // tloc_max*CP
  float *w_s = conv_s + g.tloc_max * CP;                  // tloc_max
  float *dwt_s = w_s + g.tloc_max;                        // tloc_max
  float *de_s = dwt_s + g.tloc_max;                       // tloc_max
  float *dcv_p = de_s + g.tloc_max;                       // 2*tloc_max*CP
  float *dcvT = dcv_p + 2 * g.tloc_max * CP;              // CP*Thp   (filled by every CTA of the cluster)
  float *dWatt_s = dcvT + CP * g.Thp;                     // A*CP
  float *ddp_s = dWatt_s + A * CP;                        // A
  float *dgv_s = ddp_s + A;                               // A
  float *xch = dgv_s + A;                                 // 4

  if (tid == 0) {
    for (int i = 0; i < g.ns; ++i) mbar_init(&full[i], 1);
    mbar_fence_init();
    const int first = total < g.ns ? total : g.ns;
    for (int q = 0; q < first; ++q) issue_chunk(q, gl, p.enc, D, p.pre, A, b, Th, t0, t1, stages, full);
  }
  for (int i = tid; i < Th; i += kThreads) ap_s[i] = __ldg(p.att_prev + (size_t)b * Th + i);
  for (int i = tid; i < C * K; i += kThreads) wc_s[i] = __ldg(p.W_conv + i);
  for (int i = tid; i < tloc * CP; i += kThreads) {
    const int tl = i / CP, c = i - tl * CP;
    conv_s[i] = c < C ? __ldg(p.conv + ((size_t)b * Th + t0 + tl) * C + c) : 0.0f;
  }
  for (int i = tid; i < tloc; i += kThreads) w_s[i] = __ldg(p.w + (size_t)b * Th + t0 + i);
  for (int i = tid; i < A * CP; i += kThreads) dWatt_s[i] = 0.0f;
  for (int i = tid; i < A; i += kThreads) { ddp_s[i] = 0.0f; dgv_s[i] = 0.0f; }
  float dcr[kDplMax];
#pragma unroll
  for (int j = 0; j < kDplMax; ++j)
    dcr[j] = (p.dc && lane + 32 * j < D) ? __ldg(p.dc + (size_t)b * D + lane + 32 * j) : 0.0f;
  __syncthreads();

  // ---- dwt[t] = dw[t] + enc_h[t,:] . dc      (gradient reaching w[t])
  for (int q = 0; q < nch; ++q) {
    const int st = q % g.ns;
    mbar_wait(&full[st], (uint32_t)((q / g.ns) & 1));
    const float *tile = stages + (size_t)st * g.stage_floats;
    const int rows = min(kChunkRows, tloc - kChunkRows * q);
    if (warp < rows) {
      const int tl = kChunkRows * q + warp;
      const float *row = tile + warp * D + lane;
      float dot = 0.0f;
#pragma unroll
      for (int j = 0; j < kDplMax; ++j)
        if (lane + 32 * j < D) dot = fmaf(dcr[j], row[32 * j], dot);
      dot = warp_sum(dot);
      if (lane == 0) dwt_s[tl] = dot + (p.dw ? __ldg(p.dw + (size_t)b * Th + t0 + tl) : 0.0f);
    }
    if (q + g.ns < total) {
      __syncthreads();
      if (tid == 0) issue_chunk(q + g.ns, gl, p.enc, D, p.pre, A, b, Th, t0, t1, stages, full);
    }
  }
  __syncthreads();
  // ---- softmax backward needs sum_t w[t]*dwt[t] over ALL frames: cluster reduction via DSMEM
  if (warp == 0) {
    float s = 0.0f;
    for (int tl = lane; tl < tloc; tl += 32) s = fmaf(w_s[tl], dwt_s[tl], s);
    s = warp_sum(s);
    if (lane == 0) xch[0] = s;
  }
  cluster_sync_all();
  float Stot = 0.0f;
  for (int r = 0; r < CL; ++r) Stot += dsmem_ld(dsmem_addr(xch, r));
  for (int tl = tid; tl < tloc; tl += kThreads) de_s[tl] = p.scaling * w_s[tl] * (dwt_s[tl] - Stot);
  __syncthreads();
  if (warp == 0) {
    float s = 0.0f;
    for (int tl = lane; tl < tloc; tl += 32) s += de_s[tl];
    s = warp_sum(s);
    if (lane == 0 && tloc > 0) atomicAdd(p.dgvec_b, s);
  }

  // ---- through tanh: two warps per frame, lane <-> attention channel
  float Watt[APL][CP], dWatt[APL][CP], dp[APL], gv[APL], dgv[APL], ddp[APL];
#pragma unroll
  for (int j = 0; j < APL; ++j) {
    const int a = half * (A / 2) + lane + 32 * j;
    dp[j] = __ldg(p.dec_proj + (size_t)b * A + a);
    gv[j] = __ldg(p.gvec + a);
    dgv[j] = 0.0f;
    ddp[j] = 0.0f;
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      Watt[j][c] = c < C ? __ldg(p.W_att + (size_t)a * C + c) : 0.0f;
      dWatt[j][c] = 0.0f;
    }
  }
  for (int q = nch; q < total; ++q) {
    const int st = q % g.ns;
    mbar_wait(&full[st], (uint32_t)((q / g.ns) & 1));
    float *tile = stages + (size_t)st * g.stage_floats;
    const int qq = q - nch;
    const int rows = min(kChunkRows, tloc - kChunkRows * qq);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = rw + 4 * rr;
      if (r < rows) {
        const int tl = kChunkRows * qq + r;
        const float de = de_s[tl];
        float cv[CP], dcv[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) { cv[c] = conv_s[tl * CP + c]; dcv[c] = 0.0f; }
        float *row = tile + r * A + half * (A / 2) + lane;
#pragma unroll
        for (int j = 0; j < APL; ++j) {
          float u = dp[j] + row[32 * j];
#pragma unroll
          for (int c = 0; c < CP; ++c) u = fmaf(Watt[j][c], cv[c], u);
          const float x = tanh_fast(u);
          dgv[j] = fmaf(de, x, dgv[j]);
          const float dt = de * gv[j] * (1.0f - x * x);
          ddp[j] += dt;
          row[32 * j] = dt;  // d pre, formed in place in the ring
#pragma unroll
          for (int c = 0; c < CP; ++c) {
            dWatt[j][c] = fmaf(dt, cv[c], dWatt[j][c]);
            dcv[c] = fmaf(dt, Watt[j][c], dcv[c]);
          }
        }
#pragma unroll
        for (int c = 0; c < CP; ++c) dcv[c] = warp_sum(dcv[c]);
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < CP; ++c) dcv_p[(half * g.tloc_max + tl) * CP + c] = dcv[c];
        }
      }
    }
    // hand the chunk to the TMA unit: d_pre[b, rows, :] (+)= tile
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      float *dst = p.d_pre + ((size_t)b * Th + t0 + kChunkRows * qq) * A;
      const uint32_t bytes = (uint32_t)rows * A * 4u;
      if (p.accumulate_pre) bulk_red_add_s2g(dst, tile, bytes);
      else bulk_s2g(dst, tile, bytes);
      bulk_commit();
      if (q + g.ns < total) {
        bulk_wait_read<0>();  // the store has drained the stage before it is refilled
        issue_chunk(q + g.ns, gl, p.enc, D, p.pre, A, b, Th, t0, t1, stages, full);
      }
    }
  }
  // ---- CTA-level reductions of the parameter gradients (shared-memory atomics, then one
  //      global atomic per element per CTA)
#pragma unroll
  for (int j = 0; j < APL; ++j) {
    const int a = half * (A / 2) + lane + 32 * j;
    atomicAdd(ddp_s + a, ddp[j]);
    atomicAdd(dgv_s + a, dgv[j]);
#pragma unroll
    for (int c = 0; c < CP; ++c)
      if (c < C) atomicAdd(dWatt_s + a * CP + c, dWatt[j][c]);
  }
  __syncthreads();
  if (tloc > 0) {
    for (int i = tid; i < A * C; i += kThreads) {
      const int a = i / C, c = i - a * C;
      atomicAdd(p.dW_att + i, dWatt_s[a * CP + c]);
    }
    for (int a = tid; a < A; a += kThreads) {
      atomicAdd(p.dgvec + a, dgv_s[a]);
      atomicAdd(p.d_decproj + (size_t)b * A + a, ddp_s[a]);
    }
  }
  // ---- publish d conv (this CTA's frames) to every CTA of the cluster, channel-major
  for (int item = tid; item < tloc * C; item += kThreads) {
    const int tl = item / C, c = item - tl * C;
    const float v = dcv_p[tl * CP + c] + dcv_p[(g.tloc_max + tl) * CP + c];
    for (int r = 0; r < CL; ++r) dsmem_st(dsmem_addr(dcvT + c * g.Thp + t0 + tl, r), v);
  }
  cluster_sync_all();
  // ---- d att_prev[s] = sum_c sum_k Wc[c,k] * dconv[s - k + filts, c]
  if (p.d_att_prev) {
    for (int item = tid; item < tloc * C; item += kThreads) {
      const int tl = item / C, c = item - tl * C;
      const int s = t0 + tl;
      const int klo = max(0, s + filts - Th + 1), khi = min(K, s + filts + 1);
      const float *wr = wc_s + c * K;
      const float *dr = dcvT + c * g.Thp + s + filts;  // index (s + filts - k)
      float a0 = 0.f, a1 = 0.f;
      int k = klo;
      for (; k + 1 < khi; k += 2) {
        a0 = fmaf(wr[k], dr[-k], a0);
        a1 = fmaf(wr[k + 1], dr[-k - 1], a1);
      }
      if (k < khi) a0 = fmaf(wr[k], dr[-k], a0);
      dcv_p[tl * CP + c] = a0 + a1;  // reuse as scratch (this CTA's publish is complete)
    }
    __syncthreads();
    for (int tl = tid; tl < tloc; tl += kThreads) {
      float s = 0.0f;
      for (int c = 0; c < C; ++c) s += dcv_p[tl * CP + c];
      p.d_att_prev[(size_t)b * Th + t0 + tl] = s;
    }
  }
  // ---- dWc[c,k] += sum_{t in mine} dconv[t,c] * att_prev[t + k - filts]
  for (int item = tid; item < C * K; item += kThreads) {
    const int c = item / K, k = item - c * K;
    const int lo = max(t0, filts - k), hi = min(t1, Th + filts - k);
    const float *dr = dcvT + c * g.Thp;
    const float *ar = ap_s + (k - filts);
    float a0 = 0.f, a1 = 0.f;
    int t = lo;
    for (; t + 1 < hi; t += 2) {
      a0 = fmaf(dr[t], ar[t], a0);
      a1 = fmaf(dr[t + 1], ar[t + 1], a1);
    }
    if (t < hi) a0 = fmaf(dr[t], ar[t], a0);
    if (hi > lo) atomicAdd(p.dW_conv + item, a0 + a1);
  }
  if (tid == 0) bulk_wait<0>();  // all d_pre traffic of this CTA has left shared memory / landed
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
// out[m,n] (+)= sum_k X[m,k] * W[n,k] (NT) or W[k,n] (NN);  M small (batch), one warp per output
// column, lane <-> row m.  Used for dec_proj = dec_z W_dec^T and d_dec_z = d_decproj W_dec.
template <bool NN>
__global__ void __launch_bounds__(128) skinny_gemm_kernel(const float *__restrict__ X,
                                                          const float *__restrict__ W,
                                                          float *__restrict__ out, int M, int N, int Kd,
                                                          int accumulate) {
  extern __shared__ __align__(16) float smem[];
  const int Kp = Kd | 1;                   // odd pitch: lane <-> row reads are conflict free
  float *x_s = smem;                       // [M][Kp]
  float *w_s = smem + (size_t)M * Kp;      // [4][Kd]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * 4;
  for (int i = tid; i < M * Kd; i += 128) {
    const int m = i / Kd, k = i - m * Kd;
    x_s[m * Kp + k] = __ldg(X + i);
  }
  for (int i = tid; i < 4 * Kd; i += 128) {
    int w4, k;
    if (NN) { k = i >> 2; w4 = i & 3; } else { w4 = i / Kd; k = i - w4 * Kd; }
    const int n = n0 + w4;
    float v = 0.0f;
    if (n < N) v = NN ? __ldg(W + (size_t)k * N + n) : __ldg(W + (size_t)n * Kd + k);
    w_s[w4 * Kd + k] = v;
  }
  __syncthreads();
  const int n = n0 + warp;
  if (n >= N) return;
  const float *wr = w_s + warp * Kd;
  for (int m = lane; m < M; m += 32) {
    const float *xr = x_s + m * Kp;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int k = 0;
    for (; k + 3 < Kd; k += 4) {
      a0 = fmaf(xr[k], wr[k], a0);
      a1 = fmaf(xr[k + 1], wr[k + 1], a1);
      a2 = fmaf(xr[k + 2], wr[k + 2], a2);
      a3 = fmaf(xr[k + 3], wr[k + 3], a3);
    }
    for (; k < Kd; ++k) a0 = fmaf(xr[k], wr[k], a0);
    const float v = (a0 + a1) + (a2 + a3);
    float *o = out + (size_t)m * N + n;
    *o = accumulate ? *o + v : v;
  }
}

__global__ void init_att_kernel(const int32_t *__restrict__ hlens, float *__restrict__ att, int B, int Th) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Th) return;
  const int b = i / Th, t = i - b * Th;
  const int l = __ldg(hlens + b);
  att[i] = (t < l) ? 1.0f / (float)l : 0.0f;
}

// d_enc_h[b,t,:] (+)= sum_i w_all[i,b,t] * dc_all[i,b,:]
__global__ void __launch_bounds__(kThreads)
enc_grad_kernel(const float *__restrict__ w_all, const float *__restrict__ dc_all,
                float *__restrict__ d_enc, int steps, int B, int Th, int D, int tt, int accumulate) {
  extern __shared__ __align__(16) float smem[];
  float *dc_s = smem;                    // [steps][D]
  float *w_s = smem + (size_t)steps * D;  // [steps][tt]
  const int tiles = (Th + tt - 1) / tt;
  const int b = blockIdx.x / tiles, t0 = (blockIdx.x - b * tiles) * tt;
  const int rows = min(tt, Th - t0);
  const int tid = threadIdx.x;
  for (int i = tid; i < steps * D; i += kThreads) {
    const int s = i / D, d = i - s * D;
    dc_s[i] = __ldg(dc_all + ((size_t)s * B + b) * D + d);
  }
  for (int i = tid; i < steps * tt; i += kThreads) {
    const int s = i / tt, r = i - s * tt;
    w_s[i] = r < rows ? __ldg(w_all + ((size_t)s * B + b) * Th + t0 + r) : 0.0f;
  }
  __syncthreads();
  for (int d = tid; d < D; d += kThreads) {
    for (int r = 0; r < rows; ++r) {
      float a = 0.0f;
      for (int s = 0; s < steps; ++s) a = fmaf(w_s[s * tt + r], dc_s[s * D + d], a);
      float *o = d_enc + ((size_t)b * Th + t0 + r) * D + d;
      *o = accumulate ? *o + a : a;
    }
  }
}

inline bool pick_geom(int B, int Th, int D, int A, int C, int K, int CP, bool bwd, int &CL, AttGeom &g,
                      size_t &smem) {
  // cluster size: enough CTAs to cover the SMs once, at most 8 (portable limit)
  const int sms = num_sms();
  CL = 1;
  while (CL < 8 && B * CL * 2 <= sms) CL *= 2;
  while (CL > 1 && (Th + CL - 1) / CL < kChunkRows) CL /= 2;  // tiny Th: do not over-split
  for (;; ) {
    g.tloc_max = (Th + CL - 1) / CL;
    g.nch = (g.tloc_max + kChunkRows - 1) / kChunkRows;
    g.stage_floats = kChunkRows * (A > D ? A : D);
    g.Thp = round4(Th);
    g.CKp = round4(C * K);
    size_t fixed = 256;
    if (!bwd)
      fixed += sizeof(float) * ((size_t)g.Thp + g.CKp + (size_t)g.tloc_max * (CP + 3) + (size_t)kWarps * D +
                                round4(D) + 4);
    else
      fixed += sizeof(float) * ((size_t)g.Thp + g.CKp + (size_t)g.tloc_max * (CP + 3) +
                                2 * (size_t)g.tloc_max * CP + (size_t)CP * g.Thp + (size_t)A * CP + 2 * A + 4);
    const size_t budget = 220 * 1024;
    const size_t stage_bytes = sizeof(float) * (size_t)g.stage_floats;
    if (fixed + 2 * stage_bytes <= budget) {
      int ns = (int)((budget - fixed) / stage_bytes);
      if (ns > 2 * g.nch) ns = 2 * g.nch;
      if (ns > kMaxStages) ns = kMaxStages;
      if (ns < 2) ns = 2;
      g.ns = ns;
      smem = fixed + ns * stage_bytes;
      return true;
    }
    if (CL >= 8) return false;
    CL *= 2;
  }
}

template <typename Kern, typename Params>
int launch_cluster(Kern kern, const Params &prm, int B, int CL, size_t smem, cudaStream_t st) {
  int rc0 = ensure_smem(reinterpret_cast<const void *>(kern), smem);
  if (rc0 != RE2E_OK) return rc0;
  cudaError_t e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(B * CL));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, prm);
  count_launch();
  return e == cudaSuccess ? RE2E_OK : (int)e;
}

#define ATT_DISPATCH(NAME, APLV, CPV, ...)                                              \
  switch (APLV) {                                                                       \
    case 1: rc = (CPV) == 10 ? NAME<1, 10> __VA_ARGS__ : NAME<1, 16> __VA_ARGS__; break; \
    case 2: rc = (CPV) == 10 ? NAME<2, 10> __VA_ARGS__ : NAME<2, 16> __VA_ARGS__; break; \
    case 4: rc = (CPV) == 10 ? NAME<4, 10> __VA_ARGS__ : NAME<4, 16> __VA_ARGS__; break; \
    case 5: rc = (CPV) == 10 ? NAME<5, 10> __VA_ARGS__ : NAME<5, 16> __VA_ARGS__; break; \
    case 8: rc = (CPV) == 10 ? NAME<8, 10> __VA_ARGS__ : NAME<8, 16> __VA_ARGS__; break; \
    default: rc = RE2E_E_UNSUPPORTED;                                                   \
  }

template <int APL, int CP>
int run_fwd(const AttFwdParams &prm, int CL, size_t smem, cudaStream_t st) {
  return launch_cluster(attloc_fwd_kernel<APL, CP>, prm, prm.B, CL, smem, st);
}
template <int APL, int CP>
int run_bwd(const AttBwdParams &prm, int CL, size_t smem, cudaStream_t st) {
  return launch_cluster(attloc_bwd_kernel<APL, CP>, prm, prm.B, CL, smem, st);
}

inline int check_dims(int B, int Th, int D, int A, int C, int K) {
  if (B <= 0 || Th <= 0 || D <= 0 || A <= 0 || C <= 0 || K <= 0) return RE2E_E_ARG;
  if ((K & 1) == 0) return RE2E_E_ARG;
  if (A % 64 != 0 || A > 512 || (D & 3) || D > 32 * kDplMax || C > 16) return RE2E_E_UNSUPPORTED;
  const int apl = A / 64;
  if (!(apl == 1 || apl == 2 || apl == 4 || apl == 5 || apl == 8)) return RE2E_E_UNSUPPORTED;
  return RE2E_OK;
}

int skinny(bool nn, const float *X, const float *W, float *out, int M, int N, int Kd, int accumulate,
           cudaStream_t st) {
  const size_t smem = sizeof(float) * ((size_t)M * (Kd | 1) + 4 * (size_t)Kd);
  if (smem > 200 * 1024) return RE2E_E_UNSUPPORTED;
  int rc0;
  if (nn) {
    if ((rc0 = ensure_smem(reinterpret_cast<const void *>(skinny_gemm_kernel<true>), smem)) != RE2E_OK) return rc0;
    skinny_gemm_kernel<true><<<(N + 3) / 4, 128, smem, st>>>(X, W, out, M, N, Kd, accumulate);
  } else {
    if ((rc0 = ensure_smem(reinterpret_cast<const void *>(skinny_gemm_kernel<false>), smem)) != RE2E_OK) return rc0;
    skinny_gemm_kernel<false><<<(N + 3) / 4, 128, smem, st>>>(X, W, out, M, N, Kd, accumulate);
  }
  count_launch();
  return launch_status();
}

}  // namespace
}  // namespace re2e

using namespace re2e;

extern "C" int re2e_attloc_init_att(const int32_t *hlens, float *att_prev, int B, int Th, void *stream) {
  RE2E_CHECK_ARG(hlens && att_prev && B > 0 && Th > 0);
  init_att_kernel<<<(B * Th + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(hlens, att_prev, B, Th);
  count_launch();
  return launch_status();
}

extern "C" int re2e_attloc_step_fwd(const float *pre, const float *enc_h, const float *dec_z,
                                    const float *att_prev, const float *W_dec, const float *W_att,
                                    const float *W_conv, const float *gvec, const float *gvec_b,
                                    float scaling, float *c, float *w, float *dec_proj, float *conv, int B,
                                    int Th, int D, int A, int Z, int C, int K, void *stream) {
  RE2E_CHECK_ARG(pre && enc_h && att_prev && W_dec && W_att && W_conv && gvec && gvec_b && c && w && dec_proj);
  int rc = check_dims(B, Th, D, A, C, K);
  if (rc != RE2E_OK) return rc;
  RE2E_CHECK_ARG(Z > 0 && aligned16(pre) && aligned16(enc_h));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dec_z) {
    rc = skinny(false, dec_z, W_dec, dec_proj, B, A, Z, 0, st);
    if (rc != RE2E_OK) return rc;
  } else {
    RE2E_CUDA(cudaMemsetAsync(dec_proj, 0, sizeof(float) * (size_t)B * A, st));
  }
  const int CP = C == 10 ? 10 : 16;
  AttFwdParams prm;
  prm.pre = pre; prm.enc = enc_h; prm.att_prev = att_prev; prm.dec_proj = dec_proj; prm.W_att = W_att;
  prm.W_conv = W_conv; prm.gvec = gvec; prm.gvec_b = gvec_b; prm.scaling = scaling; prm.c = c; prm.w = w;
  prm.conv = conv; prm.B = B; prm.Th = Th; prm.D = D; prm.A = A; prm.C = C; prm.K = K;
  int CL;
  size_t smem;
  if (!pick_geom(B, Th, D, A, C, K, CP, false, CL, prm.g, smem)) return RE2E_E_UNSUPPORTED;
  ATT_DISPATCH(run_fwd, A / 64, CP, (prm, CL, smem, st));
  return rc;
}

extern "C" int re2e_attloc_step_bwd(const float *dc, const float *dw, const float *pre, const float *enc_h,
                                    const float *att_prev, const float *w, const float *dec_proj,
                                    const float *conv, const float *W_att, const float *W_conv,
                                    const float *gvec, float scaling, float *d_pre, int accumulate_pre,
                                    float *d_decproj, float *d_att_prev, float *dW_att, float *dW_conv,
                                    float *dgvec, float *dgvec_b, int B, int Th, int D, int A, int C, int K,
                                    void *stream) {
  RE2E_CHECK_ARG(pre && enc_h && att_prev && w && dec_proj && conv && W_att && W_conv && gvec);
  RE2E_CHECK_ARG(d_pre && d_decproj && dW_att && dW_conv && dgvec && dgvec_b);
  int rc = check_dims(B, Th, D, A, C, K);
  if (rc != RE2E_OK) return rc;
  RE2E_CHECK_ARG(aligned16(pre) && aligned16(enc_h) && aligned16(d_pre));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RE2E_CUDA(cudaMemsetAsync(d_decproj, 0, sizeof(float) * (size_t)B * A, st));
  const int CP = C == 10 ? 10 : 16;
  AttBwdParams prm;
  prm.dc = dc; prm.dw = dw; prm.pre = pre; prm.enc = enc_h; prm.att_prev = att_prev; prm.w = w;
  prm.dec_proj = dec_proj; prm.conv = conv; prm.W_att = W_att; prm.W_conv = W_conv; prm.gvec = gvec;
  prm.scaling = scaling; prm.d_pre = d_pre; prm.d_decproj = d_decproj; prm.d_att_prev = d_att_prev;
  prm.dW_att = dW_att; prm.dW_conv = dW_conv; prm.dgvec = dgvec; prm.dgvec_b = dgvec_b;
  prm.accumulate_pre = accumulate_pre;
  prm.B = B; prm.Th = Th; prm.D = D; prm.A = A; prm.C = C; prm.K = K;
  int CL;
  size_t smem;
  if (!pick_geom(B, Th, D, A, C, K, CP, true, CL, prm.g, smem)) return RE2E_E_UNSUPPORTED;
  ATT_DISPATCH(run_bwd, A / 64, CP, (prm, CL, smem, st));
  return rc;
}

extern "C" int re2e_attloc_enc_grad(const float *w_all, const float *dc_all, float *d_enc_h, int steps,
                                    int B, int Th, int D, int accumulate, void *stream) {
  RE2E_CHECK_ARG(w_all && dc_all && d_enc_h && steps > 0 && B > 0 && Th > 0 && D > 0);
  const int tt = 8;
  const size_t smem = sizeof(float) * ((size_t)steps * D + (size_t)steps * tt);
  if (smem > 200 * 1024) return RE2E_E_UNSUPPORTED;
  {
    int rc0 = ensure_smem(reinterpret_cast<const void *>(enc_grad_kernel), smem);
    if (rc0 != RE2E_OK) return rc0;
  }
  const int tiles = (Th + tt - 1) / tt;
  enc_grad_kernel<<<B * tiles, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(w_all, dc_all, d_enc_h,
                                                                                   steps, B, Th, D, tt, accumulate);
  count_launch();
  return launch_status();
}

// out[M,N] (+)= X[M,K] @ W[N,K]^T   /   X[M,K] @ W[K,N]     (M = batch-sized)
extern "C" int re2e_skinny_nt(const float *X, const float *W, float *out, int M, int N, int K,
                              int accumulate, void *stream) {
  RE2E_CHECK_ARG(X && W && out && M > 0 && N > 0 && K > 0);
  return skinny(false, X, W, out, M, N, K, accumulate, static_cast<cudaStream_t>(stream));
}
extern "C" int re2e_skinny_nn(const float *X, const float *W, float *out, int M, int N, int K,
                              int accumulate, void *stream) {
  RE2E_CHECK_ARG(X && W && out && M > 0 && N > 0 && K > 0);
  return skinny(true, X, W, out, M, N, K, accumulate, static_cast<cudaStream_t>(stream));
}
