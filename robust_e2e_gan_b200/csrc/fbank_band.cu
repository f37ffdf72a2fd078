// Front-end for a BANDED filter bank (sm_100a): mask tail of EnhanceModel.forward (model/enhance_model.py:157-164)
// + FbankModel.forward (model/feat_model.py:118-135) and its backward, streaming at HBM rate.
//
// A triangular mel bank -- the reference's frozen 80-filter table (model/feat_model.py:15-33: 501 non-zeros, every FFT
// bin feeds at most 2 filters) and the generic 40-filter bank alike -- is banded: filter m only sees a short run of
// bins, bin f only feeds a couple of neighbouring filters.  The dense 257 x M projection (10 K MACs per frame on the
// tensor cores, whose operand traffic through shared memory was the measured limiter of the dense kernel) collapses to
// ~450 MACs per frame: the kernel is a pure stream.  The caller detects the structure once from `fc` and passes it as
// two small tables (see re2e_fbank_band_fwd in the header); a trained, dense `fc` keeps using the tcgen05 kernels.
//
// Rows are F = 257 floats = 1028 B (not 16 B aligned), but a span of 8 rows is 8224 B = 514 x 16 B: every tile of 8
// frames is ONE contiguous, aligned run per input, fetched by the TMA unit (cp.async.bulk -> UBLKCP) into a ring of
// shared-memory stages, each signalled on its own mbarrier; 4 stages (up to 25 KB each) are in flight per CTA and two
// CTAs share an SM, so ~200 KB per SM are outstanding -- enough for the HBM latency-bandwidth product.  Up to THREE
// outputs per launch (joint_train.py:158-161: enhance_feat from mask x mix, mix_feat from mix, clean_feat from clean):
// `mix` is read once for two of them.
//   forward : elementwise x^2 in place in the stage (sigmoid, length mask), then thread <-> (frame, filter) with the
//             filter's <= 32 weights in registers; Y / G of a tile are 8*M contiguous floats: fully coalesced stores
//   backward: dP = dY*G per tile in shared memory, thread <-> elements of the (frame, bin) tile with the bin's <= 4
//             weights from a shared table; the d_in tile is formed in place and handed back to the TMA unit (bulk store)
// Algorithmic bytes: forward 4N(2F + 2M) masked (+ 4N(F + M) per extra plain output); backward 4N(3F + 2M).
#include "common.cuh"

namespace re2e {
namespace {

constexpr int kBR = 8;         // frames per tile (8 * 1028 B = 514 * 16 B)
constexpr int kBS = 4;         // ring stages
constexpr int kBandW = 32;     // bins per filter window
constexpr int kBinW = 4;       // filters per bin window

struct BandFwdParams {
  const float *mask, *mag, *mag2, *fw, *cmvn;
  const int32_t *flo, *lens;
  float *Y_enh, *G, *Y_plain, *Y2;
  int mask_is_logit, N, T, F, M, ntiles;
};

struct BandBwdParams {
  const float *dY, *G, *mask, *mag, *bw;
  const int32_t *mlo, *lens;
  float *d_in;
  int mask_is_logit, N, T, F, M, ntiles;
};

__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// NBUF = inputs per stage: mask (if any), mag, mag2 (if any)
template <int NBUF>
__global__ void __launch_bounds__(640, 1) fbank_band_fwd_kernel(const BandFwdParams p) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int F = p.F, M = p.M, T = p.T;
  const int NT = blockDim.x, tid = threadIdx.x;
  const int tile_f = kBR * F;                       // floats per input tile
  const bool has_mask = p.mask != nullptr, has2 = p.mag2 != nullptr;
  uint64_t *full = reinterpret_cast<uint64_t *>(smraw);          // [kBS]
  int *valid_s = reinterpret_cast<int *>(smraw + 64);             // [2][kBR] frame t < lens[b], for this tile / the next
  float *ring = reinterpret_cast<float *>(smraw + 128);           // kBS * NBUF * tile_f
  const int my_tiles = p.ntiles > (int)blockIdx.x ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  // frame validity of tile #i of this CTA (threads 0..7, one frame each): the only place that divides by T
  auto mark_valid = [&](int i) {
    const size_t n = ((size_t)blockIdx.x + (size_t)i * gridDim.x) * kBR + tid;
    const int b = (int)(n / (size_t)T), t = (int)(n - (size_t)b * T);
    valid_s[(i & 1) * kBR + tid] = (!p.lens || t < __ldg(p.lens + b)) ? 1 : 0;
  };

  auto issue = [&](int i) {   // tile #i of this CTA -> stage i % kBS   (one thread)
    const int st = i % kBS;
    const size_t row0 = ((size_t)blockIdx.x + (size_t)i * gridDim.x) * kBR;
    const uint32_t bytes = (uint32_t)tile_f * 4u;
    float *dst = ring + (size_t)st * NBUF * tile_f;
    mbar_expect_tx(&full[st], bytes * NBUF);
    int slot = 0;
    if (has_mask) bulk_g2s(dst + (slot++) * tile_f, p.mask + row0 * F, bytes, &full[st]);
    bulk_g2s(dst + (slot++) * tile_f, p.mag + row0 * F, bytes, &full[st]);
    if (has2) bulk_g2s(dst + (slot++) * tile_f, p.mag2 + row0 * F, bytes, &full[st]);
  };
  // The projection reads 16 B-aligned windows that may extend up to 12 B past a frame: past the last frame of a buffer
  // that is the head of the NEXT buffer (or the slack after the ring).  Those elements meet zero weights, but 0 x NaN is
  // NaN, so what they can hold before the first bulk copy lands there must be finite: zero the head of every buffer and
  // the slack once, and order those generic-proxy writes before the first bulk copies.
  for (int i = tid; i < kBS * NBUF + 1; i += NT)
    *reinterpret_cast<float4 *>(ring + (size_t)i * tile_f) = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async_smem();
  if (tid == 0) {
    for (int s = 0; s < kBS; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    const int first = my_tiles < kBS ? my_tiles : kBS;
    for (int i = 0; i < first; ++i) issue(i);
  }
  // this thread's (frame, filter): the filter's weights live in registers, laid out on the 16 B-ALIGNED window of the
  // tile that contains the filter's bins for this frame (rows are 257 floats: the alignment of a filter's first bin
  // depends on the frame), so the projection reads the tile with 128-bit loads
  const int r = tid / M, m = tid - r * M;
  const bool worker = tid < kBR * M;
  constexpr int kWQ = kBandW / 4 + 1;              // quads that can hold a shifted window
  float4 w4[kWQ];
  int base4 = 0, nq = 0;                           // aligned window start (in quads), quads with a non-zero weight
  float c0 = 0.0f, c1 = 1.0f;
#pragma unroll
  for (int q = 0; q < kWQ; ++q) w4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (worker) {
    const int flo = __ldg(p.flo + m);
    const int first = r * F + flo, sh = first & 3;
    base4 = (first - sh) >> 2;
    const float *wr = p.fw + (size_t)m * kBandW;
#pragma unroll
    for (int q = 0; q < kWQ; ++q) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = 4 * q + j - sh;
        v[j] = (k >= 0 && k < kBandW) ? __ldg(wr + k) : 0.0f;
      }
      w4[q] = make_float4(v[0], v[1], v[2], v[3]);
      if (v[0] != 0.0f || v[1] != 0.0f || v[2] != 0.0f || v[3] != 0.0f) nq = q + 1;
    }
    if (p.cmvn) { c0 = __ldg(p.cmvn + m); c1 = __ldg(p.cmvn + M + m); }
  }
  // the quads this thread squares in place (same positions in every tile): frame of the quad's first element and how
  // many of its 4 elements still belong to that frame -- computed once, no division inside the tile loop
  int qr0[2], qnb[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int e0 = 4 * (tid + k * NT);
    qr0[k] = e0 / F;
    qnb[k] = (qr0[k] + 1) * F - e0;
  }
  if (tid < kBR && my_tiles > 0) mark_valid(0);
  __syncthreads();

  for (int i = 0; i < my_tiles; ++i) {
    const int st = i % kBS;
    const size_t row0 = ((size_t)blockIdx.x + (size_t)i * gridDim.x) * kBR;
    float *buf = ring + (size_t)st * NBUF * tile_f;
    float *b_mask = buf, *b_mag = has_mask ? buf + tile_f : buf, *b_mag2 = b_mag + tile_f;
    mbar_wait(&full[st], (uint32_t)(i / kBS) & 1u);
    // ---- powers in place: mask slot <- (act(mask) * valid * mag)^2, mag slot <- mag^2, mag2 slot <- mag2^2
    const bool sq_mag = !has_mask || p.Y_plain != nullptr;
    const int *vld = valid_s + (i & 1) * kBR;
    // the tile is one flat, 16 B aligned run of kBR*F floats: 128-bit accesses; a quad may straddle two frames
    auto square_quad = [&](int q, int r0, int nb) {
      const int e0 = 4 * q;
      float4 x = *reinterpret_cast<const float4 *>(b_mag + e0);
      if (has_mask) {
        const bool ok0 = vld[r0] != 0, ok1 = nb < 4 ? vld[r0 + 1] != 0 : ok0;
        const float4 mk = *reinterpret_cast<const float4 *>(b_mask + e0);
        float4 e;
        e.x = ok0 ? (p.mask_is_logit ? sigmoid_fast(mk.x) : mk.x) * x.x : 0.0f;
        e.y = (nb > 1 ? ok0 : ok1) ? (p.mask_is_logit ? sigmoid_fast(mk.y) : mk.y) * x.y : 0.0f;
        e.z = (nb > 2 ? ok0 : ok1) ? (p.mask_is_logit ? sigmoid_fast(mk.z) : mk.z) * x.z : 0.0f;
        e.w = (nb > 3 ? ok0 : ok1) ? (p.mask_is_logit ? sigmoid_fast(mk.w) : mk.w) * x.w : 0.0f;
        *reinterpret_cast<float4 *>(b_mask + e0) = make_float4(e.x * e.x, e.y * e.y, e.z * e.z, e.w * e.w);
      }
      if (sq_mag) *reinterpret_cast<float4 *>(b_mag + e0) = make_float4(x.x * x.x, x.y * x.y, x.z * x.z, x.w * x.w);
      if (has2) {
        const float4 y = *reinterpret_cast<const float4 *>(b_mag2 + e0);
        *reinterpret_cast<float4 *>(b_mag2 + e0) = make_float4(y.x * y.x, y.y * y.y, y.z * y.z, y.w * y.w);
      }
    };
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (tid + k * NT < tile_f / 4) square_quad(tid + k * NT, qr0[k], qnb[k]);
    for (int q = tid + 2 * NT; q < tile_f / 4; q += NT) {   // (only for CTAs with fewer than tile_f/8 threads)
      const int r0 = 4 * q / F;
      square_quad(q, r0, (r0 + 1) * F - 4 * q);
    }
    __syncthreads();
    if (tid < kBR && i + 1 < my_tiles) mark_valid(i + 1);   // ordered before its use by the barrier that ends this tile
    // ---- banded projection + log + CMVN: thread <-> (frame, filter)
    if (worker) {
      const size_t o = row0 * M + tid;                 // the tile's outputs are 8*M contiguous floats
      auto project = [&](const float *P) {
        const float4 *P4 = reinterpret_cast<const float4 *>(P) + base4;
        float a0 = 0.0f, a1 = 0.0f;                    // two chains: the FMA latency overlaps
#pragma unroll
        for (int q = 0; q < kWQ; ++q)
          if (q < nq) {
            const float4 x = P4[q];
            a0 = fmaf(w4[q].x, x.x, a0); a1 = fmaf(w4[q].y, x.y, a1);
            a0 = fmaf(w4[q].z, x.z, a0); a1 = fmaf(w4[q].w, x.w, a1);
          }
        return a0 + a1;
      };
      if (has_mask) {
        const float P = project(b_mask);
        const bool ok = P > 1e-7f;
        p.Y_enh[o] = (__logf(ok ? P : 1e-7f) + c0) * c1;
        if (p.G) p.G[o] = ok ? __fdividef(c1, P) : 0.0f;
      }
      if (!has_mask) {                                 // single-input form: the plain input IS output 0
        const float P = project(b_mag);
        const bool ok = P > 1e-7f;
        p.Y_enh[o] = (__logf(ok ? P : 1e-7f) + c0) * c1;
        if (p.G) p.G[o] = ok ? __fdividef(c1, P) : 0.0f;
      } else if (p.Y_plain) {
        const float P = project(b_mag);
        p.Y_plain[o] = (__logf(P > 1e-7f ? P : 1e-7f) + c0) * c1;
      }
      if (has2) {
        const float P = project(b_mag2);
        p.Y2[o] = (__logf(P > 1e-7f ? P : 1e-7f) + c0) * c1;
      }
    }
    __syncthreads();   // stage consumed
    if (tid == 0 && i + kBS < my_tiles) issue(i + kBS);
  }
}

template <bool HAS_MASK>
__global__ void __launch_bounds__(512, 1) fbank_band_bwd_kernel(const BandBwdParams p) {
  extern __shared__ __align__(128) unsigned char smraw[];
  constexpr int NBUF = HAS_MASK ? 2 : 1;
  const int F = p.F, M = p.M, T = p.T;
  const int NT = blockDim.x, tid = threadIdx.x;
  const int tile_f = kBR * F, tile_m = kBR * M;
  uint64_t *full = reinterpret_cast<uint64_t *>(smraw);          // [kBS]
  float *btab = reinterpret_cast<float *>(smraw + 128);           // F * kBinW weights
  int *mlo_s = reinterpret_cast<int *>(btab + F * kBinW);         // round4(F)
  float *ring = reinterpret_cast<float *>(mlo_s + ((F + 3) & ~3));   // kBS * (NBUF*tile_f + 2*tile_m)
  int *valid_s = reinterpret_cast<int *>(smraw + 64);             // [2][kBR]
  const int stage_f = NBUF * tile_f + 2 * tile_m;
  const int my_tiles = p.ntiles > (int)blockIdx.x ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  auto mark_valid = [&](int i) {
    const size_t n = ((size_t)blockIdx.x + (size_t)i * gridDim.x) * kBR + tid;
    const int b = (int)(n / (size_t)T), t = (int)(n - (size_t)b * T);
    valid_s[(i & 1) * kBR + tid] = (!p.lens || t < __ldg(p.lens + b)) ? 1 : 0;
  };
  auto issue = [&](int i) {
    const int st = i % kBS;
    const size_t row0 = ((size_t)blockIdx.x + (size_t)i * gridDim.x) * kBR;
    float *dst = ring + (size_t)st * stage_f;
    const uint32_t bf = (uint32_t)tile_f * 4u, bm = (uint32_t)tile_m * 4u;
    mbar_expect_tx(&full[st], bf * NBUF + 2u * bm);
    if (HAS_MASK) bulk_g2s(dst, p.mask + row0 * F, bf, &full[st]);
    bulk_g2s(dst + (NBUF - 1) * tile_f, p.mag + row0 * F, bf, &full[st]);
    bulk_g2s(dst + NBUF * tile_f, p.dY + row0 * M, bm, &full[st]);
    bulk_g2s(dst + NBUF * tile_f + tile_m, p.G + row0 * M, bm, &full[st]);
  };
  // The projection reads 16 B-aligned windows that may extend up to 12 B past a frame: past the last frame of a buffer
  // that is the head of the NEXT buffer (or the slack after the ring).  Those elements meet zero weights, but 0 x NaN is
  // NaN, so what they can hold before the first bulk copy lands there must be finite: zero the head of every buffer and
  // the slack once, and order those generic-proxy writes before the first bulk copies.
  for (int i = tid; i < kBS * NBUF + 1; i += NT)
    *reinterpret_cast<float4 *>(ring + (size_t)i * tile_f) = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async_smem();
  if (tid == 0) {
    for (int s = 0; s < kBS; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    const int first = my_tiles < kBS ? my_tiles : kBS;
    for (int i = 0; i < first; ++i) issue(i);
  }
  for (int i = tid; i < F * kBinW; i += NT) btab[i] = __ldg(p.bw + i);
  for (int i = tid; i < F; i += NT) mlo_s[i] = __ldg(p.mlo + i);
  if (tid < kBR && my_tiles > 0) mark_valid(0);
  __syncthreads();

  for (int i = 0; i < my_tiles; ++i) {
    const int st = i % kBS;
    const size_t row0 = ((size_t)blockIdx.x + (size_t)i * gridDim.x) * kBR;
    float *buf = ring + (size_t)st * stage_f;
    float *b_out = buf;                                   // mask slot (or the mag slot): d_in is formed in place
    float *b_mag = buf + (NBUF - 1) * tile_f;
    float *dP = buf + NBUF * tile_f;                      // dY slot <- dY * G
    const float *Gs = dP + tile_m;
    mbar_wait(&full[st], (uint32_t)(i / kBS) & 1u);
    for (int idx = tid; idx < tile_m; idx += NT) dP[idx] *= Gs[idx];
    __syncthreads();
    if (tid < kBR && i + 1 < my_tiles) mark_valid(i + 1);
    const int *vld = valid_s + (i & 1) * kBR;
    // thread <-> bin (two frames' worth of threads): the bin's table entry is read once for the 4 frames it serves
    for (int f = tid & 255; f < F; f += 256) {
      const float4 wv = *reinterpret_cast<const float4 *>(btab + f * kBinW);
      const int ml = mlo_s[f];
#pragma unroll
      for (int q = 0; q < kBR / 2; ++q) {
        const int rr = 2 * q + (tid >> 8);
        const int idx = rr * F + f;
        const float *dp = dP + rr * M + ml;
        const float dsq = fmaf(wv.x, dp[0], fmaf(wv.y, dp[1], fmaf(wv.z, dp[2], wv.w * dp[3])));
        const float mg = b_mag[idx];
        float out;
        if (HAS_MASK) {
          out = 0.0f;
          if (vld[rr]) {
            const float mk = b_out[idx];
            if (p.mask_is_logit) {
              const float s = sigmoid_fast(mk);
              out = 2.0f * s * mg * dsq * mg * s * (1.0f - s);    // d/d logit of (s*mag)^2 . dsq
            } else {
              out = 2.0f * mk * mg * dsq * mg;
            }
          }
        } else {
          out = 2.0f * mg * dsq;
        }
        b_out[idx] = out;
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(p.d_in + row0 * F, b_out, (uint32_t)tile_f * 4u);
      bulk_commit();
      // the stage of the PREVIOUS tile is refilled once its store has finished reading shared memory
      if (i >= 1) {
        bulk_wait_read<1>();
        if (i - 1 + kBS < my_tiles) issue(i - 1 + kBS);
      }
    }
  }
  if (tid == 0) bulk_wait<0>();
}

inline bool band_shape_ok(int B, int T, int F, int M) {
  const long long N = (long long)B * T;
  return N % kBR == 0 && F >= kBandW && M >= kBinW && M <= 80 && N / kBR <= 0x7fffffff;
}

}  // namespace
}  // namespace re2e

using namespace re2e;

extern "C" int re2e_fbank_band_supported(int B, int T, int F, int M) { return band_shape_ok(B, T, F, M) ? 1 : 0; }

extern "C" int re2e_fbank_band_fwd(const float *mask, int mask_is_logit, const float *mag, const float *mag2,
                                   const int32_t *flo, const float *fw, const float *cmvn, const int32_t *lens,
                                   float *Y_enh, float *G, float *Y_plain, float *Y2, int B, int T, int F, int M,
                                   void *stream) {
  RE2E_CHECK_ARG(mag && flo && fw && B > 0 && T > 0 && F > 0 && M > 0);
  RE2E_CHECK_ARG(Y_enh && (mask || !Y_plain));   // no mask: the plain input is output 0 (Y_enh, G)
  RE2E_CHECK_ARG((mag2 != nullptr) == (Y2 != nullptr));
  if (!band_shape_ok(B, T, F, M)) return RE2E_E_UNSUPPORTED;
  RE2E_CHECK_ARG(aligned16(mag) && (!mask || aligned16(mask)) && (!mag2 || aligned16(mag2)));
  BandFwdParams prm;
  prm.mask = mask; prm.mag = mag; prm.mag2 = mag2; prm.fw = fw; prm.cmvn = cmvn; prm.flo = flo; prm.lens = lens;
  prm.Y_enh = Y_enh; prm.G = G; prm.Y_plain = Y_plain; prm.Y2 = Y2; prm.mask_is_logit = mask_is_logit;
  prm.N = B * T; prm.T = T; prm.F = F; prm.M = M; prm.ntiles = prm.N / kBR;
  const int nbuf = (mask ? 1 : 0) + 1 + (mag2 ? 1 : 0);
  const size_t smem = 128 + sizeof(float) * (size_t)kBS * nbuf * kBR * F + 64;   // + slack: aligned windows may read 12 B past the ring
  const int threads = (kBR * M + 31) & ~31;
  const int per_sm = smem * 2 <= 220 * 1024 && threads * 2 <= 2048 ? 2 : 1;
  int grid = num_sms() * per_sm;
  if (grid > prm.ntiles) grid = prm.ntiles;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  if (nbuf == 1) {
    if ((rc = ensure_smem(reinterpret_cast<const void *>(fbank_band_fwd_kernel<1>), smem)) != RE2E_OK) return rc;
    fbank_band_fwd_kernel<1><<<grid, threads, smem, st>>>(prm);
  } else if (nbuf == 2) {
    if ((rc = ensure_smem(reinterpret_cast<const void *>(fbank_band_fwd_kernel<2>), smem)) != RE2E_OK) return rc;
    fbank_band_fwd_kernel<2><<<grid, threads, smem, st>>>(prm);
  } else {
    if ((rc = ensure_smem(reinterpret_cast<const void *>(fbank_band_fwd_kernel<3>), smem)) != RE2E_OK) return rc;
    fbank_band_fwd_kernel<3><<<grid, threads, smem, st>>>(prm);
  }
  count_launch();
  return launch_status();
}

extern "C" int re2e_fbank_band_bwd(const float *dY, const float *G, const float *mask, int mask_is_logit,
                                   const float *mag, const int32_t *mlo, const float *bw, const int32_t *lens,
                                   float *d_in, int B, int T, int F, int M, void *stream) {
  RE2E_CHECK_ARG(dY && G && mag && mlo && bw && d_in && B > 0 && T > 0 && F > 0 && M > 0);
  if (!band_shape_ok(B, T, F, M)) return RE2E_E_UNSUPPORTED;
  RE2E_CHECK_ARG(aligned16(mag) && (!mask || aligned16(mask)) && aligned16(d_in) && aligned16(dY) && aligned16(G));
  BandBwdParams prm;
  prm.dY = dY; prm.G = G; prm.mask = mask; prm.mag = mag; prm.bw = bw; prm.mlo = mlo; prm.lens = lens; prm.d_in = d_in;
  prm.mask_is_logit = mask_is_logit; prm.N = B * T; prm.T = T; prm.F = F; prm.M = M; prm.ntiles = prm.N / kBR;
  const int nbuf = mask ? 2 : 1;
  const size_t smem = 128 + sizeof(float) * ((size_t)F * kBinW + ((F + 3) & ~3) +
                                             (size_t)kBS * ((size_t)nbuf * kBR * F + 2 * (size_t)kBR * M));
  const int per_sm = smem * 2 <= 220 * 1024 ? 2 : 1;
  int grid = num_sms() * per_sm;
  if (grid > prm.ntiles) grid = prm.ntiles;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  if (mask) {
    if ((rc = ensure_smem(reinterpret_cast<const void *>(fbank_band_bwd_kernel<true>), smem)) != RE2E_OK) return rc;
    fbank_band_bwd_kernel<true><<<grid, 512, smem, st>>>(prm);
  } else {
    if ((rc = ensure_smem(reinterpret_cast<const void *>(fbank_band_bwd_kernel<false>), smem)) != RE2E_OK) return rc;
    fbank_band_bwd_kernel<false><<<grid, 512, smem, st>>>(prm);
  }
  count_launch();
  return launch_status();
}
