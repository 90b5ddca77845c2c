// Fused backward of the BASELINE-shape STFT (fp32, fft_length 512, real output formats, constant padding):
// d/dx of  y = format(|rfft(window * frame(x))|^2 + eps)   (stft.py:237-241, spec.py:152-178, fftr.py:145).
//
// stft512.cu's analysis front and istft512.cu's synthesis back in ONE kernel, nothing saved by the forward pass:
//   * a CTA (8 warps, two per SM) owns a tile of consecutive waveform samples and the <= 32 frames that overlap it;
//     the tile's sample span is staged once in shared memory (zero outside the utterance = constant padding);
//   * every half-warp recomputes the spectrum of a frame pair in registers (packed radix-16 x 16, fft16.cuh),
//     forms G = g format'(s) X for the bins it holds (k and 256 - k), turns G back into the half-length
//     spectrum of the real gradient signal (the adjoint of the real FFT is n * irfft of the doubled-edge
//     spectrum), returns the mirrored half to its owner lane by shuffle, and runs the same forward
//     butterflies again: x~[2m] = Re R[m], x~[2m+1] = -Im R[m], R = FFT_256(conj(E' + i O'));
//   * window-weighted frame gradients are parked in shared memory; after one CTA barrier every thread sums the
//     <= ceil(L / P) frames that touch its samples in frame order and writes d/dx once, coalesced -- no
//     atomics, no memset, deterministic.
// Algorithmic bytes per frame: 320 B waveform + 1 028 B output gradient read, 320 B written.
#include <algorithm>
#include <cstdlib>

#include "fft16.cuh"

namespace dsb200 {
namespace {

using namespace fft16_detail;

constexpr int kBW = 8;                 // warps per CTA
constexpr int kBT = kBW * 32;
constexpr int kBFrames = 4 * kBW;      // frames a CTA can hold: one quad per warp

struct BArgs {
  const float* x;       // [batch, T]
  const float* w;       // [L]
  const float* gy;      // [batch, N, 257]
  const float2* tw512;  // W512^k, 512 entries
  float* gx;            // [batch, T]
  int64_t batch, N, T, tiles_per_utt;
  int L, P, left, tile, alias, span;
  int fbuf_off;         // float2 units from the second plane array to the frame rows (0 when they alias)
  int c, m;             // (L - 1) / P and (L - 1) % P
  int fmt;
  float eps;
};

__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }

// d format(s) / d s for the real output formats (spec.py:123-130); the factor 2 of d|X|^2 is folded downstream.
__device__ __forceinline__ float2 fmt_grad(float2 g, float2 s, int fmt) {
  switch (fmt) {
    case DSB200_SPEC_DB: return make_float2(g.x * 4.342944819032518f / s.x, g.y * 4.342944819032518f / s.y);
    case DSB200_SPEC_LOGMAG: return make_float2(g.x * 0.5f / s.x, g.y * 0.5f / s.y);
    case DSB200_SPEC_MAGNITUDE: return make_float2(g.x * 0.5f * rsqrtf(s.x), g.y * 0.5f * rsqrtf(s.y));
    default: return g;
  }
}

template <int NJ, bool MASK_ALL>
__global__ void __launch_bounds__(kBT, 2) stft512_bwd_kernel(const BArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int l = lane & 15, h = lane >> 4;
  const int partner = (lane & 16) | ((16 - l) & 15);   // lane that holds the mirrored bins
  const int L = A.L;

  float2* tws = reinterpret_cast<float2*>(smem_raw);           // [256] W512^k
  float* win = reinterpret_cast<float*>(tws + 256);            // [512] window (zero beyond L)
  float* xs = win + 512;                                       // [span] waveform samples of the tile's frames
  float2* planes1 = reinterpret_cast<float2*>(xs + A.span);    // [kBW][2 kPlane] exchange planes of half-warp 1
  float2* planes0 = planes1 + kBW * 2 * kPlane;                // [kBW][2 kPlane] (only when they cannot alias)
  float* fbuf = reinterpret_cast<float*>(planes0 + A.fbuf_off);                              // [32][L]

  for (int i = tid; i < 256; i += kBT) tws[i] = A.tw512[i];
  for (int i = tid; i < 512; i += kBT) win[i] = i < L ? A.w[i] : 0.0f;
  float twr[16], twi[16];
#pragma unroll
  for (int k2 = 1; k2 < 16; ++k2) {
    const float2 v = A.tw512[2 * l * k2];
    twr[k2] = v.x;
    twi[k2] = v.y;
  }
  float2* xr = h ? planes1 + warp * 2 * kPlane
                 : (A.alias ? reinterpret_cast<float2*>(fbuf + static_cast<size_t>(4 * warp) * L)
                            : planes0 + warp * 2 * kPlane);
  float2* xi = xr + kPlane;
  __syncthreads();

  const int64_t n_tiles = A.batch * A.tiles_per_utt;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t b = tile / A.tiles_per_utt;
    const int64_t t0 = (tile - b * A.tiles_per_utt) * A.tile;
    const int64_t t1 = (t0 + A.tile < A.T) ? t0 + A.tile : A.T;
    const int64_t q0 = t0 + A.left, q1 = t1 - 1 + A.left;     // positions in the padded waveform
    const int64_t n_lo = (q0 - L + 1 <= 0) ? 0 : (q0 - L + A.P) / A.P;
    int64_t n_hi = q1 / A.P;
    if (n_hi > A.N - 1) n_hi = A.N - 1;
    const int nf = static_cast<int>(n_hi - n_lo + 1);          // <= kBFrames by the choice of A.tile (may be <= 0)

    if (tid == 0 && tile + gridDim.x < n_tiles) {              // pull the next tile's inputs into the L2 now
      const int64_t tn = tile + gridDim.x, bn = tn / A.tiles_per_utt;
      const int64_t qn = (tn - bn * A.tiles_per_utt) * A.tile + A.left;
      const int64_t nl = (qn - L + 1 <= 0) ? 0 : (qn - L + A.P) / A.P;
      prefetch_l2(A.gy, static_cast<size_t>(A.batch) * A.N * 257 * sizeof(float), A.gy + (bn * A.N + nl) * 257,
                  static_cast<size_t>(kBFrames) * 257 * sizeof(float));
      prefetch_l2(A.x, static_cast<size_t>(A.batch) * A.T * sizeof(float), A.x + bn * A.T + nl * A.P - A.left,
                  static_cast<size_t>(A.span) * sizeof(float));
    }
    // stage the samples of frames n_lo .. n_lo + 31 (constant padding outside the utterance)
    {
      const float* xb = A.x + b * A.T;
      const int64_t p0 = n_lo * A.P - A.left;
      for (int i = tid; i < A.span; i += kBT) {
        const int64_t p = p0 + i;
        xs[i] = (p >= 0 && p < A.T) ? xb[p] : 0.0f;
      }
    }
    __syncthreads();

    const int fA = 4 * warp + 2 * h;                           // this half-warp's frames within the tile
    if (4 * warp < nf) {                                       // warp-uniform
      // ---- analysis: X = rfft(window * frame), exactly as stft512.cu ---------------------------------------
      const float* pa = xs + fA * A.P + 2 * l;
      const float* pb = pa + A.P;
      C2 a[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j < NJ) {
          float2 xa = *reinterpret_cast<const float2*>(pa + 32 * j);
          float2 xb2 = *reinterpret_cast<const float2*>(pb + 32 * j);
          const float2 wv = *reinterpret_cast<const float2*>(win + 2 * l + 32 * j);
          if (MASK_ALL || j == NJ - 1) {
            const int p0 = 2 * l + 32 * j;
            if (p0 >= L) { xa.x = 0.0f; xb2.x = 0.0f; }
            if (p0 + 1 >= L) { xa.y = 0.0f; xb2.y = 0.0f; }
          }
          a[j].re = make_float2(xa.x * wv.x, xb2.x * wv.x);
          a[j].im = make_float2(xa.y * wv.y, xb2.y * wv.y);
        } else {
          a[j].re = make_float2(0.0f, 0.0f);
          a[j].im = make_float2(0.0f, 0.0f);
        }
      }
      fft16<NJ>(a);
#pragma unroll
      for (int k2 = 1; k2 < 16; ++k2) a[dig(k2)] = cmul_s(a[dig(k2)], twr[k2], twi[k2]);
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        xr[k2 * kXRow + l] = a[dig(k2)].re;
        xi[k2 * kXRow + l] = a[dig(k2)].im;
      }
      __syncwarp();
#pragma unroll
      for (int m1 = 0; m1 < 16; ++m1) {
        a[m1].re = xr[l * kXRow + m1];
        a[m1].im = xi[l * kXRow + m1];
      }
      __syncwarp();
      fft16<16>(a);   // a[dig(k1)] = Z[16 k1 + l]
      C2 r[8];        // r[j] = Z[256 - (16 (7 - j) + l)] from the partner lane (lane 0: its own registers)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        C2 s = a[dig(8 + j)];
        if (l == 0) s = (j < 7) ? a[dig(9 + j)] : a[dig(0)];
        r[j].re.x = __shfl_sync(0xffffffffu, s.re.x, partner);
        r[j].re.y = __shfl_sync(0xffffffffu, s.re.y, partner);
        r[j].im.x = __shfl_sync(0xffffffffu, s.im.x, partner);
        r[j].im.y = __shfl_sync(0xffffffffu, s.im.y, partner);
      }
      // ---- G = g format'(s) X, then back to the half-length spectrum conj(E' + i O') ----------------------
      // Frames past the end of a partial tile take the gradient row of the tile's last valid frame and park their
      // result in their own (never read) rows of fbuf: no predicate per load / store (packed lanes are independent,
      // the overlap-add only visits frames < nf).
      const float* ga = A.gy + (b * A.N + n_lo + (fA < nf ? fA : nf - 1)) * 257;
      const float* gb = A.gy + (b * A.N + n_lo + (fA + 1 < nf ? fA + 1 : nf - 1)) * 257;
      const float2 eps2 = make_float2(A.eps, A.eps);
      C2 c128;        // lane 0: the self-mirrored bin 128
      {
        const float2 Xr = a[dig(8)].re, Xi = neg2(a[dig(8)].im);                 // X[128] = conj(Z[128])
        float2 g = make_float2(0.0f, 0.0f);
        if (l == 0) g = make_float2(__ldg(ga + 128), __ldg(gb + 128));
        const float2 d = fmt_grad(g, fma2(Xr, Xr, fma2(Xi, Xi, eps2)), A.fmt);
        c128.re = mul2s(__fmul2_rn(d, Xr), 2.0f);
        c128.im = mul2s(__fmul2_rn(d, Xi), 2.0f);
      }
#pragma unroll
      for (int k1 = 0; k1 < 8; ++k1) {
        const int k = 16 * k1 + l;
        const C2 z = a[dig(k1)], m = r[7 - k1];
        const float2 wv = tws[k];                                     // W512^k
        // X[k] = E + T, X[256 - k] = conj(E - T)  (real-input split, as stft512.cu)
        const float2 sr = add2(z.re, m.re), dr = sub2(z.re, m.re);
        const float2 si = add2(z.im, m.im), di = sub2(z.im, m.im);
        const float2 tr = mul2s(fma2s(dr, wv.y, mul2s(si, wv.x)), 0.5f);
        const float2 ti = mul2s(fma2s(dr, -wv.x, mul2s(si, wv.y)), 0.5f);
        const float2 Xr = fma2s(sr, 0.5f, tr), Xi = fma2s(di, 0.5f, ti);
        const float2 Mr = fma2s(sr, 0.5f, neg2(tr)), Mi = fma2s(di, -0.5f, ti);
        const float2 gk = make_float2(__ldg(ga + k), __ldg(gb + k));
        const float2 gm = make_float2(__ldg(ga + 256 - k), __ldg(gb + 256 - k));
        const float2 dk = fmt_grad(gk, fma2(Xr, Xr, fma2(Xi, Xi, eps2)), A.fmt);
        const float2 dm = fmt_grad(gm, fma2(Mr, Mr, fma2(Mi, Mi, eps2)), A.fmt);
        float2 Gr = __fmul2_rn(dk, Xr), Gi = __fmul2_rn(dk, Xi);     // Y'[k]
        float2 Hr = __fmul2_rn(dm, Mr), Hi = __fmul2_rn(dm, Mi);     // Y'[256 - k]
        if (k1 == 0 && l == 0) {                                     // DC and Nyquist: doubled, real
          Gr = mul2s(Gr, 2.0f); Gi = make_float2(0.0f, 0.0f);
          Hr = mul2s(Hr, 2.0f); Hi = make_float2(0.0f, 0.0f);
        }
        const float2 Er = add2(Gr, Hr), Ei = sub2(Gi, Hi), Dr = sub2(Gr, Hr), Di = add2(Gi, Hi);
        const float2 Or = fma2s(Di, wv.y, mul2s(Dr, wv.x));           // O' = D' conj(W512^k)
        const float2 Oi = fma2s(Dr, -wv.y, mul2s(Di, wv.x));
        a[dig(k1)].re = sub2(Er, Oi);                                 // c[k]       = conj(E' + i O')
        a[dig(k1)].im = neg2(add2(Ei, Or));
        r[7 - k1].re = add2(Er, Oi);                                  // c[256 - k] = E' - i O'
        r[7 - k1].im = sub2(Ei, Or);
      }
      // ---- mirrored halves back to their owner lanes; inputs in natural order c[16 j + l] ------------------
      C2 cc[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) cc[j] = a[dig(j)];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        C2 s = r[j];                                                  // c[256 - (16 (7 - j) + l)]
        if (l == 0) s = (j == 0) ? c128 : r[j - 1 < 0 ? 0 : j - 1];   // lane 0 owns c[128], c[16 (8 + j)]
        cc[8 + j].re.x = __shfl_sync(0xffffffffu, s.re.x, partner);
        cc[8 + j].re.y = __shfl_sync(0xffffffffu, s.re.y, partner);
        cc[8 + j].im.x = __shfl_sync(0xffffffffu, s.im.x, partner);
        cc[8 + j].im.y = __shfl_sync(0xffffffffu, s.im.y, partner);
      }
      // ---- synthesis: R = FFT_256(c); x~[2m] = Re R[m], x~[2m+1] = -Im R[m] --------------------------------
      fft16<16>(cc);
#pragma unroll
      for (int k2 = 1; k2 < 16; ++k2) cc[dig(k2)] = cmul_s(cc[dig(k2)], twr[k2], twi[k2]);
      __syncwarp();
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        xr[k2 * kXRow + l] = cc[dig(k2)].re;
        xi[k2 * kXRow + l] = cc[dig(k2)].im;
      }
      __syncwarp();
#pragma unroll
      for (int m1 = 0; m1 < 16; ++m1) {
        cc[m1].re = xr[l * kXRow + m1];
        cc[m1].im = xi[l * kXRow + m1];
      }
      __syncwarp();
      fft16<16>(cc);
      float* rowA = fbuf + static_cast<size_t>(fA) * L;
      float* rowB = rowA + L;
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) {
        const int s = 32 * k1 + 2 * l;
        if (s < L) {   // L is even: s + 1 < L too
          const C2 v = cc[dig(k1)];
          const float2 wv = *reinterpret_cast<const float2*>(win + s);
          *reinterpret_cast<float2*>(rowA + s) = make_float2(v.re.x * wv.x, -v.im.x * wv.y);
          *reinterpret_cast<float2*>(rowB + s) = make_float2(v.re.y * wv.x, -v.im.y * wv.y);
        }
      }
    }
    __syncthreads();
    // ---- adjoint of pad + unfold: plain overlap-add of the frame gradients, frame order ----------------------
    {
      const int P = A.P, Nm1 = static_cast<int>(A.N) - 1, nlo = static_cast<int>(n_lo);
      const int cnt = static_cast<int>(t1 - t0);
      const int q_first = static_cast<int>(q0) + tid;
      int ne = q_first / P, rr = q_first - ne * P;
      const int dq = kBT / P, drm = kBT - dq * P;
      float* outp = A.gx + b * A.T + t0;
      for (int i = tid; i < cnt; i += kBT) {
        int na = ne - A.c + (rr > A.m ? 1 : 0);
        if (na < 0) na = 0;
        const int nb = ne < Nm1 ? ne : Nm1;
        int n = ne;
        const float* fp = fbuf + (ne - nlo) * L + rr;
        float num = 0.0f;
        for (int u = 0; u <= A.c; ++u, --n, fp -= L - P)      // same trip count in every lane, frames predicated
          if (n >= na && n <= nb) num += *fp;
        outp[i] = num;
        ne += dq;
        rr += drm;
        if (rr >= P) { rr -= P; ++ne; }
      }
    }
    __syncthreads();
  }
}

}  // namespace

// DSB200_E_UNSUPPORTED outside the envelope; the caller then runs the general backward kernel.
int stft512_bwd_try(const float* x, const float* window, const float* gy, float* gx, int64_t batch, int64_t T_len,
                    const dsb200_stft_params* p, int device, cudaStream_t stream) {
  const dsb200_frame_params& f = p->frame;
  const int L = f.frame_length, P = f.frame_period;
  if (p->spec.fft_length != 512 || L > 512 || L < 2 || (L & 1) || (P & 1)) return DSB200_E_UNSUPPORTED;
  if (p->spec.out_format == DSB200_SPEC_COMPLEX || p->spec.has_relative_floor) return DSB200_E_UNSUPPORTED;
  if (f.pad_mode != DSB200_PAD_CONSTANT || f.zmean) return DSB200_E_UNSUPPORTED;
  if (T_len + L >= (int64_t{1} << 30)) return DSB200_E_UNSUPPORTED;
  const int tile = (kBFrames * P - L + 1) & ~3;
  if (tile < 8 * P) return DSB200_E_UNSUPPORTED;
  const void* tw = twiddle_table(device, 512, false, stream);
  if (tw == nullptr) return fail(DSB200_E_CUDA, "could not build the twiddle table for fft_length=512");
  BArgs A{};
  A.x = x;
  A.w = window;
  A.gy = gy;
  A.gx = gx;
  A.tw512 = static_cast<const float2*>(tw);
  A.batch = batch;
  A.T = T_len;
  A.N = dsb200_num_frames(T_len, P);
  A.L = L;
  A.P = P;
  A.left = f.center ? L / 2 : 0;
  A.tile = tile;
  A.tiles_per_utt = (T_len + tile - 1) / tile;
  A.c = (L - 1) / P;
  A.m = (L - 1) % P;
  A.fmt = p->spec.out_format;
  A.eps = static_cast<float>(p->spec.eps);
  A.span = ((kBFrames - 1) * P + 512 + 3) & ~3;
  A.alias = (static_cast<size_t>(4) * L * sizeof(float) >= 2 * kPlane * sizeof(float2)) ? 1 : 0;
  A.fbuf_off = A.alias ? 0 : kBW * 2 * kPlane;
  const size_t smem = 256 * sizeof(float2) + 512 * sizeof(float) + static_cast<size_t>(A.span) * sizeof(float) +
                      static_cast<size_t>(kBW) * 2 * kPlane * sizeof(float2) * (A.alias ? 1 : 2) +
                      static_cast<size_t>(kBFrames) * L * sizeof(float);
  if (smem > static_cast<size_t>(max_dynamic_smem(device))) return DSB200_E_UNSUPPORTED;
  const int64_t n_tiles = batch * A.tiles_per_utt;
  const int blocks = static_cast<int>(std::min<int64_t>(n_tiles, static_cast<int64_t>(sm_count(device)) * 2));
  auto launch = [&](auto kern) -> int {
    DSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<blocks, kBT, smem, stream>>>(A);
    return DSB200_OK;
  };
  int rc;
  if ((L + 31) / 32 == 13) rc = launch(stft512_bwd_kernel<13, false>);
  else rc = launch(stft512_bwd_kernel<16, true>);
  if (rc != DSB200_OK) return rc;
  return after_launch("stft512_bwd_kernel");
}

}  // namespace dsb200
