// Fused inverse STFT for fft_length = 512 (fp32, sm_100a): the mirror image of stft512.cu.
//
// Reference: diffsptk/modules/istft.py:186-193 = unframe(ifftr(Y)[..., :L]) (ifftr.py:138-140,
// unframe.py:164-211).  One CTA (8 warps, two CTAs per SM) owns a tile of consecutive output samples of one
// utterance and the <= 32 frames that overlap it:
//   * every half-warp inverse-transforms a PAIR of frames in registers with the forward kernel's packed
//     radix-16 x 16 machinery (fft16.cuh): the half-length spectrum conj(E + i O) is formed straight from the
//     global rows (each lane reads the 16 bins k = 16 j + l and their mirrors 256 - k, 128-byte coalesced),
//     x[2m] + i x[2m+1] = conj(FFT_256(conj(E + i O))) / 512, so no inverse butterflies are needed;
//   * the samples are multiplied by the synthesis window and parked in shared memory (the [B, N, L] frame tensor
//     never exists in HBM); after one CTA barrier each thread sums the <= ceil(L / P) contributions of its output
//     samples in frame order and divides by sum w^2 + 1e-16, writing the waveform once, coalesced.
// Algorithmic bytes: 2 056 B read + P * 4 B written per frame.
#include <algorithm>

#include "fft16.cuh"

namespace dsb200 {
namespace {

using namespace fft16_detail;

constexpr int kIW = 8;                 // warps per CTA
constexpr int kIT = kIW * 32;
constexpr int kTileFrames = 4 * kIW;   // frames a CTA can hold: one quad per warp

struct IArgs {
  const float2* Y;      // [batch, N, 257]
  const float* w;       // [L]
  const float2* tw512;  // W512^k, 512 entries
  float* out;           // [batch, T_out]
  int64_t batch, N, T_out, tiles_per_utt;
  int L, P, s, tile, alias;
  int fbuf_off;         // float2 units from the second plane array to the frame rows (0 when they alias)
  int c, m;             // (L - 1) / P and (L - 1) % P
};

__global__ void __launch_bounds__(kIT, 2) istft512_kernel(const IArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int l = lane & 15, h = lane >> 4;
  const int L = A.L;

  float2* tws = reinterpret_cast<float2*>(smem_raw);           // [256] W512^k
  float* wsm = reinterpret_cast<float*>(tws + 256);            // [512] window / 512 (zero beyond L)
  float* w2 = wsm + 512;                                       // [512] window^2
  float* dtab = w2 + 512;                                      // [512] sum of window^2 over a full set of frames
  float2* planes1 = reinterpret_cast<float2*>(dtab + 512);     // [kIW][2 kPlane] exchange planes of half-warp 1
  float2* planes0 = planes1 + kIW * 2 * kPlane;                // [kIW][2 kPlane] (only when they cannot alias)
  float* fbuf = reinterpret_cast<float*>(planes0 + A.fbuf_off);                              // [32][L]

  for (int i = tid; i < 256; i += kIT) tws[i] = A.tw512[i];
  for (int i = tid; i < 512; i += kIT) {
    const float v = i < L ? A.w[i] : 0.0f;
    wsm[i] = v * (1.0f / 512.0f);
    w2[i] = v * v;
  }
  // per-lane inter-pass twiddles W256^(l k2), as in the forward kernel
  float twr[16], twi[16];
#pragma unroll
  for (int k2 = 1; k2 < 16; ++k2) {
    const float2 v = A.tw512[2 * l * k2];
    twr[k2] = v.x;
    twi[k2] = v.y;
  }
  __syncthreads();
  // Away from the utterance ends every output sample with the same phase r = q mod P is covered by the same set
  // of frames, so its normaliser sum_u w^2[r + u P] is a table (summed in the order of the general loop below).
  for (int r = tid; r < A.P; r += kIT) {
    float d = 0.0f;
    for (int j = r; j < L; j += A.P) d += w2[j];
    dtab[r] = d;
  }
  // this half-warp's exchange planes: half-warp 0 borrows the first rows of the warp's own frames in fbuf
  // (dead until the samples are written, after the last plane read); half-warp 1 has private planes
  float2* xr = h ? planes1 + warp * 2 * kPlane
                 : (A.alias ? reinterpret_cast<float2*>(fbuf + static_cast<size_t>(4 * warp) * L)
                            : planes0 + warp * 2 * kPlane);
  float2* xi = xr + kPlane;
  __syncthreads();

  const int64_t n_tiles = A.batch * A.tiles_per_utt;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t b = tile / A.tiles_per_utt;
    const int64_t t0 = (tile - b * A.tiles_per_utt) * A.tile;
    const int64_t t1 = (t0 + A.tile < A.T_out) ? t0 + A.tile : A.T_out;
    const int64_t q0 = t0 + A.s, q1 = t1 - 1 + A.s;
    const int64_t n_lo = (q0 - L + 1 <= 0) ? 0 : (q0 - L + A.P) / A.P;
    int64_t n_hi = q1 / A.P;
    if (n_hi > A.N - 1) n_hi = A.N - 1;
    const int nf = static_cast<int>(n_hi - n_lo + 1);          // <= kTileFrames by the choice of A.tile
    if (tid == 0 && tile + gridDim.x < n_tiles) {              // pull the next tile's spectra into the L2 now
      const int64_t tn = tile + gridDim.x, bn = tn / A.tiles_per_utt;
      const int64_t qn = (tn - bn * A.tiles_per_utt) * A.tile + A.s;
      const int64_t nl = (qn - L + 1 <= 0) ? 0 : (qn - L + A.P) / A.P;
      prefetch_l2(A.Y, static_cast<size_t>(A.batch) * A.N * 257 * sizeof(float2), A.Y + (bn * A.N + nl) * 257,
                  static_cast<size_t>(kTileFrames) * 257 * sizeof(float2));
    }

    const int fA = 4 * warp + 2 * h;                           // this half-warp's frames within the tile
    if (4 * warp < nf) {                                       // warp-uniform: some frame of the quad exists
      // Frames past the end of a partial tile are computed from the tile's last valid row and parked in their own
      // (never read) rows of fbuf: no predicate or branch per load / store, the lanes of a packed operation are
      // independent and the overlap-add only visits frames < nf.
      const int rA = fA < nf ? fA : nf - 1, rB = fA + 1 < nf ? fA + 1 : nf - 1;
      const float2* ya = A.Y + (b * A.N + n_lo + rA) * 257;
      const float2* yb = A.Y + (b * A.N + n_lo + rB) * 257;
      C2 a[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int k = 16 * j + l;
        float2 pa = __ldg(ya + k), qa = __ldg(ya + 256 - k);   // X[k], X[256 - k] of frames A, B
        float2 pb = __ldg(yb + k), qb = __ldg(yb + 256 - k);
        if (k == 0) { pa.y = 0.0f; qa.y = 0.0f; pb.y = 0.0f; qb.y = 0.0f; }   // irfft ignores these
        // E = X[k] + conj X[256-k], D = X[k] - conj X[256-k], O = D conj(W512^k); input conj(E + i O)
        // scalar adds write the (frame A, frame B) halves in place: cheaper than packing the loaded values first
        const float2 Er = make_float2(pa.x + qa.x, pb.x + qb.x), Ei = make_float2(pa.y - qa.y, pb.y - qb.y);
        const float2 Dr = make_float2(pa.x - qa.x, pb.x - qb.x), Di = make_float2(pa.y + qa.y, pb.y + qb.y);
        const float2 wv = tws[k];
        const float2 Or = fma2s(Di, wv.y, mul2s(Dr, wv.x));
        const float2 Oi = fma2s(Dr, -wv.y, mul2s(Di, wv.x));
        a[j].re = sub2(Er, Oi);
        a[j].im = make_float2(-(Ei.x + Or.x), -(Ei.y + Or.y));
      }
      fft16<16>(a);
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
      fft16<16>(a);   // a[dig(k1)] = R[16 k1 + l];  x[2m] = Re R[m] / 512, x[2m+1] = -Im R[m] / 512
      float* rowA = fbuf + static_cast<size_t>(fA) * L;
      float* rowB = rowA + L;
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) {
        const int s = 32 * k1 + 2 * l;
        if (s < L) {
          const C2 v = a[dig(k1)];
          const float w0 = wsm[s], w1 = wsm[s + 1];
          if (s + 1 < L) {
            *reinterpret_cast<float2*>(rowA + s) = make_float2(v.re.x * w0, -v.im.x * w1);
            *reinterpret_cast<float2*>(rowB + s) = make_float2(v.re.y * w0, -v.im.y * w1);
          } else {
            rowA[s] = v.re.x * w0;
            rowB[s] = v.re.y * w0;
          }
        }
      }
    }
    __syncthreads();
    // Overlap-add in frame order, then the sum-of-squares normalisation (unframe.py:204-206).  All index
    // arithmetic is 32-bit and incremental: one division per thread and tile, none per sample.
    {
      const int P = A.P, Nm1 = static_cast<int>(A.N) - 1, nlo = static_cast<int>(n_lo);
      const int cnt = static_cast<int>(t1 - t0);
      const int q_first = static_cast<int>(q0) + tid;
      int ne = q_first / P, r = q_first - ne * P;          // frame whose start is at or before q, offset in it
      const int dq = kIT / P, dr = kIT - dq * P;           // advance of (ne, r) per kIT samples
      float* outp = A.out + b * A.T_out + t0;
      for (int i = tid; i < cnt; i += kIT) {
        // frames n with 0 <= q - n P <= L - 1:  n in [ne - c + (r > m), ne], c = (L-1) / P, m = (L-1) % P.
        // The loop runs c + 1 times for every lane (no divergence); out-of-range frames are predicated off.
        const int first = ne - A.c + (r > A.m ? 1 : 0);
        const float* fp = fbuf + (ne - nlo) * L + r;
        float num = 0.0f, den;
        if (first >= 0 && ne <= Nm1) {
          // interior sample: all of its frames exist; only the numerator is summed
          const int terms = ne - first + 1;
#pragma unroll 5
          for (int u = 0; u < terms; ++u, fp -= L - P) num += *fp;
          den = dtab[r];
        } else {
          const int na = first < 0 ? 0 : first;
          const int nb = ne < Nm1 ? ne : Nm1;
          int n = ne, j = r;
          den = 0.0f;
          for (int u = 0; u <= A.c; ++u, --n, fp -= L - P, j += P) {
            if (n >= na && n <= nb) {
              num += *fp;
              den += w2[j];
            }
          }
        }
        outp[i] = num / (den + 1e-16f);
        ne += dq;
        r += dr;
        if (r >= P) { r -= P; ++ne; }
      }
    }
    __syncthreads();
  }
}

}  // namespace

// DSB200_E_UNSUPPORTED outside the envelope (fft_length 512, even frame_length in [.., 512], long enough hop).
int istft512_try(const float* Y, const float* w, float* out, int64_t batch, int64_t N, int64_t T_out, int L, int P,
                 int n, int center, int device, cudaStream_t stream) {
  if (n != 512 || L > 512 || (L & 1) || L < 2) return DSB200_E_UNSUPPORTED;
  if (T_out + L >= (int64_t{1} << 30) || N >= (int64_t{1} << 30)) return DSB200_E_UNSUPPORTED;   // 32-bit sample indices
  int tile = (kTileFrames * P - L + 1) & ~3;   // largest tile whose overlapping frames number <= kTileFrames
  if (tile < 8 * P) return DSB200_E_UNSUPPORTED;
  const void* tw = twiddle_table(device, 512, false, stream);
  if (tw == nullptr) return fail(DSB200_E_CUDA, "could not build the twiddle table for fft_length=512");
  IArgs A{};
  A.Y = reinterpret_cast<const float2*>(Y);
  A.w = w;
  A.tw512 = static_cast<const float2*>(tw);
  A.out = out;
  A.batch = batch;
  A.N = N;
  A.T_out = T_out;
  A.L = L;
  A.P = P;
  A.s = center ? L / 2 : 0;
  A.tile = tile;
  A.c = (L - 1) / P;
  A.m = (L - 1) % P;
  A.tiles_per_utt = (T_out + tile - 1) / tile;
  A.alias = (static_cast<size_t>(4) * L * sizeof(float) >= 2 * kPlane * sizeof(float2)) ? 1 : 0;
  A.fbuf_off = A.alias ? 0 : kIW * 2 * kPlane;
  const size_t smem = 256 * sizeof(float2) + 3 * 512 * sizeof(float) +
                      static_cast<size_t>(kIW) * 2 * kPlane * sizeof(float2) * (A.alias ? 1 : 2) +
                      static_cast<size_t>(kTileFrames) * L * sizeof(float);
  if (smem > static_cast<size_t>(max_dynamic_smem(device))) return DSB200_E_UNSUPPORTED;
  DSB_CUDA(cudaFuncSetAttribute(istft512_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int64_t n_tiles = batch * A.tiles_per_utt;
  const int blocks = static_cast<int>(std::min<int64_t>(n_tiles, static_cast<int64_t>(sm_count(device)) * 2));
  istft512_kernel<<<blocks, kIT, smem, stream>>>(A);
  return after_launch("istft512_kernel");
}

}  // namespace dsb200
