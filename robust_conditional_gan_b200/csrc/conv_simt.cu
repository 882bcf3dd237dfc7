// CUDA-core implicit-GEMM convolution family (fprop / dgrad / wgrad), NHWC x HWIO.
//
// This is (a) the fp32 "parity mode" path for every conv / deconv / linear of the
// reference (tf.nn.conv2d mnist/ops.py:62, conv2d.py:181-187; tf.nn.conv2d_transpose
// mnist/ops.py:78; tf.matmul ops.py:114-116, linear.py:163-173) and (b) the path for
// the degenerate layers in bf16 mode (cin=1, cout=1, cout=10 ...: arithmetic intensity
// < 30 flop/B) that do not belong on the tensor pipe.
//
// GEMM view (fp32 accumulate, 64x64x16 or 256x16x16 CTA tile, 256 threads):
//   fprop: M = n*ho*wo, N = cout, K = kh*kw*cin   A = im2col(x)      B = w[K][cout]
//   dgrad: M = n*h*w,   N = cin,  K = kh*kw*cout  A = col2im^T(dy)   B = w^T
//   wgrad: M = kh*kw*cin, N = cout, K = n*ho*wo   A = im2col(x)^T    B = dy     (split-K)
#include "common.cuh"

namespace {

enum { MODE_FPROP = 0, MODE_DGRAD = 1, MODE_WGRAD = 2 };
constexpr int BK = 16;

struct ConvP {
  int n, h, w, cin, ho, wo, cout, kh, kw, stride, pad_t, pad_l, ldx, ldy;
  int M, N, K;
  const float* bias;
  int act;
  float leak;
  int accumulate;
  int klen;  // K range per blockIdx.z (wgrad split-K), multiple of BK
};

template <int MODE, typename T, typename TO, int BM, int BN>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvP p, const T* __restrict__ act_in,  // x (fprop/wgrad) or dy (dgrad)
                                                        const float* __restrict__ wt,           // weights (fprop/dgrad)
                                                        const T* __restrict__ dy_in,            // dy (wgrad)
                                                        void* __restrict__ out_) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int A_PER_T = BM * BK / 256, B_PER_T = BN * BK / 256 > 0 ? BN * BK / 256 : 1;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];

  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * p.klen;
  const int kend = min(p.K, kbeg + p.klen);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

  // ---- per-thread A row bookkeeping
  // fprop/dgrad: k-fast mapping: k_l = t%16, rows m_l = t/16 + 16*i
  // wgrad     : m-fast mapping: m_l = idx % BM, k_l = idx / BM
  int rowbase[A_PER_T];  // element offset of (nb, 0, 0, 0) for fprop/dgrad rows
  int rowy[A_PER_T], rowx[A_PER_T];
  bool rowok[A_PER_T];
  if (MODE != MODE_WGRAD) {
#pragma unroll
    for (int i = 0; i < A_PER_T; i++) {
      int m = m0 + (t >> 4) + 16 * i;
      rowok[i] = m < p.M;
      int mm = rowok[i] ? m : 0;
      if (MODE == MODE_FPROP) {
        int ox = mm % p.wo, r = mm / p.wo, oy = r % p.ho, nb = r / p.ho;
        rowbase[i] = nb * p.h * p.w;
        rowy[i] = oy * p.stride - p.pad_t;
        rowx[i] = ox * p.stride - p.pad_l;
      } else {
        int ix = mm % p.w, r = mm / p.w, iy = r % p.h, nb = r / p.h;
        rowbase[i] = nb * p.ho * p.wo;
        rowy[i] = iy + p.pad_t;
        rowx[i] = ix + p.pad_l;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < A_PER_T; i++) {
      int idx = t + 256 * i;
      int m = m0 + idx % BM;
      rowok[i] = m < p.M;
      int mm = rowok[i] ? m : 0;
      int ci = mm % p.cin, tap = mm / p.cin;
      rowbase[i] = ci;
      rowy[i] = tap / p.kw - p.pad_t;  // ky - pad_t
      rowx[i] = tap % p.kw - p.pad_l;
    }
  }

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ------------------------------------------------ A tile
    if (MODE == MODE_FPROP) {
      int k = k0 + (t & 15);
      bool kok = k < kend;
      int kk = kok ? k : 0;
      int ci = kk % p.cin, tap = kk / p.cin;
      int ky = tap / p.kw, kx = tap % p.kw;
#pragma unroll
      for (int i = 0; i < A_PER_T; i++) {
        int iy = rowy[i] + ky, ix = rowx[i] + kx;
        float v = 0.f;
        if (kok && rowok[i] && iy >= 0 && iy < p.h && ix >= 0 && ix < p.w)
          v = to_f(act_in[(size_t)(rowbase[i] + iy * p.w + ix) * p.ldx + ci]);
        As[t & 15][(t >> 4) + 16 * i] = v;
      }
    } else if (MODE == MODE_DGRAD) {
      int k = k0 + (t & 15);
      bool kok = k < kend;
      int kk = kok ? k : 0;
      int co = kk % p.cout, tap = kk / p.cout;
      int ky = tap / p.kw, kx = tap % p.kw;
#pragma unroll
      for (int i = 0; i < A_PER_T; i++) {
        int ty_ = rowy[i] - ky, tx_ = rowx[i] - kx;
        float v = 0.f;
        if (kok && rowok[i] && ty_ >= 0 && tx_ >= 0) {
          int oy = ty_ / p.stride, ox = tx_ / p.stride;
          if (oy * p.stride == ty_ && ox * p.stride == tx_ && oy < p.ho && ox < p.wo)
            v = to_f(act_in[(size_t)(rowbase[i] + oy * p.wo + ox) * p.ldy + co]);
        }
        As[t & 15][(t >> 4) + 16 * i] = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < A_PER_T; i++) {
        int idx = t + 256 * i;
        int k = k0 + idx / BM;
        float v = 0.f;
        if (k < kend && rowok[i]) {
          int ox = k % p.wo, r = k / p.wo, oy = r % p.ho, nb = r / p.ho;
          int iy = oy * p.stride + rowy[i], ix = ox * p.stride + rowx[i];
          if (iy >= 0 && iy < p.h && ix >= 0 && ix < p.w)
            v = to_f(act_in[(size_t)((nb * p.h + iy) * p.w + ix) * p.ldx + rowbase[i]]);
        }
        As[idx / BM][idx % BM] = v;
      }
    }
    // ------------------------------------------------ B tile
    if (MODE == MODE_FPROP) {
#pragma unroll
      for (int i = 0; i < B_PER_T; i++) {
        int idx = t + 256 * i;
        if (idx < BN * BK) {
          int nl = idx % BN, kl = idx / BN;
          int k = k0 + kl, nn = n0 + nl;
          Bs[kl][nl] = (k < kend && nn < p.N) ? wt[(size_t)k * p.cout + nn] : 0.f;
        }
      }
    } else if (MODE == MODE_DGRAD) {
#pragma unroll
      for (int i = 0; i < B_PER_T; i++) {
        int idx = t + 256 * i;
        if (idx < BN * BK) {
          int kl = idx % BK, nl = idx / BK;
          int k = k0 + kl, ci = n0 + nl;
          float v = 0.f;
          if (k < kend && ci < p.N) {
            int co = k % p.cout, tap = k / p.cout;
            v = wt[((size_t)tap * p.cin + ci) * p.cout + co];
          }
          Bs[kl][nl] = v;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < B_PER_T; i++) {
        int idx = t + 256 * i;
        if (idx < BN * BK) {
          int nl = idx % BN, kl = idx / BN;
          int k = k0 + kl, nn = n0 + nl;
          Bs[kl][nl] = (k < kend && nn < p.N) ? to_f(dy_in[(size_t)k * p.ldy + nn]) : 0.f;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i++) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; j++) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ------------------------------------------------ epilogue
  if (MODE == MODE_WGRAD) {
    float* out = reinterpret_cast<float*>(out_) + (size_t)blockIdx.z * p.M * p.N;
#pragma unroll
    for (int i = 0; i < TM; i++) {
      int m = m0 + ty * TM + i;
      if (m >= p.M) continue;
#pragma unroll
      for (int j = 0; j < TN; j++) {
        int nn = n0 + tx * TN + j;
        if (nn < p.N) out[(size_t)m * p.N + nn] = acc[i][j];
      }
    }
  } else {
    TO* out = reinterpret_cast<TO*>(out_);
    const int ld = MODE == MODE_FPROP ? p.ldy : p.ldx;
#pragma unroll
    for (int i = 0; i < TM; i++) {
      int m = m0 + ty * TM + i;
      if (m >= p.M) continue;
#pragma unroll
      for (int j = 0; j < TN; j++) {
        int nn = n0 + tx * TN + j;
        if (nn >= p.N) continue;
        float v = acc[i][j];
        if (p.bias) v += p.bias[nn];
        v = act_fwd(v, p.act, p.leak);
        size_t o = (size_t)m * ld + nn;
        if (p.accumulate) v += to_f(out[o]);
        out[o] = from_f<TO>(v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Narrow-output path: N <= 4 output channels (g_h3's cout = 1, G.Output's cout = 3, the dgrad into a 1- or 3-channel
// image).  These are bandwidth-bound gather-dots (AI < 30 flop/B), not GEMMs: 8 lanes share one output
// pixel, each lane streams 16-byte channel chunks of the source pixels selected by the filter taps (coalesced 128 B
// per pixel) against the weights staged in shared memory as [tap][n][c]; a 3-step shuffle finishes the dot product.
template <int MODE, int NOUT, typename T, typename TO>
__global__ void __launch_bounds__(256) conv_narrow_kernel(ConvP p, const T* __restrict__ src, const float* __restrict__ wt,
                                                          TO* __restrict__ out) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  constexpr int V = 16 / sizeof(T);                 // elements per 16-byte chunk
  extern __shared__ float wsm[];                    // [taps][NOUT][Cp], Cp = C rounded up to V, zero padded
  const int taps = p.kh * p.kw;
  const int C = MODE == MODE_FPROP ? p.cin : p.cout;   // contraction channels
  const int nchunk = (C + V - 1) / V;
  const int Cp = nchunk * V;
  const int ldsrc = MODE == MODE_FPROP ? p.ldx : p.ldy;
  for (int i = threadIdx.x; i < taps * NOUT * Cp; i += 256) {
    int c = i % Cp, r = i / Cp, n = r % NOUT, tap = r / NOUT;
    // w is [tap][cin][cout]: fprop contracts cin (n = cout index), dgrad contracts cout (n = cin index)
    float v = 0.f;
    if (c < C) v = MODE == MODE_FPROP ? wt[((size_t)tap * p.cin + c) * p.cout + n] : wt[((size_t)tap * p.cin + n) * p.cout + c];
    wsm[i] = v;
  }
  __syncthreads();
  const int sub = threadIdx.x & 7;                  // lane within the 8-lane group
  const int SH = MODE == MODE_FPROP ? p.h : p.ho, SW = MODE == MODE_FPROP ? p.w : p.wo;
  const int OHh = MODE == MODE_FPROP ? p.ho : p.h, OWw = MODE == MODE_FPROP ? p.wo : p.w;
  const int st = p.stride;
  for (int mb = blockIdx.x * 32; mb < p.M; mb += gridDim.x * 32) {   // block-uniform trip count
    const int m = mb + (threadIdx.x >> 3);
    const bool valid = m < p.M;
    const int mm = valid ? m : 0;
    const int b = mm % OWw;
    const int r = mm / OWw;
    const int a = r % OHh, n = r / OHh;
    float acc[NOUT];
#pragma unroll
    for (int j = 0; j < NOUT; j++) acc[j] = 0.f;
    // fprop: every tap, source = a*stride - pad + k.   dgrad: only taps with (a + pad - k) % stride == 0, i.e.
    // k = k0 + stride*i, source = (a + pad - k) / stride (exact), so the loop steps by the stride and never divides
    const int ky0 = MODE == MODE_FPROP ? 0 : (a + p.pad_t) % st, kx0 = MODE == MODE_FPROP ? 0 : (b + p.pad_l) % st;
    const int kstep = MODE == MODE_FPROP ? 1 : st;
    const int sy0 = MODE == MODE_FPROP ? a * st - p.pad_t : (a + p.pad_t - ky0) / st;
    const int sx0 = MODE == MODE_FPROP ? b * st - p.pad_l : (b + p.pad_l - kx0) / st;
    if (valid) {
      for (int ky = ky0, iy = 0; ky < p.kh; ky += kstep, iy++) {
        const int sy = MODE == MODE_FPROP ? sy0 + ky : sy0 - iy;
        if (sy < 0 || sy >= SH) continue;
        for (int kx = kx0, ix = 0; kx < p.kw; kx += kstep, ix++) {
          const int sx = MODE == MODE_FPROP ? sx0 + kx : sx0 - ix;
          if (sx < 0 || sx >= SW) continue;
          const T* sp = src + ((size_t)(n * SH + sy) * SW + sx) * ldsrc;
          const float* wp = wsm + (size_t)(ky * p.kw + kx) * NOUT * Cp;
          for (int ck0 = sub; ck0 < nchunk; ck0 += 32) {
            // up to 4 chunks of this lane are fetched before any of them is consumed (loads overlap)
            uint4 q[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const int ck = ck0 + 8 * u;
              q[u] = ck < nchunk ? *reinterpret_cast<const uint4*>(sp + ck * V) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const int ck = ck0 + 8 * u;
              if (ck >= nchunk) break;
              float xv[V];
              if (sizeof(T) == 2) {
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q[u]);
#pragma unroll
                for (int e = 0; e < V / 2; e++) { float2 f = __bfloat1622float2(h[e]); xv[2 * e] = f.x; xv[2 * e + 1] = f.y; }
              } else {
                const float* f = reinterpret_cast<const float*>(&q[u]);
#pragma unroll
                for (int e = 0; e < V; e++) xv[e] = f[e];
              }
#pragma unroll
              for (int j = 0; j < NOUT; j++) {
                const float4* w4 = reinterpret_cast<const float4*>(wp + j * Cp + ck * V);
#pragma unroll
                for (int e = 0; e < V / 4; e++) {
                  float4 wv = w4[e];
                  acc[j] = fmaf(xv[4 * e], wv.x, acc[j]);
                  acc[j] = fmaf(xv[4 * e + 1], wv.y, acc[j]);
                  acc[j] = fmaf(xv[4 * e + 2], wv.z, acc[j]);
                  acc[j] = fmaf(xv[4 * e + 3], wv.w, acc[j]);
                }
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NOUT; j++) {
      acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
      acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 2);
      acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 4);
    }
    if (valid && sub < NOUT) {
      float v = acc[0];
#pragma unroll
      for (int j = 1; j < NOUT; j++)
        if (sub == j) v = acc[j];
      if (p.bias) v += p.bias[sub];
      v = act_fwd(v, p.act, p.leak);
      size_t o = (size_t)m * (MODE == MODE_FPROP ? p.ldy : p.ldx) + sub;
      if (p.accumulate) v += to_f(out[o]);
      out[o] = from_f<TO>(v);
    }
  }
}

template <int MODE, int NOUT, typename T, typename TO>
void launch_narrow_n(const ConvP& p, const void* src, const float* w, void* out, size_t shb, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(conv_narrow_kernel<MODE, NOUT, T, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr_done = true;
  }
  long groups = ((long)p.M + 31) / 32;
  int grid = (int)(groups < 1 ? 1 : (groups > RCGAN_NUM_SMS * 16 ? RCGAN_NUM_SMS * 16 : groups));
  launch_pdl(conv_narrow_kernel<MODE, NOUT, T, TO>, grid, 256, shb, st, p, (const T*)src, w, (TO*)out);
}

template <int MODE, typename T, typename TO>
bool launch_narrow(const ConvP& p, const void* src, const float* w, void* out, cudaStream_t st) {
  const int C = MODE == MODE_FPROP ? p.cin : p.cout, ldsrc = MODE == MODE_FPROP ? p.ldx : p.ldy;
  const int V = 16 / (int)sizeof(T);
  const int Cp = ((C + V - 1) / V) * V;
  size_t shb = (size_t)p.kh * p.kw * p.N * Cp * sizeof(float);
  if (p.N > 4 || C < 32 || ldsrc % V != 0 || Cp > ldsrc || shb > 96 * 1024) return false;
  switch (p.N) {
    case 1: launch_narrow_n<MODE, 1, T, TO>(p, src, w, out, shb, st); break;
    case 2: launch_narrow_n<MODE, 2, T, TO>(p, src, w, out, shb, st); break;
    case 3: launch_narrow_n<MODE, 3, T, TO>(p, src, w, out, shb, st); break;
    default: launch_narrow_n<MODE, 4, T, TO>(p, src, w, out, shb, st); break;
  }
  return true;
}

// ------------------------------------------------------------------------------------------------
// Narrow wgrad for cout <= 4 (G.Output: 256 -> 3 channels; measured 2.3 ms vs 13.3 ms on the split-K path at batch 512).
// The mirrored case (cin <= 4: D.Block.1, g_h3, d_h0_conv) was measured 3-5x SLOWER than split-K -- its per-tap index math
// is replicated across the channel warps -- and therefore stays on the split-K kernel.
// dW[tap][ci][co] = sum_pixels x[pixel shifted by tap][ci] * dy[pixel][co] is then a bandwidth-bound reduction: one
// thread per channel of the WIDE operand (coalesced), the <= 4 narrow values are warp-broadcast loads, taps*narrow
// accumulators live in registers over a contiguous slab of pixels, one red.global.add per accumulator per block.
template <bool XWIDE, int KS, int CN, typename T>
__global__ void __launch_bounds__(256) conv_wgrad_narrow_kernel(ConvP p, const T* __restrict__ x, const T* __restrict__ dy,
                                                                float* __restrict__ dw, int pix_per_block) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  constexpr int TAPS = KS * KS;
  const int cw = XWIDE ? p.cin : p.cout;       // wide channel count; CN = narrow channel count (<= 4)
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  const bool cok = c < cw;
  float acc[TAPS][CN];
#pragma unroll
  for (int t = 0; t < TAPS; t++)
#pragma unroll
    for (int j = 0; j < CN; j++) acc[t][j] = 0.f;
  const long Mpix = (long)p.n * p.ho * p.wo;
  const long m0 = (long)blockIdx.x * pix_per_block, m1 = min(Mpix, m0 + pix_per_block);
  int ox = (int)(m0 % p.wo);
  long r = m0 / p.wo;
  int oy = (int)(r % p.ho), n = (int)(r / p.ho);
  for (long m = m0; m < m1; m++) {
    float wv = 0.f, nv[CN];
    if (XWIDE) {
#pragma unroll
      for (int j = 0; j < CN; j++) nv[j] = to_f(dy[(size_t)m * p.ldy + j]);
    } else {
      if (cok) wv = to_f(dy[(size_t)m * p.ldy + c]);
    }
    const int iy0 = oy * p.stride - p.pad_t, ix0 = ox * p.stride - p.pad_l;
#pragma unroll
    for (int ky = 0; ky < KS; ky++) {
      const int iy = iy0 + ky;
      if (iy < 0 || iy >= p.h) continue;                                  // block-uniform branches
#pragma unroll
      for (int kx = 0; kx < KS; kx++) {
        const int ix = ix0 + kx;
        if (ix < 0 || ix >= p.w) continue;
        const size_t xoff = ((size_t)(n * p.h + iy) * p.w + ix) * p.ldx;
        if (XWIDE) {
          const float xv = cok ? to_f(x[xoff + c]) : 0.f;
#pragma unroll
          for (int j = 0; j < CN; j++) acc[ky * KS + kx][j] = fmaf(xv, nv[j], acc[ky * KS + kx][j]);
        } else {
#pragma unroll
          for (int j = 0; j < CN; j++) acc[ky * KS + kx][j] = fmaf(to_f(x[xoff + j]), wv, acc[ky * KS + kx][j]);
        }
      }
    }
    if (++ox == p.wo) { ox = 0; if (++oy == p.ho) { oy = 0; n++; } }
  }
  if (!cok) return;
#pragma unroll
  for (int t = 0; t < TAPS; t++)
#pragma unroll
    for (int j = 0; j < CN; j++) {
      const int ci = XWIDE ? c : j, co = XWIDE ? j : c;
      atomicAdd(&dw[((size_t)t * p.cin + ci) * p.cout + co], acc[t][j]);
    }
}

template <bool XW, int KS, typename T>
void launch_wgrad_narrow_cn(const ConvP& p, int cn, const void* x, const void* dy, float* dw, dim3 grid, int bx, int ppb,
                            cudaStream_t st) {
  switch (cn) {
    case 1: launch_pdl(conv_wgrad_narrow_kernel<XW, KS, 1, T>, grid, bx, 0, st, p, (const T*)x, (const T*)dy, dw, ppb); break;
    case 2: launch_pdl(conv_wgrad_narrow_kernel<XW, KS, 2, T>, grid, bx, 0, st, p, (const T*)x, (const T*)dy, dw, ppb); break;
    case 3: launch_pdl(conv_wgrad_narrow_kernel<XW, KS, 3, T>, grid, bx, 0, st, p, (const T*)x, (const T*)dy, dw, ppb); break;
    default: launch_pdl(conv_wgrad_narrow_kernel<XW, KS, 4, T>, grid, bx, 0, st, p, (const T*)x, (const T*)dy, dw, ppb); break;
  }
}

template <typename T>
bool launch_wgrad_narrow(const ConvP& p, const void* x, const void* dy, float* dw, int accumulate, cudaStream_t st) {
  const int taps = p.kh * p.kw;
  const bool xwide = p.cout <= 4 && p.cin >= 32, dywide = false;
  if ((!xwide && !dywide) || p.kh != p.kw || (p.kh != 1 && p.kh != 3 && p.kh != 5)) return false;
  const int cn = xwide ? p.cout : p.cin, cw = xwide ? p.cin : p.cout;
  if (p.kh == 5 && cn > 2) return false;        // 25 taps x 4 accumulators would spill
  if (!accumulate) cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)taps * p.cin * p.cout, st);
  const int bx = cw >= 256 ? 256 : ((cw + 31) / 32) * 32;
  const int gy = (cw + bx - 1) / bx;
  const long Mpix = (long)p.n * p.ho * p.wo;
  long want = (4L * RCGAN_NUM_SMS + gy - 1) / gy;
  long ppb = (Mpix + want - 1) / want;
  if (ppb < 64) ppb = 64;
  dim3 grid((unsigned)((Mpix + ppb - 1) / ppb), gy);
  if (xwide) {
    if (p.kh == 1) launch_wgrad_narrow_cn<true, 1, T>(p, cn, x, dy, dw, grid, bx, (int)ppb, st);
    else if (p.kh == 3) launch_wgrad_narrow_cn<true, 3, T>(p, cn, x, dy, dw, grid, bx, (int)ppb, st);
    else launch_wgrad_narrow_cn<true, 5, T>(p, cn, x, dy, dw, grid, bx, (int)ppb, st);
  } else {
    if (p.kh == 1) launch_wgrad_narrow_cn<false, 1, T>(p, cn, x, dy, dw, grid, bx, (int)ppb, st);
    else if (p.kh == 3) launch_wgrad_narrow_cn<false, 3, T>(p, cn, x, dy, dw, grid, bx, (int)ppb, st);
    else launch_wgrad_narrow_cn<false, 5, T>(p, cn, x, dy, dw, grid, bx, (int)ppb, st);
  }
  return true;
}

// dw (=|+=) sum over splits
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, long mn, int nsplit, int accumulate) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mn) return;
  float s = 0.f;
  for (int z = 0; z < nsplit; z++) s += ws[(size_t)z * mn + i];
  dw[i] = accumulate ? dw[i] + s : s;
}

int check_desc(const rcgan_conv_desc* d, const char* who) {
  RCGAN_CHECK_ARG(d, "%s: null desc", who);
  RCGAN_CHECK_ARG(d->n > 0 && d->h > 0 && d->w > 0 && d->cin > 0 && d->cout > 0 && d->ho > 0 && d->wo > 0,
                  "%s: non-positive dims", who);
  RCGAN_CHECK_ARG(d->kh > 0 && d->kw > 0 && d->stride > 0 && d->pad_t >= 0 && d->pad_l >= 0, "%s: bad filter", who);
  RCGAN_CHECK_ARG(d->ldx >= d->cin && d->ldy >= d->cout, "%s: ld smaller than channels", who);
  RCGAN_CHECK_ARG(d->dtype == RCGAN_F32 || d->dtype == RCGAN_BF16, "%s: bad dtype %d", who, d->dtype);
  RCGAN_CHECK_ARG((long)d->n * d->h * d->w * d->ldx < 2147483647L && (long)d->n * d->ho * d->wo * d->ldy < 2147483647L,
                  "%s: tensor exceeds 2^31 elements", who);
  return 0;
}

ConvP make_p(const rcgan_conv_desc* d) {
  ConvP p;
  p.n = d->n; p.h = d->h; p.w = d->w; p.cin = d->cin; p.ho = d->ho; p.wo = d->wo; p.cout = d->cout;
  p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad_t = d->pad_t; p.pad_l = d->pad_l;
  p.ldx = d->ldx; p.ldy = d->ldy;
  p.bias = nullptr; p.act = 0; p.leak = 0.f; p.accumulate = 0; p.klen = 0;
  p.M = p.N = p.K = 0;
  return p;
}

int wgrad_splits(const ConvP& p) {
  long tiles = (long)ceil_div(p.M, 64) * ceil_div(p.N, 64);
  long want = (2L * RCGAN_NUM_SMS + tiles - 1) / tiles;
  long maxs = (p.K + 8 * BK - 1) / (8 * BK);  // at least 128 k per split
  long s = want < 1 ? 1 : want;
  if (s > maxs) s = maxs;
  if (s > 512) s = 512;
  if (s < 1) s = 1;
  return (int)s;
}

// Skinny linear layers (4 < N <= 16 outputs, e.g. the MNIST classifier 784 -> 10 at batch 1024, mnist/model.py:759-768):
// a 256x16 tile grid would be 4 blocks.  One warp per row instead: lanes stride over K against the smem-resident
// weights (row pitch 17 floats: conflict-free), 16 register accumulators, shuffle reduction, lane j writes output j.
// K is walked in chunks of SKINNY_KC rows of the weight matrix staged in shared memory (the CIFAR permutation classifier is
// 3072 -> 10, gan_resnet.py:458-466: its 209 KB of padded weights do not fit at once, and the generic 256x16 tile kernel ran it
// on ONE or TWO blocks: 433 us per call, 2.6 ms per RCGAN-U iteration).  A block owns 8 rows (one per warp) per pass.
constexpr int SKINNY_KC = 1024;
template <typename T, typename TO>
__global__ void __launch_bounds__(256) linear_skinny_kernel(ConvP p, const T* __restrict__ x, const float* __restrict__ wt,
                                                            TO* __restrict__ out) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  extern __shared__ float wsk[];   // [min(K, SKINNY_KC)][17]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int m0 = blockIdx.x * 8; m0 < p.M; m0 += gridDim.x * 8) {
    const int m = m0 + warp;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = 0.f;
    for (int k0 = 0; k0 < p.K; k0 += SKINNY_KC) {
      const int kc = min(SKINNY_KC, p.K - k0);
      __syncthreads();           // the previous chunk (or pass) is consumed
      for (int i = threadIdx.x; i < kc * 16; i += 256) {
        const int k = i >> 4, j = i & 15;
        wsk[k * 17 + j] = j < p.N ? wt[(size_t)(k0 + k) * p.N + j] : 0.f;
      }
      __syncthreads();
      if (m < p.M) {
        const T* xr = x + (size_t)m * p.ldx + k0;
        for (int k = lane; k < kc; k += 32) {
          const float xv = to_f(xr[k]);
          const float* wr = wsk + k * 17;
#pragma unroll
          for (int j = 0; j < 16; j++) acc[j] = fmaf(xv, wr[j], acc[j]);
        }
      }
    }
    if (m < p.M) {
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < 16; j++) {
        const float sj = warp_sum(acc[j]);
        if (lane == j) v = sj;
      }
      if (lane < p.N) {
        if (p.bias) v += p.bias[lane];
        out[(size_t)m * p.ldy + lane] = from_f<TO>(act_fwd(v, p.act, p.leak));
      }
    }
  }
}

template <typename T, typename TO>
bool launch_skinny(const ConvP& p, const void* x, const float* w, void* out, cudaStream_t st) {
  const size_t shb = (size_t)(p.K < SKINNY_KC ? p.K : SKINNY_KC) * 17 * sizeof(float);
  if (p.kh != 1 || p.kw != 1 || p.stride != 1 || p.N <= 4 || p.N > 16 || p.M < 256) return false;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(linear_skinny_kernel<T, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, SKINNY_KC * 17 * (int)sizeof(float));
    attr_done = true;
  }
  int grid = ceil_div(p.M, 8);
  if (grid > RCGAN_NUM_SMS * 2) grid = RCGAN_NUM_SMS * 2;
  launch_pdl(linear_skinny_kernel<T, TO>, grid, 256, shb, st, p, (const T*)x, w, (TO*)out);
  return true;
}

template <int MODE, typename T, typename TO>
void launch(const ConvP& p, const void* a, const float* w, const void* dy, void* out, int nz, cudaStream_t st) {
  if (p.N <= 16 && MODE != MODE_WGRAD) {
    dim3 grid(ceil_div(p.M, 256), ceil_div(p.N, 16), nz);
    launch_pdl(conv_simt_kernel<MODE, T, TO, 256, 16>, grid, 256, 0, st, p, (const T*)a, w, (const T*)dy, out);
  } else {
    dim3 grid(ceil_div(p.M, 64), ceil_div(p.N, 64), nz);
    launch_pdl(conv_simt_kernel<MODE, T, TO, 64, 64>, grid, 256, 0, st, p, (const T*)a, w, (const T*)dy, out);
  }
}

// operand dtype x output dtype: (f32,f32) (bf16,bf16) (bf16,f32 -- tensors that feed a batch norm stay fp32)
template <int MODE>
int launch_io(const ConvP& p, int dtype, int out_dtype, const void* a, const float* w, void* out, cudaStream_t st) {
  if (dtype == RCGAN_F32 && out_dtype == RCGAN_F32) {
    if (!launch_narrow<MODE, float, float>(p, a, w, out, st) && !(MODE == MODE_FPROP && launch_skinny<float, float>(p, a, w, out, st)))
      launch<MODE, float, float>(p, a, w, nullptr, out, 1, st);
  } else if (dtype == RCGAN_BF16 && out_dtype == RCGAN_BF16) {
    if (!launch_narrow<MODE, bf16, bf16>(p, a, w, out, st) && !(MODE == MODE_FPROP && launch_skinny<bf16, bf16>(p, a, w, out, st)))
      launch<MODE, bf16, bf16>(p, a, w, nullptr, out, 1, st);
  } else if (dtype == RCGAN_BF16 && out_dtype == RCGAN_F32) {
    if (!launch_narrow<MODE, bf16, float>(p, a, w, out, st) && !(MODE == MODE_FPROP && launch_skinny<bf16, float>(p, a, w, out, st)))
      launch<MODE, bf16, float>(p, a, w, nullptr, out, 1, st);
  }
  else { rcgan_set_error("conv: unsupported dtype pair (%d -> %d)", dtype, out_dtype); return RCGAN_EUNSUPPORTED; }
  return 0;
}

}  // namespace

// Implemented in conv_tc.cu (tcgen05 path); return 1 when they handled the call.
int rcgan_tc_fprop(const rcgan_conv_desc* d, const void* x, const void* wpack, const float* bias, const void* res, void* y,
                   int out_dtype, int act, float leak, cudaStream_t st, int* handled, const rcgan_conv_epilogue* ep);
int rcgan_tc_dgrad(const rcgan_conv_desc* d, const void* dy, const void* wpack, const float* bias, void* dx, int out_dtype,
                   int act, float leak, int accumulate, cudaStream_t st, int* handled, const rcgan_conv_epilogue* ep);

int rcgan_tc_wgrad(const rcgan_conv_desc* d, const void* x, const void* dy, float* dw, int accumulate, cudaStream_t st,
                   int* handled);

static int conv2d_fprop_impl(const rcgan_conv_desc* d, const void* x, const float* w, const void* wpack, const float* bias,
                            const void* res, void* y, int out_dtype, int act, float leak, void* stream) {
  if (int e = check_desc(d, "conv2d_fprop")) return e;
  if (wpack) {
    int handled = 0;
    int e = rcgan_tc_fprop(d, x, wpack, bias, res, y, out_dtype, act, leak, as_stream(stream), &handled, nullptr);
    if (e || handled) return e;
  }
  RCGAN_CHECK_ARG(w, "conv2d_fprop: null fp32 weights");
  ConvP p = make_p(d);
  p.M = d->n * d->ho * d->wo; p.N = d->cout; p.K = d->kh * d->kw * d->cin;
  p.bias = bias; p.act = act; p.leak = leak; p.klen = ((p.K + BK - 1) / BK) * BK;
  if (int e = launch_io<MODE_FPROP>(p, d->dtype, out_dtype, x, w, y, as_stream(stream))) return e;
  RCGAN_LAUNCH_CHECK("conv2d_fprop");
  rcgan_set_conv_variant("conv_simt");
  if (res) {
    // CUDA-core path (fp32 parity mode, odd shapes): the residual is a separate in-place add over the dense output
    RCGAN_CHECK_ARG(d->ldy == d->cout, "conv2d_fprop_res: the CUDA-core path needs a dense output (ldy == cout)");
    return rcgan_add(y, res, y, (long)d->n * d->ho * d->wo * d->cout, out_dtype, stream);
  }
  return 0;
}

extern "C" int rcgan_conv2d_fprop(const rcgan_conv_desc* d, const void* x, const float* w, const void* wpack,
                                  const float* bias, void* y, int out_dtype, int act, float leak, void* stream) {
  return conv2d_fprop_impl(d, x, w, wpack, bias, nullptr, y, out_dtype, act, leak, stream);
}

extern "C" int rcgan_conv2d_fprop_res(const rcgan_conv_desc* d, const void* x, const float* w, const void* wpack,
                                      const float* bias, const void* res, void* y, int out_dtype, int act, float leak,
                                      void* stream) {
  RCGAN_CHECK_ARG(res, "conv2d_fprop_res: null residual");
  return conv2d_fprop_impl(d, x, w, wpack, bias, res, y, out_dtype, act, leak, stream);
}

extern "C" int rcgan_conv2d_dgrad(const rcgan_conv_desc* d, const void* dy, const float* w, const void* wpack,
                                  const float* bias, void* dx, int out_dtype, int act, float leak, int accumulate,
                                  void* stream) {
  if (int e = check_desc(d, "conv2d_dgrad")) return e;
  if (wpack) {
    int handled = 0;
    int e = rcgan_tc_dgrad(d, dy, wpack, bias, dx, out_dtype, act, leak, accumulate, as_stream(stream), &handled, nullptr);
    if (e || handled) return e;
  }
  RCGAN_CHECK_ARG(w, "conv2d_dgrad: null fp32 weights");
  ConvP p = make_p(d);
  p.M = d->n * d->h * d->w; p.N = d->cin; p.K = d->kh * d->kw * d->cout;
  p.bias = bias; p.act = act; p.leak = leak; p.accumulate = accumulate; p.klen = ((p.K + BK - 1) / BK) * BK;
  if (int e = launch_io<MODE_DGRAD>(p, d->dtype, out_dtype, dy, w, dx, as_stream(stream))) return e;
  RCGAN_LAUNCH_CHECK("conv2d_dgrad");
  rcgan_set_conv_variant("conv_simt");
  return 0;
}

// tensor-core only: the fused epilogues exist in the tcgen05 kernels (rcgan_conv_epilogue_supported tells the caller beforehand)
extern "C" int rcgan_conv2d_fprop_ex(const rcgan_conv_desc* d, const void* x, const void* wpack, const float* bias, void* y,
                                     int out_dtype, int act, float leak, const rcgan_conv_epilogue* ep, void* stream) {
  if (int e = check_desc(d, "conv2d_fprop_ex")) return e;
  RCGAN_CHECK_ARG(wpack && x && y, "conv2d_fprop_ex: null argument");
  int handled = 0;
  int e = rcgan_tc_fprop(d, x, wpack, bias, nullptr, y, out_dtype, act, leak, as_stream(stream), &handled, ep);
  if (e) return e;
  if (!handled) { rcgan_set_error("conv2d_fprop_ex: shape has no tensor-core path"); return RCGAN_EUNSUPPORTED; }
  return 0;
}
extern "C" int rcgan_conv2d_dgrad_ex(const rcgan_conv_desc* d, const void* dy, const void* wpack, const float* bias, void* dx,
                                     int out_dtype, int act, float leak, int accumulate, const rcgan_conv_epilogue* ep,
                                     void* stream) {
  if (int e = check_desc(d, "conv2d_dgrad_ex")) return e;
  RCGAN_CHECK_ARG(wpack && dy && dx, "conv2d_dgrad_ex: null argument");
  int handled = 0;
  int e = rcgan_tc_dgrad(d, dy, wpack, bias, dx, out_dtype, act, leak, accumulate, as_stream(stream), &handled, ep);
  if (e) return e;
  if (!handled) { rcgan_set_error("conv2d_dgrad_ex: shape has no tensor-core path"); return RCGAN_EUNSUPPORTED; }
  return 0;
}

extern "C" size_t rcgan_conv2d_wgrad_workspace(const rcgan_conv_desc* d) {
  if (!d) return 0;
  ConvP p = make_p(d);
  p.M = d->kh * d->kw * d->cin; p.N = d->cout; p.K = d->n * d->ho * d->wo;
  return (size_t)wgrad_splits(p) * p.M * p.N * sizeof(float);
}

extern "C" int rcgan_conv2d_wgrad(const rcgan_conv_desc* d, const void* x, const void* dy, float* dw, int accumulate,
                                  void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_desc(d, "conv2d_wgrad")) return e;
  {
    int handled = 0;
    int e = rcgan_tc_wgrad(d, x, dy, dw, accumulate, as_stream(stream), &handled);
    if (e || handled) return e;
  }
  ConvP p = make_p(d);
  p.M = d->kh * d->kw * d->cin; p.N = d->cout; p.K = d->n * d->ho * d->wo;
  if (d->dtype == RCGAN_F32 ? launch_wgrad_narrow<float>(p, x, dy, dw, accumulate, as_stream(stream))
                            : launch_wgrad_narrow<bf16>(p, x, dy, dw, accumulate, as_stream(stream))) {
    RCGAN_LAUNCH_CHECK("conv2d_wgrad_narrow");
    rcgan_set_conv_variant("conv_wgrad_narrow");
    return 0;
  }
  int ns = wgrad_splits(p);
  RCGAN_CHECK_ARG(ws && ws_bytes >= (size_t)ns * p.M * p.N * sizeof(float), "conv2d_wgrad: workspace too small");
  int klen = (p.K + ns - 1) / ns;
  klen = ((klen + BK - 1) / BK) * BK;
  ns = (p.K + klen - 1) / klen;
  p.klen = klen;
  if (d->dtype == RCGAN_F32) launch<MODE_WGRAD, float, float>(p, x, nullptr, dy, ws, ns, as_stream(stream));
  else launch<MODE_WGRAD, bf16, float>(p, x, nullptr, dy, ws, ns, as_stream(stream));
  RCGAN_LAUNCH_CHECK("conv2d_wgrad");
  long mn = (long)p.M * p.N;
  launch_pdl(splitk_reduce_kernel, ceil_div(mn, 256), 256, 0, as_stream(stream), (const float*)ws, dw, mn, ns, accumulate);
  RCGAN_LAUNCH_CHECK("conv2d_wgrad_reduce");
  rcgan_set_conv_variant("conv_simt");
  return 0;
}
