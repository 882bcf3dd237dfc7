// Batch norm (tf.contrib.layers.batch_norm, mnist/ops.py:30-44) and conditional batch norm
// (cond_batchnorm, cifar10/common/ops/normalization.py:27-59) with the following activation fused.
// x is [rows = samples*hw, c] channels-last.  HBM-bound; ideal traffic fwd = 2 reads + 1 write of x
// (stats pass + apply pass; the stats pass of a tensor that fits the 126 MB L2 is served from L2),
// bwd = reads of dy, x, y for the reduction + dy, x, y again for dx.
//
// Types: x is TX, y/dy/dx are TY.  (f32,f32) is parity mode, (bf16,bf16) plain bf16, and (f32,bf16)
// the bf16 training layout: the tensor feeding a norm is kept in fp32 because the norm's backward
// dx = istd*(scale*g - mean(scale*g) - xhat*mean(scale*g*xhat)) cancels catastrophically and amplifies
// the bf16 rounding of x ~50x (measured on the reference's discriminator at initialisation).
//
//   stats   : per row-chunk shifted sums (pivot = first row of the chunk) -> (mean, M2) partials,
//             merged with Chan's formula in the finalize kernel (no E[x^2]-E[x]^2 cancellation)
//   apply   : y = act((x-mean)*invstd*scale[label] + offset[label]), 4-element vectors
//   bwd     : g = dy*act'(y); partial sums of g and g*xhat per chunk; finalize folds them into
//             dscale/doffset[label] and the two per-channel projections; dx elementwise.
#include "common.cuh"

namespace {

template <typename T, int V> __device__ __forceinline__ void load_vec(const T* p, float* out) {
  if (V == 1) { out[0] = to_f(*p); return; }
  if (sizeof(T) == 4) {
    float4 v = *reinterpret_cast<const float4*>(p);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
  } else if (V == 8) {
    uint4 v = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; i++) { float2 f = __bfloat1622float2(h[i]); out[2 * i] = f.x; out[2 * i + 1] = f.y; }
  } else {
    uint2 v = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
    float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
  }
}
template <typename T, int V> __device__ __forceinline__ void store_vec(T* p, const float* in) {
  if (V == 1) { *p = from_f<T>(in[0]); return; }
  if (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(in[0], in[1], in[2], in[3]);
  } else if (V == 8) {
    uint4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = v;
  } else {
    uint2 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
    h[0] = __floats2bfloat162_rn(in[0], in[1]);
    h[1] = __floats2bfloat162_rn(in[2], in[3]);
    *reinterpret_cast<uint2*>(p) = v;
  }
}

struct Geo {
  int rows, c, hw;
  int ldy;         // row stride of y / dy (the normalised side): c, or the padded width of a fused label concat
  int V;           // channels per thread-vector (4 or 1)
  int cg;          // channel vectors = c / V
  int LC, LR;      // thread layout: LC channel lanes x LR row lanes = 256
  int gx;          // ceil(cg / LC)
  int chunk_rows, nchunk;
};

Geo make_geo(int samples, int hw, int c, bool per_sample, int vmax = 4) {
  Geo g;
  g.rows = samples * hw; g.c = c; g.hw = hw; g.ldy = c;
  g.V = (vmax == 8 && c % 8 == 0) ? 8 : ((c % 4 == 0) ? 4 : 1);   // 16-byte vectors when both x and y are bf16
  g.cg = c / g.V;
  int lc = 1;
  while (lc < 32 && lc * 2 <= g.cg) lc *= 2;
  g.LC = lc; g.LR = 256 / lc;
  g.gx = (g.cg + lc - 1) / lc;
  // blocks per SM wanted: 2, or 4 for large unconditional norms (measured, tools/bn_bench.py: [1024*196, 128] bf16
  // fwd 72 -> 53 us, bwd 117 -> 96 us; the conditional per-sample chunking and small tensors are better off with 2)
  int waves = (!per_sample && (long)g.rows * c * 2 >= (32L << 20)) ? 4 : 2;
  { const char* e = getenv("RCGAN_BN_WAVES"); if (e && atoi(e) > 0) waves = atoi(e); }
  long want = ((long)waves * RCGAN_NUM_SMS + g.gx - 1) / g.gx;  // chunks wanted
  if (per_sample) {
    // chunks never straddle a sample: chunk_rows = hw / 2^j
    int cr = hw;
    while ((long)samples * (hw / cr) < want && cr % 2 == 0 && cr / 2 >= 16) cr /= 2;
    g.chunk_rows = cr;
  } else {
    long cr = (g.rows + want - 1) / want;
    if (cr < 32) cr = 32;
    if (cr > g.rows) cr = g.rows;
    g.chunk_rows = (int)cr;
  }
  g.nchunk = (g.rows + g.chunk_rows - 1) / g.chunk_rows;
  return g;
}

// ws layout (floats): [0, nchunk*c) = P1 ; [nchunk*c, 2*nchunk*c) = P2 ; then 2*c floats (A, B) for bwd
template <typename TX, int V>
__global__ void __launch_bounds__(256) bn_stats_partial_kernel(const TX* __restrict__ x, Geo g, float* __restrict__ ws) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  extern __shared__ float sh[];  // [LR][LC*V] x 2
  const int lc = threadIdx.x % g.LC, lr = threadIdx.x / g.LC;
  const int cv = blockIdx.x * g.LC + lc;
  const int chunk = blockIdx.y;
  const int r0 = chunk * g.chunk_rows, r1 = min(g.rows, r0 + g.chunk_rows);
  float piv[V], s[V], ss[V];
#pragma unroll
  for (int i = 0; i < V; i++) { s[i] = 0.f; ss[i] = 0.f; piv[i] = 0.f; }
  if (cv < g.cg) {
    const TX* base = x + (size_t)cv * V;
    load_vec<TX, V>(base + (size_t)r0 * g.c, piv);
#pragma unroll 4
    for (int r = r0 + lr; r < r1; r += g.LR) {
      float v[V];
      load_vec<TX, V>(base + (size_t)r * g.c, v);
#pragma unroll
      for (int i = 0; i < V; i++) { float d = v[i] - piv[i]; s[i] += d; ss[i] = fmaf(d, d, ss[i]); }
    }
  }
  float* S = sh;
  float* SS = sh + 256 * V;
#pragma unroll
  for (int i = 0; i < V; i++) { S[(lr * g.LC + lc) * V + i] = s[i]; SS[(lr * g.LC + lc) * V + i] = ss[i]; }
  __syncthreads();
  if (lr == 0 && cv < g.cg) {
    float n = (float)(r1 - r0);
#pragma unroll
    for (int i = 0; i < V; i++) {
      float a = 0.f, b = 0.f;
      for (int k = 0; k < g.LR; k++) { a += S[(k * g.LC + lc) * V + i]; b += SS[(k * g.LC + lc) * V + i]; }
      int ch = cv * V + i;
      ws[(size_t)chunk * g.c + ch] = piv[i] + a / n;                          // chunk mean
      ws[(size_t)(g.nchunk + chunk) * g.c + ch] = fmaxf(b - a * a / n, 0.f);  // chunk M2
    }
  }
}

// one warp per channel: lanes stride over the chunks, Chan-merge their partials, then merge across lanes by shuffle
__global__ void __launch_bounds__(256) bn_stats_finalize_kernel(Geo g, const float* __restrict__ ws, float eps, float decay,
                                                               float* mm, float* mv, float* __restrict__ save) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  const int ch = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (ch >= g.c) return;
  float n = 0.f, mean = 0.f, m2 = 0.f;
  for (int k = lane; k < g.nchunk; k += 32) {
    int r0 = k * g.chunk_rows;
    float nb = (float)(min(g.rows, r0 + g.chunk_rows) - r0);
    float mb = ws[(size_t)k * g.c + ch], m2b = ws[(size_t)(g.nchunk + k) * g.c + ch];
    float nt = n + nb, delta = mb - mean;
    mean += delta * (nb / nt);
    m2 += m2b + delta * delta * (n * nb / nt);
    n = nt;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float nb = __shfl_xor_sync(0xffffffffu, n, o), mb = __shfl_xor_sync(0xffffffffu, mean, o),
          m2b = __shfl_xor_sync(0xffffffffu, m2, o);
    float nt = n + nb;
    if (nt > 0.f) {
      float delta = mb - mean;
      mean += delta * (nb / nt);
      m2 += m2b + delta * delta * (n * nb / nt);
    }
    n = nt;
  }
  if (lane == 0) {
    float var = m2 / n;
    save[ch] = mean;
    save[g.c + ch] = rsqrtf(var + eps);
    if (mm) {
      float unbiased = var * (n / fmaxf(n - 1.f, 1.f));
      mm[ch] = decay * mm[ch] + (1.f - decay) * mean;
      mv[ch] = decay * mv[ch] + (1.f - decay) * unbiased;
    }
  }
}

// statistics from the per-CTA partial sums a producing conv accumulated in its epilogue (rcgan_conv_epilogue.colstats:
// [RCGAN_NUM_SMS][2][c] sums / sums of squares, then [RCGAN_NUM_SMS] row counts): one warp per channel, every partial is turned
// into (n, mean, M2) in double and Chan-merged -- the subtraction sumsq - sum^2/n only ever spans one CTA's rows
__global__ void __launch_bounds__(256) bn_stats_from_partials_kernel(int c, const float* __restrict__ parts, float eps, float decay,
                                                                    float* mm, float* mv, float* __restrict__ save) {
  pdl_sync();
  const int ch = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (ch >= c) return;
  const float* counts = parts + (size_t)RCGAN_NUM_SMS * 2 * c;
  double n = 0.0, mean = 0.0, m2 = 0.0;
  for (int k = lane; k < RCGAN_NUM_SMS; k += 32) {
    const double nb = (double)counts[k];
    if (nb <= 0.0) continue;
    const double sb = (double)parts[(size_t)k * 2 * c + ch], qb = (double)parts[(size_t)k * 2 * c + c + ch];
    const double mb = sb / nb, m2b = fmax(qb - sb * mb, 0.0);
    const double nt = n + nb, delta = mb - mean;
    mean += delta * (nb / nt);
    m2 += m2b + delta * delta * (n * nb / nt);
    n = nt;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double nb = __shfl_xor_sync(0xffffffffu, n, o), mb = __shfl_xor_sync(0xffffffffu, mean, o),
                 m2b = __shfl_xor_sync(0xffffffffu, m2, o);
    const double nt = n + nb;
    if (nt > 0.0) {
      const double delta = mb - mean;
      mean += delta * (nb / nt);
      m2 += m2b + delta * delta * (n * nb / nt);
    }
    n = nt;
  }
  if (lane == 0) {
    const float var = (float)(m2 / n), fm = (float)mean, fn = (float)n;
    save[ch] = fm;
    save[c + ch] = rsqrtf(var + eps);
    if (mm) {
      const float unbiased = var * (fn / fmaxf(fn - 1.f, 1.f));
      mm[ch] = decay * mm[ch] + (1.f - decay) * fm;
      mv[ch] = decay * mv[ch] + (1.f - decay) * unbiased;
    }
  }
}

__global__ void bn_infer_stats_kernel(int c, const float* __restrict__ mm, const float* __restrict__ mv, float eps,
                                      float* __restrict__ save) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  save[ch] = mm[ch];
  save[c + ch] = rsqrtf(mv[ch] + eps);
}

// apply / dx kernels: a block owns one row chunk (inside one sample when the norm is conditional, so the label is a block
// constant) and a thread one channel vector: the per-(label, channel) coefficients are folded once into registers and
// the row loop is pure streaming -- y = act(a*x + b).  (The first version re-derived them per element from four scalar
// table loads and ran at 1.8 TB/s.)
template <typename TX, typename TY, int V>
__global__ void __launch_bounds__(256) bn_apply_kernel(const TX* __restrict__ x, TY* __restrict__ y, Geo g,
                                                       const float* __restrict__ scale, const float* __restrict__ offset,
                                                       const int* __restrict__ labels, const float* __restrict__ save,
                                                       int act, float leak, const float* __restrict__ yb, int c2) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  const int lc = threadIdx.x % g.LC, lr = threadIdx.x / g.LC;
  const int cv = blockIdx.x * g.LC + lc;
  const int r0 = blockIdx.y * g.chunk_rows, r1 = min(g.rows, r0 + g.chunk_rows);
  if (yb && blockIdx.x == 0) {
    // fused conv_cond_concat (mnist/ops.py:46-51): channels [c, c + c2) of every output row carry the sample's label vector
    // (the padding up to ldy stays zero from allocation)
    for (int i = threadIdx.x; i < (r1 - r0) * c2; i += 256) {
      const int r = r0 + i / c2, j = i - (i / c2) * c2;
      y[(size_t)r * g.ldy + g.c + j] = from_f<TY>(yb[(size_t)(r / g.hw) * c2 + j]);
    }
  }
  if (cv >= g.cg) return;
  const int ch = cv * V;
  const size_t tab = (size_t)(labels ? labels[r0 / g.hw] : 0) * g.c + ch;
  float a[V], b[V];
#pragma unroll
  for (int k = 0; k < V; k++) {
    a[k] = save[g.c + ch + k] * scale[tab + k];
    b[k] = fmaf(-save[ch + k], a[k], offset[tab + k]);
  }
  const TX* xp = x + ch;
  TY* yp = y + ch;
  if (act == RCGAN_ACT_RELU) {
#pragma unroll 4
    for (int r = r0 + lr; r < r1; r += g.LR) {
      float v[V];
      load_vec<TX, V>(xp + (size_t)r * g.c, v);
#pragma unroll
      for (int k = 0; k < V; k++) v[k] = fmaxf(fmaf(v[k], a[k], b[k]), 0.f);
      store_vec<TY, V>(yp + (size_t)r * g.ldy, v);
    }
  } else {
#pragma unroll 4
    for (int r = r0 + lr; r < r1; r += g.LR) {
      float v[V];
      load_vec<TX, V>(xp + (size_t)r * g.c, v);
#pragma unroll
      for (int k = 0; k < V; k++) v[k] = act_fwd(fmaf(v[k], a[k], b[k]), act, leak);
      store_vec<TY, V>(yp + (size_t)r * g.ldy, v);
    }
  }
}

template <typename TX, typename TY, int V, bool FROM_X>
__global__ void __launch_bounds__(256, 4) bn_bwd_partial_kernel(const TY* __restrict__ dy, const TX* __restrict__ x,
                                                             const TY* __restrict__ y, Geo g, const float* __restrict__ save,
                                                             int act, float leak, float* __restrict__ ws,
                                                             const float* __restrict__ scale, const float* __restrict__ offset,
                                                             const int* __restrict__ labels) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  extern __shared__ float sh[];
  const int lc = threadIdx.x % g.LC, lr = threadIdx.x / g.LC;
  const int cv = blockIdx.x * g.LC + lc;
  const int chunk = blockIdx.y;
  const int r0 = chunk * g.chunk_rows, r1 = min(g.rows, r0 + g.chunk_rows);
  float s1[V], s2[V], mean[V], istd[V], ca[V], cb[V];
#pragma unroll
  for (int i = 0; i < V; i++) { s1[i] = 0.f; s2[i] = 0.f; mean[i] = 0.f; istd[i] = 0.f; ca[i] = 0.f; cb[i] = 0.f; }
  // relu / lrelu: the sign of the forward pre-activation a*x + b (same expression as bn_apply_kernel) IS the mask y > 0, so
  // the output tensor is not read back (2 of the 7 tensor passes of the backward)
  constexpr bool from_x = FROM_X;       // compile-time: the two forms keep separate (lean) register allocations
  if (cv < g.cg) {
#pragma unroll
    for (int i = 0; i < V; i++) { mean[i] = save[cv * V + i]; istd[i] = save[g.c + cv * V + i]; }
    if (from_x) {
      const size_t tab = (size_t)(labels ? labels[r0 / g.hw] : 0) * g.c + (size_t)cv * V;
#pragma unroll
      for (int i = 0; i < V; i++) { ca[i] = istd[i] * scale[tab + i]; cb[i] = fmaf(-mean[i], ca[i], offset[tab + i]); }
    }
#pragma unroll 4
    for (int r = r0 + lr; r < r1; r += g.LR) {
      float vd[V], vx[V], vy[V];
      const size_t o = (size_t)r * g.c + (size_t)cv * V, oy = (size_t)r * g.ldy + (size_t)cv * V;
      load_vec<TY, V>(dy + oy, vd);
      load_vec<TX, V>(x + o, vx);
      if (from_x) {
#pragma unroll
        for (int i = 0; i < V; i++) vy[i] = fmaf(vx[i], ca[i], cb[i]);
      } else if (act != RCGAN_ACT_NONE) load_vec<TY, V>(y + oy, vy);
#pragma unroll
      for (int i = 0; i < V; i++) {
        float gg = vd[i];
        if (act != RCGAN_ACT_NONE) gg *= act_bwd_from_y(vy[i], act, leak);
        s1[i] += gg;
        s2[i] = fmaf(gg, (vx[i] - mean[i]) * istd[i], s2[i]);
      }
    }
  }
  float* S = sh;
  float* SS = sh + 256 * V;
#pragma unroll
  for (int i = 0; i < V; i++) { S[(lr * g.LC + lc) * V + i] = s1[i]; SS[(lr * g.LC + lc) * V + i] = s2[i]; }
  __syncthreads();
  if (lr == 0 && cv < g.cg) {
#pragma unroll
    for (int i = 0; i < V; i++) {
      float a = 0.f, b = 0.f;
      for (int k = 0; k < g.LR; k++) { a += S[(k * g.LC + lc) * V + i]; b += SS[(k * g.LC + lc) * V + i]; }
      int ch = cv * V + i;
      ws[(size_t)chunk * g.c + ch] = a;
      ws[(size_t)(g.nchunk + chunk) * g.c + ch] = b;
    }
  }
}

// one warp per channel.  Unconditional norm (labels == NULL): lanes stride over the chunks and shuffle-reduce.
// Conditional norm: the per-label sums need a scatter, so lane l owns label l (n_labels <= 32) and scans the chunks.
__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(Geo g, float* __restrict__ ws, const float* __restrict__ scale,
                                                             const int* __restrict__ labels, int n_labels,
                                                             float* __restrict__ dscale, float* __restrict__ doffset,
                                                             int accumulate_param) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  const int ch = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (ch >= g.c) return;
  float A = 0.f, B = 0.f;
  if (!labels) {
    float s1 = 0.f, s2 = 0.f;
    for (int k = lane; k < g.nchunk; k += 32) {
      s1 += ws[(size_t)k * g.c + ch];
      s2 += ws[(size_t)(g.nchunk + k) * g.c + ch];
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    const float sc = scale[ch];
    A = sc * s1; B = sc * s2;
    if (lane == 0) {
      doffset[ch] = accumulate_param ? doffset[ch] + s1 : s1;
      dscale[ch] = accumulate_param ? dscale[ch] + s2 : s2;
    }
  } else {
    // lanes stride over the chunks (a chunk never straddles two samples) and keep per-label partial sums in registers
    // through a compile-time select chain, LG labels per pass; a shuffle tree then reduces each label's pair.
    // (the first version gave lane l label l and let it scan every chunk serially: 354 us per call in the G step)
    constexpr int LG = 16;
    for (int l0 = 0; l0 < n_labels; l0 += LG) {
      float a1[LG], a2[LG];
#pragma unroll
      for (int j = 0; j < LG; j++) { a1[j] = 0.f; a2[j] = 0.f; }
      for (int k = lane; k < g.nchunk; k += 32) {
        const int lab = labels[((long)k * g.chunk_rows) / g.hw] - l0;
        const float v1 = ws[(size_t)k * g.c + ch], v2 = ws[(size_t)(g.nchunk + k) * g.c + ch];
#pragma unroll
        for (int j = 0; j < LG; j++) {
          a1[j] += (lab == j) ? v1 : 0.f;
          a2[j] += (lab == j) ? v2 : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < LG; j++) {
        const float s1 = warp_sum(a1[j]), s2 = warp_sum(a2[j]);
        const int l = l0 + j;
        if (l < n_labels && lane == 0) {
          const size_t o = (size_t)l * g.c + ch;
          const float sc = scale[o];
          A = fmaf(sc, s1, A); B = fmaf(sc, s2, B);
          doffset[o] = accumulate_param ? doffset[o] + s1 : s1;
          dscale[o] = accumulate_param ? dscale[o] + s2 : s2;
        }
      }
    }
  }
  if (lane == 0) {
    float inv = 1.f / (float)g.rows;
    float* AB = ws + (size_t)2 * g.nchunk * g.c;
    AB[ch] = A * inv;
    AB[g.c + ch] = B * inv;
  }
}

// dx = istd*(scale[label]*g - A - xhat*B) = c1*g + c2*(x - mean) + c3 with the coefficients in registers (see bn_apply_kernel)
// (4 blocks per SM: at 78 registers only 3 fit, and the 512 per-sample chunks of the largest generator norm then need a second,
// 15 % full wave -- ncu: 224 us = 3.5 TB/s)
template <typename TX, typename TY, int V, bool FROM_X>
__global__ void __launch_bounds__(256, 4) bn_bwd_dx_kernel(const TY* __restrict__ dy, const TX* __restrict__ x,
                                                        const TY* __restrict__ y, TY* __restrict__ dx, Geo g,
                                                        const float* __restrict__ scale, const int* __restrict__ labels,
                                                        const float* __restrict__ save, const float* __restrict__ AB,
                                                        int act, float leak, int accumulate, const float* __restrict__ offset) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  const int lc = threadIdx.x % g.LC, lr = threadIdx.x / g.LC;
  const int cv = blockIdx.x * g.LC + lc;
  if (cv >= g.cg) return;
  const int r0 = blockIdx.y * g.chunk_rows, r1 = min(g.rows, r0 + g.chunk_rows);
  const int ch = cv * V;
  const size_t tab = (size_t)(labels ? labels[r0 / g.hw] : 0) * g.c + ch;
  float c1[V], c2[V], c3[V], mean[V], cb[V];
  constexpr bool from_x = FROM_X;   // see bn_bwd_partial_kernel
#pragma unroll
  for (int k = 0; k < V; k++) {
    const float istd = save[g.c + ch + k];
    mean[k] = save[ch + k];
    c1[k] = istd * scale[tab + k];
    cb[k] = from_x ? fmaf(-mean[k], c1[k], offset[tab + k]) : 0.f;
    c2[k] = -istd * istd * AB[g.c + ch + k];
    c3[k] = -istd * AB[ch + k];
  }
#pragma unroll 2
  for (int r = r0 + lr; r < r1; r += g.LR) {
    const size_t off = (size_t)r * g.c + ch, offy = (size_t)r * g.ldy + ch;
    float vd[V], vx[V], vy[V], o[V];
    load_vec<TY, V>(dy + offy, vd);
    load_vec<TX, V>(x + off, vx);
    if (from_x) {
#pragma unroll
      for (int k = 0; k < V; k++) vy[k] = fmaf(vx[k], c1[k], cb[k]);
    } else if (act != RCGAN_ACT_NONE) load_vec<TY, V>(y + offy, vy);
    if (accumulate) load_vec<TY, V>(dx + off, o);
#pragma unroll
    for (int k = 0; k < V; k++) {
      float gg = vd[k];
      if (act == RCGAN_ACT_RELU) gg = vy[k] > 0.f ? gg : 0.f;
      else if (act != RCGAN_ACT_NONE) gg *= act_bwd_from_y(vy[k], act, leak);
      const float v = fmaf(c1[k], gg, fmaf(c2[k], vx[k] - mean[k], c3[k]));
      o[k] = accumulate ? o[k] + v : v;
    }
    store_vec<TY, V>(dx + off, o);
  }
}

inline int ew_grid(long work) {
  long gsz = (work + 255) / 256;
  long cap = (long)RCGAN_NUM_SMS * 16;
  return (int)(gsz < 1 ? 1 : (gsz > cap ? cap : gsz));
}

inline int check_types(int xd, int yd, const char* who) {
  bool ok = (xd == RCGAN_F32 && yd == RCGAN_F32) || (xd == RCGAN_BF16 && yd == RCGAN_BF16) ||
            (xd == RCGAN_F32 && yd == RCGAN_BF16);
  if (!ok) { rcgan_set_error("%s: unsupported dtype pair x=%d y=%d", who, xd, yd); return RCGAN_EUNSUPPORTED; }
  return 0;
}

// experiment switch: RCGAN_BN_STATS_V=8 -> 16-byte bf16 vectors in the reduction kernels too
int stats_vmax() {
  const char* e = getenv("RCGAN_BN_STATS_V");
  return (e && atoi(e) == 8) ? 8 : 4;
}

}  // namespace

extern "C" size_t rcgan_bn_workspace(int samples, int hw, int c) {
  int nc = 0;
  for (int vmax = 4; vmax <= 8; vmax += 4)
    for (int ps = 0; ps < 2; ps++) {
      Geo g = make_geo(samples, hw, c, ps != 0, vmax);
      if (g.nchunk > nc) nc = g.nchunk;
    }
  return ((size_t)2 * nc * c + 2 * c) * sizeof(float);
}

// expands `...` with TX, TY, VV bound for the (xdtype, ydtype, V) triple
#define BN_DISPATCH(xd, yd, V, ...)                                                                      \
  if ((xd) == RCGAN_F32 && (yd) == RCGAN_F32) {                                                          \
    typedef float TX; typedef float TY;                                                                  \
    if ((V) == 4) { constexpr int VV = 4; __VA_ARGS__; } else { constexpr int VV = 1; __VA_ARGS__; }     \
  } else if ((xd) == RCGAN_BF16) {                                                                       \
    typedef bf16 TX; typedef bf16 TY;                                                                    \
    if ((V) == 8) { constexpr int VV = 8; __VA_ARGS__; }                                                 \
    else if ((V) == 4) { constexpr int VV = 4; __VA_ARGS__; } else { constexpr int VV = 1; __VA_ARGS__; } \
  } else {                                                                                               \
    typedef float TX; typedef bf16 TY;                                                                   \
    if ((V) == 4) { constexpr int VV = 4; __VA_ARGS__; } else { constexpr int VV = 1; __VA_ARGS__; }     \
  }

static int bn_fwd_impl(const void* x, void* y, int samples, int hw, int c, int xdtype, int ydtype, const float* scale,
                       const float* offset, const int* labels, float eps, int act, float leak, int train, float decay,
                       float* moving_mean, float* moving_var, float* save, void* ws, size_t ws_bytes, void* stream, int ldy,
                       const float* yb, int c2, const float* colstats = nullptr) {
  RCGAN_CHECK_ARG(samples > 0 && hw > 0 && c > 0, "bn_fwd: bad shape");
  RCGAN_CHECK_ARG(ldy >= c + c2 && (ldy == c || ldy % 8 == 0) && c2 >= 0 && (c2 == 0 || yb), "bn_fwd: bad concat geometry");
  if (int e = check_types(xdtype, ydtype, "bn_fwd")) return e;
  RCGAN_CHECK_ARG(x && y && scale && offset && save, "bn_fwd: null pointer");
  RCGAN_CHECK_ARG((long)samples * hw * c < 2147483647L, "bn_fwd: too large");
  cudaStream_t st = as_stream(stream);
  Geo g = make_geo(samples, hw, c, labels != nullptr, xdtype == RCGAN_BF16 ? stats_vmax() : 4);
  if (train && colstats) {
    launch_pdl(bn_stats_from_partials_kernel, ceil_div(c, 8), 256, 0, st, c, colstats, eps, decay, moving_mean, moving_var, save);
    RCGAN_LAUNCH_CHECK("bn_stats_from_partials");
  } else if (train) {
    RCGAN_CHECK_ARG(ws && ws_bytes >= ((size_t)2 * g.nchunk * c + 2 * c) * sizeof(float), "bn_fwd: workspace too small");
    dim3 grid(g.gx, g.nchunk);
    size_t shb = (size_t)2 * 256 * g.V * sizeof(float);
    BN_DISPATCH(xdtype, ydtype, g.V, launch_pdl(bn_stats_partial_kernel<TX, VV>, grid, 256, shb, st, (const TX*)x, g, (float*)ws));
    RCGAN_LAUNCH_CHECK("bn_stats_partial");
    launch_pdl(bn_stats_finalize_kernel, ceil_div(c, 8), 256, 0, st, g, (const float*)ws, eps, decay, moving_mean, moving_var, save);
    RCGAN_LAUNCH_CHECK("bn_stats_finalize");
  } else {
    RCGAN_CHECK_ARG(moving_mean && moving_var, "bn_fwd: inference needs moving statistics");
    launch_pdl(bn_infer_stats_kernel, ceil_div(c, 128), 128, 0, st, c, moving_mean, moving_var, eps, save);
    RCGAN_LAUNCH_CHECK("bn_infer_stats");
  }
  // streaming geometry: 16-byte vectors when x and y are both bf16
  Geo ga = make_geo(samples, hw, c, labels != nullptr, xdtype == RCGAN_BF16 ? 8 : 4);
  ga.ldy = ldy;
  BN_DISPATCH(xdtype, ydtype, ga.V, launch_pdl(bn_apply_kernel<TX, TY, VV>, dim3(ga.gx, ga.nchunk), 256, 0, st, (const TX*)x, (TY*)y, ga, scale, offset, labels, save, act, leak, c2 > 0 ? yb : (const float*)nullptr, c2));
  RCGAN_LAUNCH_CHECK("bn_apply");
  return 0;
}

extern "C" int rcgan_bn_fwd(const void* x, void* y, int samples, int hw, int c, int xdtype, int ydtype, const float* scale,
                            const float* offset, const int* labels, float eps, int act, float leak, int train,
                            float decay, float* moving_mean, float* moving_var, float* save, void* ws, size_t ws_bytes,
                            void* stream) {
  return bn_fwd_impl(x, y, samples, hw, c, xdtype, ydtype, scale, offset, labels, eps, act, leak, train, decay, moving_mean,
                     moving_var, save, ws, ws_bytes, stream, c, nullptr, 0);
}

extern "C" int rcgan_bn_fwd_prestats(const void* x, void* y, int samples, int hw, int c, int xdtype, int ydtype, const float* scale,
                                     const float* offset, const int* labels, float eps, int act, float leak, float decay,
                                     float* moving_mean, float* moving_var, float* save, const float* colstats, void* stream) {
  RCGAN_CHECK_ARG(colstats, "bn_fwd_prestats: null statistics");
  return bn_fwd_impl(x, y, samples, hw, c, xdtype, ydtype, scale, offset, labels, eps, act, leak, 1, decay, moving_mean,
                     moving_var, save, nullptr, 0, stream, c, nullptr, 0, colstats);
}

extern "C" int rcgan_bn_fwd_cat(const void* x, void* y, int ldy, const float* yb, int c2, int samples, int hw, int c, int xdtype,
                                int ydtype, const float* scale, const float* offset, const int* labels, float eps, int act,
                                float leak, int train, float decay, float* moving_mean, float* moving_var, float* save,
                                void* ws, size_t ws_bytes, void* stream) {
  return bn_fwd_impl(x, y, samples, hw, c, xdtype, ydtype, scale, offset, labels, eps, act, leak, train, decay, moving_mean,
                     moving_var, save, ws, ws_bytes, stream, ldy, yb, c2);
}

static int bn_bwd_impl(const void* dy, const void* x, const void* y, void* dx, int samples, int hw, int c, int xdtype,
                       int ydtype, const float* scale, const int* labels, int n_labels, const float* save, int act,
                       float leak, float* dscale, float* doffset, int accumulate_dx, int accumulate_param, void* ws,
                       size_t ws_bytes, void* stream, int ldy, const float* offset) {
  RCGAN_CHECK_ARG(samples > 0 && hw > 0 && c > 0 && n_labels > 0, "bn_bwd: bad shape");
  RCGAN_CHECK_ARG(ldy >= c && (ldy == c || ldy % 8 == 0), "bn_bwd: bad dy/y row stride");
  if (int e = check_types(xdtype, ydtype, "bn_bwd")) return e;
  RCGAN_CHECK_ARG(dy && x && dx && scale && save && dscale && doffset, "bn_bwd: null pointer");
  RCGAN_CHECK_ARG(act == RCGAN_ACT_NONE || y || (offset && (act == RCGAN_ACT_RELU || act == RCGAN_ACT_LRELU)),
                  "bn_bwd: the activation's backward needs y, or offset for relu / lrelu");
  cudaStream_t st = as_stream(stream);
  Geo g = make_geo(samples, hw, c, labels != nullptr, xdtype == RCGAN_BF16 ? stats_vmax() : 4);
  g.ldy = ldy;
  RCGAN_CHECK_ARG(ws && ws_bytes >= ((size_t)2 * g.nchunk * c + 2 * c) * sizeof(float), "bn_bwd: workspace too small");
  dim3 grid(g.gx, g.nchunk);
  size_t shb = (size_t)2 * 256 * g.V * sizeof(float);
  const bool from_x = offset != nullptr && (act == RCGAN_ACT_RELU || act == RCGAN_ACT_LRELU);
  BN_DISPATCH(xdtype, ydtype, g.V, launch_pdl(from_x ? bn_bwd_partial_kernel<TX, TY, VV, true> : bn_bwd_partial_kernel<TX, TY, VV, false>, grid, 256, shb, st, (const TY*)dy, (const TX*)x, (const TY*)y, g, save, act, leak, (float*)ws, scale, offset, labels));
  RCGAN_LAUNCH_CHECK("bn_bwd_partial");
  launch_pdl(bn_bwd_finalize_kernel, ceil_div(c, 8), 256, 0, st, g, (float*)ws, scale, labels, n_labels, dscale, doffset,
                                                            accumulate_param);
  RCGAN_LAUNCH_CHECK("bn_bwd_finalize");
  const float* AB = (const float*)ws + (size_t)2 * g.nchunk * c;
  Geo ga = make_geo(samples, hw, c, labels != nullptr, xdtype == RCGAN_BF16 ? 8 : 4);
  ga.ldy = ldy;
  BN_DISPATCH(xdtype, ydtype, ga.V, launch_pdl(from_x ? bn_bwd_dx_kernel<TX, TY, VV, true> : bn_bwd_dx_kernel<TX, TY, VV, false>, dim3(ga.gx, ga.nchunk), 256, 0, st, (const TY*)dy, (const TX*)x, (const TY*)y, (TY*)dx, ga, scale, labels, save, AB, act,
                                        leak, accumulate_dx, offset));
  RCGAN_LAUNCH_CHECK("bn_bwd_dx");
  return 0;
}

extern "C" int rcgan_bn_bwd(const void* dy, const void* x, const void* y, void* dx, int samples, int hw, int c, int xdtype,
                            int ydtype, const float* scale, const int* labels, int n_labels, const float* save, int act,
                            float leak, float* dscale, float* doffset, int accumulate_dx, int accumulate_param, void* ws,
                            size_t ws_bytes, const float* offset, void* stream) {
  return bn_bwd_impl(dy, x, y, dx, samples, hw, c, xdtype, ydtype, scale, labels, n_labels, save, act, leak, dscale, doffset,
                     accumulate_dx, accumulate_param, ws, ws_bytes, stream, c, offset);
}

extern "C" int rcgan_bn_bwd_cat(const void* dy, const void* x, const void* y, int ldy, void* dx, int samples, int hw, int c,
                                int xdtype, int ydtype, const float* scale, const int* labels, int n_labels, const float* save,
                                int act, float leak, float* dscale, float* doffset, int accumulate_dx, int accumulate_param,
                                void* ws, size_t ws_bytes, const float* offset, void* stream) {
  return bn_bwd_impl(dy, x, y, dx, samples, hw, c, xdtype, ydtype, scale, labels, n_labels, save, act, leak, dscale, doffset,
                     accumulate_dx, accumulate_param, ws, ws_bytes, stream, ldy, offset);
}

extern "C" int rcgan_bn_infer_bwd(const void* dy, const void* y, int ldy, void* dx, int samples, int hw, int c, int dtype,
                                  const float* scale, const int* labels, const float* save, int act, float leak,
                                  int accumulate_dx, void* ws, size_t ws_bytes, void* stream) {
  RCGAN_CHECK_ARG(samples > 0 && hw > 0 && c > 0 && ldy >= c && (ldy == c || ldy % 8 == 0), "bn_infer_bwd: bad shape");
  RCGAN_CHECK_ARG(dtype == RCGAN_F32 || dtype == RCGAN_BF16, "bn_infer_bwd: bad dtype");
  RCGAN_CHECK_ARG(dy && dx && scale && save && (act == RCGAN_ACT_NONE || y), "bn_infer_bwd: null pointer");
  RCGAN_CHECK_ARG(ws && ws_bytes >= (size_t)2 * c * sizeof(float), "bn_infer_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  // the statistics are constants, so the two batch-projection terms vanish: the training dx kernel with A = B = 0
  // (c2 = c3 = 0; its x operand is only multiplied by 0 -- dy stands in for it)
  cudaError_t e = cudaMemsetAsync(ws, 0, (size_t)2 * c * sizeof(float), st);
  if (e != cudaSuccess) { rcgan_set_error("bn_infer_bwd: memset failed: %s", cudaGetErrorString(e)); return RCGAN_ECUDA; }
  Geo ga = make_geo(samples, hw, c, labels != nullptr, dtype == RCGAN_BF16 ? 8 : 4);
  ga.ldy = ldy;
  BN_DISPATCH(dtype, dtype, ga.V, launch_pdl(bn_bwd_dx_kernel<TX, TY, VV, false>, dim3(ga.gx, ga.nchunk), 256, 0, st, (const TY*)dy,
                                             (const TX*)dy, (const TY*)y, (TY*)dx, ga, scale, labels, save, (const float*)ws, act,
                                             leak, accumulate_dx, (const float*)nullptr));
  RCGAN_LAUNCH_CHECK("bn_infer_bwd");
  return 0;
}
