// Spectral normalisation, forward and the reference's FULL backward (gradient through the
// power iteration): spectral_normed_weight, mnist/sn.py:17-75 == cifar10/common/ops/sn.py.
//
//   a = W u0 (row dots)     na = |a|      v = a/(na+eps)
//   b = v W  (col sums)     n  = |b|      u' = b/(n+eps)      sigma = v W u'^T = n^2/(n+eps)
//   W_bar = W / sigma
// backward, G = dL/dW_bar (verified against autograd to 4e-16):
//   g_sigma = -sum(G.W)/sigma^2 ; g_b = g_sigma (n+2eps)/(n+eps)^2 b ; t = W b (row dots)
//   g_a[i] = g_sigma k ( t_i/(na+eps) - (t.a) a_i / (na (na+eps)^2) ),  k = (n+2eps)/(n+eps)^2
//   dW = G/sigma + v^T g_b + g_a^T u0
//
// HBM-bound on W (fp32): fwd reads W twice (pass 1: a and the partial column sums in ONE sweep;
// pass 2: scale + t), writes W_bar once.  Weights are <= 9.4 MB, i.e. L2-resident between passes.
// save layout (floats): [0]=sigma [1]=1/sigma [2]=na [3]=n [4]=t.a  then a[m], b[c], t[m].
#include "common.cuh"

namespace {

constexpr float SN_EPS = 1e-12f;
constexpr int SN_ROWS_PER_BLOCK = 32;

__host__ __device__ inline int sn_nblk(int m) { return (m + SN_ROWS_PER_BLOCK - 1) / SN_ROWS_PER_BLOCK; }

// pass 1: block handles 32 rows.  warp w computes a_i for rows w, w+8, ...; then the block forms the
// partial (unnormalised) column sums  sum_i a_i W[i,:]  and partial |a|^2.   ws: [nblk][c+1]
__device__ __forceinline__ void sn_pass1_body(const float* __restrict__ W, const float* __restrict__ u, int m, int c,
                                              float* __restrict__ a_out, float* __restrict__ ws, int bx) {
  __shared__ float a_sh[SN_ROWS_PER_BLOCK];
  const int r0 = bx * SN_ROWS_PER_BLOCK;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int rr = wid; rr < SN_ROWS_PER_BLOCK; rr += 8) {
    int r = r0 + rr;
    float acc = 0.f;
    if (r < m)
      for (int j = lane; j < c; j += 32) acc = fmaf(W[(size_t)r * c + j], u[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      a_sh[rr] = acc;
      if (r < m) a_out[r] = acc;
    }
  }
  __syncthreads();
  const int nr = min(SN_ROWS_PER_BLOCK, m - r0);
  for (int j = threadIdx.x; j < c; j += 256) {
    float acc = 0.f;
    for (int rr = 0; rr < nr; rr++) acc = fmaf(a_sh[rr], W[(size_t)(r0 + rr) * c + j], acc);
    ws[(size_t)bx * (c + 1) + j] = acc;
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int rr = 0; rr < nr; rr++) s = fmaf(a_sh[rr], a_sh[rr], s);
    ws[(size_t)bx * (c + 1) + c] = s;
  }
}

// pass 2 (one block): reduce partials -> b, n, u', sigma
__device__ __forceinline__ void sn_finalize_body(int m, int c, int nblk, const float* __restrict__ ws,
                                                 float* __restrict__ u_new, float* __restrict__ save) {
  __shared__ float red[33];
  float na2 = 0.f;
  for (int k = threadIdx.x; k < nblk; k += 256) na2 += ws[(size_t)k * (c + 1) + c];
  na2 = block_sum(na2, red);
  float na = sqrtf(na2);
  float inv_na = 1.f / (na + SN_EPS);
  float* b = save + 5 + m;
  float n2 = 0.f;
  for (int j = threadIdx.x; j < c; j += 256) {
    float s = 0.f;
    for (int k = 0; k < nblk; k++) s += ws[(size_t)k * (c + 1) + j];
    s *= inv_na;
    b[j] = s;
    n2 = fmaf(s, s, n2);
  }
  n2 = block_sum(n2, red);
  float n = sqrtf(n2);
  float inv_n = 1.f / (n + SN_EPS);
  for (int j = threadIdx.x; j < c; j += 256) u_new[j] = b[j] * inv_n;
  if (threadIdx.x == 0) {
    float sigma = n2 * inv_n;  // v W u'^T = b.u' = n^2/(n+eps)
    save[0] = sigma;
    save[1] = 1.f / sigma;
    save[2] = na;
    save[3] = n;
  }
}

// pass 3: W_bar = W/sigma (optional), t_i = <W[i,:], b>, partial t.a -> ws2[nblk]
__device__ __forceinline__ void sn_pass3_body(const float* __restrict__ W, int m, int c, float* __restrict__ w_bar,
                                              float* __restrict__ save, float* __restrict__ ws2, int bx) {
  __shared__ float ta_sh[8];
  const float* a = save + 5;
  const float* b = save + 5 + m;
  float* tt = save + 5 + m + c;
  const float inv_sigma = save[1];
  const int r0 = bx * SN_ROWS_PER_BLOCK;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float ta = 0.f;
  for (int rr = wid; rr < SN_ROWS_PER_BLOCK; rr += 8) {
    int r = r0 + rr;
    if (r >= m) break;
    float acc = 0.f;
    for (int j = lane; j < c; j += 32) {
      float w = W[(size_t)r * c + j];
      acc = fmaf(w, b[j], acc);
      if (w_bar) w_bar[(size_t)r * c + j] = w * inv_sigma;
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      tt[r] = acc;
      ta = fmaf(acc, a[r], ta);
    }
  }
  if (lane == 0) ta_sh[wid] = ta;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < 8; k++) s += ta_sh[k];
    ws2[bx] = s;
  }
}

__device__ __forceinline__ void sn_ta_body(int nblk, const float* __restrict__ ws2, float* __restrict__ save) {
  __shared__ float red[33];
  float s = 0.f;
  for (int k = threadIdx.x; k < nblk; k += blockDim.x) s += ws2[k];
  s = block_sum(s, red);
  if (threadIdx.x == 0) save[4] = s;
}

// backward pass 1: partial sums of G.W per block -> ws[nblk]
__device__ __forceinline__ void sn_bwd_dot_body(const float* __restrict__ W, const float* __restrict__ G, long numel,
                                                float* __restrict__ ws, int bx, int nb) {
  __shared__ float red[33];
  float s = 0.f;
  for (long i = (long)bx * 256 + threadIdx.x; i < numel; i += (long)nb * 256) s = fmaf(G[i], W[i], s);
  s = block_sum(s, red);
  if (threadIdx.x == 0) ws[bx] = s;
}

// backward pass 2: dW elementwise (every block re-reduces the <= 296 partials)
__device__ __forceinline__ void sn_bwd_dw_body(const float* __restrict__ G, const float* __restrict__ u, int m, int c,
                                               const float* __restrict__ save, const float* __restrict__ ws,
                                               int nparts, float* __restrict__ dW, int accumulate, int bx, int nb) {
  __shared__ float red[33];
  float s = 0.f;
  for (int k = threadIdx.x; k < nparts; k += 256) s += ws[k];
  s = block_sum(s, red);
  const float sigma = save[0], inv_sigma = save[1], na = save[2], n = save[3], ta = save[4];
  const float* a = save + 5;
  const float* b = save + 5 + m;
  const float* tt = save + 5 + m + c;
  const float g_sigma = -s * inv_sigma * inv_sigma;
  const float kk = (n + 2.f * SN_EPS) / ((n + SN_EPS) * (n + SN_EPS));
  const float gk = g_sigma * kk;
  const float inv_na = 1.f / (na + SN_EPS);
  const float c2 = ta / (na * (na + SN_EPS) * (na + SN_EPS));
  (void)sigma;
  long numel = (long)m * c;
  for (long i = (long)bx * 256 + threadIdx.x; i < numel; i += (long)nb * 256) {
    int r = (int)(i / c), j = (int)(i - (long)r * c);
    float v_r = a[r] * inv_na;
    float g_a = gk * (tt[r] * inv_na - c2 * a[r]);
    float val = G[i] * inv_sigma + v_r * (gk * b[j]) + g_a * u[j];
    dW[i] = accumulate ? dW[i] + val : val;
  }
}

__global__ void __launch_bounds__(256) sn_pass1_kernel(const float* __restrict__ W, const float* __restrict__ u, int m, int c,
                                                       float* __restrict__ a_out, float* __restrict__ ws) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  sn_pass1_body(W, u, m, c, a_out, ws, blockIdx.x);
}
__global__ void __launch_bounds__(256) sn_finalize_kernel(int m, int c, int nblk, const float* __restrict__ ws,
                                                          float* __restrict__ u_new, float* __restrict__ save) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  sn_finalize_body(m, c, nblk, ws, u_new, save);
}
__global__ void __launch_bounds__(256) sn_pass3_kernel(const float* __restrict__ W, int m, int c, float* __restrict__ w_bar,
                                                       float* __restrict__ save, float* __restrict__ ws2) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  sn_pass3_body(W, m, c, w_bar, save, ws2, blockIdx.x);
}
__global__ void __launch_bounds__(256) sn_ta_kernel(int nblk, const float* __restrict__ ws2, float* __restrict__ save) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  sn_ta_body(nblk, ws2, save);
}
__global__ void __launch_bounds__(256) sn_bwd_dot_kernel(const float* __restrict__ W, const float* __restrict__ G, long numel,
                                                         float* __restrict__ ws) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  sn_bwd_dot_body(W, G, numel, ws, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(256) sn_bwd_dw_kernel(const float* __restrict__ G, const float* __restrict__ u, int m, int c,
                                                        const float* __restrict__ save, const float* __restrict__ ws,
                                                        int nparts, float* __restrict__ dW, int accumulate) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  sn_bwd_dw_body(G, u, m, c, save, ws, nparts, dW, accumulate, blockIdx.x, gridDim.x);
}

// ---- batched variants: every spectrally-normalised weight of a program in ONE launch per pass (blockIdx.y = weight).
// The D step of the CIFAR SN-ResNet normalises 16 weights; as 6 launches per weight that was 96 launches of 3-8 us each.
constexpr int SN_MAX_BATCH = 32;
struct SnItem {
  const float* W; const float* u; float* w_bar; float* u_new; float* save; float* ws; float* ws2;
  const float* G; float* dW;
  int m, c, nblk, accumulate, nparts, g2;
};
struct SnBatch { SnItem it[SN_MAX_BATCH]; };

__global__ void __launch_bounds__(256) sn_pass1_batched(const __grid_constant__ SnBatch b) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  const SnItem& t = b.it[blockIdx.y];
  if ((int)blockIdx.x < t.nblk) sn_pass1_body(t.W, t.u, t.m, t.c, t.save + 5, t.ws, blockIdx.x);
}
__global__ void __launch_bounds__(256) sn_finalize_batched(const __grid_constant__ SnBatch b) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  const SnItem& t = b.it[blockIdx.x];
  sn_finalize_body(t.m, t.c, t.nblk, t.ws, t.u_new, t.save);
}
__global__ void __launch_bounds__(256) sn_pass3_batched(const __grid_constant__ SnBatch b) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  const SnItem& t = b.it[blockIdx.y];
  if ((int)blockIdx.x < t.nblk) sn_pass3_body(t.W, t.m, t.c, t.w_bar, t.save, t.ws2, blockIdx.x);
}
__global__ void __launch_bounds__(256) sn_ta_batched(const __grid_constant__ SnBatch b) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  const SnItem& t = b.it[blockIdx.x];
  sn_ta_body(t.nblk, t.ws2, t.save);
}
__global__ void __launch_bounds__(256) sn_bwd_dot_batched(const __grid_constant__ SnBatch b) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  const SnItem& t = b.it[blockIdx.y];
  if ((int)blockIdx.x < t.nparts) sn_bwd_dot_body(t.W, t.G, (long)t.m * t.c, t.ws, blockIdx.x, t.nparts);
}
__global__ void __launch_bounds__(256) sn_bwd_dw_batched(const __grid_constant__ SnBatch b) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  const SnItem& t = b.it[blockIdx.y];
  if ((int)blockIdx.x < t.g2) sn_bwd_dw_body(t.G, t.u, t.m, t.c, t.save, t.ws, t.nparts, t.dW, t.accumulate, blockIdx.x, t.g2);
}

inline size_t sn_ws_floats(int m, int c) {
  size_t a = (size_t)sn_nblk(m) * (c + 1) + sn_nblk(m);
  size_t b = 2 * RCGAN_NUM_SMS;
  return ((a > b ? a : b) + 3) / 4 * 4;
}

}  // namespace

extern "C" size_t rcgan_sn_save_floats(int m, int c) { return (size_t)5 + 2 * (size_t)m + c; }
extern "C" size_t rcgan_sn_workspace(int m, int c) { return sn_ws_floats(m, c) * sizeof(float); }

extern "C" int rcgan_sn_fwd(const float* W, const float* u, int m, int c, float* w_bar, float* u_new, float* save,
                            void* ws, size_t ws_bytes, void* stream) {
  RCGAN_CHECK_ARG(W && u && u_new && save && m > 0 && c > 0, "sn_fwd: bad args");
  RCGAN_CHECK_ARG(ws && ws_bytes >= rcgan_sn_workspace(m, c), "sn_fwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  int nblk = sn_nblk(m);
  float* wsf = (float*)ws;
  float* ws2 = wsf + (size_t)nblk * (c + 1);
  launch_pdl(sn_pass1_kernel, nblk, 256, 0, st, W, u, m, c, save + 5, wsf);
  RCGAN_LAUNCH_CHECK("sn_pass1");
  launch_pdl(sn_finalize_kernel, 1, 256, 0, st, m, c, nblk, wsf, u_new, save);
  RCGAN_LAUNCH_CHECK("sn_finalize");
  launch_pdl(sn_pass3_kernel, nblk, 256, 0, st, W, m, c, w_bar, save, ws2);
  RCGAN_LAUNCH_CHECK("sn_pass3");
  launch_pdl(sn_ta_kernel, 1, 256, 0, st, nblk, ws2, save);
  RCGAN_LAUNCH_CHECK("sn_ta");
  return 0;
}

extern "C" int rcgan_sn_bwd(const float* W, const float* u, const float* G, int m, int c, const float* save, float* dW,
                            int accumulate, void* ws, size_t ws_bytes, void* stream) {
  RCGAN_CHECK_ARG(W && u && G && save && dW && m > 0 && c > 0, "sn_bwd: bad args");
  RCGAN_CHECK_ARG(ws && ws_bytes >= rcgan_sn_workspace(m, c), "sn_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  long numel = (long)m * c;
  int nparts = (int)((numel + 255) / 256);
  if (nparts > 2 * RCGAN_NUM_SMS) nparts = 2 * RCGAN_NUM_SMS;
  launch_pdl(sn_bwd_dot_kernel, nparts, 256, 0, st, W, G, numel, (float*)ws);
  RCGAN_LAUNCH_CHECK("sn_bwd_dot");
  int g2 = (int)((numel + 255) / 256);
  if (g2 > 4 * RCGAN_NUM_SMS) g2 = 4 * RCGAN_NUM_SMS;
  launch_pdl(sn_bwd_dw_kernel, g2, 256, 0, st, G, u, m, c, save, (const float*)ws, nparts, dW, accumulate);
  RCGAN_LAUNCH_CHECK("sn_bwd_dw");
  return 0;
}

extern "C" size_t rcgan_sn_workspace_batched(int count, const int* m, const int* c) {
  size_t n = 0;
  for (int i = 0; i < count; i++) n += sn_ws_floats(m[i], c[i]);
  return n * sizeof(float);
}

static int sn_fill(SnBatch& b, int i0, int n, const float* const* W, const float* const* u, const int* m, const int* c,
                   float* const* w_bar, float* const* u_new, float* const* save, const float* const* G, float* const* dW,
                   const int* accumulate, float* ws, int* max_nblk, int* max_parts, int* max_g2) {
  *max_nblk = *max_parts = *max_g2 = 1;
  for (int k = 0; k < n; k++) {
    const int i = i0 + k;
    if (!(W[i] && u[i] && save[i] && m[i] > 0 && c[i] > 0)) { rcgan_set_error("sn batched: bad item %d", i); return RCGAN_EBADSHAPE; }
    SnItem& t = b.it[k];
    t.W = W[i]; t.u = u[i]; t.w_bar = w_bar ? w_bar[i] : nullptr; t.u_new = u_new ? u_new[i] : nullptr; t.save = save[i];
    t.G = G ? G[i] : nullptr; t.dW = dW ? dW[i] : nullptr; t.accumulate = accumulate ? accumulate[i] : 0;
    t.m = m[i]; t.c = c[i]; t.nblk = sn_nblk(m[i]);
    t.ws = ws; t.ws2 = ws + (size_t)t.nblk * (c[i] + 1);
    ws += sn_ws_floats(m[i], c[i]);
    const long numel = (long)m[i] * c[i];
    long np = (numel + 255) / 256;
    t.nparts = (int)(np > 2 * RCGAN_NUM_SMS ? 2 * RCGAN_NUM_SMS : np);
    t.g2 = (int)(np > 4 * RCGAN_NUM_SMS ? 4 * RCGAN_NUM_SMS : np);
    if (t.nblk > *max_nblk) *max_nblk = t.nblk;
    if (t.nparts > *max_parts) *max_parts = t.nparts;
    if (t.g2 > *max_g2) *max_g2 = t.g2;
  }
  return 0;
}

extern "C" int rcgan_sn_fwd_batched(int count, const float* const* W, const float* const* u, const int* m, const int* c,
                                    float* const* w_bar, float* const* u_new, float* const* save, void* ws, size_t ws_bytes,
                                    void* stream) {
  RCGAN_CHECK_ARG(count > 0 && W && u && m && c && u_new && save, "sn_fwd_batched: bad args");
  RCGAN_CHECK_ARG(ws && ws_bytes >= rcgan_sn_workspace_batched(count, m, c), "sn_fwd_batched: workspace too small");
  cudaStream_t st = as_stream(stream);
  float* wsf = (float*)ws;
  for (int i0 = 0; i0 < count; i0 += SN_MAX_BATCH) {
    const int n = count - i0 < SN_MAX_BATCH ? count - i0 : SN_MAX_BATCH;
    SnBatch b;
    int mb, mp, mg;
    if (int e = sn_fill(b, i0, n, W, u, m, c, w_bar, u_new, save, nullptr, nullptr, nullptr, wsf, &mb, &mp, &mg)) return e;
    for (int k = 0; k < n; k++) {
      RCGAN_CHECK_ARG(b.it[k].u_new, "sn_fwd_batched: null u_new");
      wsf += sn_ws_floats(m[i0 + k], c[i0 + k]);
    }
    launch_pdl(sn_pass1_batched, dim3(mb, n), 256, 0, st, b);
    RCGAN_LAUNCH_CHECK("sn_pass1_batched");
    launch_pdl(sn_finalize_batched, n, 256, 0, st, b);
    RCGAN_LAUNCH_CHECK("sn_finalize_batched");
    launch_pdl(sn_pass3_batched, dim3(mb, n), 256, 0, st, b);
    RCGAN_LAUNCH_CHECK("sn_pass3_batched");
    launch_pdl(sn_ta_batched, n, 256, 0, st, b);
    RCGAN_LAUNCH_CHECK("sn_ta_batched");
  }
  return 0;
}

extern "C" int rcgan_sn_bwd_batched(int count, const float* const* W, const float* const* u, const float* const* G, const int* m,
                                    const int* c, float* const* save, float* const* dW, const int* accumulate, void* ws,
                                    size_t ws_bytes, void* stream) {
  RCGAN_CHECK_ARG(count > 0 && W && u && G && m && c && save && dW && accumulate, "sn_bwd_batched: bad args");
  RCGAN_CHECK_ARG(ws && ws_bytes >= rcgan_sn_workspace_batched(count, m, c), "sn_bwd_batched: workspace too small");
  cudaStream_t st = as_stream(stream);
  float* wsf = (float*)ws;
  for (int i0 = 0; i0 < count; i0 += SN_MAX_BATCH) {
    const int n = count - i0 < SN_MAX_BATCH ? count - i0 : SN_MAX_BATCH;
    SnBatch b;
    int mb, mp, mg;
    if (int e = sn_fill(b, i0, n, W, u, m, c, nullptr, nullptr, save, G, dW, accumulate, wsf, &mb, &mp, &mg)) return e;
    for (int k = 0; k < n; k++) {
      RCGAN_CHECK_ARG(b.it[k].G && b.it[k].dW, "sn_bwd_batched: null gradient pointer");
      wsf += sn_ws_floats(m[i0 + k], c[i0 + k]);
    }
    launch_pdl(sn_bwd_dot_batched, dim3(mp, n), 256, 0, st, b);
    RCGAN_LAUNCH_CHECK("sn_bwd_dot_batched");
    launch_pdl(sn_bwd_dw_batched, dim3(mg, n), 256, 0, st, b);
    RCGAN_LAUNCH_CHECK("sn_bwd_dw_batched");
  }
  return 0;
}
