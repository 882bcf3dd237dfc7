// RCGAN noisy-channel losses.
//
// rcgan_channel_loss: the unified projection-discriminator loss of SURVEY appendix B --
//   l[b,j] = psi[b] + <h[b,:], V[j,:]>,  L = scale * sum_b sum_j wgt[b,j] * phi(l[b,j])
// covering mnist/model.py:150-207 + :679-686 and cifar10/gan_resnet.py:588-606, 649-685, 751-760:
//   wgt = onehot(noisy label)            -> known-C RCGAN / biased (bit-identical to gathering V[label])
//   wgt = onehot(y_gen) . softmax(Lambda) -> learned-C RCGAN-U (one trunk evaluation instead of the
//                                           reference's 10 label-wise discriminator calls)
//   wgt = row of C^-1                     -> unbiased
// forward and all four gradients (dh, dpsi, dV, dwgt) in ONE pass over h: one warp per sample,
// V staged in shared memory, warp-shuffle reductions over d.  HBM-bound: B*(d+1+k) elements in,
// the same out.
#include "common.cuh"

namespace {

constexpr int KMAX = 16;

__device__ __forceinline__ float softplus_f(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void phi(int mode, float l, float& val, float& der) {
  switch (mode) {
    case RCGAN_HINGE_D_REAL: val = fmaxf(1.f - l, 0.f); der = (1.f - l > 0.f) ? -1.f : 0.f; break;
    case RCGAN_HINGE_D_FAKE: val = fmaxf(1.f + l, 0.f); der = (1.f + l > 0.f) ? 1.f : 0.f; break;
    case RCGAN_HINGE_G: val = -l; der = -1.f; break;
    case RCGAN_CE_D_FAKE: val = softplus_f(l); der = sigmoid_f(l); break;
    default: /* CE_D_REAL, CE_G: sCE(l, 1) */ val = softplus_f(-l); der = sigmoid_f(l) - 1.f; break;
  }
}

template <typename T, int DPL>  // DPL = d / 32 columns per lane
__global__ void __launch_bounds__(256) channel_loss_kernel(const T* __restrict__ h, const float* __restrict__ psi,
                                                           const float* __restrict__ V, const float* __restrict__ wgt, int B,
                                                           int d, int k, int mode, float scale, float* loss_acc,
                                                           float* __restrict__ logits, T* __restrict__ dh, int accumulate_dh,
                                                           float* __restrict__ dpsi, float* dV, float* __restrict__ dwgt) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  extern __shared__ float sh[];  // V[k*d] then dV accumulators [k*d] then 8 loss partials
  float* Vs = sh;
  float* dVs = sh + k * d;
  float* lsh = sh + 2 * k * d;
  for (int i = threadIdx.x; i < k * d; i += 256) { Vs[i] = V[i]; dVs[i] = 0.f; }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float dv_acc[KMAX][DPL];
#pragma unroll
  for (int j = 0; j < KMAX; j++)
#pragma unroll
    for (int q = 0; q < DPL; q++) dv_acc[j][q] = 0.f;
  float loss = 0.f;
  for (int b = blockIdx.x * 8 + wid; b < B; b += gridDim.x * 8) {
    float hv[DPL];
#pragma unroll
    for (int q = 0; q < DPL; q++) hv[q] = to_f(h[(size_t)b * d + lane + 32 * q]);
    const float ps = psi ? psi[b] : 0.f;
    float dhv[DPL];
#pragma unroll
    for (int q = 0; q < DPL; q++) dhv[q] = 0.f;
    float dps = 0.f;
#pragma unroll
    for (int j = 0; j < KMAX; j++) {
      if (j < k) {
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < DPL; q++) acc = fmaf(hv[q], Vs[j * d + lane + 32 * q], acc);
        float l = warp_sum(acc) + ps;
        float w = wgt[(size_t)b * k + j];
        float val, der;
        phi(mode, l, val, der);
        float g = w * der * scale;
        loss = fmaf(w, val, loss);  // identical on every lane; lane 0's copy is used
        dps += g;
#pragma unroll
        for (int q = 0; q < DPL; q++) {
          dhv[q] = fmaf(g, Vs[j * d + lane + 32 * q], dhv[q]);
          dv_acc[j][q] = fmaf(g, hv[q], dv_acc[j][q]);
        }
        if (lane == 0) {
          if (logits) logits[(size_t)b * k + j] = l;
          if (dwgt) dwgt[(size_t)b * k + j] = val * scale;
        }
      }
    }
    if (dh) {
#pragma unroll
      for (int q = 0; q < DPL; q++) {
        size_t o = (size_t)b * d + lane + 32 * q;
        float v = dhv[q];
        if (accumulate_dh) v += to_f(dh[o]);
        dh[o] = from_f<T>(v);
      }
    }
    if (dpsi && lane == 0) dpsi[b] = dps;
  }
  if (dV) {
#pragma unroll
    for (int j = 0; j < KMAX; j++)
      if (j < k)
#pragma unroll
        for (int q = 0; q < DPL; q++) atomicAdd(&dVs[j * d + lane + 32 * q], dv_acc[j][q]);
  }
  if (lane == 0) lsh[wid] = loss;
  __syncthreads();
  if (dV)
    for (int i = threadIdx.x; i < k * d; i += 256) atomicAdd(&dV[i], dVs[i]);
  if (threadIdx.x == 0 && loss_acc) {
    float s = 0.f;
    for (int w = 0; w < 8; w++) s += lsh[w];
    atomicAdd(loss_acc, s * scale);
  }
}

__global__ void __launch_bounds__(256) sigmoid_ce_kernel(const float* __restrict__ logits, const float* __restrict__ targets,
                                                         long numel, float scale, float* loss_acc, float* __restrict__ dl) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  __shared__ float red[33];
  float s = 0.f;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < numel; i += (long)gridDim.x * 256) {
    float x = logits[i], z = targets[i];
    s += fmaxf(x, 0.f) - x * z + log1pf(expf(-fabsf(x)));
    if (dl) dl[i] = scale * (sigmoid_f(x) - z);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0 && loss_acc) atomicAdd(loss_acc, s * scale);
}

__global__ void __launch_bounds__(256) logit_loss_kernel(const float* __restrict__ logits, long B, int mode, float scale,
                                                        float* loss_acc, float* __restrict__ dl) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  __shared__ float red[33];
  float s = 0.f;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < B; i += (long)gridDim.x * 256) {
    float val, der;
    phi(mode, logits[i], val, der);
    s += val;
    if (dl) dl[i] = scale * der;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0 && loss_acc) atomicAdd(loss_acc, s * scale);
}

__global__ void softmax_rows_fwd_kernel(const float* __restrict__ L, float* __restrict__ C, int rows, int k) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float mx = -INFINITY;
  for (int j = 0; j < k; j++) mx = fmaxf(mx, L[r * k + j]);
  float s = 0.f;
  for (int j = 0; j < k; j++) s += expf(L[r * k + j] - mx);
  for (int j = 0; j < k; j++) C[r * k + j] = expf(L[r * k + j] - mx) / s;
}

__global__ void softmax_rows_bwd_kernel(const float* __restrict__ C, const float* __restrict__ dC, float* __restrict__ dL,
                                        int rows, int k, int accumulate) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float dot = 0.f;
  for (int j = 0; j < k; j++) dot = fmaf(dC[r * k + j], C[r * k + j], dot);
  for (int j = 0; j < k; j++) {
    float v = C[r * k + j] * (dC[r * k + j] - dot);
    dL[r * k + j] = accumulate ? dL[r * k + j] + v : v;
  }
}

__global__ void gather_rows_fwd_kernel(const float* __restrict__ C, const int* __restrict__ y, float* __restrict__ wgt, int B,
                                       int k) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * k) return;
  int b = i / k, j = i - b * k;
  wgt[i] = C[y[b] * k + j];
}

// deterministic: one block per destination row, thread j < k scans the batch
__global__ void gather_rows_bwd_kernel(const float* __restrict__ dwgt, const int* __restrict__ y, float* __restrict__ dC,
                                       int B, int k, int accumulate) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  __shared__ float red[33];
  int r = blockIdx.x;
  for (int j = 0; j < k; j++) {
    float s = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x)
      if (y[b] == r) s += dwgt[(size_t)b * k + j];
    s = block_sum(s, red);
    if (threadIdx.x == 0) dC[r * k + j] = accumulate ? dC[r * k + j] + s : s;
  }
}

}  // namespace

extern "C" int rcgan_channel_loss(const void* h, const float* psi, const float* V, const float* wgt, int B, int d, int k,
                                  int dtype, int mode, float scale, float* loss_acc, float* logits, void* dh,
                                  int accumulate_dh, float* dpsi, float* dV, float* dwgt, void* stream) {
  RCGAN_CHECK_ARG(h && V && wgt && B > 0, "channel_loss: bad args");
  RCGAN_CHECK_ARG(k > 0 && k <= KMAX, "channel_loss: k=%d unsupported (1..%d)", k, KMAX);
  RCGAN_CHECK_ARG(d == 32 || d == 64 || d == 128 || d == 256, "channel_loss: d=%d unsupported (32/64/128/256)", d);
  RCGAN_CHECK_ARG(mode >= 0 && mode <= RCGAN_CE_G, "channel_loss: bad mode");
  cudaStream_t st = as_stream(stream);
  int grid = ceil_div(B, 8);
  if (grid > RCGAN_NUM_SMS * 2) grid = RCGAN_NUM_SMS * 2;
  size_t shb = ((size_t)2 * k * d + 8) * sizeof(float);
#define CL_LAUNCH(T, DPL)                                                                                             \
  launch_pdl(channel_loss_kernel<T, DPL>, grid, 256, shb, st, (const T*)h, psi, V, wgt, B, d, k, mode, scale, loss_acc, logits, \
                                                      (T*)dh, accumulate_dh, dpsi, dV, dwgt)
  if (dtype == RCGAN_F32) {
    if (d == 32) CL_LAUNCH(float, 1); else if (d == 64) CL_LAUNCH(float, 2); else if (d == 128) CL_LAUNCH(float, 4); else CL_LAUNCH(float, 8);
  } else if (dtype == RCGAN_BF16) {
    if (d == 32) CL_LAUNCH(bf16, 1); else if (d == 64) CL_LAUNCH(bf16, 2); else if (d == 128) CL_LAUNCH(bf16, 4); else CL_LAUNCH(bf16, 8);
  } else {
    rcgan_set_error("channel_loss: bad dtype");
    return RCGAN_EBADSHAPE;
  }
#undef CL_LAUNCH
  RCGAN_LAUNCH_CHECK("channel_loss");
  return 0;
}

extern "C" int rcgan_sigmoid_ce(const float* logits, const float* targets, long numel, float scale, float* loss_acc,
                                float* dlogits, void* stream) {
  RCGAN_CHECK_ARG(logits && targets && numel > 0, "sigmoid_ce: bad args");
  int grid = (int)((numel + 255) / 256);
  if (grid > RCGAN_NUM_SMS) grid = RCGAN_NUM_SMS;
  launch_pdl(sigmoid_ce_kernel, grid, 256, 0, as_stream(stream), logits, targets, numel, scale, loss_acc, dlogits);
  RCGAN_LAUNCH_CHECK("sigmoid_ce");
  return 0;
}

extern "C" int rcgan_logit_loss(const float* logits, long B, int mode, float scale, float* loss_acc, float* dlogits,
                                void* stream) {
  RCGAN_CHECK_ARG(logits && B > 0 && mode >= 0 && mode <= RCGAN_CE_G, "logit_loss: bad args");
  int grid = (int)((B + 255) / 256);
  if (grid > RCGAN_NUM_SMS) grid = RCGAN_NUM_SMS;
  launch_pdl(logit_loss_kernel, grid, 256, 0, as_stream(stream), logits, B, mode, scale, loss_acc, dlogits);
  RCGAN_LAUNCH_CHECK("logit_loss");
  return 0;
}

extern "C" int rcgan_softmax_rows_fwd(const float* logits, float* C, int rows, int k, void* stream) {
  RCGAN_CHECK_ARG(logits && C && rows > 0 && k > 0, "softmax_rows_fwd: bad args");
  launch_pdl(softmax_rows_fwd_kernel, ceil_div(rows, 32), 32, 0, as_stream(stream), logits, C, rows, k);
  RCGAN_LAUNCH_CHECK("softmax_rows_fwd");
  return 0;
}
extern "C" int rcgan_softmax_rows_bwd(const float* C, const float* dC, float* dlogits, int rows, int k, int accumulate,
                                      void* stream) {
  RCGAN_CHECK_ARG(C && dC && dlogits && rows > 0 && k > 0, "softmax_rows_bwd: bad args");
  launch_pdl(softmax_rows_bwd_kernel, ceil_div(rows, 32), 32, 0, as_stream(stream), C, dC, dlogits, rows, k, accumulate);
  RCGAN_LAUNCH_CHECK("softmax_rows_bwd");
  return 0;
}
extern "C" int rcgan_gather_rows_fwd(const float* C, const int* y, float* wgt, int B, int k, void* stream) {
  RCGAN_CHECK_ARG(C && y && wgt && B > 0 && k > 0, "gather_rows_fwd: bad args");
  launch_pdl(gather_rows_fwd_kernel, ceil_div((long)B * k, 256), 256, 0, as_stream(stream), C, y, wgt, B, k);
  RCGAN_LAUNCH_CHECK("gather_rows_fwd");
  return 0;
}
extern "C" int rcgan_gather_rows_bwd(const float* dwgt, const int* y, float* dC, int B, int k, int rows, int accumulate,
                                     void* stream) {
  RCGAN_CHECK_ARG(dwgt && y && dC && B > 0 && k > 0 && rows > 0, "gather_rows_bwd: bad args");
  launch_pdl(gather_rows_bwd_kernel, rows, 128, 0, as_stream(stream), dwgt, y, dC, B, k, accumulate);
  RCGAN_LAUNCH_CHECK("gather_rows_bwd");
  return 0;
}

// ------------------------------------------------------------------------------------------------ label recovery
namespace {
// one block per (r, j): sq = mean_p (actual[r] - sample[r*k+j])^2, then the sample gradient in a second sweep (L1/L2 hits)
__global__ void __launch_bounds__(128) recover_mse_kernel(const float* __restrict__ sample, const float* __restrict__ actual,
                                                          const float* __restrict__ y_rec, int R, int k, int npix,
                                                          float* __restrict__ loss_acc, float* __restrict__ sq,
                                                          float* __restrict__ dsample, float* __restrict__ dyrec) {
  pdl_sync();
  __shared__ float red[33];
  const int rj = blockIdx.x, r = rj / k;
  const float* s = sample + (size_t)rj * npix;
  const float* a = actual + (size_t)r * npix;
  float acc = 0.f;
  for (int p = threadIdx.x; p < npix; p += 128) { const float d = a[p] - s[p]; acc = fmaf(d, d, acc); }
  acc = block_sum(acc, red);
  const float m = acc / (float)npix, w = y_rec[rj], invR = 1.f / (float)R;
  if (threadIdx.x == 0) {
    if (sq) sq[rj] = m;
    if (dyrec) dyrec[rj] = m * invR;
    if (loss_acc) atomicAdd(loss_acc, m * w * invR);
  }
  if (dsample) {
    const float g = 2.f / (float)npix * w * invR;
    float* ds = dsample + (size_t)rj * npix;
    for (int p = threadIdx.x; p < npix; p += 128) ds[p] = g * (s[p] - a[p]);
  }
}
}  // namespace

extern "C" int rcgan_recover_mse(const float* sample, const float* actual, const float* y_rec, int R, int k, int npix,
                                 float* loss_acc, float* sq, float* dsample, float* dyrec, void* stream) {
  RCGAN_CHECK_ARG(sample && actual && y_rec && R > 0 && k > 0 && npix > 0, "recover_mse: bad args");
  launch_pdl(recover_mse_kernel, R * k, 128, 0, as_stream(stream), sample, actual, y_rec, R, k, npix, loss_acc, sq, dsample, dyrec);
  RCGAN_LAUNCH_CHECK("recover_mse");
  return 0;
}
