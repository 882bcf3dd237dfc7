// tf.train.AdamOptimizer over one flat fp32 arena (mnist/model.py:250-262; cifar10/gan_resnet.py:802-817),
// fused with the 1/world_size gradient scaling that follows the NCCL allreduce-sum and with the
// max-norm variable constraint clip_by_value(-1,1) of mnist/ops.py:101-111.
// HBM-bound: 16 B read + 12 B written per parameter (28 B/param).
#include "common.cuh"

namespace {
struct ClipRanges { long lo[8]; long hi[8]; int n; };

__global__ void __launch_bounds__(256) adam_tf_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                      float* __restrict__ v, long numel, float lr_t, const float* __restrict__ lr_t_dev,
                                                      float b1, float b2, float eps, float gs, ClipRanges cr) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  if (lr_t_dev) lr_t = *lr_t_dev;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < numel; i += (long)gridDim.x * 256) {
    float gi = g[i] * gs;
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    float pi = p[i] - lr_t * mi / (sqrtf(vi) + eps);
    for (int k = 0; k < cr.n; k++)
      if (i >= cr.lo[k] && i < cr.hi[k]) pi = fminf(fmaxf(pi, -1.f), 1.f);
    m[i] = mi; v[i] = vi; p[i] = pi;
  }
}
}  // namespace

extern "C" int rcgan_adam_tf(float* p, const float* g, float* m, float* v, long numel, float lr_t, const float* lr_t_dev,
                             float b1, float b2, float eps, float grad_scale, const long* clip_lo, const long* clip_hi, int n_clip, void* stream) {
  RCGAN_CHECK_ARG(p && g && m && v && numel > 0, "adam_tf: bad args");
  RCGAN_CHECK_ARG(n_clip >= 0 && n_clip <= 8, "adam_tf: at most 8 clip ranges");
  ClipRanges cr;
  cr.n = n_clip;
  for (int k = 0; k < n_clip; k++) { cr.lo[k] = clip_lo[k]; cr.hi[k] = clip_hi[k]; }
  int grid = (int)((numel + 255) / 256);
  if (grid > RCGAN_NUM_SMS * 8) grid = RCGAN_NUM_SMS * 8;
  launch_pdl(adam_tf_kernel, grid, 256, 0, as_stream(stream), p, g, m, v, numel, lr_t, lr_t_dev, b1, b2, eps, grad_scale, cr);
  RCGAN_LAUNCH_CHECK("adam_tf");
  return 0;
}

extern "C" int rcgan_zero(void* ptr, size_t bytes, void* stream) {
  RCGAN_CHECK_ARG(ptr || bytes == 0, "zero: null pointer");
  if (bytes == 0) return 0;
  cudaError_t e = cudaMemsetAsync(ptr, 0, bytes, as_stream(stream));
  if (e != cudaSuccess) { rcgan_set_error("zero: %s", cudaGetErrorString(e)); return RCGAN_ECUDA; }
  return 0;
}

namespace {
__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, const float* __restrict__ g, long numel, float step) {
  pdl_sync();
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < numel; i += (long)gridDim.x * 256) p[i] -= step * g[i];
}
}  // namespace

extern "C" int rcgan_sgd(float* p, const float* g, long numel, float lr, float grad_scale, void* stream) {
  RCGAN_CHECK_ARG(p && g && numel > 0, "sgd: bad args");
  long grid = (numel + 255) / 256;
  if (grid > 8L * RCGAN_NUM_SMS) grid = 8L * RCGAN_NUM_SMS;
  launch_pdl(sgd_kernel, (int)grid, 256, 0, as_stream(stream), p, g, numel, lr * grad_scale);
  RCGAN_LAUNCH_CHECK("sgd");
  return 0;
}
