// Bandwidth-bound rows x channels helpers: bias/activation, label concat, spatial mean,
// 2x2 mean-pool, nearest-neighbour upsample, residual add, casts, column sums.
// (mnist/ops.py:46-51,94-95; mnist/model.py:678,714-728; cifar10/gan_resnet.py:231-272,328,405-407)
// All are grid-stride, coalesced along the channel dimension.
#include "common.cuh"

namespace {

inline int grid_for(long work, int block) {
  long g = (work + block - 1) / block;
  long cap = (long)RCGAN_NUM_SMS * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <typename T>
__global__ void bias_act_fwd_kernel(const T* __restrict__ x, const float* __restrict__ bias, T* __restrict__ y, long rows,
                                    int c, int ldx, int ldy, int act, float leak) {
  long total = rows * c;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i / c;
    int ch = (int)(i - r * c);
    float v = to_f(x[r * ldx + ch]);
    if (bias) v += bias[ch];
    y[r * ldy + ch] = from_f<T>(act_fwd(v, act, leak));
  }
}

template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, long rows, int c,
                               int ld_dy, int ld_y, int ld_dx, int act, float leak, int accumulate) {
  long total = rows * c;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i / c;
    int ch = (int)(i - r * c);
    float g = to_f(dy[r * ld_dy + ch]) * act_bwd_from_y(to_f(y[r * ld_y + ch]), act, leak);
    long o = r * ld_dx + ch;
    if (accumulate) g += to_f(dx[o]);
    dx[o] = from_f<T>(g);
  }
}

template <typename T>
__global__ void concat_label_kernel(const T* __restrict__ a, int lda, const float* __restrict__ yb, T* __restrict__ out,
                                    int ldo, long rows, int rows_per_sample, int c1, int c2) {
  long total = rows * ldo;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i / ldo;
    int ch = (int)(i - r * ldo);
    T v;
    if (ch < c1) v = a[r * lda + ch];
    else if (ch < c1 + c2) v = from_f<T>(yb[(r / rows_per_sample) * c2 + (ch - c1)]);
    else v = from_f<T>(0.f);
    out[i] = v;
  }
}

template <typename T>
__global__ void slice_bwd_kernel(const T* __restrict__ dout, int ldo, T* __restrict__ da, int lda, long rows, int c1,
                                 int accumulate) {
  long total = rows * c1;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i / c1;
    int ch = (int)(i - r * c1);
    float g = to_f(dout[r * ldo + ch]);
    long o = r * lda + ch;
    if (accumulate) g += to_f(da[o]);
    da[o] = from_f<T>(g);
  }
}

// one thread per (sample, channel); loops hw (coalesced over channels)
template <typename T>
__global__ void meanhw_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int samples, int hw, int c, int relu) {
  long total = (long)samples * c;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long s = i / c;
    int ch = (int)(i - s * c);
    const T* p = x + s * hw * c + ch;
    float acc = 0.f;
    for (int k = 0; k < hw; k++) {
      float v = to_f(p[(long)k * c]);
      acc += relu ? fmaxf(v, 0.f) : v;
    }
    y[i] = from_f<T>(acc / hw);
  }
}

template <typename T>
__global__ void meanhw_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, T* __restrict__ dx, int samples,
                                  int hw, int c, int relu, int accumulate) {
  long total = (long)samples * hw * c;
  float inv = 1.f / hw;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int ch = (int)(i % c);
    long s = i / ((long)hw * c);
    float g = to_f(dy[s * c + ch]) * inv;
    if (relu && !(to_f(x[i]) > 0.f)) g = 0.f;
    if (accumulate) g += to_f(dx[i]);
    dx[i] = from_f<T>(g);
  }
}

template <typename T>
__global__ void avgpool2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int n, int h, int w, int c) {
  int ho = h / 2, wo = w / 2;
  long total = (long)n * ho * wo * c;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int ch = (int)(i % c);
    long r = i / c;
    int ox = (int)(r % wo);
    r /= wo;
    int oy = (int)(r % ho);
    long nb = r / ho;
    const T* p = x + ((nb * h + 2 * oy) * w + 2 * ox) * c + ch;
    // add_n([x[::2,::2], x[1::2,::2], x[::2,1::2], x[1::2,1::2]]) / 4  (gan_resnet.py:239-240)
    float v = ((to_f(p[0]) + to_f(p[(long)w * c])) + to_f(p[c])) + to_f(p[(long)w * c + c]);
    y[i] = from_f<T>(v * 0.25f);
  }
}

template <typename T>
__global__ void avgpool2_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int n, int h, int w, int c,
                                    int accumulate) {
  int ho = h / 2, wo = w / 2;
  long total = (long)n * h * w * c;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int ch = (int)(i % c);
    long r = i / c;
    int x_ = (int)(r % w);
    r /= w;
    int y_ = (int)(r % h);
    long nb = r / h;
    float g = 0.25f * to_f(dy[((nb * ho + y_ / 2) * wo + x_ / 2) * c + ch]);
    if (accumulate) g += to_f(dx[i]);
    dx[i] = from_f<T>(g);
  }
}

template <typename T>
__global__ void upsample2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int n, int h, int w, int c) {
  long total = (long)n * (2 * h) * (2 * w) * c;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int ch = (int)(i % c);
    long r = i / c;
    int ox = (int)(r % (2 * w));
    r /= 2 * w;
    int oy = (int)(r % (2 * h));
    long nb = r / (2 * h);
    y[i] = x[((nb * h + oy / 2) * w + ox / 2) * c + ch];
  }
}

template <typename T>
__global__ void upsample2_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int n, int h, int w, int c,
                                     int accumulate) {
  long total = (long)n * h * w * c;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int ch = (int)(i % c);
    long r = i / c;
    int x_ = (int)(r % w);
    r /= w;
    int y_ = (int)(r % h);
    long nb = r / h;
    const T* p = dy + ((nb * 2 * h + 2 * y_) * (2 * w) + 2 * x_) * c + ch;
    float g = to_f(p[0]) + to_f(p[c]) + to_f(p[(long)2 * w * c]) + to_f(p[(long)2 * w * c + c]);
    if (accumulate) g += to_f(dx[i]);
    dx[i] = from_f<T>(g);
  }
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long numel) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (long)gridDim.x * blockDim.x)
    out[i] = from_f<T>(to_f(a[i]) + to_f(b[i]));
}

template <typename T>
__global__ void copy_acc_kernel(const T* __restrict__ src, T* __restrict__ dst, long numel, int accumulate) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (long)gridDim.x * blockDim.x) {
    float v = to_f(src[i]);
    if (accumulate) v += to_f(dst[i]);
    dst[i] = from_f<T>(v);
  }
}

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, D* __restrict__ dst, long numel) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (long)gridDim.x * blockDim.x)
    dst[i] = from_f<D>(to_f(src[i]));
}

// column sums: block = 32 channels x 8 row-lanes, grid.x over channel groups, grid.y over row slabs;
// slab partials are combined with fp32 atomics only when grid.y > 1 (db pre-zeroed by a memset node).
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ dy, int rows, int c, int ld, float* __restrict__ db, int accumulate,
                              int use_atomic) {
  __shared__ float sh[8][33];
  int ch = blockIdx.x * 32 + threadIdx.x;
  int rows_per = (rows + gridDim.y - 1) / gridDim.y;
  int r0 = blockIdx.y * rows_per, r1 = min(rows, r0 + rows_per);
  float acc = 0.f;
  if (ch < c)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) acc += to_f(dy[(size_t)r * ld + ch]);
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && ch < c) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) s += sh[k][threadIdx.x];
    if (use_atomic) atomicAdd(&db[ch], s);
    else db[ch] = accumulate ? db[ch] + s : s;
  }
}

template <typename T>
__global__ void im2col_kernel(const T* __restrict__ x, T* __restrict__ P, int n, int h, int w, int cin, int ho, int wo, int kh,
                              int kw, int stride, int pad_t, int pad_l, int ldx, int ldp) {
  const int K = kh * kw * cin;
  long total = (long)n * ho * wo * ldp;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int k = (int)(i % ldp);
    long m = i / ldp;
    T v = from_f<T>(0.f);
    if (k < K) {
      int ci = k % cin, tap = k / cin;
      int ox = (int)(m % wo);
      long r = m / wo;
      int oy = (int)(r % ho);
      long nb = r / ho;
      int iy = oy * stride - pad_t + tap / kw, ix = ox * stride - pad_l + tap % kw;
      if (iy >= 0 && iy < h && ix >= 0 && ix < w) v = x[((nb * h + iy) * w + ix) * ldx + ci];
    }
    P[i] = v;
  }
}

__global__ void preprocess_cifar_kernel(const int32_t* __restrict__ chw, const float* __restrict__ noise, void* out_,
                                        int n, int is_bf16) {
  long total = (long)n * 3072;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    // i indexes the NHWC output: ((nb*32 + y)*32 + x)*3 + ch ; source is CHW
    int ch = (int)(i % 3);
    long r = i / 3;
    int pix = (int)(r % 1024);
    long nb = r / 1024;
    // reference (gan_resnet.py:548-552): 2*((int/256) - .5) + U(0,1/128), noise added BEFORE the transpose
    long src = nb * 3072 + ch * 1024 + pix;
    float v = 2.f * ((float)chw[src] / 256.f - 0.5f);
    if (noise) v += noise[src];
    if (is_bf16) reinterpret_cast<bf16*>(out_)[i] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(out_)[i] = v;
  }
}

}  // namespace

#define DISPATCH_T(dtype, ...)                                   \
  if ((dtype) == RCGAN_F32) { typedef float T; __VA_ARGS__; }     \
  else if ((dtype) == RCGAN_BF16) { typedef bf16 T; __VA_ARGS__; } \
  else { rcgan_set_error("bad dtype %d", (int)(dtype)); return RCGAN_EBADSHAPE; }

extern "C" int rcgan_bias_act_fwd(const void* x, const float* bias, void* y, long rows, int c, int ldx, int ldy,
                                  int dtype, int act, float leak, void* stream) {
  RCGAN_CHECK_ARG(rows >= 0 && c > 0 && ldx >= c && ldy >= c, "bias_act_fwd: bad shape");
  if (rows == 0) return 0;
  DISPATCH_T(dtype, bias_act_fwd_kernel<T><<<grid_for(rows * c, 256), 256, 0, as_stream(stream)>>>(
                        (const T*)x, bias, (T*)y, rows, c, ldx, ldy, act, leak));
  RCGAN_LAUNCH_CHECK("bias_act_fwd");
  return 0;
}

extern "C" int rcgan_act_bwd(const void* dy, const void* y, void* dx, long rows, int c, int ld_dy, int ld_y, int ld_dx,
                             int dtype, int act, float leak, int accumulate, void* stream) {
  RCGAN_CHECK_ARG(rows >= 0 && c > 0, "act_bwd: bad shape");
  if (rows == 0) return 0;
  DISPATCH_T(dtype, act_bwd_kernel<T><<<grid_for(rows * c, 256), 256, 0, as_stream(stream)>>>(
                        (const T*)dy, (const T*)y, (T*)dx, rows, c, ld_dy, ld_y, ld_dx, act, leak, accumulate));
  RCGAN_LAUNCH_CHECK("act_bwd");
  return 0;
}

extern "C" int rcgan_concat_label_fwd(const void* a, int lda, const float* yb, void* out, int ldo, long rows,
                                      int rows_per_sample, int c1, int c2, int dtype, void* stream) {
  RCGAN_CHECK_ARG(rows > 0 && rows_per_sample > 0 && c1 >= 0 && c2 >= 0 && ldo >= c1 + c2 && lda >= c1,
                  "concat_label_fwd: bad shape");
  DISPATCH_T(dtype, concat_label_kernel<T><<<grid_for(rows * ldo, 256), 256, 0, as_stream(stream)>>>(
                        (const T*)a, lda, yb, (T*)out, ldo, rows, rows_per_sample, c1, c2));
  RCGAN_LAUNCH_CHECK("concat_label_fwd");
  return 0;
}

extern "C" int rcgan_slice_bwd(const void* dout, int ldo, void* da, int lda, long rows, int c1, int dtype,
                               int accumulate, void* stream) {
  RCGAN_CHECK_ARG(rows > 0 && c1 > 0 && ldo >= c1 && lda >= c1, "slice_bwd: bad shape");
  DISPATCH_T(dtype, slice_bwd_kernel<T><<<grid_for(rows * c1, 256), 256, 0, as_stream(stream)>>>(
                        (const T*)dout, ldo, (T*)da, lda, rows, c1, accumulate));
  RCGAN_LAUNCH_CHECK("slice_bwd");
  return 0;
}

extern "C" int rcgan_meanhw_fwd(const void* x, void* y, int samples, int hw, int c, int dtype, int relu, void* stream) {
  RCGAN_CHECK_ARG(samples > 0 && hw > 0 && c > 0, "meanhw_fwd: bad shape");
  DISPATCH_T(dtype, meanhw_fwd_kernel<T><<<grid_for((long)samples * c, 128), 128, 0, as_stream(stream)>>>(
                        (const T*)x, (T*)y, samples, hw, c, relu));
  RCGAN_LAUNCH_CHECK("meanhw_fwd");
  return 0;
}

extern "C" int rcgan_meanhw_bwd(const void* dy, const void* x, void* dx, int samples, int hw, int c, int dtype, int relu,
                                int accumulate, void* stream) {
  RCGAN_CHECK_ARG(samples > 0 && hw > 0 && c > 0, "meanhw_bwd: bad shape");
  DISPATCH_T(dtype, meanhw_bwd_kernel<T><<<grid_for((long)samples * hw * c, 256), 256, 0, as_stream(stream)>>>(
                        (const T*)dy, (const T*)x, (T*)dx, samples, hw, c, relu, accumulate));
  RCGAN_LAUNCH_CHECK("meanhw_bwd");
  return 0;
}

extern "C" int rcgan_avgpool2_fwd(const void* x, void* y, int n, int h, int w, int c, int dtype, void* stream) {
  RCGAN_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && h % 2 == 0 && w % 2 == 0, "avgpool2_fwd: bad shape");
  DISPATCH_T(dtype, avgpool2_fwd_kernel<T><<<grid_for((long)n * h * w * c / 4, 256), 256, 0, as_stream(stream)>>>(
                        (const T*)x, (T*)y, n, h, w, c));
  RCGAN_LAUNCH_CHECK("avgpool2_fwd");
  return 0;
}
extern "C" int rcgan_avgpool2_bwd(const void* dy, void* dx, int n, int h, int w, int c, int dtype, int accumulate,
                                  void* stream) {
  RCGAN_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && h % 2 == 0 && w % 2 == 0, "avgpool2_bwd: bad shape");
  DISPATCH_T(dtype, avgpool2_bwd_kernel<T><<<grid_for((long)n * h * w * c, 256), 256, 0, as_stream(stream)>>>(
                        (const T*)dy, (T*)dx, n, h, w, c, accumulate));
  RCGAN_LAUNCH_CHECK("avgpool2_bwd");
  return 0;
}
extern "C" int rcgan_upsample2_fwd(const void* x, void* y, int n, int h, int w, int c, int dtype, void* stream) {
  RCGAN_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0, "upsample2_fwd: bad shape");
  DISPATCH_T(dtype, upsample2_fwd_kernel<T><<<grid_for((long)n * h * w * c * 4, 256), 256, 0, as_stream(stream)>>>(
                        (const T*)x, (T*)y, n, h, w, c));
  RCGAN_LAUNCH_CHECK("upsample2_fwd");
  return 0;
}
extern "C" int rcgan_upsample2_bwd(const void* dy, void* dx, int n, int h, int w, int c, int dtype, int accumulate,
                                   void* stream) {
  RCGAN_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0, "upsample2_bwd: bad shape");
  DISPATCH_T(dtype, upsample2_bwd_kernel<T><<<grid_for((long)n * h * w * c, 256), 256, 0, as_stream(stream)>>>(
                        (const T*)dy, (T*)dx, n, h, w, c, accumulate));
  RCGAN_LAUNCH_CHECK("upsample2_bwd");
  return 0;
}
extern "C" int rcgan_add(const void* a, const void* b, void* out, long numel, int dtype, void* stream) {
  RCGAN_CHECK_ARG(numel > 0, "add: bad shape");
  DISPATCH_T(dtype, add_kernel<T><<<grid_for(numel, 256), 256, 0, as_stream(stream)>>>((const T*)a, (const T*)b, (T*)out, numel));
  RCGAN_LAUNCH_CHECK("add");
  return 0;
}
extern "C" int rcgan_copy_acc(const void* src, void* dst, long numel, int dtype, int accumulate, void* stream) {
  RCGAN_CHECK_ARG(numel > 0, "copy_acc: bad shape");
  DISPATCH_T(dtype, copy_acc_kernel<T><<<grid_for(numel, 256), 256, 0, as_stream(stream)>>>((const T*)src, (T*)dst, numel, accumulate));
  RCGAN_LAUNCH_CHECK("copy_acc");
  return 0;
}
extern "C" int rcgan_cast(const void* src, int src_dtype, void* dst, int dst_dtype, long numel, void* stream) {
  RCGAN_CHECK_ARG(numel > 0, "cast: bad shape");
  int g = grid_for(numel, 256);
  cudaStream_t st = as_stream(stream);
  if (src_dtype == RCGAN_F32 && dst_dtype == RCGAN_BF16) cast_kernel<float, bf16><<<g, 256, 0, st>>>((const float*)src, (bf16*)dst, numel);
  else if (src_dtype == RCGAN_BF16 && dst_dtype == RCGAN_F32) cast_kernel<bf16, float><<<g, 256, 0, st>>>((const bf16*)src, (float*)dst, numel);
  else if (src_dtype == RCGAN_F32 && dst_dtype == RCGAN_F32) cast_kernel<float, float><<<g, 256, 0, st>>>((const float*)src, (float*)dst, numel);
  else if (src_dtype == RCGAN_BF16 && dst_dtype == RCGAN_BF16) cast_kernel<bf16, bf16><<<g, 256, 0, st>>>((const bf16*)src, (bf16*)dst, numel);
  else { rcgan_set_error("cast: bad dtypes"); return RCGAN_EBADSHAPE; }
  RCGAN_LAUNCH_CHECK("cast");
  return 0;
}

extern "C" int rcgan_colsum(const void* dy, int rows, int c, int ld, int dtype, float* db, int accumulate, void* stream) {
  RCGAN_CHECK_ARG(rows > 0 && c > 0 && ld >= c, "colsum: bad shape");
  int gx = ceil_div(c, 32);
  int gy = 1;
  if (rows >= 1024) {
    gy = (2 * RCGAN_NUM_SMS + gx - 1) / gx;
    int maxy = rows / 256;
    if (gy > maxy) gy = maxy;
    if (gy < 1) gy = 1;
  }
  cudaStream_t st = as_stream(stream);
  if (gy > 1 && !accumulate) {
    cudaError_t e = cudaMemsetAsync(db, 0, sizeof(float) * c, st);
    if (e != cudaSuccess) { rcgan_set_error("colsum: memset failed: %s", cudaGetErrorString(e)); return RCGAN_ECUDA; }
  }
  dim3 grid(gx, gy), block(32, 8);
  DISPATCH_T(dtype, colsum_kernel<T><<<grid, block, 0, st>>>((const T*)dy, rows, c, ld, db, accumulate, gy > 1));
  RCGAN_LAUNCH_CHECK("colsum");
  return 0;
}

extern "C" int rcgan_im2col(const rcgan_conv_desc* d, const void* x, void* patches, int ldp, void* stream) {
  RCGAN_CHECK_ARG(d && x && patches && ldp >= d->kh * d->kw * d->cin, "im2col: bad args");
  long total = (long)d->n * d->ho * d->wo * ldp;
  DISPATCH_T(d->dtype, im2col_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
                           (const T*)x, (T*)patches, d->n, d->h, d->w, d->cin, d->ho, d->wo, d->kh, d->kw, d->stride, d->pad_t,
                           d->pad_l, d->ldx, ldp));
  RCGAN_LAUNCH_CHECK("im2col");
  return 0;
}

extern "C" int rcgan_preprocess_cifar(const int32_t* chw, const float* noise, void* out, int n, int dtype, void* stream) {
  RCGAN_CHECK_ARG(n > 0 && (dtype == RCGAN_F32 || dtype == RCGAN_BF16), "preprocess_cifar: bad args");
  preprocess_cifar_kernel<<<grid_for((long)n * 3072, 256), 256, 0, as_stream(stream)>>>(chw, noise, out, n, dtype == RCGAN_BF16);
  RCGAN_LAUNCH_CHECK("preprocess_cifar");
  return 0;
}
