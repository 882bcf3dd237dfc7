// tcgen05 / TMEM / TMA implicit-GEMM convolution (bf16 operands, fp32 accumulation in tensor memory).
//
// Covers the GEMM-shaped layers of the reference in bf16 mode -- tf.nn.conv2d fprop and its dgrad
// (mnist/ops.py:62, cifar10/common/ops/conv2d.py:181-187), tf.nn.conv2d_transpose as a forward op
// (mnist/ops.py:78, = dgrad; stride 2 runs as four output-parity classes, each a dense small-filter conv, so only
// useful flops are issued) and the tf.matmul linears (1x1, h=w=1).
//
//   GEMM view:  D[128 pixels x BN channels] += A[128 x 64] * B[BN x 64]^T   per K block (one filter tap x 64 channels)
//   A (activations): TMA im2col (cp.async.bulk.tensor.4d...im2col: 128 pixels x 64 channels per load, the SAME-padding halo
//       and the channel tail zero-filled by the bounding box) into the UMMA canonical K-major SWIZZLE_128B layout; when the
//       im2col map is not encodable, 4 producer warps gather the same layout with 16-byte cp.async (one-tile kernel only)
//   B (weights)    : TMA 3-D tiled load (cp.async.bulk.tensor, SWIZZLE_128B) from the packed bf16 weight copy
//   MMA            : one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) x4 per K block,
//                    accumulator in TMEM; tcgen05.commit releases the smem stage / signals the epilogue
//   epilogue       : 4 warps (8 in the persistent kernels) read TMEM with tcgen05.ld.32x32b, add bias, apply the activation,
//                    stage whole rows in shared memory and store bf16 or fp32 NHWC rows (optionally accumulating, adding a
//                    residual, masking with an activation derivative, emitting a second relu output / column statistics)
// Three kernels share these operands and the epilogue:
//   conv_tc_kernel         one 128 x BN tile per CTA, 3-4 stage ring, <= 96 KB smem so two CTAs share an SM (one CTA's epilogue
//                          overlaps the other's MMAs): small problems
//   conv_tc_persist_kernel one CTA per SM loops over tiles; accumulators double buffered in TMEM; 10 warps (two epilogue groups)
//   conv_tc_pair_kernel    the persistent kernel on a CTA pair: tcgen05.mma.cta_group::2 (M = 256 over two SMs), each CTA stages
//                          half of the weight tile; default for the 256-wide tiles
// and the filter gradient runs as wgrad_tc_kernel (one 128 x 128 tile per CTA, split-K) or wgrad_tc_pair_kernel (256 x 256 per pair).
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;       // pixels per CTA tile (UMMA M)
constexpr int BK = 64;        // bf16 channels per K block = one 128-byte swizzle row
constexpr int LAG = 2;        // cp.async groups in flight per producer thread
constexpr int MAX_TAPS = 25;

struct TcParams {
  const bf16* src;            // gathered operand (x for fprop, dy for dgrad), NHWC with channel stride ld_src
  int SH, SW, ld_src, cvalid; // source image dims; channels readable as 16-byte chunks (multiple of 8, <= ld_src)
  int MH, MW, M;              // GEMM rows: m -> (n, a, b), a < MH, b < MW
  int by_mul, by_add, bx_mul, bx_add;   // source base coordinate of a row: (a*by_mul + by_add, b*bx_mul + bx_add)
  int ntaps, kb_per_tap;
  short tdy[MAX_TAPS], tdx[MAX_TAPS], twi[MAX_TAPS];   // per tap: source offset, weight-pack tap index
  void* out;                  // output NHWC
  int out_f32, ld_out, OH, OW, oy_mul, oy_add, ox_mul, ox_add, N;
  const float* bias;
  int act;
  float leak;
  int accumulate;
  const void* res;            // fused residual: out = act(conv + bias) + res, res laid out exactly like out (NULL = none)
  // further epilogue fusions (rcgan_conv_epilogue in the header; bf16 outputs only)
  int res_up, ld_res;         // res is [n, OH/2, OW/2, N] with channel stride ld_res, read through a nearest-neighbour 2x upsampling
  const void* mask;           // out = value * act'(mask): the backward of the activation whose OUTPUT is mask (laid out like out)
  int mask_act;
  float mask_leak;
  void* out2;                 // second output act2(final value), laid out like out
  int out2_act;
  float* colstats;            // persistent kernel: per-CTA partial column sums / sums of squares of the stored output (batch-norm statistics)
  // TMA im2col A loader (one cp.async.bulk.tensor.im2col per K block instead of 1024 cp.async):
  // base pixel of GEMM row (n,a,b) = (im_h_lo + a*im_sh, im_w_lo + b*im_sw); tap t adds (toffh[t], toffw[t])
  int im_w_lo, im_h_lo, im_sw, im_sh;
  unsigned short toffw[MAX_TAPS], toffh[MAX_TAPS];
  int bn_eff;   // persistent kernel: N-tile width actually computed (multiple of 16, <= the template's BN)
  int dbg;   // experiments only: 1 = skip MMA issue, 2 = skip TMA loads, 4 = skip A load, 8 = skip B load
};

// up to 4 problems that share operands, N and the epilogue but differ in geometry / taps (the 4 output parity classes of a
// stride-2 transposed conv), scheduled as one persistent launch
struct TcMulti {
  TcParams p[4];
  int tile_start[5];          // first (TM x bn) tile of each problem; tile_start[nprob] = total
  int nprob;
};
struct TcMaps { CUtensorMap a[4]; };

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded spin: a protocol bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it)
    if (it > (1u << 24)) __trap();
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int n,
                                                   unsigned short off_w, unsigned short off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], "
      "{%7, %8};" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: LBO = 1 (unused), SBO = 1024 B (8 rows x 128 B), version 1
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                             // leading byte offset (>>4), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset (>>4), bits [32,46)
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                             // layout type SWIZZLE_128B
  return d;
}
// instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- CTA pair (cta_group::2): two CTAs of a 2-cluster on one TPC run ONE M = 256 UMMA; each stages its own 128 A rows and its
// own half of the B (weight) columns, the tensor cores of both SMs read both halves.  Per SM that is 4 KB (A) + N/2 x 32 B (B)
// of shared-memory operand reads per K = 16 step instead of 4 KB + N x 32 B, and half the weight bytes through TMA.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (release at CTA scope), as CUTLASS's ClusterBarrier::arrive(cta_id): the waiter only needs this warp's
  // tcgen05.ld's to have completed, which tcgen05.wait::ld + fence::before_thread_sync in program order guarantee; a
  // .release.cluster here costs a MEMBAR per tile (8 % of the epilogue warps' samples in ncu)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a barrier of THIS CTA that threads of the peer CTA arrive on
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  for (uint32_t it = 0;; ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (it > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair once all prior MMAs of this thread retire
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
// TMA loads of a CTA pair: the transaction bytes are counted on the LEADER's (rank 0) barrier -- the peer bit (bit 24) of the
// shared::cluster barrier address is cleared, as CUTLASS's SM100_TMA_2SM_LOAD does
constexpr uint32_t PAIR_PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar & PAIR_PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int n,
                                                        unsigned short off_w, unsigned short off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2], {%7, %8};" ::"r"(dst),
      "l"(map), "r"(bar & PAIR_PEER_BIT_MASK), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}

template <int BN, int ST> struct Cfg {
  static constexpr int STAGES = ST;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Epilogue of one 128 x BN accumulator tile.  A thread owns one TMEM lane = one GEMM row, so writing straight from
// registers scatters 16-byte pieces over 32 different output rows per store instruction (measured: 400 of 508 us of a
// 1x1 256->256 conv at 512x32x32 were those stores).  Instead each warp stages its 32 rows in shared memory (row pitch
// padded by 16 B: conflict-free 16-byte accesses) and then writes whole rows, consecutive lanes on consecutive 16-byte
// pieces.  Only the owning warp touches its slab, so __syncwarp() orders the two phases.
template <int BN, int AVAIL, typename TO, int U = 4>
__device__ __forceinline__ void tc_epilogue(const TcParams& p, uint8_t* smem, uint32_t tmem_base, int warp, int lane, int m0,
                                            int n0, uint32_t tempty_bar = 0, int bn_lim = BN, float* st_acc = nullptr,
                                            int* st_cnt = nullptr, bool tempty_cluster = false) {
  constexpr int VEC = 16 / (int)sizeof(TO);
  constexpr int PITCH = BN * (int)sizeof(TO) + 16;
  static_assert(4 * 32 * PITCH + 2048 <= AVAIL, "staging must fit the drained pipeline buffers");
  uint8_t* slab = smem + warp * (32 * PITCH);
  unsigned long long* rowoff = reinterpret_cast<unsigned long long*>(smem + 4 * 32 * PITCH) + warp * 32;
  unsigned long long* rowoff_res = rowoff + 4 * 32;      // offsets into an upsampled-on-the-fly residual (res_up)
  const int m = m0 + warp * 32 + lane;
  unsigned long long ob = ~0ull, orr = 0;
  if (m < p.M) {
    int b = m % p.MW, r = m / p.MW, a = r % p.MH, n = r / p.MH;
    const int oy = a * p.oy_mul + p.oy_add, ox = b * p.ox_mul + p.ox_add;
    ob = ((unsigned long long)(n * p.OH + oy) * p.OW + ox) * p.ld_out;
    if (p.res_up) orr = ((unsigned long long)(n * (p.OH >> 1) + (oy >> 1)) * (p.OW >> 1) + (ox >> 1)) * p.ld_res;
  }
  rowoff[lane] = ob;
  if (p.res_up) rowoff_res[lane] = orr;
  const int ncols = min(bn_lim, p.N - n0);
#pragma unroll 1
  for (int c0 = 0; c0 < bn_lim; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);   // warp-collective: all lanes execute
    if (c0 >= ncols || (p.dbg & 16)) continue;
    // bias and activation with the branches hoisted out of the element loops: the 4 epilogue warps run one per
    // scheduler, so every per-element branch/constant load is exposed latency (ncu: 25 instr/element before, ~3 now)
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
    if (p.bias) {
      const float* bp = p.bias + n0 + c0;
      if (c0 + 32 <= ncols && (reinterpret_cast<uintptr_t>(bp) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + j));
          v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; j++)
          if (c0 + j < ncols) v[j] += bp[j];
      }
    }
    switch (p.act) {
      case RCGAN_ACT_RELU:
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], 0.f);
        break;
      case RCGAN_ACT_LRELU: {
        const float leak = p.leak;
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], leak * v[j]);
        break;
      }
      case RCGAN_ACT_SIGMOID:
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = 1.f / (1.f + expf(-v[j]));
        break;
      case RCGAN_ACT_TANH:
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = tanhf(v[j]);
        break;
      default: break;
    }
    uint8_t* dst = slab + lane * PITCH + c0 * (int)sizeof(TO);
#pragma unroll
    for (int j = 0; j < 32; j += VEC) {
      uint4 q;
      if (sizeof(TO) == 4) {
        q = make_uint4(__float_as_uint(v[j]), __float_as_uint(v[j + 1]), __float_as_uint(v[j + 2]), __float_as_uint(v[j + 3]));
      } else {
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
        for (int e = 0; e < 4; e++) h[e] = __floats2bfloat162_rn(v[(j + 2 * e) % 32], v[(j + 2 * e + 1) % 32]);
      }
      *reinterpret_cast<uint4*>(dst + (j / VEC) * 16) = q;
    }
  }
  // persistent kernel: the accumulator is drained -> hand it back to the MMA warp before the (slow) global stores
  if (tempty_bar) {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (tempty_cluster) mbar_arrive_cluster(tempty_bar);     // CTA pair: the leader's barrier, a shared::cluster address
      else mbar_arrive(tempty_bar);
    }
  }
  __syncwarp();
  if (p.dbg & 16) return;
  TO* out = reinterpret_cast<TO*>(p.out);
  // the tile is added onto `addsrc` (same layout as out): out itself when accumulating, the residual input when fusing the
  // ResidualBlock's shortcut add (gan_resnet.py:328); both round like the separate add kernel did (bf16 + bf16 in fp32 -> bf16)
  const TO* addsrc = p.accumulate ? out : reinterpret_cast<const TO*>(p.res);
  const bool up = !p.accumulate && p.res && p.res_up;       // residual read through a 2x nearest-neighbour upsampling
  const TO* mask = reinterpret_cast<const TO*>(p.mask);
  TO* out2 = reinterpret_cast<TO*>(p.out2);
  const bool fast = p.ld_out % VEC == 0 && (reinterpret_cast<uintptr_t>(out + n0) & 15) == 0 && ncols >= VEC &&
                    (reinterpret_cast<uintptr_t>(addsrc) & 15) == 0 && (!up || p.ld_res % VEC == 0) &&
                    (reinterpret_cast<uintptr_t>(mask) & 15) == 0 && (reinterpret_cast<uintptr_t>(out2) & 15) == 0;
  if (sizeof(TO) == 2 && (mask || out2 || up || st_acc)) {
    // fused variant of the store loop (bf16 only): mask -> residual / accumulate -> store -> second output
    const float mleak = p.mask_act == RCGAN_ACT_LRELU ? p.mask_leak : 0.f;
    if (fast) {
      // The extra operands (mask, residual / old value) are global loads in the store loop: U pieces per lane are put in flight
      // before the first one is consumed (a dependent load-use chain per piece made this epilogue slower than the tile's main
      // loop: 25.8 vs 24.2 ms per iteration with the separate passes)
      const int lpr = ncols / VEC;
      const int drow = 32 / lpr, dpiece = 32 - drow * lpr;
      int row = lane / lpr, piece = lane - row * lpr;
      const int niter = lpr;                       // idx = lane + 32 * it < 32 * lpr
      for (int it0 = 0; it0 < niter; it0 += U) {
        uint4 q[U], mq[U], aq[U];
        unsigned long long off[U];
        bool live[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          live[u] = false;
          if (it0 + u < niter) {
            if (piece >= lpr) { piece -= lpr; row++; }
            const unsigned long long o = rowoff[row];
            if (o != ~0ull) {
              live[u] = true;
              const size_t col = (size_t)n0 + piece * VEC;
              off[u] = o + col;
              if (mask) mq[u] = __ldg(reinterpret_cast<const uint4*>(mask + o + col));
              if (addsrc) aq[u] = *reinterpret_cast<const uint4*>(addsrc + (up ? rowoff_res[row] : o) + col);
              q[u] = *reinterpret_cast<const uint4*>(slab + row * PITCH + piece * 16);
            }
            row += drow; piece += dpiece;
          }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          if (!live[u]) continue;
          __nv_bfloat162* c = reinterpret_cast<__nv_bfloat162*>(&q[u]);
          // fp32 throughout, ONE rounding at the end: exactly what rcgan_act_bwd (dx (=|+=) dy * act'(y)) and rcgan_add compute
          const __nv_bfloat162* mm = reinterpret_cast<const __nv_bfloat162*>(&mq[u]);
          const __nv_bfloat162* a = reinterpret_cast<const __nv_bfloat162*>(&aq[u]);
#pragma unroll
          for (int e = 0; e < 4; e++) {
            float2 fc = __bfloat1622float2(c[e]);
            float2 sl = make_float2(1.f, 1.f);
            if (mask) {
              const float2 fm = __bfloat1622float2(mm[e]);
              sl.x = fm.x > 0.f ? 1.f : mleak; sl.y = fm.y > 0.f ? 1.f : mleak;
            }
            if (addsrc) {       // explicit fma / mul: the same operation rcgan_act_bwd performs (bit-identical results)
              const float2 fa = __bfloat1622float2(a[e]);
              fc.x = __fmaf_rn(fc.x, sl.x, fa.x); fc.y = __fmaf_rn(fc.y, sl.y, fa.y);
            } else {
              fc.x = __fmul_rn(fc.x, sl.x); fc.y = __fmul_rn(fc.y, sl.y);
            }
            c[e] = __floats2bfloat162_rn(fc.x, fc.y);
          }
          *reinterpret_cast<uint4*>(out + off[u]) = q[u];
          if (st_acc) {
            // batch-norm statistics of the tensor being stored (the bf16 values a separate statistics pass would read back):
            // a lane owns the same 8 columns in every row it stores (ncols / 8 is a power of two <= 32), so the running
            // sums stay in registers across all tiles of this persistent CTA
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const float2 f = __bfloat1622float2(c[e]);
              st_acc[2 * e] += f.x; st_acc[2 * e + 1] += f.y;
              st_acc[8 + 2 * e] = fmaf(f.x, f.x, st_acc[8 + 2 * e]); st_acc[8 + 2 * e + 1] = fmaf(f.y, f.y, st_acc[8 + 2 * e + 1]);
            }
            (*st_cnt)++;
          }
          if (out2) {
            const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
            for (int e = 0; e < 4; e++) c[e] = __hmax2(c[e], z);
            *reinterpret_cast<uint4*>(out2 + off[u]) = q[u];
          }
        }
      }
    }
    const int tail0 = fast ? (ncols / VEC) * VEC : 0, ntail = ncols - tail0;
    if (ntail > 0) {
      for (int idx = lane; idx < 32 * ntail; idx += 32) {
        const int row = idx / ntail, cc = tail0 + idx - row * ntail;
        const unsigned long long o = rowoff[row];
        if (o == ~0ull) continue;
        float x = to_f(reinterpret_cast<const TO*>(slab + row * PITCH)[cc]);
        const float sl = (mask && !(to_f(mask[o + n0 + cc]) > 0.f)) ? mleak : 1.f;
        x = addsrc ? __fmaf_rn(x, sl, to_f(addsrc[(up ? rowoff_res[row] : o) + n0 + cc])) : __fmul_rn(x, sl);
        const TO v = from_f<TO>(x);
        out[o + n0 + cc] = v;
        if (out2) out2[o + n0 + cc] = from_f<TO>(fmaxf(to_f(v), 0.f));
      }
    }
    return;
  }
  if (fast) {
    // whole 16-byte pieces of every row, consecutive lanes on consecutive pieces; a ragged tail (ncols % VEC columns,
    // e.g. 138 = 17 pieces + 2) is finished element-wise below
    const int lpr = ncols / VEC;
    // (row, piece) of idx = lane + 32*i advanced incrementally: no division in the loop (the 4 epilogue warps run one per
    // scheduler, every instruction of this loop is exposed latency when K is short)
    const int drow = 32 / lpr, dpiece = 32 - drow * lpr;
    int row = lane / lpr, piece = lane - row * lpr;
#pragma unroll 4
    for (int idx = lane; idx < 32 * lpr; idx += 32, row += drow, piece += dpiece) {
      if (piece >= lpr) { piece -= lpr; row++; }
      const unsigned long long o = rowoff[row];
      if (o == ~0ull) continue;
      uint4 q = *reinterpret_cast<const uint4*>(slab + row * PITCH + piece * 16);
      TO* g = out + o + n0 + piece * VEC;
      if (addsrc) {
        const uint4 old = *reinterpret_cast<const uint4*>(addsrc + o + n0 + piece * VEC);
        if (sizeof(TO) == 4) {
          const float* a = reinterpret_cast<const float*>(&old);
          float* c = reinterpret_cast<float*>(&q);
#pragma unroll
          for (int e = 0; e < 4; e++) c[e] += a[e];
        } else {
          const __nv_bfloat162* a = reinterpret_cast<const __nv_bfloat162*>(&old);
          __nv_bfloat162* c = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const float2 fa = __bfloat1622float2(a[e]), fc = __bfloat1622float2(c[e]);
            c[e] = __floats2bfloat162_rn(fa.x + fc.x, fa.y + fc.y);
          }
        }
      }
      *reinterpret_cast<uint4*>(g) = q;
    }
    const int tail0 = lpr * VEC, ntail = ncols - tail0;
    if (ntail > 0) {
      for (int idx = lane; idx < 32 * ntail; idx += 32) {
        const int row = idx / ntail, c = tail0 + idx - row * ntail;
        const unsigned long long o = rowoff[row];
        if (o == ~0ull) continue;
        TO* g = out + o + n0 + c;
        float x = to_f(reinterpret_cast<const TO*>(slab + row * PITCH)[c]);
        if (addsrc) x += to_f(addsrc[o + n0 + c]);
        *g = from_f<TO>(x);
      }
    }
  } else {
    // ragged tile (N not a multiple of the tile, odd strides): element-wise, lanes along the row
    for (int row = 0; row < 32; row++) {
      const unsigned long long o = rowoff[row];
      if (o == ~0ull) continue;
      const TO* srow = reinterpret_cast<const TO*>(slab + row * PITCH);
      for (int c = lane; c < ncols; c += 32) {
        TO* g = out + o + n0 + c;
        float x = to_f(srow[c]);
        if (addsrc) x += to_f(addsrc[o + n0 + c]);
        *g = from_f<TO>(x);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ kernel
template <int BN, int ST, bool IM2COL>
__global__ void __launch_bounds__(192, 1) conv_tc_kernel(const __grid_constant__ TcParams p,
                                                         const __grid_constant__ CUtensorMap wmap,
                                                         const __grid_constant__ CUtensorMap amap) {
  using C = Cfg<BN, ST>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned; stays a provable __shared__ pointer (LDS/STS, not generic LD/ST)
  uint8_t* smA = smem;
  uint8_t* smB = smem + STAGES * C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + STAGES * C::B_BYTES);
  uint64_t* full = bars;                  // [STAGES]  128 producer arrivals + 1 expect_tx arrival
  uint64_t* empty = bars + STAGES;        // [STAGES]  1 arrival (tcgen05.commit)
  uint64_t* accum_full = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 1-D grid, N tiles of one M tile adjacent: the A rows they share are fetched from DRAM once and hit in L2
  const int n_tiles = (p.N + BN - 1) / BN;
  const int m0 = (blockIdx.x / n_tiles) * BM, n0 = (blockIdx.x % n_tiles) * BN;
  const int nkb = p.ntaps * p.kb_per_tap;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(smem_u32(&full[s]), IM2COL ? 1 : 128 + 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    mbar_init(smem_u32(accum_full), 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(tmem_slot), BN);
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
      if (IM2COL) asm volatile("prefetch.tensormap [%0];" ::"l"(&amap) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();   // PDL: the prologue above overlapped the previous kernel's tail; its results are touched only from here on

  if (warp < 4) {
    // =========================================================== A producers (then epilogue)
    const int t = threadIdx.x;
    if (!IM2COL) {
    const int chunk = t & 7;               // 16-byte chunk within the 128-byte row
    // per handled row: element offset of its base source pixel and a bit mask of the taps that land inside the image
    // (SAME-padding halo, rows past M) -- the K loop then costs one shift/and + one add per 16-byte cp.async
    int roff[8];
    uint32_t rmask[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int m = m0 + (t >> 3) + 16 * i;
      const bool rok = m < p.M;
      int mm = rok ? m : 0;
      int b = mm % p.MW, r = mm / p.MW, a = r % p.MH, n = r / p.MH;
      const int ry = a * p.by_mul + p.by_add, rx = b * p.bx_mul + p.bx_add;
      roff[i] = ((n * p.SH + ry) * p.SW + rx) * p.ld_src + chunk * 8;
      uint32_t msk = 0;
      if (rok)
        for (int tp = 0; tp < p.ntaps; tp++) {
          const int sy = ry + p.tdy[tp], sx = rx + p.tdx[tp];
          if (sy >= 0 && sy < p.SH && sx >= 0 && sx < p.SW) msk |= 1u << tp;
        }
      rmask[i] = msk;
    }
    for (int kb = 0; kb < nkb; kb++) {
      const int s = kb % STAGES;
      mbar_wait(smem_u32(&empty[s]), ((kb / STAGES) & 1) ^ 1);
      const int tap = kb / p.kb_per_tap;
      const int c0 = (kb - tap * p.kb_per_tap) * BK;
      const int toff = (p.tdy[tap] * p.SW + p.tdx[tap]) * p.ld_src + c0;
      const bool chok = c0 + chunk * 8 < p.cvalid;
      const uint32_t abase = smem_u32(smA + s * C::A_BYTES) + ((t >> 3) * 128);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int row = (t >> 3) + 16 * i;
        const bool ok = chok && ((rmask[i] >> tap) & 1u);
        const bf16* src = p.src + (ok ? (ptrdiff_t)roff[i] + toff : 0);
        cp_async16(abase + i * 2048 + ((chunk ^ (row & 7)) << 4), src, ok ? 16u : 0u);
      }
      cp_async_commit();
      if (kb >= LAG) {
        cp_async_wait<LAG>();
        fence_proxy_async();
        mbar_arrive(smem_u32(&full[(kb - LAG) % STAGES]));
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (int kb = (nkb > LAG ? nkb - LAG : 0); kb < nkb; kb++) mbar_arrive(smem_u32(&full[kb % STAGES]));
    }

    // =========================================================== epilogue: TMEM -> registers -> smem -> coalesced global rows
    mbar_wait(smem_u32(accum_full), 0);
    tc_fence_after();
    // every TMA/cp.async load has landed and every MMA has retired: the pipeline buffers are free for staging
    if (p.out_f32) tc_epilogue<BN, C::STAGES * (C::A_BYTES + C::B_BYTES), float>(p, smem, tmem_base, warp, lane, m0, n0);
    else tc_epilogue<BN, C::STAGES * (C::A_BYTES + C::B_BYTES), bf16>(p, smem, tmem_base, warp, lane, m0, n0);
    tc_fence_before();
  } else if (warp == 4) {
    // =========================================================== B (and, with IM2COL, A) producer (TMA)
    if (lane == 0) {
      const int mb = m0 % p.MW, mr = m0 / p.MW;
      const int im_w = p.im_w_lo + mb * p.im_sw, im_h = p.im_h_lo + (mr % p.MH) * p.im_sh, im_n = mr / p.MH;
      for (int kb = 0; kb < nkb; kb++) {
        const int s = kb % STAGES;
        mbar_wait(smem_u32(&empty[s]), ((kb / STAGES) & 1) ^ 1);
        const int tap = kb / p.kb_per_tap;
        const int k0 = (kb - tap * p.kb_per_tap) * BK;
        if (p.dbg & 14) {
          const bool la = IM2COL && !(p.dbg & 6), lb = !(p.dbg & 10);
          mbar_arrive_expect_tx(smem_u32(&full[s]), (lb ? C::B_BYTES : 0) + (la ? C::A_BYTES : 0));
          if (la) tma_load_im2col_4d(smem_u32(smA + s * C::A_BYTES), &amap, smem_u32(&full[s]), k0, im_w, im_h, im_n, p.toffw[tap],
                                     p.toffh[tap]);
          if (lb) tma_load_3d(smem_u32(smB + s * C::B_BYTES), &wmap, smem_u32(&full[s]), k0, n0, p.twi[tap]);
          continue;
        }
        mbar_arrive_expect_tx(smem_u32(&full[s]), C::B_BYTES + (IM2COL ? C::A_BYTES : 0));
        if (IM2COL)
          tma_load_im2col_4d(smem_u32(smA + s * C::A_BYTES), &amap, smem_u32(&full[s]), k0, im_w, im_h, im_n, p.toffw[tap],
                             p.toffh[tap]);
        tma_load_3d(smem_u32(smB + s * C::B_BYTES), &wmap, smem_u32(&full[s]), k0, n0, p.twi[tap]);
      }
    }
  } else {
    // =========================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      for (int kb = 0; kb < nkb; kb++) {
        const int s = kb % STAGES;
        mbar_wait(smem_u32(&full[s]), (kb / STAGES) & 1);
        tc_fence_after();
        const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smA + s * C::A_BYTES));
        const uint64_t db = umma_desc_kmajor_sw128(smem_u32(smB + s * C::B_BYTES));
        if (!(p.dbg & 1)) {
#pragma unroll
        for (int k = 0; k < BK / 16; k++)   // advance 16 bf16 = 32 bytes inside the swizzle row: +2 in the (>>4) address field
          umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
        }
        umma_commit(smem_u32(&empty[s]));   // implies tcgen05.fence::before_thread_sync
      }
      umma_commit(smem_u32(accum_full));
    }
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ------------------------------------------------------------------------------------------------ persistent kernel
// One CTA per SM loops over output tiles (N tiles of one M tile adjacent).  The TMA producer and the MMA issuer run
// ahead across tile boundaries through the same smem ring; the accumulator is double buffered in TMEM (2 x BN columns),
// so the epilogue of tile i (TMEM -> registers -> private staging smem -> global) overlaps the main loop of tile i+1,
// and TMEM allocation / barrier init / descriptor prefetch are paid once per SM instead of once per tile.
// (Measured on the one-tile-per-CTA kernel above at 256x32x32x128 3x3: 89 us, of which 45 us were per-CTA prologue +
// handshake skeleton and 18 us the exposed epilogue.)   TMA im2col operand only.
// The persistent kernels run 10 warps: 0-3 and 6-9 are two epilogue groups (warp w reads TMEM lanes 32*(w%4)..+31, the hardware's
// rule; group 0 owns the left half of the tile's columns, group 1 the right half), warp 4 is the TMA producer, warp 5 issues
// the MMAs.  With one group a 128x256 tile took ~7.7 us to drain (ncu: the four warps, one per scheduler, spend it on exposed
// instruction latency), longer than the main loop of the K = 1024 transposed convs and of every conv with a fused epilogue.
constexpr int EPI_GROUPS = 2;
constexpr int PERSIST_THREADS = 64 + 128 * EPI_GROUPS;

template <int BN>
__device__ __forceinline__ void epilogue_stats_flush(const TcParams& p0, uint8_t* epi, const float* st_acc, int st_cnt, int quarter,
                                                     int eg, int lane, int bn) {
  constexpr int NT = 128 * EPI_GROUPS;
  float* S = reinterpret_cast<float*>(epi);
  int* Cn = reinterpret_cast<int*>(S + 2 * BN);
  const int tid = (eg * 4 + quarter) * 32 + lane, hb = bn >> 1, lpr = hb / 8;      // a lane owns 8 columns of its group's half
  asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
  for (int i = tid; i < 2 * BN + 1; i += NT) S[i] = 0.f;
  asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
  const int col0 = eg * hb + (lane % lpr) * 8;
#pragma unroll
  for (int j = 0; j < 8; j++) { atomicAdd(&S[col0 + j], st_acc[j]); atomicAdd(&S[BN + col0 + j], st_acc[8 + j]); }
  atomicAdd(Cn, st_cnt);
  asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
  float* dst = p0.colstats + (size_t)blockIdx.x * 2 * p0.N;
  for (int i = tid; i < p0.N; i += NT) { dst[i] = S[i]; dst[p0.N + i] = S[BN + i]; }
  float* counts = p0.colstats + (size_t)RCGAN_NUM_SMS * 2 * p0.N;
  if (tid == 0) {
    counts[blockIdx.x] = (float)(*Cn / (bn / 8));       // st_cnt counts 16-byte pieces: bn / 8 per stored row
    for (int j = blockIdx.x + gridDim.x; j < RCGAN_NUM_SMS; j += gridDim.x) counts[j] = 0.f;
  }
}

template <int BN, int MT, int ST, typename TO>
struct PCfg {
  static constexpr int A_BYTES = MT * BM * BK * 2;     // MT stacked 128-row sub-tiles share one B tile
  static constexpr int B_BYTES = BN * BK * 2;
  // two groups of 4 epilogue warps, each draining one half of the tile's columns through its own staging region
  static constexpr int EPI_GROUP_BYTES = 4 * 32 * ((BN / 2) * (int)sizeof(TO) + 16) + 2048;
  static constexpr int EPI_BYTES = EPI_GROUPS * EPI_GROUP_BYTES;
  static constexpr int SMEM = ST * (A_BYTES + B_BYTES) + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(SMEM <= 227 * 1024, "persistent conv tile does not fit shared memory");
  static_assert(2 * MT * BN <= 512, "double-buffered accumulators must fit TMEM");
};

// Tile = (MT x 128) rows x BN columns.  The operand stream from L2 is what bounds this kernel (~50 B/clk/SM measured), so
// the tile shapes are chosen for flops per staged byte: 128x256 (N >= 256) and 256x128 (N <= 128) both move 48 KB per
// 4.2 MFLOP K block, against 32 KB per 2.1 MFLOP for 128x128.
template <int BN, int MT, int ST, typename TO, bool MULTI>
__global__ void __launch_bounds__(PERSIST_THREADS, 1) conv_tc_persist_kernel(const __grid_constant__ TcMulti mp,
                                                                 const __grid_constant__ CUtensorMap wmap,
                                                                 const __grid_constant__ TcMaps amaps) {
  using C = PCfg<BN, MT, ST, TO>;
  constexpr int TM = MT * BM;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned; stays a provable __shared__ pointer (LDS/STS, not generic LD/ST)
  uint8_t* smA = smem;
  uint8_t* smB = smem + ST * C::A_BYTES;
  uint8_t* epi = smB + ST * C::B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + C::EPI_BYTES);
  uint64_t* full = bars;               // [ST] 1 arrival (expect_tx) + TMA bytes
  uint64_t* empty = bars + ST;         // [ST] 1 arrival (tcgen05.commit)
  uint64_t* tfull = bars + 2 * ST;     // [2]  accumulator ready (tcgen05.commit)
  uint64_t* tempty = bars + 2 * ST + 2;   // [2]  accumulator drained (4 * EPI_GROUPS epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * ST + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TcParams& p0 = mp.p[0];                          // N, operands, output, bias/act are common to all problems
  const int bn = p0.bn_eff;                              // N-tile width of this launch: multiple of 16, <= BN
  const int n_tiles_n = (p0.N + bn - 1) / bn;
  const int n_tiles = mp.tile_start[mp.nprob];
  // tile -> (problem q, tile inside q): the stride-2 transposed conv's 4 output parity classes are 4 problems of ONE launch
  // (MULTI = false: one problem, q is the compile-time constant 0 and every parameter access keeps its immediate offset)
  auto locate = [&](int tile, int& q) {
    q = 0;
    if (!MULTI) return tile;
    while (q + 1 < mp.nprob && tile >= mp.tile_start[q + 1]) q++;
    return tile - mp.tile_start[q];
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < ST; s++) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(smem_u32(&tfull[a]), 1);
      mbar_init(smem_u32(&tempty[a]), 4 * EPI_GROUPS);
    }
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(tmem_slot), 2 * MT * BN);
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
      for (int q = 0; q < mp.nprob; q++) asm volatile("prefetch.tensormap [%0];" ::"l"(&amaps.a[q]) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();   // PDL: the prologue above overlapped the previous kernel's tail; its results are touched only from here on

  if (warp < 4 || warp >= 6) {
    // =========================================================== epilogue warps: group eg drains columns [eg * hb, +hb) of the tile
    const int quarter = warp & 3, eg = warp >= 6 ? 1 : 0, hb = bn >> 1;
    uint8_t* my_epi = epi + eg * C::EPI_GROUP_BYTES;
    int ti = 0;
    float st_acc[16];
    int st_cnt = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) st_acc[j] = 0.f;
    const bool stats = sizeof(TO) == 2 && p0.colstats != nullptr;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ti++) {
      const int a = ti & 1;
      int q;
      const int lt = locate(tile, q);
      const TcParams& p = mp.p[q];
      const int m0 = (lt / n_tiles_n) * TM, n0 = (lt % n_tiles_n) * bn + eg * hb;
      mbar_wait(smem_u32(&tfull[a]), (ti >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < MT; j++)
        tc_epilogue<BN / 2, C::EPI_GROUP_BYTES, TO, 4>(p, my_epi, tmem_base + (uint32_t)((a * MT + j) * BN + eg * hb), quarter, lane,
                                                       m0 + j * BM, n0, j == MT - 1 ? smem_u32(&tempty[a]) : 0u, hb,
                                                       stats ? st_acc : nullptr, &st_cnt);
    }
    if (stats) epilogue_stats_flush<BN>(p0, epi, st_acc, st_cnt, quarter, eg, lane, bn);
    tc_fence_before();
  } else if (warp == 4) {
    // =========================================================== TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int q;
        const int lt = locate(tile, q);
        const TcParams& p = mp.p[q];
        const CUtensorMap* amap = &amaps.a[q];
        const int nkb = p.ntaps * p.kb_per_tap;
        const int m0 = (lt / n_tiles_n) * TM, n0 = (lt % n_tiles_n) * bn;
        int im_w[MT], im_h[MT], im_n[MT];
        int nsub = 0;
#pragma unroll
        for (int j = 0; j < MT; j++) {
          const int mj = m0 + j * BM;
          if (mj < p.M) nsub = j + 1;
          const int mb = mj % p.MW, mr = mj / p.MW;
          im_w[j] = p.im_w_lo + mb * p.im_sw; im_h[j] = p.im_h_lo + (mr % p.MH) * p.im_sh; im_n[j] = mr / p.MH;
        }
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int s = it % ST;
          mbar_wait(smem_u32(&empty[s]), ((it / ST) & 1) ^ 1);
          const int tap = kb / p.kb_per_tap;
          const int k0 = (kb - tap * p.kb_per_tap) * BK;
          const bool la = !(p.dbg & 6), lb = !(p.dbg & 10);
          mbar_arrive_expect_tx(smem_u32(&full[s]), (lb ? bn * BK * 2 : 0) + (la ? nsub * (BM * BK * 2) : 0));
          if (la) {
#pragma unroll
            for (int j = 0; j < MT; j++)
              if (j < nsub)      // sub-tiles past the last row are not loaded (their accumulator rows are never stored)
                tma_load_im2col_4d(smem_u32(smA + s * C::A_BYTES + j * (BM * BK * 2)), amap, smem_u32(&full[s]), k0, im_w[j],
                                   im_h[j], im_n[j], p.toffw[tap], p.toffh[tap]);
          }
          if (lb) tma_load_3d(smem_u32(smB + s * C::B_BYTES), &wmap, smem_u32(&full[s]), k0, n0, p.twi[tap]);
        }
      }
    }
  } else {
    // =========================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BM, bn);
      uint32_t it = 0;
      int ti = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ti++) {
        const int a = ti & 1;
        int q;
        locate(tile, q);
        const TcParams& p = mp.p[q];
        const int nkb = p.ntaps * p.kb_per_tap;
        mbar_wait(smem_u32(&tempty[a]), ((ti >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(a * MT * BN);
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int s = it % ST;
          mbar_wait(smem_u32(&full[s]), (it / ST) & 1);
          tc_fence_after();
          const uint64_t db = umma_desc_kmajor_sw128(smem_u32(smB + s * C::B_BYTES));
          if (!(p.dbg & 1)) {
#pragma unroll
            for (int j = 0; j < MT; j++) {
              const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smA + s * C::A_BYTES + j * (BM * BK * 2)));
#pragma unroll
              for (int k = 0; k < BK / 16; k++) umma_bf16(tacc + (uint32_t)(j * BN), da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            }
          }
          umma_commit(smem_u32(&empty[s]));
        }
        umma_commit(smem_u32(&tfull[a]));
      }
    }
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * MT * BN);
  }
}

// ------------------------------------------------------------------------------------------------ CTA-pair persistent kernel
// The persistent kernel above with cta_group::2: a cluster of two CTAs (one TPC) owns a (2 x MT x 128) x bn tile.  CTA r of the pair
// stages rows [m0 + r*MT*128, +MT*128) of A and columns [n0 + r*bn/2, +bn/2) of B; the leader's (rank 0) elected thread issues
// M = 256 UMMAs that read both CTAs' shared memory and write 128 accumulator lanes into each CTA's TMEM; every CTA drains its own
// lanes with the same epilogue as above.  Barriers: full[s] lives on the leader (both CTAs' TMA bytes are counted there), the
// commit that frees a stage / publishes an accumulator is multicast to both CTAs, tempty[a] lives on the leader and collects
// the 8 epilogue warps of the pair.
template <int BN, int MT, int ST, typename TO>
struct PairCfg {
  static constexpr int A_BYTES = MT * BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;    // this CTA's half of the weight tile
  static constexpr int EPI_GROUP_BYTES = 4 * 32 * ((BN / 2) * (int)sizeof(TO) + 16) + 2048;
  static constexpr int EPI_BYTES = EPI_GROUPS * EPI_GROUP_BYTES;
  static constexpr int SMEM = ST * (A_BYTES + B_BYTES) + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(SMEM <= 227 * 1024, "pair conv tile does not fit shared memory");
  static_assert(2 * MT * BN <= 512, "double-buffered accumulators must fit TMEM");
};

template <int BN, int MT, int ST, typename TO, bool MULTI>
__global__ void __launch_bounds__(PERSIST_THREADS, 1) conv_tc_pair_kernel(const __grid_constant__ TcMulti mp,
                                                              const __grid_constant__ CUtensorMap wmap,
                                                              const __grid_constant__ TcMaps amaps) {
  using C = PairCfg<BN, MT, ST, TO>;
  constexpr int TM = MT * BM;          // rows per CTA; the pair's tile has 2 * TM
  extern __shared__ uint8_t smem_raw[];
  // the dynamic shared window starts at the same offset in both CTAs of the pair, so every carved address below is the same
  // shared::cta offset in both (the UMMA descriptors, the multicast commits and tcgen05.alloc rely on that)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned; stays a provable __shared__ pointer (LDS/STS, not generic LD/ST)
  uint8_t* smA = smem;
  uint8_t* smB = smem + ST * C::A_BYTES;
  uint8_t* epi = smB + ST * C::B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + C::EPI_BYTES);
  uint64_t* full = bars;               // [ST] leader: 1 arrival (expect_tx) + both CTAs' TMA bytes
  uint64_t* empty = bars + ST;         // [ST] each CTA: 1 arrival (multicast tcgen05.commit)
  uint64_t* tfull = bars + 2 * ST;     // [2]  each CTA: accumulator ready (multicast tcgen05.commit)
  uint64_t* tempty = bars + 2 * ST + 2;   // [2]  leader: accumulator drained (4 * EPI_GROUPS epilogue warps of each CTA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * ST + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const TcParams& p0 = mp.p[0];
  const int bn = p0.bn_eff;                              // N-tile width: multiple of 32, <= BN
  const int n_tiles_n = (p0.N + bn - 1) / bn;
  const int n_tiles = mp.tile_start[mp.nprob];
  auto locate = [&](int tile, int& q) {
    q = 0;
    if (!MULTI) return tile;
    while (q + 1 < mp.nprob && tile >= mp.tile_start[q + 1]) q++;
    return tile - mp.tile_start[q];
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < ST; s++) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(smem_u32(&tfull[a]), 1);
      mbar_init(smem_u32(&tempty[a]), 2 * 4 * EPI_GROUPS);
    }
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc_pair(smem_u32(tmem_slot), 2 * MT * BN);   // the same warp of both CTAs, same destination offset
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
      for (int q = 0; q < mp.nprob; q++) asm volatile("prefetch.tensormap [%0];" ::"l"(&amaps.a[q]) : "memory");
    }
  }
  tc_fence_before();
  cluster_sync_all();      // the peer's barriers are initialised before anything (TMA bytes, commits, arrives) can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();

  if (warp < 4 || warp >= 6) {
    // =========================================================== epilogue warps (both CTAs, own 128 x MT rows; two column groups)
    const int quarter = warp & 3, eg = warp >= 6 ? 1 : 0, hb = bn >> 1;
    uint8_t* my_epi = epi + eg * C::EPI_GROUP_BYTES;
    int ti = 0;
    float st_acc[16];
    int st_cnt = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) st_acc[j] = 0.f;
    const bool stats = sizeof(TO) == 2 && p0.colstats != nullptr;
    const uint32_t tempty_leader0 = mapa_u32(smem_u32(&tempty[0]), 0), tempty_leader1 = mapa_u32(smem_u32(&tempty[1]), 0);
    for (int tile = pair; tile < n_tiles; tile += npairs, ti++) {
      const int a = ti & 1;
      int q;
      const int lt = locate(tile, q);
      const TcParams& p = mp.p[q];
      const int m0 = (lt / n_tiles_n) * (2 * TM) + (int)rank * TM, n0 = (lt % n_tiles_n) * bn + eg * hb;
      mbar_wait(smem_u32(&tfull[a]), (ti >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < MT; j++)
        tc_epilogue<BN / 2, C::EPI_GROUP_BYTES, TO, 4>(p, my_epi, tmem_base + (uint32_t)((a * MT + j) * BN + eg * hb), quarter, lane,
                                                       m0 + j * BM, n0, j == MT - 1 ? (a ? tempty_leader1 : tempty_leader0) : 0u, hb,
                                                       stats ? st_acc : nullptr, &st_cnt, true);
    }
    if (stats) epilogue_stats_flush<BN>(p0, epi, st_acc, st_cnt, quarter, eg, lane, bn);
    tc_fence_before();
  } else if (warp == 4) {
    // =========================================================== TMA producer (both CTAs: own A rows, own half of B)
    if (lane == 0) {
      uint32_t it = 0;
      const int hb = bn >> 1;
      for (int tile = pair; tile < n_tiles; tile += npairs) {
        int q;
        const int lt = locate(tile, q);
        const TcParams& p = mp.p[q];
        const CUtensorMap* amap = &amaps.a[q];
        const int nkb = p.ntaps * p.kb_per_tap;
        const int mpair = (lt / n_tiles_n) * (2 * TM), n0 = (lt % n_tiles_n) * bn;
        const int m0 = mpair + (int)rank * TM;
        int im_w[MT], im_h[MT], im_n[MT];
        int nsub = 0, nsub_pair = 0;         // sub-tiles with rows < M: of this CTA, of both CTAs
#pragma unroll
        for (int j = 0; j < MT; j++) {
          const int mj = m0 + j * BM;
          if (mj < p.M) nsub = j + 1;
          const int mb = mj % p.MW, mr = mj / p.MW;
          im_w[j] = p.im_w_lo + mb * p.im_sw; im_h[j] = p.im_h_lo + (mr % p.MH) * p.im_sh; im_n[j] = mr / p.MH;
        }
#pragma unroll
        for (int j = 0; j < 2 * MT; j++)
          if (mpair + j * BM < p.M) nsub_pair++;
        const uint32_t tx = (uint32_t)(bn * BK * 2 + nsub_pair * (BM * BK * 2));
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int s = it % ST;
          mbar_wait(smem_u32(&empty[s]), ((it / ST) & 1) ^ 1);
          const int tap = kb / p.kb_per_tap;
          const int k0 = (kb - tap * p.kb_per_tap) * BK;
          if (rank == 0) mbar_arrive_expect_tx(smem_u32(&full[s]), tx);
#pragma unroll
          for (int j = 0; j < MT; j++)
            if (j < nsub)
              tma_load_im2col_4d_pair(smem_u32(smA + s * C::A_BYTES + j * (BM * BK * 2)), amap, smem_u32(&full[s]), k0, im_w[j],
                                      im_h[j], im_n[j], p.toffw[tap], p.toffh[tap]);
          tma_load_3d_pair(smem_u32(smB + s * C::B_BYTES), &wmap, smem_u32(&full[s]), k0, n0 + (int)rank * hb, p.twi[tap]);
        }
      }
    }
  } else {
    // =========================================================== MMA issuer (leader CTA only)
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(2 * BM, bn);
      uint32_t it = 0;
      int ti = 0;
      for (int tile = pair; tile < n_tiles; tile += npairs, ti++) {
        const int a = ti & 1;
        int q;
        locate(tile, q);
        const TcParams& p = mp.p[q];
        const int nkb = p.ntaps * p.kb_per_tap;
        mbar_wait_cluster(smem_u32(&tempty[a]), ((ti >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(a * MT * BN);
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int s = it % ST;
          mbar_wait(smem_u32(&full[s]), (it / ST) & 1);
          tc_fence_after();
          const uint64_t db = umma_desc_kmajor_sw128(smem_u32(smB + s * C::B_BYTES));
#pragma unroll
          for (int j = 0; j < MT; j++) {
            const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smA + s * C::A_BYTES + j * (BM * BK * 2)));
#pragma unroll
            for (int k = 0; k < BK / 16; k++)
              umma_bf16_pair(tacc + (uint32_t)(j * BN), da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit_pair(smem_u32(&empty[s]));
        }
        umma_commit_pair(smem_u32(&tfull[a]));
      }
    }
  }
  // neither CTA may leave (or free its TMEM) while the other can still read its shared memory / write its accumulator lanes
  tc_fence_before();
  __syncthreads();         // reconverges the single-lane roles before the .aligned cluster barrier
  cluster_sync_all();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 2 * MT * BN);
  }
}

// ------------------------------------------------------------------------------------------------ wgrad
// dW[tap][ci][co] = sum over output pixels m of x[pix(m, tap)][ci] * dy[m][co]: the contraction runs over PIXELS, which
// is the strided dimension of both NHWC operands, so both tiles are MN-major for the UMMA (a_major = b_major = 1):
//   A: [K = 128 pixels][MN = 64 channels] per unit, unit = (filter tap, 64-channel block of cin); a CTA tile stacks
//      TWO units into M = 128 (for cin = 64 that is two taps side by side, for cin = 128 the two halves of one tap);
//      gathered with cp.async exactly like the fprop A tile (the swizzle is address based, so the same physical
//      layout is the canonical MN-major SWIZZLE_128B atom: 64 channels contiguous, 8 pixel rows per 1024 B group)
//   B: [K = 128 pixels][MN = 64 channels of dy] per 64-column box, plain 2-D TMA over dy (rows are consecutive pixels)
// split-K over pixel blocks across gridDim.z, fp32 partial sums are red.global.add'ed into dW.
struct WgParams {
  const bf16* x;
  int H, W, ldx, cvalid;
  int HO, WO, HW, Mpix;
  int stride, pad_t, pad_l, kw;
  int cin, cout, cblks, units;
  int kb_total, kb_per_split;
  float* dw;
};

__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // distance between 64-element MN atoms
  d |= (uint64_t)(1024 >> 4) << 32;                   // distance between 8-row K groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar & PAIR_PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}

template <int BN> struct WgCfg {
  static constexpr int STAGES = BN == 64 ? 4 : 3;
  static constexpr int UNIT_BYTES = 128 * 128;                 // 128 pixel rows x 128 B
  static constexpr int A_BYTES = 2 * UNIT_BYTES;
  static constexpr int B_BYTES = (BN / 64) * UNIT_BYTES;
  static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + 1024 + 256 + 2048 /*pixel lut*/;
};

template <int BN, bool IM2COL>
__global__ void __launch_bounds__(192, 1) wgrad_tc_kernel(const __grid_constant__ WgParams p,
                                                          const __grid_constant__ CUtensorMap dymap,
                                                          const __grid_constant__ CUtensorMap xmap) {
  using C = WgCfg<BN>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned; stays a provable __shared__ pointer (LDS/STS, not generic LD/ST)
  uint8_t* smA = smem;
  uint8_t* smB = smem + STAGES * C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + STAGES * C::B_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* accum_full = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  uint16_t* lut = reinterpret_cast<uint16_t*>(bars + 2 * STAGES + 2);   // pixel-in-image -> (oy << 8 | ox), HW <= 1024

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * 2;                 // first of the two stacked units
  const int n0 = blockIdx.y * BN;
  const int kb_beg = blockIdx.z * p.kb_per_split;
  const int kb_end = min(p.kb_total, kb_beg + p.kb_per_split);
  const int nkb = kb_end - kb_beg;

  for (int i = threadIdx.x; i < p.HW; i += 192) lut[i] = (uint16_t)(((i / p.WO) << 8) | (i % p.WO));
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(smem_u32(&full[s]), IM2COL ? 1 : 128 + 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    mbar_init(smem_u32(accum_full), 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(tmem_slot), BN);
    if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&dymap) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();   // PDL: the prologue above overlapped the previous kernel's tail; its results are touched only from here on
  if (nkb <= 0) {            // empty split (cannot happen with the host's split choice; keep teardown well-formed)
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, BN);
    return;
  }

  if (warp < 4) {
    const int t = threadIdx.x;
    if (!IM2COL) {
    const int chunk = t & 7;
    // the two units of this tile: tap offsets and channel bases
    int udy[2], udx[2], uch[2];
    bool uok[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
      int u = u0 + j;
      uok[j] = u < p.units;
      int uu = uok[j] ? u : 0;
      int tap = uu / p.cblks, cb = uu - tap * p.cblks;
      udy[j] = tap / p.kw - p.pad_t;
      udx[j] = tap % p.kw - p.pad_l;
      uch[j] = cb * 64 + chunk * 8;
      uok[j] = uok[j] && uch[j] < p.cvalid;
    }
    // per handled row: image index and pixel-in-image index of pixel m = kb*128 + row, advanced by 128 per K block
    int rn[8], rpi[8];
    const int adv_n = 128 / p.HW, adv_pi = 128 % p.HW;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      long m = (long)kb_beg * 128 + (t >> 3) + 16 * i;
      rn[i] = (int)(m / p.HW);
      rpi[i] = (int)(m - (long)rn[i] * p.HW);
    }
    for (int kk = 0; kk < nkb; kk++) {
      const int s = kk % STAGES;
      mbar_wait(smem_u32(&empty[s]), ((kk / STAGES) & 1) ^ 1);
      const uint32_t abase = smem_u32(smA + s * C::A_BYTES);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int row = (t >> 3) + 16 * i;
        const bool rowok = ((long)rn[i] * p.HW + rpi[i]) < p.Mpix;
        const int e = lut[rowok ? rpi[i] : 0];
        const int oy = e >> 8, ox = e & 255;
        const uint32_t soff = row * 128 + ((chunk ^ (row & 7)) << 4);
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const int iy = oy * p.stride + udy[j], ix = ox * p.stride + udx[j];
          const bool ok = uok[j] && rowok && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
          const bf16* src = ok ? p.x + ((size_t)((rn[i] * p.H + iy) * p.W + ix) * p.ldx + uch[j]) : p.x;
          cp_async16(abase + j * C::UNIT_BYTES + soff, src, ok ? 16u : 0u);
        }
        rn[i] += adv_n; rpi[i] += adv_pi;
        if (rpi[i] >= p.HW) { rpi[i] -= p.HW; rn[i] += 1; }
      }
      cp_async_commit();
      if (kk >= LAG) {
        cp_async_wait<LAG>();
        fence_proxy_async();
        mbar_arrive(smem_u32(&full[(kk - LAG) % STAGES]));
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (int kk = (nkb > LAG ? nkb - LAG : 0); kk < nkb; kk++) mbar_arrive(smem_u32(&full[kk % STAGES]));
    }

    // ---- epilogue: accumulator row r = unit (r / 64), channel (r % 64).  A thread owns a row, so reducing straight from
    // registers would issue one L2 atomic per element with 32 different lines per instruction; instead each warp stages
    // its 32 rows in the drained pipeline buffers and reduces whole rows: 32 lanes x 16 bytes = one 512-byte
    // red.global.add.v4.f32 per row (scalar, still lane-contiguous, when cout or the base is not 16-byte friendly).
    mbar_wait(smem_u32(accum_full), 0);
    tc_fence_after();
    constexpr int PITCH = BN * 4 + 16;
    static_assert(4 * 32 * PITCH <= STAGES * (C::A_BYTES + C::B_BYTES), "staging must fit the drained pipeline buffers");
    uint8_t* slab = smem + warp * (32 * PITCH);
    const int ncols = min(BN, p.cout - n0);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      if (c0 >= ncols) continue;
      uint4* dst = reinterpret_cast<uint4*>(slab + lane * PITCH + c0 * 4);
#pragma unroll
      for (int j = 0; j < 8; j++) dst[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    __syncwarp();
    // the warp's 32 rows belong to one unit: rows (tap, cb*64 + (warp&1)*32 + i), i = 0..31
    const int u = u0 + (warp >> 1);
    if (u < p.units) {
      const int tap = u / p.cblks;
      const int ci0 = (u - tap * p.cblks) * 64 + (warp & 1) * 32;
      float* obase = p.dw + ((size_t)tap * p.cin + ci0) * p.cout + n0;
      const int nrows = min(32, p.cin - ci0);
      const bool vec = (ncols % 4 == 0) && (p.cout % 4 == 0) && ((reinterpret_cast<uintptr_t>(obase) & 15) == 0);
      if (vec) {
        const int lpr = ncols / 4;                    // 16-byte pieces per row
        for (int idx = lane; idx < nrows * lpr; idx += 32) {
          const int row = idx / lpr, piece = idx - row * lpr;
          const float4 q = *reinterpret_cast<const float4*>(slab + row * PITCH + piece * 16);
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(obase + (size_t)row * p.cout + piece * 4), "f"(q.x),
                       "f"(q.y), "f"(q.z), "f"(q.w)
                       : "memory");
        }
      } else {
        for (int row = 0; row < nrows; row++) {
          const float* srow = reinterpret_cast<const float*>(slab + row * PITCH);
          for (int c = lane; c < ncols; c += 32) atomicAdd(obase + (size_t)row * p.cout + c, srow[c]);
        }
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    if (lane == 0) {
      // IM2COL: the x operand (two (tap, 64-channel) units) also comes from the TMA, pixel block by pixel block
      int utap[2], uc0[2];
      bool uok[2];
      for (int j = 0; j < 2; j++) {
        const int u = u0 + j;
        uok[j] = u < p.units;
        const int uu = uok[j] ? u : 0;
        utap[j] = uu / p.cblks;
        uc0[j] = (uu - utap[j] * p.cblks) * 64;
      }
      for (int kk = 0; kk < nkb; kk++) {
        const int s = kk % STAGES;
        mbar_wait(smem_u32(&empty[s]), ((kk / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(smem_u32(&full[s]), C::B_BYTES + (IM2COL ? C::A_BYTES : 0));
        if (IM2COL) {
          const long mpix = (long)(kb_beg + kk) * 128;
          const int n = (int)(mpix / p.HW), pi = (int)(mpix - (long)n * p.HW);
          const int oy = pi / p.WO, ox = pi - oy * p.WO;
#pragma unroll
          for (int j = 0; j < 2; j++) {
            // a unit past the end of the filter loads channel block `cin` (fully out of bounds -> zeros)
            tma_load_im2col_4d(smem_u32(smA + s * C::A_BYTES + j * C::UNIT_BYTES), &xmap, smem_u32(&full[s]),
                               uok[j] ? uc0[j] : p.cblks * 64, ox * p.stride - p.pad_l, oy * p.stride - p.pad_t, n,
                               (unsigned short)(utap[j] % p.kw), (unsigned short)(utap[j] / p.kw));
          }
        }
#pragma unroll
        for (int j = 0; j < BN / 64; j++)
          tma_load_2d(smem_u32(smB + s * C::B_BYTES + j * C::UNIT_BYTES), &dymap, smem_u32(&full[s]), n0 + j * 64,
                      (kb_beg + kk) * 128);
      }
    }
  } else {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN) | (1u << 15) | (1u << 16);   // A and B MN-major
      for (int kk = 0; kk < nkb; kk++) {
        const int s = kk % STAGES;
        mbar_wait(smem_u32(&full[s]), (kk / STAGES) & 1);
        tc_fence_after();
        const uint64_t da = umma_desc_mnmajor_sw128(smem_u32(smA + s * C::A_BYTES), C::UNIT_BYTES);
        const uint64_t db = umma_desc_mnmajor_sw128(smem_u32(smB + s * C::B_BYTES), C::UNIT_BYTES);
#pragma unroll
        for (int k = 0; k < 128 / 16; k++)   // 16 pixel rows per MMA = two 1024-byte K groups: +2048 B = +128 units
          umma_bf16(tmem_base, da + 128 * k, db + 128 * k, idesc, (kk | k) != 0);
        umma_commit(smem_u32(&empty[s]));
      }
      umma_commit(smem_u32(accum_full));
    }
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// wgrad on a CTA pair (cta_group::2): the pair's tile is 4 units (256 accumulator rows) x BN output channels.  CTA r stages
// its own two units of x (TMA im2col) and its half of the dy columns; the leader issues M = 256 UMMAs.  Per staged byte that
// is twice the flops of the one-CTA 128 x 128 tile (BN = 256: 64 KB per 8 x (256 x 256 x 16) MMAs and CTA, against 64 KB per
// 8 x (128 x 128 x 16)), and the operand stream from L2 is what bounds this kernel.  TMA im2col operand only.
template <int BN> struct WgPairCfg {
  static constexpr int STAGES = BN == 256 ? 3 : 4;
  static constexpr int UNIT_BYTES = 128 * 128;
  static constexpr int A_BYTES = 2 * UNIT_BYTES;
  static constexpr int B_BYTES = (BN / 128) * UNIT_BYTES;      // this CTA's BN/2 columns of dy, 64 per box
  static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + 1024 + 256;
  static_assert(SMEM <= 227 * 1024, "wgrad pair tile does not fit shared memory");
};

template <int BN>
__global__ void __launch_bounds__(192, 1) wgrad_tc_pair_kernel(const __grid_constant__ WgParams p,
                                                               const __grid_constant__ CUtensorMap dymap,
                                                               const __grid_constant__ CUtensorMap xmap) {
  using C = WgPairCfg<BN>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smA = smem;
  uint8_t* smB = smem + STAGES * C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + STAGES * C::B_BYTES);
  uint64_t* full = bars;                  // leader: 1 arrival (expect_tx) + both CTAs' TMA bytes
  uint64_t* empty = bars + STAGES;        // each CTA: multicast tcgen05.commit
  uint64_t* accum_full = bars + 2 * STAGES;   // each CTA: multicast tcgen05.commit
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int u0 = (blockIdx.x >> 1) * 4 + (int)rank * 2;     // this CTA's two units
  const int n0 = blockIdx.y * BN;
  const int kb_beg = blockIdx.z * p.kb_per_split;
  const int kb_end = min(p.kb_total, kb_beg + p.kb_per_split);
  const int nkb = kb_end - kb_beg;                           // >= 1 by the host's split choice

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    mbar_init(smem_u32(accum_full), 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc_pair(smem_u32(tmem_slot), BN);
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&dymap) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
    }
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();

  if (warp < 4) {
    // ---- epilogue: this CTA's 128 accumulator rows (two units x 64 channels) x BN columns, staged per warp and reduced
    // into dW one whole row piece per lane (red.global.add.v4.f32), as in the one-CTA kernel
    if (nkb > 0) {
      mbar_wait(smem_u32(accum_full), 0);
      tc_fence_after();
      constexpr int PITCH = BN * 4 + 16;
      static_assert(4 * 32 * PITCH <= STAGES * (C::A_BYTES + C::B_BYTES), "staging must fit the drained pipeline buffers");
      uint8_t* slab = smem + warp * (32 * PITCH);
      const int ncols = min(BN, p.cout - n0);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        if (c0 >= ncols) continue;
        uint4* dst = reinterpret_cast<uint4*>(slab + lane * PITCH + c0 * 4);
#pragma unroll
        for (int j = 0; j < 8; j++) dst[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      __syncwarp();
      const int u = u0 + (warp >> 1);
      if (u < p.units) {
        const int tap = u / p.cblks;
        const int ci0 = (u - tap * p.cblks) * 64 + (warp & 1) * 32;
        float* obase = p.dw + ((size_t)tap * p.cin + ci0) * p.cout + n0;
        const int nrows = min(32, p.cin - ci0);
        const bool vec = (ncols % 4 == 0) && (p.cout % 4 == 0) && ((reinterpret_cast<uintptr_t>(obase) & 15) == 0);
        if (vec) {
          const int lpr = ncols / 4;
          for (int idx = lane; idx < nrows * lpr; idx += 32) {
            const int row = idx / lpr, piece = idx - row * lpr;
            const float4 q = *reinterpret_cast<const float4*>(slab + row * PITCH + piece * 16);
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(obase + (size_t)row * p.cout + piece * 4), "f"(q.x),
                         "f"(q.y), "f"(q.z), "f"(q.w)
                         : "memory");
          }
        } else {
          for (int row = 0; row < nrows; row++) {
            const float* srow = reinterpret_cast<const float*>(slab + row * PITCH);
            for (int c = lane; c < ncols; c += 32) atomicAdd(obase + (size_t)row * p.cout + c, srow[c]);
          }
        }
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    if (lane == 0) {
      int utap[2], uc0[2];
      bool uok[2];
      for (int j = 0; j < 2; j++) {
        const int u = u0 + j;
        uok[j] = u < p.units;
        const int uu = uok[j] ? u : 0;
        utap[j] = uu / p.cblks;
        uc0[j] = (uu - utap[j] * p.cblks) * 64;
      }
      for (int kk = 0; kk < nkb; kk++) {
        const int s = kk % STAGES;
        mbar_wait(smem_u32(&empty[s]), ((kk / STAGES) & 1) ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(smem_u32(&full[s]), 2 * (C::A_BYTES + C::B_BYTES));
        const long mpix = (long)(kb_beg + kk) * 128;
        const int n = (int)(mpix / p.HW), pi = (int)(mpix - (long)n * p.HW);
        const int oy = pi / p.WO, ox = pi - oy * p.WO;
#pragma unroll
        for (int j = 0; j < 2; j++)      // a unit past the end of the filter loads channel block `cin` (out of bounds -> zeros)
          tma_load_im2col_4d_pair(smem_u32(smA + s * C::A_BYTES + j * C::UNIT_BYTES), &xmap, smem_u32(&full[s]),
                                  uok[j] ? uc0[j] : p.cblks * 64, ox * p.stride - p.pad_l, oy * p.stride - p.pad_t, n,
                                  (unsigned short)(utap[j] % p.kw), (unsigned short)(utap[j] / p.kw));
#pragma unroll
        for (int j = 0; j < BN / 128; j++)
          tma_load_2d_pair(smem_u32(smB + s * C::B_BYTES + j * C::UNIT_BYTES), &dymap, smem_u32(&full[s]),
                           n0 + (int)rank * (BN / 2) + j * 64, (kb_beg + kk) * 128);
      }
    }
  } else {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN) | (1u << 15) | (1u << 16);   // A and B MN-major
      for (int kk = 0; kk < nkb; kk++) {
        const int s = kk % STAGES;
        mbar_wait(smem_u32(&full[s]), (kk / STAGES) & 1);
        tc_fence_after();
        const uint64_t da = umma_desc_mnmajor_sw128(smem_u32(smA + s * C::A_BYTES), C::UNIT_BYTES);
        const uint64_t db = umma_desc_mnmajor_sw128(smem_u32(smB + s * C::B_BYTES), C::UNIT_BYTES);
#pragma unroll
        for (int k = 0; k < 128 / 16; k++) umma_bf16_pair(tmem_base, da + 128 * k, db + 128 * k, idesc, (kk | k) != 0);
        umma_commit_pair(smem_u32(&empty[s]));
      }
      if (nkb > 0) umma_commit_pair(smem_u32(accum_full));
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, BN);
  }
}

// weights fp32 [taps][cin][cout] (* scale) -> bf16  F: [taps][cout][kpadF]   D: [taps][cin][kpadD]
// F is a transpose of the (cin, cout) plane: 32x32 tiles through shared memory so that both the fp32 reads (cout fastest)
// and the bf16 writes (cin fastest) are coalesced (the element-per-thread version read with stride cout: 0.95 TB/s on
// g_h1_lin's 6.4 M weights).  D keeps cout fastest on both sides.  Work items: F tiles first, then 1024-element D chunks.
__device__ __forceinline__ void wpack_body(const float* __restrict__ w, const float* __restrict__ scale, bf16* __restrict__ packF,
                                           bf16* __restrict__ packD, int taps, int cin, int cout, int kpadF, int kpadD, int bid,
                                           int nblocks, float (*sm)[33]) {
  const float sc = scale ? *scale : 1.f;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int cot = (cout + 31) / 32, kt = kpadF / 32;
  const long nFt = (long)taps * cot * kt;
  const long nD = (long)taps * cin * kpadD, nDt = (nD + 1023) / 1024;
  for (long t = bid; t < nFt + nDt; t += nblocks) {
    if (t < nFt) {
      const int k0 = (int)(t % kt) * 32;
      const long r = t / kt;
      const int co0 = (int)(r % cot) * 32, tap = (int)(r / cot);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int ci = k0 + ty + 8 * j, co = co0 + tx;
        sm[ty + 8 * j][tx] = (ci < cin && co < cout) ? w[((size_t)tap * cin + ci) * cout + co] * sc : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int co = co0 + ty + 8 * j;
        if (co < cout) packF[((size_t)tap * cout + co) * kpadF + k0 + tx] = __float2bfloat16_rn(sm[tx][ty + 8 * j]);
      }
      __syncthreads();
    } else {
      const long base = (t - nFt) * 1024;
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const long j = base + e * 256 + threadIdx.x;
        if (j < nD) {
          const int k = (int)(j % kpadD);
          const long r = j / kpadD;
          const int ci = (int)(r % cin), tap = (int)(r / cin);
          packD[j] = __float2bfloat16_rn(k < cout ? w[((size_t)tap * cin + ci) * cout + k] * sc : 0.f);
        }
      }
    }
  }
}
__global__ void __launch_bounds__(256) wpack_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                                    bf16* __restrict__ packF, bf16* __restrict__ packD, int taps, int cin, int cout,
                                                    int kpadF, int kpadD) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  __shared__ float sm[32][33];
  wpack_body(w, scale, packF, packD, taps, cin, cout, kpadF, kpadD, blockIdx.x, gridDim.x, sm);
}
// every weight pack of a program in ONE launch (blockIdx.y = weight): the CIFAR iteration refreshed 153 packs in 153 launches of
// ~7 us (3.7 % of the iteration); the weights of a step are all known at its start (parameters, spectral-norm outputs, folds)
constexpr int WPACK_MAX_BATCH = 32;
struct WpackItem { const float* w; bf16* packF; bf16* packD; int taps, cin, cout, kpadF, kpadD, nblk; };
struct WpackBatch { WpackItem it[WPACK_MAX_BATCH]; };
__global__ void __launch_bounds__(256) wpack_batched_kernel(const __grid_constant__ WpackBatch b) {
  pdl_sync();
  __shared__ float sm[32][33];
  const WpackItem& t = b.it[blockIdx.y];
  if ((int)blockIdx.x < t.nblk) wpack_body(t.w, nullptr, t.packF, t.packD, t.taps, t.cin, t.cout, t.kpadF, t.kpadD, blockIdx.x, t.nblk, sm);
}

// element-per-thread variant (kept for A/B: RCGAN_WPACK_FLAT=1)
__global__ void wpack_flat_kernel(const float* __restrict__ w, const float* __restrict__ scale, bf16* __restrict__ packF,
                             bf16* __restrict__ packD, int taps, int cin, int cout, int kpadF, int kpadD) {
  pdl_sync();   // programmatic dependent launch: nothing of the previous kernel is touched before this
  const float sc = scale ? *scale : 1.f;
  long nF = (long)taps * cout * kpadF, nD = (long)taps * cin * kpadD;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nF + nD; i += (long)gridDim.x * blockDim.x) {
    if (i < nF) {
      int k = (int)(i % kpadF);
      long r = i / kpadF;
      int co = (int)(r % cout), tap = (int)(r / cout);
      packF[i] = __float2bfloat16_rn(k < cin ? w[((size_t)tap * cin + k) * cout + co] * sc : 0.f);
    } else {
      long j = i - nF;
      int k = (int)(j % kpadD);
      long r = j / kpadD;
      int ci = (int)(r % cin), tap = (int)(r / cin);
      packD[j] = __float2bfloat16_rn(k < cout ? w[((size_t)tap * cin + ci) * cout + k] * sc : 0.f);
    }
  }
}

// UpsampleConv (gan_resnet.py:259-272) = 3x3 SAME conv of the 2x nearest-neighbour upsampled input.  For output pixel
// (2a+py, 2b+px) the three filter rows read source rows {a-1: ky 0; a: ky 1,2} (py = 0) or {a: ky 0,1; a+1: ky 2} (py = 1), and
// likewise for columns: each output parity class is a 2x2-tap conv of the SMALL input with pre-summed filter taps
// -- 4/9 of the multiply-adds and no materialised upsampled tensor.  wf: [class(4)][r(2)][c(2)][cin][cout] fp32.
__global__ void upconv_fold_kernel(const float* __restrict__ w, float* __restrict__ wf, int cin, int cout) {
  pdl_sync();
  const long per = (long)cin * cout, total = 16 * per;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int t = (int)(i / per);
    const long e = i - (long)t * per;
    const int cls = t >> 2, r = (t >> 1) & 1, c = t & 1, py = cls >> 1, px = cls & 1;
    // taps of the 3x3 filter that land on source offset index r (resp. c) for parity py (resp. px)
    const int ky0 = py == 0 ? (r == 0 ? 0 : 1) : (r == 0 ? 0 : 2), ky1 = py == 0 ? (r == 0 ? 0 : 2) : (r == 0 ? 1 : 2);
    const int kx0 = px == 0 ? (c == 0 ? 0 : 1) : (c == 0 ? 0 : 2), kx1 = px == 0 ? (c == 0 ? 0 : 2) : (c == 0 ? 1 : 2);
    float acc = 0.f;
    for (int ky = ky0; ky <= ky1; ky++)
      for (int kx = kx0; kx <= kx1; kx++) acc += w[(size_t)(ky * 3 + kx) * per + e];
    wf[i] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                                   const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

EncodeIm2colFn get_encode_im2col() {
  static EncodeIm2colFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(p);
  }
  return fn;
}

// RCGAN_TC_IM2COL=0 forces the cp.async gather (A/B comparison and a fallback switch for debugging)
bool im2col_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RCGAN_TC_IM2COL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
struct PackGeo { int taps, kpadF, kpadD; size_t offD, bytes; };
PackGeo pack_geo(const rcgan_conv_desc* d) {
  PackGeo g;
  g.taps = d->kh * d->kw;
  g.kpadF = round_up(d->cin, BK);
  g.kpadD = round_up(d->cout, BK);
  g.offD = (size_t)g.taps * d->cout * g.kpadF;                       // elements
  g.bytes = (g.offD + (size_t)g.taps * d->cin * g.kpadD) * sizeof(bf16);
  return g;
}
bool fprop_ok(const rcgan_conv_desc* d) {
  // any cout: a 3-channel output (G.Output) still moves 9 x 256 input channels per pixel -- the A-operand stream, not the MMA,
  // is the cost, and the TMA im2col path streams it ~3x faster than the CUDA-core gather-dot
  // cin < 32 only for pure GEMMs (1x1 over a [rows,1,1,K] patch matrix, see rcgan_im2col)
  const bool gemm = d->kh == 1 && d->kw == 1 && d->h == 1 && d->w == 1;
  return d->dtype == RCGAN_BF16 && d->ldx % 8 == 0 && (d->cin >= 32 || gemm) && d->cout >= 1 && d->kh * d->kw <= MAX_TAPS;
}
bool wgrad_ok(const rcgan_conv_desc* d) {
  const bool gemm = d->kh == 1 && d->kw == 1 && d->h == 1 && d->w == 1;
  return d->dtype == RCGAN_BF16 && d->ldx % 8 == 0 && d->ldy % 8 == 0 && (d->cin >= 32 || gemm) && d->cout >= 32 &&
         d->ho * d->wo <= 1024 && d->wo <= 255 && d->ho <= 255;
}
bool dgrad_ok(const rcgan_conv_desc* d) {
  // narrow outputs: stride 1 streams well through the TMA im2col path; stride 2 with cin >= 8 (the 11-channel image + label
  // concat of --concat_y) runs as one merged parity launch; 1..4 channels go through the GEMM + col2im path (nnops.ScatterDgrad)
  return d->dtype == RCGAN_BF16 && d->ldy % 8 == 0 && d->cout >= 32 && (d->cin >= 8 || d->stride == 1) && d->kh * d->kw <= MAX_TAPS &&
         (d->stride == 1 || d->stride == 2);
}

int make_wmap(CUtensorMap* map, const bf16* base, int kpad, int rows, int taps, int bn) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { rcgan_set_error("conv_tc: cuTensorMapEncodeTiled unavailable"); return RCGAN_ECUDA; }
  cuuint64_t gdim[3] = {(cuuint64_t)kpad, (cuuint64_t)rows, (cuuint64_t)taps};
  cuuint64_t gstr[2] = {(cuuint64_t)kpad * 2, (cuuint64_t)kpad * 2 * rows};
  cuuint32_t box[3] = {BK, (cuuint32_t)bn, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { rcgan_set_error("conv_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return RCGAN_ECUDA; }
  return 0;
}

template <int BN, int ST, bool IM2COL>
int launch_tc(const TcParams& p, const CUtensorMap& map, const CUtensorMap& amap, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, ST, IM2COL>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN, ST>::SMEM);
    if (e != cudaSuccess) { rcgan_set_error("conv_tc: smem opt-in failed: %s", cudaGetErrorString(e)); return RCGAN_ECUDA; }
    attr_done = true;
  }
  dim3 grid(((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN));
  launch_pdl(conv_tc_kernel<BN, ST, IM2COL>, grid, 192, Cfg<BN, ST>::SMEM, st, p, map, amap);
  RCGAN_LAUNCH_CHECK("conv_tc");
  rcgan_set_conv_variant("conv_tc<%d,%d,im2col=%d>", BN, ST, (int)IM2COL);
  return 0;
}

// im2col tensor map over the gathered operand [n, SH, SW, C] (channel stride ld): 128 pixels x 64 channels per load,
// SWIZZLE_128B; the pixel bounding box is [lo, lo + (MW-1)*stride] x [lo, lo + (MH-1)*stride] base positions so that
// the TMA's own w -> h -> n traversal enumerates exactly the GEMM rows m = (n, a, b) of the tile.
bool make_amap(CUtensorMap* map, const TcParams& p, int channels, int nimg) {
  EncodeIm2colFn enc = get_encode_im2col();
  if (!enc || !im2col_enabled()) return false;
  if ((p.ld_src * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(p.src) & 15)) return false;
  const int up_w = p.im_w_lo + (p.MW - 1) * p.im_sw + 1 - p.SW, up_h = p.im_h_lo + (p.MH - 1) * p.im_sh + 1 - p.SH;
  const int lim = 120;
  if (p.im_w_lo < -lim || p.im_w_lo > lim || p.im_h_lo < -lim || p.im_h_lo > lim || up_w < -lim || up_w > lim || up_h < -lim ||
      up_h > lim)
    return false;
  if (p.im_sw < 1 || p.im_sw > 8 || p.im_sh < 1 || p.im_sh > 8) return false;
  cuuint64_t gdim[4] = {(cuuint64_t)channels, (cuuint64_t)p.SW, (cuuint64_t)p.SH, (cuuint64_t)nimg};
  cuuint64_t gstr[3] = {(cuuint64_t)p.ld_src * 2, (cuuint64_t)p.SW * p.ld_src * 2, (cuuint64_t)p.SH * p.SW * p.ld_src * 2};
  int lo[2] = {p.im_w_lo, p.im_h_lo}, up[2] = {up_w, up_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)p.im_sw, (cuuint32_t)p.im_sh, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)p.src, gdim, gstr, lo, up, BK, BM, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN, int MT, int ST, typename TO>
int launch_tc_persist(TcMulti& mp, const CUtensorMap& map, const TcMaps& amaps, cudaStream_t st) {
  using C = PCfg<BN, MT, ST, TO>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_persist_kernel<BN, MT, ST, TO, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_tc_persist_kernel<BN, MT, ST, TO, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) { rcgan_set_error("conv_tc: smem opt-in failed: %s", cudaGetErrorString(e)); return RCGAN_ECUDA; }
    attr_done = true;
  }
  const int bn = mp.p[0].bn_eff, n_tiles_n = (mp.p[0].N + bn - 1) / bn;
  mp.tile_start[0] = 0;
  for (int q = 0; q < mp.nprob; q++)
    mp.tile_start[q + 1] = mp.tile_start[q] + ((mp.p[q].M + MT * BM - 1) / (MT * BM)) * n_tiles_n;
  const int n_tiles = mp.tile_start[mp.nprob];
  const int grid = n_tiles < RCGAN_NUM_SMS ? n_tiles : RCGAN_NUM_SMS;
  if (mp.nprob > 1) launch_pdl(conv_tc_persist_kernel<BN, MT, ST, TO, true>, grid, PERSIST_THREADS, C::SMEM, st, mp, map, amaps);
  else launch_pdl(conv_tc_persist_kernel<BN, MT, ST, TO, false>, grid, PERSIST_THREADS, C::SMEM, st, mp, map, amaps);
  RCGAN_LAUNCH_CHECK("conv_tc_persist");
  rcgan_set_conv_variant("conv_tc_persist<%d,%d,%d,%s,multi=%d>", BN, MT, ST, sizeof(TO) == 4 ? "f32" : "bf16", mp.nprob > 1 ? 1 : 0);
  return 0;
}

template <int BN, int MT, int ST, typename TO>
int launch_tc_pair(TcMulti& mp, const CUtensorMap& map, const TcMaps& amaps, cudaStream_t st) {
  using C = PairCfg<BN, MT, ST, TO>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_pair_kernel<BN, MT, ST, TO, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_tc_pair_kernel<BN, MT, ST, TO, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) { rcgan_set_error("conv_tc_pair: smem opt-in failed: %s", cudaGetErrorString(e)); return RCGAN_ECUDA; }
    attr_done = true;
  }
  const int bn = mp.p[0].bn_eff, n_tiles_n = (mp.p[0].N + bn - 1) / bn;
  mp.tile_start[0] = 0;
  for (int q = 0; q < mp.nprob; q++)
    mp.tile_start[q + 1] = mp.tile_start[q] + ((mp.p[q].M + 2 * MT * BM - 1) / (2 * MT * BM)) * n_tiles_n;
  const int n_tiles = mp.tile_start[mp.nprob];
  const int npairs = n_tiles < RCGAN_NUM_SMS / 2 ? n_tiles : RCGAN_NUM_SMS / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * npairs); cfg.blockDim = dim3(PERSIST_THREADS); cfg.dynamicSmemBytes = C::SMEM; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = rcgan_pdl_enabled() ? 2 : 1;
  if (mp.nprob > 1) cudaLaunchKernelEx(&cfg, conv_tc_pair_kernel<BN, MT, ST, TO, true>, mp, map, amaps);
  else cudaLaunchKernelEx(&cfg, conv_tc_pair_kernel<BN, MT, ST, TO, false>, mp, map, amaps);
  RCGAN_LAUNCH_CHECK("conv_tc_pair");
  rcgan_set_conv_variant("conv_tc_pair<%d,%d,%d,%s,multi=%d>", BN, MT, ST, sizeof(TO) == 4 ? "f32" : "bf16", mp.nprob > 1 ? 1 : 0);
  return 0;
}

// RCGAN_TC_PAIR: 0 = persistent launches stay on the one-CTA (cta_group::1) kernel, 1 = CTA pairs for the 256-wide tiles
// (and wgrad with cout % 256 == 0), 2 = CTA pairs wherever the N tile is a multiple of 32 (wgrad: cout % 128 == 0),
// 3 = fprop / dgrad as 2, wgrad as 1
int pair_mode() {
  const char* e = getenv("RCGAN_TC_PAIR");
  return e ? atoi(e) : 1;
}

// RCGAN_TC_PERSIST=0 selects the one-tile-per-CTA kernel everywhere (A/B comparisons); =2 forces the persistent kernel
int persist_mode() {
  const char* e = getenv("RCGAN_TC_PERSIST");
  return e ? atoi(e) : 1;
}

// persistent launch of nprob problems (common N / operands); tile shape by output width and by how many tiles there are
int run_tc_persist(TcMulti& mp, const TcMaps& amaps, const bf16* wbase, int kpad, int rows, int taps, cudaStream_t st) {
  TcParams& p = mp.p[0];
  // fp32 staging of a 128x256 tile does not fit next to the ring: fp32 outputs (and N <= 128) use (1|2) x (128 x <=128)
  const bool wide = p.N > 128 && !p.out_f32;
  const int cap = wide ? 256 : 128;
  const int bn_eff = p.N >= cap ? cap : round_up(p.N, 16);   // e.g. N = 138 (g_h2's 128 + 10 label channels) -> one 144-wide tile
  long tiles256 = 0;
  for (int q = 0; q < mp.nprob; q++) { mp.p[q].bn_eff = bn_eff; tiles256 += (long)((mp.p[q].M + 255) / 256) * ((p.N + bn_eff - 1) / bn_eff); }
  CUtensorMap map;
  if (bn_eff % 32 == 0 && (pair_mode() >= 2 || (pair_mode() == 1 && wide))) {   // 3: as 2 here
    // CTA pairs: each CTA loads half of the weight tile's columns
    if (int e = make_wmap(&map, wbase, kpad, rows, taps, bn_eff / 2)) return e;
    if (wide) return launch_tc_pair<256, 1, 4, bf16>(mp, map, amaps, st);
    if (tiles256 >= 2 * RCGAN_NUM_SMS)
      return p.out_f32 ? launch_tc_pair<128, 2, 3, float>(mp, map, amaps, st) : launch_tc_pair<128, 2, 4, bf16>(mp, map, amaps, st);
    return p.out_f32 ? launch_tc_pair<128, 1, 5, float>(mp, map, amaps, st) : launch_tc_pair<128, 1, 6, bf16>(mp, map, amaps, st);
  }
  if (int e = make_wmap(&map, wbase, kpad, rows, taps, bn_eff)) return e;
  if (wide) return launch_tc_persist<256, 1, 3, bf16>(mp, map, amaps, st);
  if (tiles256 >= 2 * RCGAN_NUM_SMS)     // enough work for 256-row tiles (one B tile feeds two MMAs)
    return p.out_f32 ? launch_tc_persist<128, 2, 3, float>(mp, map, amaps, st) : launch_tc_persist<128, 2, 3, bf16>(mp, map, amaps, st);
  return p.out_f32 ? launch_tc_persist<128, 1, 4, float>(mp, map, amaps, st) : launch_tc_persist<128, 1, 5, bf16>(mp, map, amaps, st);
}

int run_tc(TcParams& p, const bf16* wbase, int kpad, int rows, int taps, int channels, int nimg, cudaStream_t st) {
  { const char* e = getenv("RCGAN_TC_DBG"); p.dbg = e ? atoi(e) : 0; }
  CUtensorMap map, amap;
  const bool im2col = make_amap(&amap, p, channels, nimg);
  if (!im2col && im2col_enabled()) {
    // the cp.async gather is ~2x slower than the TMA im2col operand path: never take it silently (first occurrence per process)
    static bool warned = false;
    if (!warned) {
      warned = true;
      fprintf(stderr, "rcgan_b200: conv %dx%d source, ld %d, %d rows: im2col tensor map not encodable -> cp.async gather path "
                      "(about half the throughput); further occurrences are not reported\n", p.SH, p.SW, p.ld_src, p.M);
    }
  }
  // persistent big-tile kernel when it has at least ~3 tiles per SM to pipeline; below that the one-tile-per-CTA kernel
  // with two resident CTAs per SM fills the machine better
  const int pm = persist_mode();
  const long big_tiles = (p.N > 128 && !p.out_f32) ? (long)((p.M + 127) / 128) * ((p.N + 255) / 256)
                                                   : (long)((p.M + 255) / 256) * ((p.N + 127) / 128);
  if (p.colstats && !im2col) { rcgan_set_error("conv_tc: column statistics need the TMA im2col path"); return RCGAN_EUNSUPPORTED; }
  const bool persist = im2col && pm != 0 && (pm == 2 || p.colstats || big_tiles >= 3 * RCGAN_NUM_SMS);
  if (persist) {
    TcMulti mp;
    TcMaps amaps;
    mp.p[0] = p; mp.nprob = 1; amaps.a[0] = amap;
    return run_tc_persist(mp, amaps, wbase, kpad, rows, taps, st);
  }
  const int bn = p.N <= 64 ? 64 : 128;
  if (int e = make_wmap(&map, wbase, kpad, rows, taps, bn)) return e;
  if (im2col) return bn == 64 ? launch_tc<64, 4, true>(p, map, amap, st) : launch_tc<128, 3, true>(p, map, amap, st);
  return bn == 64 ? launch_tc<64, 4, false>(p, map, map, st) : launch_tc<128, 3, false>(p, map, map, st);
}

}  // namespace

extern "C" size_t rcgan_colstats_floats(int n) { return (size_t)RCGAN_COLSTATS_PARTS * (2 * (size_t)n + 1); }
static_assert(RCGAN_COLSTATS_PARTS == RCGAN_NUM_SMS, "one partial per persistent CTA");

extern "C" size_t rcgan_conv_wpack_bytes(const rcgan_conv_desc* d) {
  if (!d || !(fprop_ok(d) || dgrad_ok(d))) return 0;
  return pack_geo(d).bytes;
}

extern "C" int rcgan_conv_uses_tensor_cores(const rcgan_conv_desc* d, int direction) {
  if (!d) return 0;
  if (direction == 0) return fprop_ok(d) ? 1 : 0;
  if (direction == 1) return dgrad_ok(d) ? 1 : 0;
  if (direction == 2) return wgrad_ok(d) ? 1 : 0;
  return 0;
}

extern "C" int rcgan_conv_wpack(const rcgan_conv_desc* d, const float* w, const float* scale_dev, void* pack, void* stream) {
  RCGAN_CHECK_ARG(d && w && pack, "conv_wpack: null argument");
  if (!(fprop_ok(d) || dgrad_ok(d))) { rcgan_set_error("conv_wpack: shape has no tensor-core pack"); return RCGAN_EUNSUPPORTED; }
  PackGeo g = pack_geo(d);
  bf16* pk = reinterpret_cast<bf16*>(pack);
  static int flat = -1;
  if (flat < 0) { const char* e = getenv("RCGAN_WPACK_FLAT"); flat = (e && e[0] == '1') ? 1 : 0; }
  if (flat) {
    long n = (long)(g.bytes / sizeof(bf16));
    int fg = (int)((n + 255) / 256);
    if (fg > RCGAN_NUM_SMS * 8) fg = RCGAN_NUM_SMS * 8;
    launch_pdl(wpack_flat_kernel, fg, 256, 0, as_stream(stream), w, scale_dev, pk, pk + g.offD, g.taps, d->cin, d->cout, g.kpadF, g.kpadD);
    RCGAN_LAUNCH_CHECK("conv_wpack");
    return 0;
  }
  const long items = (long)g.taps * ((d->cout + 31) / 32) * (g.kpadF / 32) + ((long)g.taps * d->cin * g.kpadD + 1023) / 1024;
  int grid = (int)(items < RCGAN_NUM_SMS * 8 ? items : RCGAN_NUM_SMS * 8);
  launch_pdl(wpack_kernel, grid, 256, 0, as_stream(stream), w, scale_dev, pk, pk + g.offD, g.taps, d->cin, d->cout, g.kpadF, g.kpadD);
  RCGAN_LAUNCH_CHECK("conv_wpack");
  return 0;
}

extern "C" int rcgan_conv_wpack_batched(int count, const rcgan_conv_desc* const* descs, const float* const* w, void* const* packs,
                                       void* stream) {
  RCGAN_CHECK_ARG(count > 0 && descs && w && packs, "conv_wpack_batched: bad args");
  for (int i0 = 0; i0 < count; i0 += WPACK_MAX_BATCH) {
    const int n = count - i0 < WPACK_MAX_BATCH ? count - i0 : WPACK_MAX_BATCH;
    WpackBatch b;
    int max_blk = 1;
    for (int k = 0; k < n; k++) {
      const rcgan_conv_desc* d = descs[i0 + k];
      RCGAN_CHECK_ARG(d && w[i0 + k] && packs[i0 + k], "conv_wpack_batched: null item %d", i0 + k);
      if (!(fprop_ok(d) || dgrad_ok(d))) { rcgan_set_error("conv_wpack_batched: item %d has no tensor-core pack", i0 + k); return RCGAN_EUNSUPPORTED; }
      PackGeo g = pack_geo(d);
      bf16* pk = reinterpret_cast<bf16*>(packs[i0 + k]);
      const long items = (long)g.taps * ((d->cout + 31) / 32) * (g.kpadF / 32) + ((long)g.taps * d->cin * g.kpadD + 1023) / 1024;
      WpackItem& t = b.it[k];
      t.w = w[i0 + k]; t.packF = pk; t.packD = pk + g.offD; t.taps = g.taps; t.cin = d->cin; t.cout = d->cout;
      t.kpadF = g.kpadF; t.kpadD = g.kpadD;
      t.nblk = (int)(items < 8 * RCGAN_NUM_SMS ? items : 8 * RCGAN_NUM_SMS);
      if (t.nblk > max_blk) max_blk = t.nblk;
    }
    launch_pdl(wpack_batched_kernel, dim3(max_blk, n), 256, 0, as_stream(stream), b);
    RCGAN_LAUNCH_CHECK("conv_wpack_batched");
  }
  return 0;
}

// optional epilogue fusions -> kernel parameters (bf16 outputs only; the caller has validated the combination)
static void set_epilogue(TcParams& p, const rcgan_conv_epilogue* ep) {
  p.res_up = 0; p.ld_res = 0; p.mask = nullptr; p.mask_act = 0; p.mask_leak = 0.f; p.out2 = nullptr; p.out2_act = 0;
  p.colstats = nullptr;
  if (!ep) return;
  p.colstats = ep->colstats;
  if (ep->res) { p.res = ep->res; p.res_up = ep->res_up; p.ld_res = ep->ld_res; }
  p.mask = ep->mask; p.mask_act = ep->mask_act; p.mask_leak = ep->mask_leak;
  p.out2 = ep->out2; p.out2_act = ep->out2_act;
}
static bool epilogue_ok(const rcgan_conv_epilogue* ep, int out_dtype, int accumulate, int oh, int ow) {
  if (!ep) return true;
  if (out_dtype != RCGAN_BF16) return false;
  if (ep->res && accumulate) return false;
  if (ep->res && ep->res_up && (oh % 2 || ow % 2 || ep->ld_res <= 0)) return false;
  if (ep->mask && ep->mask_act != RCGAN_ACT_RELU && ep->mask_act != RCGAN_ACT_LRELU) return false;
  if (ep->out2 && ep->out2_act != RCGAN_ACT_RELU) return false;
  return true;
}
// column statistics need one N tile whose 16-byte pieces map to fixed lanes: N in {64, 128, 256}, dense rows, persistent kernel
static bool colstats_ok(const rcgan_conv_epilogue* ep, int n, int ld_out) {
  if (!ep || !ep->colstats) return true;
  return (n == 64 || n == 128 || n == 256) && ld_out == n;
}

int rcgan_tc_fprop(const rcgan_conv_desc* d, const void* x, const void* wpack, const float* bias, const void* res, void* y,
                   int out_dtype, int act, float leak, cudaStream_t st, int* handled, const rcgan_conv_epilogue* ep) {
  *handled = 0;
  if (!fprop_ok(d) || (out_dtype != RCGAN_BF16 && out_dtype != RCGAN_F32)) return 0;
  if (!epilogue_ok(ep, out_dtype, 0, d->ho, d->wo) || !colstats_ok(ep, d->cout, d->ldy)) {
    rcgan_set_error("conv2d_fprop_ex: unsupported epilogue combination"); return RCGAN_EUNSUPPORTED;
  }
  PackGeo g = pack_geo(d);
  TcParams p;
  p.src = reinterpret_cast<const bf16*>(x);
  p.SH = d->h; p.SW = d->w; p.ld_src = d->ldx; p.cvalid = round_up(d->cin, 8);
  p.MH = d->ho; p.MW = d->wo; p.M = d->n * d->ho * d->wo;
  p.by_mul = d->stride; p.by_add = -d->pad_t; p.bx_mul = d->stride; p.bx_add = -d->pad_l;
  p.ntaps = g.taps; p.kb_per_tap = g.kpadF / BK;
  for (int t = 0; t < g.taps; t++) { p.tdy[t] = (short)(t / d->kw); p.tdx[t] = (short)(t % d->kw); p.twi[t] = (short)t; }
  p.out = y; p.out_f32 = out_dtype == RCGAN_F32; p.ld_out = d->ldy; p.OH = d->ho; p.OW = d->wo;
  p.oy_mul = 1; p.oy_add = 0; p.ox_mul = 1; p.ox_add = 0; p.N = d->cout;
  p.bias = bias; p.act = act; p.leak = leak; p.accumulate = 0; p.res = res;
  set_epilogue(p, ep);
  p.im_w_lo = -d->pad_l; p.im_h_lo = -d->pad_t; p.im_sw = d->stride; p.im_sh = d->stride;
  for (int t = 0; t < g.taps; t++) { p.toffh[t] = (unsigned short)(t / d->kw); p.toffw[t] = (unsigned short)(t % d->kw); }
  if (int e = run_tc(p, reinterpret_cast<const bf16*>(wpack), g.kpadF, d->cout, g.taps, d->cin, d->n, st)) return e;
  *handled = 1;
  return 0;
}

int rcgan_tc_dgrad(const rcgan_conv_desc* d, const void* dy, const void* wpack, const float* bias, void* dx, int out_dtype,
                   int act, float leak, int accumulate, cudaStream_t st, int* handled, const rcgan_conv_epilogue* ep) {
  *handled = 0;
  if (!dgrad_ok(d) || (out_dtype != RCGAN_BF16 && out_dtype != RCGAN_F32)) return 0;
  if (!epilogue_ok(ep, out_dtype, accumulate, d->h, d->w) || !colstats_ok(ep, d->cin, d->ldx) ||
      (ep && ep->colstats && d->stride != 2 && d->stride != 1)) {
    rcgan_set_error("conv2d_dgrad_ex: unsupported epilogue combination"); return RCGAN_EUNSUPPORTED;
  }
  PackGeo g = pack_geo(d);
  const bf16* wD = reinterpret_cast<const bf16*>(wpack) + g.offD;
  const int s = d->stride;
  TcMulti mp;
  TcMaps amaps;
  mp.nprob = 0;
  bool all_im2col = s == 2 && persist_mode() != 0;     // the 4 parity classes as ONE persistent launch when all are encodable
  for (int py = 0; py < s; py++)
    for (int px = 0; px < s; px++) {
      TcParams p;
      p.src = reinterpret_cast<const bf16*>(dy);
      p.SH = d->ho; p.SW = d->wo; p.ld_src = d->ldy; p.cvalid = round_up(d->cout, 8);
      p.MH = (d->h - py + s - 1) / s; p.MW = (d->w - px + s - 1) / s;
      if (p.MH <= 0 || p.MW <= 0) continue;
      p.M = d->n * p.MH * p.MW;
      // input pixel iy = s*a + py receives tap ky iff (iy + pad_t - ky) % s == 0:  ky = kpar + s*j,  oy = a + cy - j
      const int kpy = (py + d->pad_t) % s, kpx = (px + d->pad_l) % s;
      const int cy = (py + d->pad_t - kpy) / s, cx = (px + d->pad_l - kpx) / s;
      p.by_mul = 1; p.by_add = cy; p.bx_mul = 1; p.bx_add = cx;
      int nt = 0;
      const int njy = (d->kh - kpy + s - 1) / s, njx = (d->kw - kpx + s - 1) / s;   // taps of this parity class per dim
      for (int ky = kpy, jy = 0; ky < d->kh; ky += s, jy++)
        for (int kx = kpx, jx = 0; kx < d->kw; kx += s, jx++) {
          p.tdy[nt] = (short)(-jy); p.tdx[nt] = (short)(-jx); p.twi[nt] = (short)(ky * d->kw + kx);
          p.toffh[nt] = (unsigned short)(njy - 1 - jy); p.toffw[nt] = (unsigned short)(njx - 1 - jx);
          nt++;
        }
      p.im_h_lo = cy - (njy - 1); p.im_w_lo = cx - (njx - 1); p.im_sh = 1; p.im_sw = 1;
      p.ntaps = nt; p.kb_per_tap = g.kpadD / BK;
      p.out = dx; p.out_f32 = out_dtype == RCGAN_F32; p.ld_out = d->ldx; p.OH = d->h; p.OW = d->w;
      p.oy_mul = s; p.oy_add = py; p.ox_mul = s; p.ox_add = px; p.N = d->cin;
      p.bias = bias; p.act = act; p.leak = leak; p.accumulate = accumulate; p.res = nullptr;
      set_epilogue(p, ep);
      if (nt == 0) { rcgan_set_error("conv_tc dgrad: parity class without taps"); return RCGAN_EUNSUPPORTED; }
      if (s == 2) {
        { const char* e = getenv("RCGAN_TC_DBG"); p.dbg = e ? atoi(e) : 0; }
        all_im2col = all_im2col && mp.nprob < 4 && make_amap(&amaps.a[mp.nprob], p, d->cout, d->n);
        if (mp.nprob < 4) mp.p[mp.nprob++] = p;
      } else {
        if (int e = run_tc(p, wD, g.kpadD, d->cin, g.taps, d->cout, d->n, st)) return e;
      }
    }
  if (s == 2) {
    if (all_im2col && mp.nprob > 0) {
      if (int e = run_tc_persist(mp, amaps, wD, g.kpadD, d->cin, g.taps, st)) return e;
    } else {
      for (int q = 0; q < mp.nprob; q++)
        if (int e = run_tc(mp.p[q], wD, g.kpadD, d->cin, g.taps, d->cout, d->n, st)) return e;
    }
  }
  *handled = 1;
  return 0;
}

extern "C" size_t rcgan_upconv2d_pack_bytes(const rcgan_conv_desc* d) {
  // d: the 3x3 stride-1 conv at the OUTPUT resolution (h x w); folded filter = a 16-tap pack
  if (!d || d->kh != 3 || d->kw != 3 || d->stride != 1 || d->pad_t != 1 || d->pad_l != 1 || d->dtype != RCGAN_BF16 || d->h % 2 ||
      d->w % 2 || d->cin < 32 || d->cin % 8 || d->ldx % 8 || d->cout < 32)
    return 0;
  rcgan_conv_desc f = *d;
  f.kh = 4; f.kw = 4;
  return pack_geo(&f).bytes;
}

extern "C" int rcgan_upconv2d_fold(const rcgan_conv_desc* d, const float* w, float* wfold, void* pack, void* stream) {
  RCGAN_CHECK_ARG(d && w && wfold && pack && rcgan_upconv2d_pack_bytes(d) > 0, "upconv2d_fold: unsupported shape");
  const long total = 16L * d->cin * d->cout;
  int grid = (int)((total + 255) / 256);
  if (grid > RCGAN_NUM_SMS * 8) grid = RCGAN_NUM_SMS * 8;
  launch_pdl(upconv_fold_kernel, grid, 256, 0, as_stream(stream), w, wfold, d->cin, d->cout);
  RCGAN_LAUNCH_CHECK("upconv_fold");
  rcgan_conv_desc f = *d;
  f.kh = 4; f.kw = 4;
  return rcgan_conv_wpack(&f, wfold, nullptr, pack, stream);
}

extern "C" int rcgan_upconv2d_fprop(const rcgan_conv_desc* d, const void* x_small, const void* pack, const float* bias, void* y,
                                    int out_dtype, int act, float leak, void* stream) {
  RCGAN_CHECK_ARG(d && x_small && pack && y && rcgan_upconv2d_pack_bytes(d) > 0, "upconv2d_fprop: unsupported shape");
  RCGAN_CHECK_ARG(out_dtype == RCGAN_BF16 || out_dtype == RCGAN_F32, "upconv2d_fprop: bad output dtype");
  rcgan_conv_desc f = *d;
  f.kh = 4; f.kw = 4;
  PackGeo g = pack_geo(&f);
  const int hs = d->h / 2, ws = d->w / 2;
  TcMulti mp;
  TcMaps amaps;
  mp.nprob = 0;
  for (int py = 0; py < 2; py++)
    for (int px = 0; px < 2; px++) {
      TcParams p;
      p.src = reinterpret_cast<const bf16*>(x_small);
      p.SH = hs; p.SW = ws; p.ld_src = d->ldx; p.cvalid = round_up(d->cin, 8);
      p.MH = hs; p.MW = ws; p.M = d->n * hs * ws;
      const int oy0 = py == 0 ? -1 : 0, ox0 = px == 0 ? -1 : 0;       // first source offset of this parity class
      p.by_mul = 1; p.by_add = 0; p.bx_mul = 1; p.bx_add = 0;
      int nt = 0;
      for (int r = 0; r < 2; r++)
        for (int c = 0; c < 2; c++) {
          p.tdy[nt] = (short)(oy0 + r); p.tdx[nt] = (short)(ox0 + c);
          p.twi[nt] = (short)((py * 2 + px) * 4 + r * 2 + c);
          p.toffh[nt] = (unsigned short)r; p.toffw[nt] = (unsigned short)c;
          nt++;
        }
      p.im_h_lo = oy0; p.im_w_lo = ox0; p.im_sh = 1; p.im_sw = 1;
      p.ntaps = nt; p.kb_per_tap = g.kpadF / BK;
      p.out = y; p.out_f32 = out_dtype == RCGAN_F32; p.ld_out = d->ldy; p.OH = d->h; p.OW = d->w;
      p.oy_mul = 2; p.oy_add = py; p.ox_mul = 2; p.ox_add = px; p.N = d->cout;
      p.bias = bias; p.act = act; p.leak = leak; p.accumulate = 0; p.res = nullptr;
      set_epilogue(p, nullptr);
      { const char* e = getenv("RCGAN_TC_DBG"); p.dbg = e ? atoi(e) : 0; }
      if (!make_amap(&amaps.a[mp.nprob], p, d->cin, d->n)) {
        rcgan_set_error("upconv2d_fprop: im2col tensor map not encodable for this shape");
        return RCGAN_EUNSUPPORTED;
      }
      mp.p[mp.nprob++] = p;
    }
  return run_tc_persist(mp, amaps, reinterpret_cast<const bf16*>(pack), g.kpadF, d->cout, g.taps, as_stream(stream));
}

template <int BN, bool IM2COL>
static int launch_wgrad_tc(const WgParams& p, const CUtensorMap& map, const CUtensorMap& xmap, dim3 grid, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<BN, IM2COL>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg<BN>::SMEM);
    if (e != cudaSuccess) { rcgan_set_error("wgrad_tc: smem opt-in failed: %s", cudaGetErrorString(e)); return RCGAN_ECUDA; }
    attr_done = true;
  }
  launch_pdl(wgrad_tc_kernel<BN, IM2COL>, grid, 192, WgCfg<BN>::SMEM, st, p, map, xmap);
  RCGAN_LAUNCH_CHECK("wgrad_tc");
  rcgan_set_conv_variant("wgrad_tc<%d,im2col=%d>", BN, (int)IM2COL);
  return 0;
}

template <int BN>
static int launch_wgrad_tc_pair(const WgParams& p, const CUtensorMap& map, const CUtensorMap& xmap, dim3 grid, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgPairCfg<BN>::SMEM);
    if (e != cudaSuccess) { rcgan_set_error("wgrad_tc_pair: smem opt-in failed: %s", cudaGetErrorString(e)); return RCGAN_ECUDA; }
    attr_done = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = WgPairCfg<BN>::SMEM; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = rcgan_pdl_enabled() ? 2 : 1;
  cudaLaunchKernelEx(&cfg, wgrad_tc_pair_kernel<BN>, p, map, xmap);
  RCGAN_LAUNCH_CHECK("wgrad_tc_pair");
  rcgan_set_conv_variant("wgrad_tc_pair<%d>", BN);
  return 0;
}

// dw (=|+=) x^T dy on the tensor cores.  dw must be zero-initialised by the caller path when !accumulate.
int rcgan_tc_wgrad(const rcgan_conv_desc* d, const void* x, const void* dy, float* dw, int accumulate, cudaStream_t st,
                   int* handled) {
  *handled = 0;
  if (!wgrad_ok(d)) return 0;
  EncodeTiledFn enc = get_encode();
  if (!enc) { rcgan_set_error("wgrad_tc: cuTensorMapEncodeTiled unavailable"); return RCGAN_ECUDA; }
  WgParams p;
  p.x = reinterpret_cast<const bf16*>(x);
  p.H = d->h; p.W = d->w; p.ldx = d->ldx; p.cvalid = round_up(d->cin, 8);
  p.HO = d->ho; p.WO = d->wo; p.HW = d->ho * d->wo; p.Mpix = d->n * p.HW;
  p.stride = d->stride; p.pad_t = d->pad_t; p.pad_l = d->pad_l; p.kw = d->kw;
  p.cin = d->cin; p.cout = d->cout; p.cblks = (d->cin + 63) / 64; p.units = d->kh * d->kw * p.cblks;
  p.kb_total = (p.Mpix + 127) / 128;
  p.dw = dw;
  const int bn = d->cout <= 64 ? 64 : 128;
  const int tiles = ((p.units + 1) / 2) * ((d->cout + bn - 1) / bn);
  int waves_x2 = 2;                                            // experiment: RCGAN_WG_WAVES_X2 = CTAs per SM x 2
  { const char* e = getenv("RCGAN_WG_WAVES_X2"); if (e) waves_x2 = atoi(e); }
  int splits = (waves_x2 * RCGAN_NUM_SMS / 2 + (waves_x2 > 2 ? tiles - 1 : 0)) / tiles;
  int max_splits = (p.kb_total + 3) / 4;                        // at least 4 K blocks (512 pixels) per split
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  CUtensorMap map;
  cuuint64_t gdim[2] = {(cuuint64_t)d->ldy, (cuuint64_t)p.Mpix};
  cuuint64_t gstr[1] = {(cuuint64_t)d->ldy * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(dy), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { rcgan_set_error("wgrad_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return RCGAN_ECUDA; }
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->kh * d->kw * d->cin * d->cout, st);
    if (e != cudaSuccess) { rcgan_set_error("wgrad_tc: memset failed: %s", cudaGetErrorString(e)); return RCGAN_ECUDA; }
  }
  dim3 grid((p.units + 1) / 2, (d->cout + bn - 1) / bn, splits);
  // x operand through the TMA im2col path when encodable (same geometry as the fprop A operand)
  TcParams ap;
  ap.src = p.x; ap.SH = d->h; ap.SW = d->w; ap.ld_src = d->ldx; ap.MH = d->ho; ap.MW = d->wo;
  ap.im_w_lo = -d->pad_l; ap.im_h_lo = -d->pad_t; ap.im_sw = d->stride; ap.im_sh = d->stride;
  CUtensorMap xmap;
  const bool im2col = make_amap(&xmap, ap, d->cin, d->n);
  // CTA pairs for 256-multiple cout (measured 512x32x32 256->256 k3: 580 -> 398 us, 1.07 -> 1.56 PFLOP/s; with 128 output channels
  // the pair tile stages 48 KB per 8 MMAs against 64 KB and came out even or slower: 43.6 -> 48.0 us at 512x16x16 128->128 k3;
  // RCGAN_TC_PAIR=2 takes that path too)
  if (im2col && p.units >= 4 && ((pair_mode() >= 1 && d->cout % 256 == 0) || (pair_mode() == 2 && d->cout % 128 == 0))) {
    // 4 units x (256 | 128) channels per pair; split K so that the pairs fill the 74 TPCs in whole waves
    const int pbn = d->cout % 256 == 0 ? 256 : 128;
    const int ptiles = ((p.units + 3) / 4) * (d->cout / pbn);
    long best = -1;
    int best_s = 1;
    for (int sp = 1; sp <= max_splits && sp <= 4 * RCGAN_NUM_SMS; sp++) {
      const int per = (p.kb_total + sp - 1) / sp;
      const int real = (p.kb_total + per - 1) / per;
      const long waves = ((long)ptiles * real + RCGAN_NUM_SMS / 2 - 1) / (RCGAN_NUM_SMS / 2);
      const long cost = waves * (per + 6);               // + ~6 K blocks of prologue / epilogue per CTA
      if (best < 0 || cost < best) { best = cost; best_s = sp; }
    }
    p.kb_per_split = (p.kb_total + best_s - 1) / best_s;
    const int psplits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
    dim3 pgrid(2 * ((p.units + 3) / 4), d->cout / pbn, psplits);
    if (int e = (pbn == 256 ? launch_wgrad_tc_pair<256>(p, map, xmap, pgrid, st) : launch_wgrad_tc_pair<128>(p, map, xmap, pgrid, st)))
      return e;
    *handled = 1;
    return 0;
  }
  if (im2col) {
    if (int e = (bn == 64 ? launch_wgrad_tc<64, true>(p, map, xmap, grid, st) : launch_wgrad_tc<128, true>(p, map, xmap, grid, st)))
      return e;
  } else {
    if (int e = (bn == 64 ? launch_wgrad_tc<64, false>(p, map, map, grid, st) : launch_wgrad_tc<128, false>(p, map, map, grid, st)))
      return e;
  }
  *handled = 1;
  return 0;
}
