// fp32-accurate tensor-core GEMMs for the fp32 tier (BASELINE config 3: h = 600, depth 6, fp32): every product is
// evaluated as THREE tcgen05.mma.kind::tf32 passes on an error-free split of both operands,
//
//     x = x_hi + x_lo,   x_hi = rna_tf32(x),  x_lo = rna_tf32(x - x_hi)   (x - x_hi is exact in f32)
//     a . b  ~=  a_lo . b_hi + a_hi . b_lo + a_hi . b_hi                  (the a_lo . b_lo term, 2^-22 relative, is dropped)
//
// with f32 accumulation in tensor memory: 2^-21-class relative error per product instead of tf32's 2^-11, i.e. the
// accuracy of an f32 FMA chain (the reference's ATen sgemm), at tensor-core speed.  SURVEY.md section 7, hard part 2.
//
//   dmpnn_linear_x3   C[r, 0:N] = act(A[r, 0:K] . W^T + bias + res[r, 0:N])        A, C, res f32 row-major
//        replaces chemprop/nn/message_passing/base.py:135-141 (W_h), :180-182 (W_o) and the dX GEMMs of their autograd
//   dmpnn_wgrad_x3    dW[n, k] (+)= sum_r dY[r, n] X[r, k]                          autograd of the same nn.Linear layers
//
// k_linear_x3 -- persistent, one CTA per SM, work item = (256-row pair of tiles, 128-column pass over N):
//   warp 0       cp.async.bulk producer of the W stages: pre-split, pre-swizzled {W_hi, W_lo} blocks of one 32-wide k slab
//   warp 1       tcgen05.mma issuer (converged warp, elected lane): per stage 2 tiles x 4 k steps x 3 products, M = 128, N <= 128
//   warp 2       TMEM allocator: 512 columns = 2 buffers x 2 tiles x 128 accumulator columns (epilogue of item i overlaps item i+1)
//   warps 4-7    epilogue: tcgen05.ld -> + bias + residual -> act -> f32 rows straight to global memory (64-byte segments)
//   warps 8-15   A producers, thread = tile row: 128-byte row segment of the slab by LDG.128 (optionally through a row
//                index), split into hi / lo in registers, written as K-major SWIZZLE_128B operand tiles (generic proxy ->
//                fence.proxy.async); the loads of slab s+1 are in flight while slab s waits for its stage
// Every W stage (32 KB) serves two A tiles: L2 -> SM operand traffic per flop is half that of a 128-row item, which keeps
// the kernel on the tensor pipe rather than on L2 bandwidth (the pipe needs ~100 flop / L2 byte at 3 passes of tf32).
//
// k_wgrad_x3 -- CTA (n tile of 128, k tile of <= 256, row slot): D[128 x 256] (TMEM) += dY_blk^T . X_blk over 32-row
// blocks; both operands are MN-major (the contraction runs over rows), written by the producers into the one layout the
// tensor core transposes 32-bit operands from (SWIZZLE_128B_BASE32B); partial sums per CTA, fixed-order reduce (deterministic).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace dmpnn {
namespace x3 {

using namespace dmpnn::tc;

// ------------------------------------------------------------------------------------------------------------------
// shared pieces
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rna_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
// x -> (hi, lo), both exactly representable in tf32; hi + lo == x up to 2^-22 |x|
__device__ __forceinline__ void split1(float x, uint32_t& hi, uint32_t& lo) {
  hi = rna_tf32(x);
  lo = rna_tf32(__fsub_rn(x, __uint_as_float(hi)));
}
__device__ __forceinline__ void split4(const float4 v, uint4& hi, uint4& lo) {
  split1(v.x, hi.x, lo.x);
  split1(v.y, hi.y, lo.y);
  split1(v.z, hi.z, lo.z);
  split1(v.w, hi.w, lo.w);
}
// c = F32, a = b = TF32 (format 2), optional MN-major operands, N >> 3 at [17, 23), M >> 4 at [24, 29)
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N, bool mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (mn_major ? ((1u << 15) | (1u << 16)) : 0u) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// MN-major tf32 operand.  32-bit operands transposed by the tensor core exist in ONE shared-memory layout only,
// SWIZZLE_128B_BASE32B (descriptor layout type 1): rows of the contraction index at a 128-byte pitch (32 tf32 along MN),
// atoms of FOUR rows (512 B) inside which the 32-byte chunk index is XORed with (row & 3) -- Swizzle<2,5,2> on byte
// addresses; atoms along K `sbo` = 512 B apart, 32-element blocks along MN `lbo` bytes apart.  One k step of the
// instruction (8 tf32) covers two atoms.
__device__ __forceinline__ uint64_t desc_mn_sw128_32b(uint32_t saddr, uint32_t lbo) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
// byte offset of 16-byte chunk `c` (0..7) of contraction row `r` inside one 32-element MN block of that layout
__device__ __forceinline__ uint32_t mn32b_off(int r, int c) {
  return (uint32_t)(r * 128 + ((((c >> 1) ^ (r & 3)) << 5) | ((c & 1) << 4)));
}

// ------------------------------------------------------------------------------------------------------------------
// linear
// ------------------------------------------------------------------------------------------------------------------
namespace lin {
constexpr int kTileM = 128;
constexpr int kPairM = 256;
constexpr int kNP = 128;                      // output columns per pass
constexpr int kABytes = kTileM * 128;         // one tile x one 32-wide k slab x one of {hi, lo}: 16 KB
constexpr int kWBytes = kNP * 128;            // one of {W_hi, W_lo} of a k slab: 16 KB
constexpr int kOffAhi = 0, kOffAlo = 2 * kABytes, kOffW = 4 * kABytes;
constexpr int kStageBytes = 4 * kABytes + 2 * kWBytes;   // 96 KB
constexpr int kStages = 2;
constexpr int kThreads = 512;
constexpr int kProducerWarps = 8;
constexpr int kTmemCols = 512;
constexpr int kOffBar = kStages * kStageBytes;
constexpr int kOffTmem = kOffBar + 16 * 8;
constexpr int kSmemBytes = kOffTmem + 16;
constexpr int kSmemAlloc = kSmemBytes + 1024;
static_assert(kStageBytes % 1024 == 0, "SWIZZLE_128B operand tiles need 1024-byte alignment");
static_assert(kSmemAlloc <= 232448, "exceeds shared memory");

enum { B_FULL = 0, B_EMPTY = 2, B_ACCFULL = 4, B_ACCFREE = 6 };

struct Params {
  const float* A;
  const int32_t* idx;      // optional row gather: A row of output row r is idx[r]
  int64_t lda;
  const uint8_t* Wpk;
  const float* bias;
  const float* res;
  int64_t ldres;
  float* C;
  int64_t ldc;
  int64_t R;
  int K, N, Npad, ldc_pad, nslab, n_pass, n_pairs;
  int act;
  float act_param;
};

__global__ void __launch_bounds__(kThreads, 1) k_linear_x3(Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t sBar = sbase + kOffBar;
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + kOffTmem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar = [&](int i) { return sBar + 8u * (uint32_t)i; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar(B_FULL + i), 1 + kProducerWarps);   // W loader's expect_tx arrive + one arrive per producer warp
      mbar_init(bar(B_EMPTY + i), 1);                   // tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(B_ACCFULL + i), 1);
      mbar_init(bar(B_ACCFREE + i), 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(const_cast<uint32_t*>(s_tmem)), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===================== W stages: {W_hi, W_lo} of (pass, slab), contiguous 32 KB in the packed image ==========
    uint32_t ks = 0;
    for (int pair = blockIdx.x; pair < p.n_pairs; pair += gridDim.x)
      for (int ps = 0; ps < p.n_pass; ++ps)
        for (int s = 0; s < p.nslab; ++s, ++ks) {
          const uint32_t st = ks % kStages, use = ks / kStages;
          mbar_wait(bar(B_EMPTY + st), (use & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx(bar(B_FULL + st), 2 * kWBytes);
            bulk_load(sbase + st * kStageBytes + kOffW, p.Wpk + ((size_t)ps * p.nslab + s) * (2 * kWBytes), 2 * kWBytes,
                      bar(B_FULL + st));
          }
          __syncwarp();
        }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    uint32_t ks = 0;
    int it = 0;
    for (int pair = blockIdx.x; pair < p.n_pairs; pair += gridDim.x)
      for (int ps = 0; ps < p.n_pass; ++ps, ++it) {
        const int np = min(kNP, p.Npad - ps * kNP);
        const uint32_t idesc = idesc_tf32(kTileM, np, false);
        const uint32_t buf = (uint32_t)it & 1u;
        mbar_wait(bar(B_ACCFREE + buf), ((it >> 1) & 1) ^ 1);     // the epilogue drained this buffer two items ago
        tc_fence_after();
        for (int s = 0; s < p.nslab; ++s, ++ks) {
          const uint32_t st = ks % kStages, use = ks / kStages;
          mbar_wait(bar(B_FULL + st), use & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t stg = sbase + st * kStageBytes;
            const uint64_t bhi = umma_desc_sw128(stg + kOffW), blo = umma_desc_sw128(stg + kOffW + kWBytes);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const uint32_t d = tmem_base + buf * 256u + (uint32_t)t * 128u;
              const uint64_t ahi = umma_desc_sw128(stg + kOffAhi + t * kABytes), alo = umma_desc_sw128(stg + kOffAlo + t * kABytes);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {          // k step = 8 tf32 = 32 B of every operand row
                const uint64_t o = (uint64_t)(2 * kk);
                umma_tf32(d, alo + o, bhi + o, idesc, (s > 0 || kk > 0) ? 1u : 0u);
                umma_tf32(d, ahi + o, blo + o, idesc, 1u);
                umma_tf32(d, ahi + o, bhi + o, idesc, 1u);
              }
            }
            umma_commit(bar(B_EMPTY + st));
            if (s == p.nslab - 1) umma_commit(bar(B_ACCFULL + buf));
          }
          __syncwarp();
        }
      }
  } else if (warp >= 4 && warp < 8) {
    // ===================== epilogue: thread = accumulator row (TMEM lane) =====================
    const int row = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    int it = 0;
    for (int pair = blockIdx.x; pair < p.n_pairs; pair += gridDim.x)
      for (int ps = 0; ps < p.n_pass; ++ps, ++it) {
        const int np = min(kNP, p.Npad - ps * kNP);
        const uint32_t buf = (uint32_t)it & 1u;
        mbar_wait(bar(B_ACCFULL + buf), (it >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          const int64_t grow = (int64_t)pair * kPairM + t * kTileM + row;
          const bool rv = grow < p.R;
          float* crow = p.C + grow * p.ldc;
          const float* rrow = p.res ? p.res + grow * p.ldres : nullptr;
#pragma unroll 1
          for (int j = 0; j < (np >> 4); ++j) {
            uint32_t v[16];
            tmem_ld16(tlane + buf * 256u + (uint32_t)t * 128u + (uint32_t)(j * 16), v);
            const int n0 = ps * kNP + j * 16;
            float r[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) r[q] = 0.f;
            if (rv && rrow != nullptr) {
              if (n0 + 16 <= p.N) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float4 x = __ldg(reinterpret_cast<const float4*>(rrow + n0 + 4 * q));
                  r[4 * q] = x.x; r[4 * q + 1] = x.y; r[4 * q + 2] = x.z; r[4 * q + 3] = x.w;
                }
              } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) if (n0 + q < p.N) r[q] = __ldg(rrow + n0 + q);
              }
            }
            tmem_wait_ld();
            if (rv) {
              float o[16];
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                const int n = n0 + q;
                float z = __uint_as_float(v[q]) + r[q];
                if (p.bias != nullptr && n < p.N) z += __ldg(p.bias + n);
                o[q] = n < p.N ? act_apply(p.act, p.act_param, z) : 0.f;
              }
              if (n0 + 16 <= p.ldc_pad) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  *reinterpret_cast<float4*>(crow + n0 + 4 * q) = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
              } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) if (n0 + q < p.ldc_pad) crow[n0 + q] = o[q];
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(bar(B_ACCFREE + buf));
      }
  } else if (warp >= 8) {
    // ===================== A producers: thread = row of the 256-row pair =====================
    const int r = threadIdx.x - 256;                  // 0..255
    const int t = r >> 7, rl = r & 127;
    uint32_t ks = 0;
    for (int pair = blockIdx.x; pair < p.n_pairs; pair += gridDim.x) {
      const int64_t grow = (int64_t)pair * kPairM + r;
      const bool rv = grow < p.R;
      const float* arow = nullptr;
      if (rv) arow = p.A + (p.idx ? (int64_t)__ldg(p.idx + grow) : grow) * p.lda;
      for (int ps = 0; ps < p.n_pass; ++ps)
        for (int s = 0; s < p.nslab; ++s, ++ks) {
          // the row segment of this slab: 8 x 16 B, all loads in flight before the stage is waited for
          float4 x[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int k0 = s * 32 + c * 4;
            x[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rv && k0 < p.K) x[c] = __ldg(reinterpret_cast<const float4*>(arow + k0));   // K % 4 == 0
          }
          const uint32_t st = ks % kStages, use = ks / kStages;
          mbar_wait(bar(B_EMPTY + st), (use & 1) ^ 1);
          const uint32_t stg = sbase + st * kStageBytes + (uint32_t)t * kABytes;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 hi, lo;
            split4(x[c], hi, lo);
            const uint32_t off = sw128_off(rl, c);
            sts128(stg + kOffAhi + off, hi);
            sts128(stg + kOffAlo + off, lo);
          }
          fence_proxy_async();                        // generic-proxy writes -> visible to the tensor core's async proxy
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(B_FULL + st));
        }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// W (f32, N x K row-major, or its transpose) -> per (pass, slab): {hi block, lo block}, each 128 rows x 128 B, K-major
// SWIZZLE_128B; rows >= N and columns >= K are zero
struct PackGeom {
  int N, K, Npad, nslab, n_pass;
};
__host__ __device__ inline PackGeom pack_geom(int64_t N, int64_t K) {
  PackGeom g;
  g.N = (int)N; g.K = (int)K;
  g.Npad = (int)((N + 15) / 16 * 16);
  g.nslab = (int)((K + 31) / 32);
  g.n_pass = (g.Npad + kNP - 1) / kNP;
  return g;
}
__global__ void k_pack_weight_x3(const float* __restrict__ W, int64_t ldw, int transpose, PackGeom g, uint8_t* __restrict__ out) {
  const int n = blockIdx.x;                 // 0 .. n_pass * 128 - 1
  const int ps = n / kNP, nl = n % kNP;
  for (int k = threadIdx.x; k < g.nslab * 32; k += blockDim.x) {
    const int s = k >> 5, kl = k & 31;
    float v = 0.f;
    if (n < g.N && k < g.K) v = transpose ? W[(int64_t)k * ldw + n] : W[(int64_t)n * ldw + k];
    uint32_t hi, lo;
    split1(v, hi, lo);
    const size_t blk = ((size_t)ps * g.nslab + s) * (2 * kWBytes);
    const size_t off = (size_t)(nl >> 3) * 1024 + (nl & 7) * 128 + (((kl >> 2) ^ (nl & 7)) << 4) + (kl & 3) * 4;
    *reinterpret_cast<uint32_t*>(out + blk + off) = hi;
    *reinterpret_cast<uint32_t*>(out + blk + kWBytes + off) = lo;
  }
}
}  // namespace lin

// ------------------------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------------------------
namespace wg {
constexpr int kRows = 32;                      // contraction rows per stage
constexpr int kMT = 128, kKT = 256;            // output tile: n x k
constexpr int kAtomBytes = kRows * 128;        // one 32-column MN block x 32 rows: 4 KB
constexpr int kABytes = (kMT / 32) * kAtomBytes;     // 16 KB per {hi, lo}
constexpr int kBBytes = (kKT / 32) * kAtomBytes;     // 32 KB per {hi, lo}
constexpr int kOffAhi = 0, kOffAlo = kABytes, kOffBhi = 2 * kABytes, kOffBlo = 2 * kABytes + kBBytes;
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;   // 96 KB
constexpr int kStages = 2;
constexpr int kThreads = 512;
constexpr int kProducerWarps = 8;
constexpr int kTmemCols = 256;
constexpr int kOffBar = kStages * kStageBytes;
constexpr int kOffTmem = kOffBar + 8 * 8;
constexpr int kSmemBytes = kOffTmem + 16;
constexpr int kSmemAlloc = kSmemBytes + 1024;
static_assert(kSmemAlloc <= 232448, "exceeds shared memory");
enum { B_FULL = 0, B_EMPTY = 2, B_ACCFULL = 4 };

struct Params {
  const float* dY;
  const float* X;
  int64_t lddy, ldx, R;
  float* partial;            // [grid][128][256]
  int N, K, n_mt, n_kt, slots, n_blocks;
};

__global__ void __launch_bounds__(kThreads, 1) k_wgrad_x3(Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t sBar = sbase + kOffBar;
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + kOffTmem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar = [&](int i) { return sBar + 8u * (uint32_t)i; };
  const int tile = blockIdx.x % (p.n_mt * p.n_kt), slot = blockIdx.x / (p.n_mt * p.n_kt);
  const int mt = tile % p.n_mt, kt = tile / p.n_mt;
  const int n_base = mt * kMT, k_base = kt * kKT;
  const int kw = min(kKT, (p.K - k_base + 15) / 16 * 16);      // accumulator columns in use (multiple of 16)

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar(B_FULL + i), kProducerWarps);
      mbar_init(bar(B_EMPTY + i), 1);
    }
    mbar_init(bar(B_ACCFULL), 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(const_cast<uint32_t*>(s_tmem)), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 1) {
    const uint32_t idesc = idesc_tf32(kMT, kw, true);
    uint32_t ks = 0;
    for (int blk = slot; blk < p.n_blocks; blk += p.slots, ++ks) {
      const uint32_t st = ks % kStages, use = ks / kStages;
      mbar_wait(bar(B_FULL + st), use & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t stg = sbase + st * kStageBytes;
#pragma unroll
        for (int kk = 0; kk < kRows / 8; ++kk) {        // k step = 8 contraction rows = one 1024-byte group of every atom
          const uint64_t ahi = desc_mn_sw128_32b(stg + kOffAhi + kk * 1024, kAtomBytes), alo = desc_mn_sw128_32b(stg + kOffAlo + kk * 1024, kAtomBytes);
          const uint64_t bhi = desc_mn_sw128_32b(stg + kOffBhi + kk * 1024, kAtomBytes), blo = desc_mn_sw128_32b(stg + kOffBlo + kk * 1024, kAtomBytes);
          umma_tf32(tmem_base, alo, bhi, idesc, (ks > 0 || kk > 0) ? 1u : 0u);
          umma_tf32(tmem_base, ahi, blo, idesc, 1u);
          umma_tf32(tmem_base, ahi, bhi, idesc, 1u);
        }
        umma_commit(bar(B_EMPTY + st));
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(bar(B_ACCFULL));
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // final drain: thread = output row n (TMEM lane)
    const int row = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    float* out = p.partial + ((size_t)blockIdx.x * kMT + row) * kKT;
    const bool have = slot < p.n_blocks;
    mbar_wait(bar(B_ACCFULL), 0);
    tc_fence_after();
    for (int j = 0; j < (kw >> 4); ++j) {
      uint32_t v[16];
      if (have) {
        tmem_ld16(tlane + (uint32_t)(j * 16), v);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = 0u;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(out + j * 16 + q * 4) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
  } else if (warp >= 8) {
    // producers: 32 rows x (128 dY columns + 256 X columns) = 3072 16-byte chunks per stage, 12 per thread,
    // consecutive threads on consecutive chunks of a row (coalesced), split into hi / lo, MN-major atom layout
    const int tid = threadIdx.x - 256;
    uint32_t ks = 0;
    for (int blk = slot; blk < p.n_blocks; blk += p.slots, ++ks) {
      float4 x[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const int q = tid + 256 * i;
        const int rr = q / 96, cc = q % 96;
        const int64_t grow = (int64_t)blk * kRows + rr;
        x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (grow < p.R) {
          if (cc < 32) {
            const int n = n_base + cc * 4;
            if (n < p.N) x[i] = __ldg(reinterpret_cast<const float4*>(p.dY + grow * p.lddy + n));     // N % 4 == 0
          } else {
            const int k = k_base + (cc - 32) * 4;
            if (k < p.K) x[i] = __ldg(reinterpret_cast<const float4*>(p.X + grow * p.ldx + k));        // K % 4 == 0
          }
        }
      }
      const uint32_t st = ks % kStages, use = ks / kStages;
      mbar_wait(bar(B_EMPTY + st), (use & 1) ^ 1);
      const uint32_t stg = sbase + st * kStageBytes;
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const int q = tid + 256 * i;
        const int rr = q / 96, cc = q % 96;
        uint4 hi, lo;
        split4(x[i], hi, lo);
        const int cb = cc < 32 ? cc : cc - 32;            // 16-byte chunk along MN inside the operand
        const uint32_t off = (uint32_t)(cb >> 3) * kAtomBytes + mn32b_off(rr, cb & 7);
        if (cc < 32) {
          sts128(stg + kOffAhi + off, hi);
          sts128(stg + kOffAlo + off, lo);
        } else {
          sts128(stg + kOffBhi + off, hi);
          sts128(stg + kOffBlo + off, lo);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_FULL + st));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// dW[n, k] (+)= sum_slot partial[(slot * n_tiles + kt * n_mt + mt)][n % 128][k % 256]   (fixed order: deterministic)
__global__ void k_wgrad_x3_reduce(const float* __restrict__ partial, int slots, int n_mt, int n_kt, int N, int K,
                                  float* __restrict__ dW, int64_t lddw, int accumulate) {
  const int n = blockIdx.x;
  const int mt = n >> 7, nl = n & 127;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const int kt = k >> 8, kl = k & 255;
    float s = 0.f;
    for (int sl = 0; sl < slots; ++sl)
      s += partial[((size_t)(sl * n_mt * n_kt + kt * n_mt + mt) * kMT + nl) * kKT + kl];
    float* o = dW + (int64_t)n * lddw + k;
    *o = accumulate ? (*o + s) : s;
  }
}

struct Geom {
  int n_mt, n_kt, slots, n_blocks, grid;
};
static Geom geom(int64_t R, int64_t N, int64_t K) {
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  Geom g;
  g.n_mt = (int)((N + kMT - 1) / kMT);
  g.n_kt = (int)((K + kKT - 1) / kKT);
  const int64_t nb = (R + kRows - 1) / kRows;
  g.n_blocks = nb > 0x7fffffff ? 0x7fffffff : (int)nb;
  int slots = sm_count / (g.n_mt * g.n_kt);
  if (slots < 1) slots = 1;
  if (slots > g.n_blocks) slots = g.n_blocks > 0 ? g.n_blocks : 1;
  g.slots = slots;
  g.grid = g.n_mt * g.n_kt * slots;
  return g;
}
// the workspace must not depend on R (callers size it once): slots <= SM count / tiles
static size_t workspace_bytes(int64_t N, int64_t K) {
  Geom g = geom((int64_t)1 << 30, N, K);     // any R large enough to give every tile its full set of row slots
  return (size_t)g.grid * kMT * kKT * sizeof(float);
}
}  // namespace wg

}  // namespace x3
}  // namespace dmpnn

using namespace dmpnn;

extern "C" int dmpnn_pack_weight_x3_bytes(int64_t N, int64_t K, size_t* bytes) {
  DMPNN_CHECK_ARG(bytes && N > 0 && K > 0 && N <= 4096 && K <= 8192, "pack_weight_x3: need 0 < N <= 4096, 0 < K <= 8192");
  x3::lin::PackGeom g = x3::lin::pack_geom(N, K);
  *bytes = (size_t)g.n_pass * g.nslab * 2 * x3::lin::kWBytes;
  return 0;
}

extern "C" int dmpnn_pack_weight_x3(const float* W, int64_t ldw, int64_t N, int64_t K, int transpose, void* Wpk, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DMPNN_CHECK_ARG(W && Wpk && N > 0 && K > 0 && N <= 4096 && K <= 8192, "pack_weight_x3: bad args");
  x3::lin::PackGeom g = x3::lin::pack_geom(N, K);
  x3::lin::k_pack_weight_x3<<<g.n_pass * x3::lin::kNP, 128, 0, st>>>(W, ldw, transpose, g, (uint8_t*)Wpk);
  DMPNN_CHECK_LAUNCH("pack_weight_x3", 1);
  return 0;
}

extern "C" int dmpnn_linear_x3(const float* A, int64_t lda, const int32_t* row_idx, int64_t R, int64_t K, const void* Wpk,
                               int64_t N, const float* bias, const float* res, int64_t ldres, int act, float act_param,
                               float* Cout, int64_t ldc, int64_t ldc_pad, void* stream_) {
  using namespace x3::lin;
  cudaStream_t st = (cudaStream_t)stream_;
  DMPNN_CHECK_ARG(R >= 0 && K > 0 && N > 0 && N <= 4096 && K <= 8192, "linear_x3: unsupported sizes K=%lld N=%lld", (long long)K, (long long)N);
  if (R == 0) return 0;
  DMPNN_CHECK_ARG(A && Wpk && Cout, "linear_x3: null pointer");
  DMPNN_CHECK_ARG(K % 4 == 0 && lda % 4 == 0 && lda >= K && (reinterpret_cast<uintptr_t>(A) & 15) == 0,
                  "linear_x3: A needs K %% 4 == 0, lda %% 4 == 0 and a 16-byte aligned base (K=%lld lda=%lld)", (long long)K, (long long)lda);
  DMPNN_CHECK_ARG(ldc % 4 == 0 && ldc >= N && ldc_pad >= N && ldc_pad <= ldc && (reinterpret_cast<uintptr_t>(Cout) & 15) == 0,
                  "linear_x3: C needs ldc %% 4 == 0, N <= ldc_pad <= ldc and a 16-byte aligned base");
  DMPNN_CHECK_ARG(res == nullptr || (ldres % 4 == 0 && ldres >= N && (reinterpret_cast<uintptr_t>(res) & 15) == 0 && res != Cout),
                  "linear_x3: residual needs ldres %% 4 == 0, 16-byte alignment, and may not alias C");
  DMPNN_CHECK_ARG((reinterpret_cast<uintptr_t>(Wpk) & 15) == 0, "linear_x3: packed weight must be 16-byte aligned");
  DMPNN_CHECK_ARG(act >= DMPNN_ACT_NONE && act <= DMPNN_ACT_ELU, "linear_x3: bad activation %d", act);
  PackGeom g = pack_geom(N, K);
  Params p;
  p.A = A; p.idx = row_idx; p.lda = lda;
  p.Wpk = (const uint8_t*)Wpk;
  p.bias = bias; p.res = res; p.ldres = ldres;
  p.C = Cout; p.ldc = ldc; p.R = R;
  p.K = (int)K; p.N = (int)N; p.Npad = g.Npad;
  p.ldc_pad = (int)ldc_pad;
  p.nslab = g.nslab; p.n_pass = g.n_pass;
  p.n_pairs = (int)((R + kPairM - 1) / kPairM);
  p.act = act; p.act_param = act_param;
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_linear_x3, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc);
    DMPNN_CHECK_ARG(e == cudaSuccess, "linear_x3: cannot configure %d B dynamic smem: %s", kSmemAlloc, cudaGetErrorString(e));
    attr_set = true;
  }
  const int grid = p.n_pairs < sm_count ? p.n_pairs : sm_count;
  k_linear_x3<<<grid, kThreads, kSmemAlloc, st>>>(p);
  DMPNN_CHECK_LAUNCH("linear_x3", 1);
  return 0;
}

extern "C" int dmpnn_wgrad_x3_workspace_bytes(int64_t N, int64_t K, size_t* bytes) {
  DMPNN_CHECK_ARG(bytes && N > 0 && K > 0 && N <= 4096 && K <= 8192, "wgrad_x3_workspace_bytes: bad args");
  *bytes = x3::wg::workspace_bytes(N, K) + 256;
  return 0;
}

extern "C" int dmpnn_wgrad_x3(const float* dY, int64_t lddy, const float* X, int64_t ldx, int64_t R, int64_t N, int64_t K,
                              float* dW, int64_t lddw, int accumulate, void* workspace, void* stream_) {
  using namespace x3::wg;
  cudaStream_t st = (cudaStream_t)stream_;
  DMPNN_CHECK_ARG(R >= 0 && N > 0 && K > 0 && N <= 4096 && K <= 8192, "wgrad_x3: unsupported sizes N=%lld K=%lld", (long long)N, (long long)K);
  DMPNN_CHECK_ARG(dW && workspace && lddw >= K, "wgrad_x3: null pointer / lddw");
  DMPNN_CHECK_ARG(R == 0 || (dY && X), "wgrad_x3: null operand");
  DMPNN_CHECK_ARG(N % 4 == 0 && K % 4 == 0 && lddy % 4 == 0 && ldx % 4 == 0 && lddy >= N && ldx >= K &&
                      (reinterpret_cast<uintptr_t>(dY) & 15) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0,
                  "wgrad_x3: N, K, lddy, ldx must be multiples of 4 and the operands 16-byte aligned");
  if (R == 0) {
    if (!accumulate) cudaMemset2DAsync(dW, (size_t)lddw * 4, 0, (size_t)K * 4, (size_t)N, st);
    return 0;
  }
  Geom g = geom(R, N, K);
  Params p;
  p.dY = dY; p.X = X; p.lddy = lddy; p.ldx = ldx; p.R = R;
  p.partial = (float*)workspace;
  p.N = (int)N; p.K = (int)K;
  p.n_mt = g.n_mt; p.n_kt = g.n_kt; p.slots = g.slots; p.n_blocks = g.n_blocks;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_wgrad_x3, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc);
    DMPNN_CHECK_ARG(e == cudaSuccess, "wgrad_x3: cannot configure %d B dynamic smem: %s", kSmemAlloc, cudaGetErrorString(e));
    attr_set = true;
  }
  k_wgrad_x3<<<g.grid, kThreads, kSmemAlloc, st>>>(p);
  k_wgrad_x3_reduce<<<(unsigned)N, 256, 0, st>>>(p.partial, g.slots, g.n_mt, g.n_kt, (int)N, (int)K, dW, lddw, accumulate);
  DMPNN_CHECK_LAUNCH("wgrad_x3", 2);
  return 0;
}
