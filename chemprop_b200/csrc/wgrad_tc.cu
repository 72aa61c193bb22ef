// Tensor-core weight gradient of the bf16 tier:  dW[n, k] (+)= sum_r dY[r, n] * X[r, k]
// (autograd mirror of the nn.Linear layers W_i / W_h / W_o: base.py:135-141, 180-182, mixins.py:8-9).
//
// The contraction runs over ROWS, so both operands are used "MN-major": a TMA box of 64 rows x 64
// columns (SWIZZLE_128B) of the row-major dY / X matrices is exactly the canonical MN-major SW128 atom
// stack (row r = K index at 128-byte pitch, 64 contiguous columns = MN index), no transposes needed.
//
//   CTA (mt, slot): output rows n in [128*mt, 128*mt+128), all k; loops over 64-row blocks slot, slot+S, ...
//     warp 0  TMA producer: 2 dY boxes + ceil(Kpad/64) X boxes per 64-row stage, 3-stage ring
//     warp 1  tcgen05.mma issuer: D[128 x Kpad] (TMEM, fp32) += dY_blk^T . X_blk   (a_major = b_major = MN)
//     warp 2  TMEM allocator
//     warps 4-7  after the last block: TMEM -> f32 partial[cta][128][Kpad] in global memory
//   a second kernel adds the partials of each mt in fixed order (deterministic) into dW.
// Up to three dY TERMS may share one launch (dW = sum_t dY_t^T X): the X boxes of a stage are loaded once and every term
// accumulates into the same TMEM tile -- the W_i gradient of the bf16 mirror, whose dH_0 is a sum of per-step terms that
// is never formed (engine.bond_backward_tc), reads X_0 once instead of once per term.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace dmpnn {
namespace wgtc {

using namespace dmpnn::tc;

constexpr int kRowsPerStage = 64;
constexpr int kBoxBytes = kRowsPerStage * 128;    // 64 rows x 64 bf16 = 8 KB
constexpr int kMaxXBoxes = 7;                     // K <= 448 (TMEM: 448 of 512 accumulator columns)
constexpr int kStageBytes = (2 + kMaxXBoxes) * kBoxBytes;   // 72 KB
constexpr int kStages = 3;
constexpr int kThreads = 256;
constexpr int kTmemCols = 512;
constexpr int kOffBar = kStages * kStageBytes;    // 196608
constexpr int kOffTmem = kOffBar + 8 * 8;
constexpr int kSmemBytes = kOffTmem + 16;
constexpr int kSmemAlloc = kSmemBytes + 1024;
static_assert(kSmemAlloc <= 232448, "exceeds shared memory");

enum { B_FULL = 0, B_EMPTY = 3, B_ACCFULL = 6 };

struct Params {
  float* partial;       // [gridDim.x][128][Kpad]
  int64_t R;
  int N, Kx, Kpad, nxb, n_mt, slots, n_blocks, n_terms;
};

// MN-major SWIZZLE_128B operand: rows (K index) at 128 B pitch, 8-row groups 1024 B apart (SBO),
// 64-column MN blocks `lbo` bytes apart (LBO)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t umma_idesc_bf16_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(kThreads, 1)
k_wgrad_tc(const __grid_constant__ CUtensorMap tmapY, const __grid_constant__ CUtensorMap tmapY1,
           const __grid_constant__ CUtensorMap tmapY2, const __grid_constant__ CUtensorMap tmapX, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t sBar = sbase + kOffBar;
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + kOffTmem);
  const int warp = threadIdx.x >> 5;
  auto bar = [&](int i) { return sBar + 8u * (uint32_t)i; };
  const int mt = blockIdx.x % p.n_mt, slot = blockIdx.x / p.n_mt;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar(B_FULL + i), 1);
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
  const int nt = p.n_terms;
  const uint32_t stage_tx = (uint32_t)(2 * nt + p.nxb) * kBoxBytes;
  const uint32_t x_off = (uint32_t)(2 * nt) * kBoxBytes;      // a stage = [2 boxes per term | X boxes]

  if (warp == 0) {
    uint32_t ks = 0;
    for (int blk = slot; blk < p.n_blocks; blk += p.slots, ++ks) {
      const uint32_t st = ks % kStages, use = ks / kStages;
      mbar_wait(bar(B_EMPTY + st), (use & 1) ^ 1);
      if (elect_one()) {
        const uint32_t base = sbase + st * kStageBytes;
        mbar_expect_tx(bar(B_FULL + st), stage_tx);
        for (int t = 0; t < nt; ++t) {
          const CUtensorMap* my = t == 0 ? &tmapY : (t == 1 ? &tmapY1 : &tmapY2);
          tma_load_2d(base + (2 * t) * kBoxBytes, my, bar(B_FULL + st), mt * 128, blk * kRowsPerStage);
          tma_load_2d(base + (2 * t + 1) * kBoxBytes, my, bar(B_FULL + st), mt * 128 + 64, blk * kRowsPerStage);
        }
        for (int xb = 0; xb < p.nxb; ++xb)
          tma_load_2d(base + x_off + xb * kBoxBytes, &tmapX, bar(B_FULL + st), xb * 64, blk * kRowsPerStage);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const int n0 = p.Kpad <= 256 ? p.Kpad : 256;
    const int n1 = p.Kpad - n0;
    const uint32_t idesc0 = umma_idesc_bf16_mn(128, n0);
    const uint32_t idesc1 = umma_idesc_bf16_mn(128, n1 > 0 ? n1 : 16);
    uint32_t ks = 0;
    bool any = false;
    for (int blk = slot; blk < p.n_blocks; blk += p.slots, ++ks) {
      const uint32_t st = ks % kStages, use = ks / kStages;
      mbar_wait(bar(B_FULL + st), use & 1);
      tc_fence_after();
      const uint32_t base = sbase + st * kStageBytes;
      if (elect_one()) {
        for (int t = 0; t < nt; ++t)
          for (int kk = 0; kk < kRowsPerStage / 16; ++kk) {
            const uint32_t acc = (ks > 0 || kk > 0 || t > 0) ? 1u : 0u;
            const uint64_t adesc = umma_desc_mn_sw128(base + (2 * t) * kBoxBytes + kk * 2048, kBoxBytes);
            umma_bf16(tmem_base, adesc, umma_desc_mn_sw128(base + x_off + kk * 2048, kBoxBytes), idesc0, acc);
            if (n1 > 0)
              umma_bf16(tmem_base + 256u, adesc, umma_desc_mn_sw128(base + x_off + 4 * kBoxBytes + kk * 2048, kBoxBytes), idesc1, acc);
          }
        umma_commit(bar(B_EMPTY + st));
      }
      __syncwarp();
      any = true;
    }
    if (elect_one()) umma_commit(bar(B_ACCFULL));
    __syncwarp();
    (void)any;
  } else if (warp >= 4) {
    // final drain: thread == output row n (TMEM lane), 16 k-columns per tcgen05.ld
    const int r = (warp & 3) * 32 + (threadIdx.x & 31);
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    float* out = p.partial + ((size_t)blockIdx.x * 128 + r) * p.Kpad;
    const bool have = slot < p.n_blocks;   // did this CTA accumulate anything?
    mbar_wait(bar(B_ACCFULL), 0);
    tc_fence_after();
    for (int j = 0; j < p.Kpad / 16; ++j) {
      uint32_t v[16];
      if (have) {
        tmem_ld16(taddr + j * 16, v);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = 0u;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(out + j * 16 + q * 4) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// dW[n, k] (+)= sum_slot partial[slot * n_mt + n/128][n % 128][k]
// block = 128 k lanes x 4 slot groups: group g sums slots g, g+4, ... (independent loads, four in flight per thread), the four
// group sums are combined in a fixed order through shared memory -- deterministic, and the ~50 serial loads per output of
// the one-thread-per-output version (9-27 us per call, six calls per step) become ~12.
__global__ void __launch_bounds__(512) k_wgrad_reduce(const float* __restrict__ partial, int slots, int n_mt, int Kpad, int N, int Kx,
                                                      float* __restrict__ dW, int64_t lddw, int accumulate) {
  __shared__ float part[4][128];
  const int n = blockIdx.x;
  const int mt = n >> 7, nl = n & 127;
  const int kx = threadIdx.x & 127, g = threadIdx.x >> 7;
  const int k = blockIdx.y * 128 + kx;
  float s0 = 0.f, s1 = 0.f;
  if (k < Kx) {
    const float* src = partial + ((size_t)mt * 128 + nl) * Kpad + k;
    const size_t step = (size_t)n_mt * 128 * Kpad;
    int sl = g;
    for (; sl + 4 < slots; sl += 8) {
      s0 += src[(size_t)sl * step];
      s1 += src[(size_t)(sl + 4) * step];
    }
    if (sl < slots) s0 += src[(size_t)sl * step];
  }
  part[g][kx] = s0 + s1;
  __syncthreads();
  if (g == 0 && k < Kx) {
    const float t = (part[0][kx] + part[1][kx]) + (part[2][kx] + part[3][kx]);
    float* o = dW + (int64_t)n * lddw + k;
    *o = accumulate ? (*o + t) : t;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
static bool encode_map(EncodeTiledFn enc, CUtensorMap* tmap, const void* base, int64_t cols, int64_t rows, int64_t ld) {
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)kRowsPerStage};
  cuuint32_t estr[2] = {1, 1};
  return enc(tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct Geom {
  int Kpad, nxb, n_mt, slots, grid;
};
static Geom geom(int64_t N, int64_t Kx) {
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  Geom g;
  g.Kpad = (int)((Kx + 15) / 16 * 16);
  g.nxb = (g.Kpad + 63) / 64;
  g.n_mt = (int)((N + 127) / 128);
  g.slots = sm_count / g.n_mt;
  if (g.slots < 1) g.slots = 1;
  g.grid = g.slots * g.n_mt;
  return g;
}

}  // namespace wgtc
}  // namespace dmpnn

using namespace dmpnn;
using namespace dmpnn::wgtc;

extern "C" int dmpnn_wgrad_tc_workspace_bytes(int64_t N, int64_t Kx, size_t* bytes) {
  DMPNN_CHECK_ARG(bytes && N > 0 && N <= 384 && Kx > 0 && Kx <= 64 * kMaxXBoxes, "wgrad_tc: need 0 < N <= 384, 0 < K <= 448");
  Geom g = geom(N, Kx);
  *bytes = (size_t)g.grid * 128 * g.Kpad * sizeof(float);
  return 0;
}

extern "C" int dmpnn_wgrad_tc_multi_bf16(const void* const* dYs, int n_terms, int64_t lddy, const void* X, int64_t ldx, int64_t R,
                                         int64_t N, int64_t Kx, float* dW, int64_t lddw, int accumulate, void* workspace,
                                         void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DMPNN_CHECK_ARG(R >= 0 && N > 0 && N <= 384 && Kx > 0 && Kx <= 64 * kMaxXBoxes, "wgrad_tc: unsupported sizes N=%lld K=%lld",
                  (long long)N, (long long)Kx);
  DMPNN_CHECK_ARG(dW && workspace && dYs, "wgrad_tc: null pointer");
  Geom g = geom(N, Kx);
  DMPNN_CHECK_ARG(n_terms >= 1 && n_terms <= 3 && 2 * n_terms + g.nxb <= 2 + kMaxXBoxes,
                  "wgrad_tc: %d terms x K=%lld do not fit a stage (2 * terms + ceil(K / 64) <= 9)", n_terms, (long long)Kx);
  DMPNN_CHECK_ARG(lddy % 8 == 0 && ldx % 8 == 0 && lddy >= N && ldx >= Kx, "wgrad_tc: lddy/ldx must be multiples of 8");
  for (int t = 0; t < n_terms; ++t)
    DMPNN_CHECK_ARG(R == 0 || (dYs[t] && (reinterpret_cast<uintptr_t>(dYs[t]) & 15) == 0), "wgrad_tc: dY operands must be 16-byte aligned");
  DMPNN_CHECK_ARG(R == 0 || (X && (reinterpret_cast<uintptr_t>(X) & 15) == 0), "wgrad_tc: operands must be 16-byte aligned");
  if (R > 0) {
    EncodeTiledFn enc = get_encode_fn();
    DMPNN_CHECK_ARG(enc != nullptr, "wgrad_tc: cuTensorMapEncodeTiled not available from the driver");
    CUtensorMap mY[3], mX;
    for (int t = 0; t < 3; ++t)
      DMPNN_CHECK_ARG(encode_map(enc, &mY[t], dYs[t < n_terms ? t : 0], N, R, lddy), "wgrad_tc: cuTensorMapEncodeTiled failed");
    DMPNN_CHECK_ARG(encode_map(enc, &mX, X, Kx, R, ldx), "wgrad_tc: cuTensorMapEncodeTiled failed");
    Params p;
    p.partial = (float*)workspace;
    p.R = R;
    p.N = (int)N;
    p.Kx = (int)Kx;
    p.Kpad = g.Kpad;
    p.nxb = g.nxb;
    p.n_mt = g.n_mt;
    p.slots = g.slots;
    p.n_blocks = (int)((R + kRowsPerStage - 1) / kRowsPerStage);
    p.n_terms = n_terms;
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc);
      DMPNN_CHECK_ARG(e == cudaSuccess, "wgrad_tc: cannot configure %d B dynamic smem: %s", kSmemAlloc, cudaGetErrorString(e));
      attr_set = true;
    }
    k_wgrad_tc<<<g.grid, kThreads, kSmemAlloc, st>>>(mY[0], mY[1], mY[2], mX, p);
  } else {
    cudaMemsetAsync(workspace, 0, (size_t)g.grid * 128 * g.Kpad * sizeof(float), st);
  }
  k_wgrad_reduce<<<dim3((unsigned)N, (unsigned)((Kx + 127) / 128)), 512, 0, st>>>((const float*)workspace, g.slots, g.n_mt, g.Kpad, (int)N, (int)Kx, dW, lddw, accumulate);
  DMPNN_CHECK_LAUNCH("wgrad_tc", 2);
  return 0;
}

extern "C" int dmpnn_wgrad_tc_bf16(const void* dY, int64_t lddy, const void* X, int64_t ldx, int64_t R, int64_t N,
                                   int64_t Kx, float* dW, int64_t lddw, int accumulate, void* workspace,
                                   void* stream_) {
  DMPNN_CHECK_ARG(R == 0 || dY, "wgrad_tc: null operand");
  const void* terms[1] = {dY};
  return dmpnn_wgrad_tc_multi_bf16(terms, 1, lddy, X, ldx, R, N, Kx, dW, lddw, accumulate, workspace, stream_);
}
