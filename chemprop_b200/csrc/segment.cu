// Segmented reductions / irregular gathers of the D-MPNN path (f32 accumulate, deterministic):
//   segment_sum   : base.py:208-211 (atom scatter-sum), mixins.py:25-30 (atom neighbour sum),
//                   agg.py:73-78 / 90-95 / 112-113 (Mean / Sum / Norm aggregation)
//   bond_message  : mixins.py:11-18  (M = A[src] - H[rev]) and its autograd mirror
//   rev_average   : base.py:202-203  (undirected)
//   act_bwd       : autograd mirror of tau in base.py:137-139, :181
// The reference builds an (E x h) int64 index with .repeat() and runs scatter_reduce_ (atomics on
// CUDA); here rows are dst-sorted so every reduction is a contiguous segment owned by one warp.
#include "common.cuh"

namespace dmpnn {

template <typename TX, typename TY>
__global__ void k_segment_sum(const TX* __restrict__ X, int64_t ldx, const int32_t* __restrict__ idx,
                              const int32_t* __restrict__ ptr, int64_t n_seg, int C, int act, float ap,
                              int scale_mode, float scale, TY* __restrict__ Y, int64_t ldy, int ldy_pad) {
  const int lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= n_seg) return;
  const int32_t a = ptr[s], b = ptr[s + 1];
  for (int c = lane; c < C; c += 32) {
    float acc = 0.f;
    for (int32_t r = a; r < b; ++r) {
      int64_t rr = idx ? (int64_t)idx[r] : (int64_t)r;
      acc += act_apply(act, ap, ld_as_float(X + rr * ldx + c));
    }
    // mean: sum then divide (scatter_reduce_ "mean" semantics, agg.py:76-78)
    float out = acc;
    if (scale_mode == DMPNN_SCALE_INV_COUNT) out = (b > a) ? acc / (float)(b - a) : 0.f;
    else if (scale_mode == DMPNN_SCALE_DIV_CONST) out = acc / scale;
    st_from_float(Y + s * ldy + c, out);
  }
  for (int c = C + lane; c < ldy_pad; c += 32) st_from_float(Y + s * ldy + c, 0.f);
}

template <typename TG, typename TY>
__global__ void k_segment_bcast(const TG* __restrict__ G, int64_t ldg, const int32_t* __restrict__ seg_of_row,
                                const int32_t* __restrict__ ptr, int64_t R, int C, int scale_mode, float scale,
                                TY* __restrict__ Y, int64_t ldy) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (r >= R || c >= C) return;
  const int32_t s = seg_of_row[r];
  float g = ld_as_float(G + (int64_t)s * ldg + c);
  float out;
  if (scale_mode == DMPNN_SCALE_INV_COUNT) out = g / (float)(ptr[s + 1] - ptr[s]);
  else if (scale_mode == DMPNN_SCALE_DIV_CONST) out = g / scale;
  else out = g;
  st_from_float(Y + r * ldy + c, out);
}

template <typename TX, typename TO>
__global__ void k_bond_message(const TX* __restrict__ X, int64_t ldx, const int32_t* __restrict__ rowptr,
                               const int32_t* __restrict__ rev_row, int64_t V, int C, int act, float ap,
                               int permute_on_read, TO* __restrict__ OUT, int64_t ldo) {
  const int lane = threadIdx.x & 31;
  const int64_t v = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (v >= V) return;
  const int32_t a = rowptr[v], b = rowptr[v + 1];
  if (b <= a) return;
  for (int c = lane; c < C; c += 32) {
    float s = 0.f;
    for (int32_t r = a; r < b; ++r) {
      int64_t rd = permute_on_read ? (int64_t)rev_row[r] : (int64_t)r;
      s += act_apply(act, ap, ld_as_float(X + rd * ldx + c));
    }
    for (int32_t r = a; r < b; ++r) {
      int64_t q = (int64_t)rev_row[r];
      int64_t rd = permute_on_read ? q : (int64_t)r;
      int64_t wr = permute_on_read ? (int64_t)r : q;
      float x = act_apply(act, ap, ld_as_float(X + rd * ldx + c));
      st_from_float(OUT + wr * ldo + c, s - x);
    }
  }
}

template <typename TX, typename TO>
__global__ void k_rev_average(const TX* __restrict__ X, int64_t ldx, const int32_t* __restrict__ rev_row,
                              int64_t R, int C, int act, float ap, TO* __restrict__ OUT, int64_t ldo) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (r >= R || c >= C) return;
  float a = act_apply(act, ap, ld_as_float(X + r * ldx + c));
  float b = act_apply(act, ap, ld_as_float(X + (int64_t)rev_row[r] * ldx + c));
  st_from_float(OUT + r * ldo + c, (a + b) / 2.f);
}

template <typename TG, typename TYA, typename TZ, typename TA>
__global__ void k_act_bwd(const TG* __restrict__ G, int64_t ldg, const int32_t* __restrict__ gidx,
                          const TYA* __restrict__ Yact, int64_t ldy, int from_preact, int act, float ap,
                          TZ* __restrict__ dZ, int64_t lddz, TA* __restrict__ ACC, int64_t ldacc,
                          int64_t R, int C) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (r >= R || c >= C) return;
  int64_t gr = gidx ? (int64_t)gidx[r] : r;
  float g = ld_as_float(G + gr * ldg + c);
  float y = ld_as_float(Yact + r * ldy + c);
  float d = from_preact ? act_grad_from_pre(act, ap, y) : act_grad_from_out(act, ap, y);
  float dz = g * d;
  if (dZ) st_from_float(dZ + r * lddz + c, dz);
  if (ACC) {
    TA* p = ACC + r * ldacc + c;
    st_from_float(p, ld_as_float(p) + dz);
  }
}


// ---- 4-wide variants (used when every pointer / row stride is 4-element aligned and C % 4 == 0) --------
template <typename TX, typename TY>
__global__ void k_segment_sum_v4(const TX* __restrict__ X, int64_t ldx, const int32_t* __restrict__ idx,
                                 const int32_t* __restrict__ ptr, int64_t n_seg, int Q, int act, float ap,
                                 int scale_mode, float scale, TY* __restrict__ Y, int64_t ldy, int Qpad) {
  const int lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= n_seg) return;
  const int32_t a = ptr[s], b = ptr[s + 1];
  for (int q = lane; q < Qpad; q += 32) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (q < Q) {
      int32_t r = a;
      for (; r + 4 <= b; r += 4) {     // four independent row loads in flight (long segments: molecules)
        float v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t rr = idx ? (int64_t)idx[r + u] : (int64_t)(r + u);
          ld4(X + rr * ldx + 4 * q, v[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] += act_apply(act, ap, v[u][i]);
      }
      for (; r < b; ++r) {
        const int64_t rr = idx ? (int64_t)idx[r] : (int64_t)r;
        float v[4];
        ld4(X + rr * ldx + 4 * q, v);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] += act_apply(act, ap, v[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (scale_mode == DMPNN_SCALE_INV_COUNT) acc[i] = (b > a) ? acc[i] / (float)(b - a) : 0.f;
        else if (scale_mode == DMPNN_SCALE_DIV_CONST) acc[i] = acc[i] / scale;
      }
    }
    st4(Y + s * ldy + 4 * q, acc);
  }
}

// flat segmented sum for SHORT segments (atoms: 2-4 in-edge rows): thread = one 8-column chunk of `par` interleaved
// segments, so the warps stay full (a warp-per-segment mapping has 3 passes of 32/32/11 lanes at h = 300) and every
// row of a segment is an independent 16-byte load.  Columns in [C, QW*8) are written as zeros.
template <bool HAS_ACT, typename TX, typename TY>
__global__ void __launch_bounds__(256)
k_segment_sum_flat8(const TX* __restrict__ X, int64_t ldx, const int32_t* __restrict__ idx,
                    const int32_t* __restrict__ ptr, int64_t n_seg, int C, int QW, int act, float ap, int scale_mode,
                    float scale, TY* __restrict__ Y, int64_t ldy, int segs_per_block) {
  const int par = blockDim.x / QW;
  const int sub = threadIdx.x / QW;
  const int c0 = 8 * (threadIdx.x - sub * QW);
  if (sub >= par) return;
  const int64_t s_end = min(n_seg, ((int64_t)blockIdx.x + 1) * segs_per_block);
  for (int64_t s = (int64_t)blockIdx.x * segs_per_block + sub; s < s_end; s += par) {
    const int32_t a = ptr[s], b = ptr[s + 1];
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    if (c0 < C) {
      int32_t r = a;
      for (; r + 4 <= b; r += 4) {
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t rr = idx ? (int64_t)idx[r + u] : (int64_t)(r + u);
          ldv<8>(X + rr * ldx + c0, v[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += HAS_ACT ? act_apply(act, ap, v[u][i]) : v[u][i];
      }
      if (r < b) {                        // 1..3 remaining rows, loaded together
        float v[3][8];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int32_t ru = min(r + u, b - 1);
          const int64_t rr = idx ? (int64_t)idx[ru] : (int64_t)ru;
          ldv<8>(X + rr * ldx + c0, v[u]);
        }
#pragma unroll
        for (int u = 0; u < 3; ++u)
          if (r + u < b) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += HAS_ACT ? act_apply(act, ap, v[u][i]) : v[u][i];
          }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (scale_mode == DMPNN_SCALE_INV_COUNT) acc[i] = (b > a) ? acc[i] / (float)(b - a) : 0.f;
        else if (scale_mode == DMPNN_SCALE_DIV_CONST) acc[i] = acc[i] / scale;
        if (c0 + i >= C) acc[i] = 0.f;    // the chunk straddling C: whatever the source padding held is dropped
      }
    }
    stv<8>(Y + s * ldy + c0, acc);
  }
}

// thread = one column quad of `rows_par` interleaved rows (one integer division per thread, not per element)
template <typename TG, typename TY>
__global__ void __launch_bounds__(256)
k_segment_bcast_v4(const TG* __restrict__ G, int64_t ldg, const int32_t* __restrict__ seg_of_row,
                   const int32_t* __restrict__ ptr, int64_t R, int Q, int scale_mode, float scale,
                   TY* __restrict__ Y, int64_t ldy, int rows_per_block) {
  const int rows_par = blockDim.x / Q;
  const int rsub = threadIdx.x / Q;
  const int q = threadIdx.x - rsub * Q;
  if (rsub >= rows_par) return;
  const int64_t r_end = min(R, ((int64_t)blockIdx.x + 1) * rows_per_block);
  for (int64_t r = (int64_t)blockIdx.x * rows_per_block + rsub; r < r_end; r += rows_par) {
    const int32_t s = seg_of_row[r];
    float v[4];
    ld4(G + (int64_t)s * ldg + 4 * q, v);
    if (scale_mode != DMPNN_SCALE_NONE) {
      const float div = (scale_mode == DMPNN_SCALE_INV_COUNT) ? (float)(ptr[s + 1] - ptr[s]) : scale;
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = v[k] / div;
    }
    st4(Y + r * ldy + 4 * q, v);
  }
}

// segment walk: one block per segment, thread = (row lane, column quad); the scaled row of G stays in registers
template <typename TG, typename TY>
__global__ void __launch_bounds__(256)
k_segment_bcast_seg_v4(const TG* __restrict__ G, int64_t ldg, const int32_t* __restrict__ ptr, int Q, int scale_mode,
                       float scale, TY* __restrict__ Y, int64_t ldy) {
  const int rows_par = blockDim.x / Q;
  const int rsub = threadIdx.x / Q;
  const int q = threadIdx.x - rsub * Q;
  if (rsub >= rows_par) return;
  const int64_t s = blockIdx.x;
  const int32_t a = ptr[s], b = ptr[s + 1];
  if (b <= a) return;
  float v[4];
  ld4(G + s * ldg + 4 * q, v);
  if (scale_mode != DMPNN_SCALE_NONE) {
    const float div = (scale_mode == DMPNN_SCALE_INV_COUNT) ? (float)(b - a) : scale;
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = v[k] / div;
  }
  for (int32_t r = a + rsub; r < b; r += rows_par) st4(Y + (int64_t)r * ldy + 4 * q, v);
}

template <int VW, typename TX, typename TO>
__global__ void k_bond_message_v4(const TX* __restrict__ X, int64_t ldx, const int32_t* __restrict__ rowptr,
                                  const int32_t* __restrict__ rev_row, int64_t V, int Q, int act, float ap,
                                  int permute_on_read, TO* __restrict__ OUT, int64_t ldo,
                                  const TO* __restrict__ Ymask, int64_t ldm, int mask_act) {
  const int lane = threadIdx.x & 31;
  const int64_t v = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (v >= V) return;
  const int32_t a = rowptr[v], b = rowptr[v + 1];
  if (b <= a) return;
  for (int q = lane; q < Q; q += 32) {
    float s[VW];
#pragma unroll
    for (int i = 0; i < VW; ++i) s[i] = 0.f;
    for (int32_t r = a; r < b; ++r) {
      const int64_t rd = permute_on_read ? (int64_t)rev_row[r] : (int64_t)r;
      float x[VW];
      ldv<VW>(X + rd * ldx + VW * q, x);
#pragma unroll
      for (int i = 0; i < VW; ++i) s[i] += act_apply(act, ap, x[i]);
    }
    for (int32_t r = a; r < b; ++r) {
      const int64_t qq = (int64_t)rev_row[r];
      const int64_t rd = permute_on_read ? qq : (int64_t)r;
      const int64_t wr = permute_on_read ? (int64_t)r : qq;
      float x[VW], o[VW];
      ldv<VW>(X + rd * ldx + VW * q, x);
#pragma unroll
      for (int i = 0; i < VW; ++i) o[i] = s[i] - act_apply(act, ap, x[i]);
      if (Ymask) {   // fused tau'(.) of the autograd mirror: OUT = (sum - x) * tau'(Y[wr])
        float y[VW];
        ldv<VW>(Ymask + wr * ldm + VW * q, y);
#pragma unroll
        for (int i = 0; i < VW; ++i) o[i] *= act_grad_from_out(mask_act, ap, y[i]);
      }
      stv<VW>(OUT + wr * ldo + VW * q, o);
    }
  }
}

// thread = one VW-element column chunk of several rows: `rows_par` rows side by side in a block (all lanes busy at h = 300:
// 38 chunks x 6 rows of 256 threads), each thread walking its rows with FOUR rows' loads in flight (the row gather through
// `gidx` and the two streams are latency-bound with one row per thread); 32-bit index arithmetic, no per-element division
template <int VW, typename TG, typename TYA, typename TZ, typename TA>
__global__ void __launch_bounds__(256)
k_act_bwd_v4(const TG* __restrict__ G, int64_t ldg, const int32_t* __restrict__ gidx,
             const TYA* __restrict__ Yact, int64_t ldy, int from_preact, int act, float ap,
             TZ* __restrict__ dZ, int64_t lddz, TA* __restrict__ ACC, int64_t ldacc, int64_t R, int Q, int rows_per_block) {
  const int rows_par = blockDim.x / Q;
  const int rsub = threadIdx.x / Q;
  const int c = VW * (threadIdx.x - rsub * Q);
  if (rsub >= rows_par) return;
  const int64_t r_end = min(R, ((int64_t)blockIdx.x + 1) * rows_per_block);
  int64_t r = (int64_t)blockIdx.x * rows_per_block + rsub;
  auto finish = [&](int64_t rr, const float (&g)[VW], const float (&y)[VW]) {
    float dz[VW];
#pragma unroll
    for (int k = 0; k < VW; ++k)
      dz[k] = g[k] * (from_preact ? act_grad_from_pre(act, ap, y[k]) : act_grad_from_out(act, ap, y[k]));
    if (dZ) stv<VW>(dZ + rr * lddz + c, dz);
    if (ACC) {
      float a[VW];
      ldv<VW>(ACC + rr * ldacc + c, a);
#pragma unroll
      for (int k = 0; k < VW; ++k) a[k] += dz[k];
      stv<VW>(ACC + rr * ldacc + c, a);
    }
  };
  for (; r + 3 * rows_par < r_end; r += 4 * rows_par) {
    float g[4][VW], y[4][VW];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t rr = r + u * rows_par;
      const int64_t gr = gidx ? (int64_t)__ldg(gidx + rr) : rr;
      ldv<VW>(G + gr * ldg + c, g[u]);
      ldv<VW>(Yact + rr * ldy + c, y[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) finish(r + u * rows_par, g[u], y[u]);
  }
  for (; r < r_end; r += rows_par) {
    float g[VW], y[VW];
    const int64_t gr = gidx ? (int64_t)__ldg(gidx + r) : r;
    ldv<VW>(G + gr * ldg + c, g);
    ldv<VW>(Yact + r * ldy + c, y);
    finish(r, g, y);
  }
}

// ReLU, all-bf16, no accumulator: dZ[r] = G[gidx ? gidx[r] : r] masked by [Y[r] > 0] -- the two instances every bf16 training
// step runs (atom rows after the read-out, edge rows with the atom -> edge broadcast).  Nothing is converted: the mask is
// taken on the packed pairs (`__hgt2_mask`: 0xFFFF per half where y > 0, false for NaN like the scalar test) and ANDed
// into the gradient bits, so a thread keeps U rows of NB-byte chunks in flight in 2 * U * NB / 4 registers.  (The generic
// kernel above spends ~200 instructions per warp and row on its run-time activation switch and conversions: 262 us for
// the 0.76 GB of the bench's edge instance, profiles/r2_glue_ncu.md.)
template <int NB> struct BytesVec;
template <> struct BytesVec<8> { using type = uint2; };
template <> struct BytesVec<16> { using type = uint4; };
__device__ __forceinline__ uint32_t relu_mask_bits(uint32_t g, uint32_t y) {
  __nv_bfloat162 y2;
  memcpy(&y2, &y, 4);
  return g & __hgt2_mask(y2, __float2bfloat162_rn(0.f));
}
__device__ __forceinline__ uint2 relu_mask_bits(uint2 g, uint2 y) { return make_uint2(relu_mask_bits(g.x, y.x), relu_mask_bits(g.y, y.y)); }
__device__ __forceinline__ uint4 relu_mask_bits(uint4 g, uint4 y) {
  return make_uint4(relu_mask_bits(g.x, y.x), relu_mask_bits(g.y, y.y), relu_mask_bits(g.z, y.z), relu_mask_bits(g.w, y.w));
}
template <int NB, int U>
__global__ void __launch_bounds__(256)
k_act_bwd_relu_bf16(const __nv_bfloat16* __restrict__ G, int64_t ldg, const int32_t* __restrict__ gidx,
                    const __nv_bfloat16* __restrict__ Y, int64_t ldy, __nv_bfloat16* __restrict__ dZ, int64_t lddz,
                    int64_t R, int Q, int rows_per_block) {
  using V = typename BytesVec<NB>::type;
  const int rows_par = blockDim.x / Q;
  const int rsub = threadIdx.x / Q;
  const int c = (NB / 2) * (threadIdx.x - rsub * Q);
  if (rsub >= rows_par) return;
  const int64_t r_end = min(R, ((int64_t)blockIdx.x + 1) * rows_per_block);
  int64_t r = (int64_t)blockIdx.x * rows_per_block + rsub;
  for (; r + (U - 1) * rows_par < r_end; r += U * rows_par) {
    int64_t gr[U];
    V g[U], y[U];
#pragma unroll
    for (int u = 0; u < U; ++u) gr[u] = gidx ? (int64_t)__ldg(gidx + r + u * rows_par) : r + u * rows_par;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      g[u] = __ldg(reinterpret_cast<const V*>(G + gr[u] * ldg + c));
      y[u] = __ldcs(reinterpret_cast<const V*>(Y + (r + u * rows_par) * ldy + c));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) *reinterpret_cast<V*>(dZ + (r + u * rows_par) * lddz + c) = relu_mask_bits(g[u], y[u]);
  }
  for (; r < r_end; r += rows_par) {
    const int64_t gr = gidx ? (int64_t)__ldg(gidx + r) : r;
    const V g = __ldg(reinterpret_cast<const V*>(G + gr * ldg + c));
    const V y = __ldcs(reinterpret_cast<const V*>(Y + r * ldy + c));
    *reinterpret_cast<V*>(dZ + r * lddz + c) = relu_mask_bits(g, y);
  }
}

template <typename TX, typename TO>
__global__ void k_rev_average_v4(const TX* __restrict__ X, int64_t ldx, const int32_t* __restrict__ rev_row, int64_t R,
                                 int Q, int act, float ap, TO* __restrict__ OUT, int64_t ldo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * Q) return;
  const int64_t r = i / Q;
  const int q = (int)(i - r * Q);
  float a[4], b[4], o[4];
  ld4(X + r * ldx + 4 * q, a);
  ld4(X + (int64_t)rev_row[r] * ldx + 4 * q, b);
#pragma unroll
  for (int k = 0; k < 4; ++k) o[k] = (act_apply(act, ap, a[k]) + act_apply(act, ap, b[k])) / 2.f;
  st4(OUT + r * ldo + 4 * q, o);
}

// OUT[r] = sum_k Z_k[r] + G[r] * tau'(Ypre[r])  -- the dH_0 total of the autograd mirror in one pass
struct SumSrcs { const void* z[8]; int n; };
template <int VW, typename TZ, typename TO>
__global__ void k_sum_act_bwd_v4(SumSrcs srcs, int64_t ldz, const TZ* __restrict__ G, int64_t ldg,
                                 const TZ* __restrict__ Ypre, int64_t ldy, int act, float ap, TO* __restrict__ OUT,
                                 int64_t ldo, int64_t R, int Q) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * Q) return;
  const int64_t r = i / Q;
  const int q = (int)(i - r * Q);
  float acc[VW];
#pragma unroll
  for (int k = 0; k < VW; ++k) acc[k] = 0.f;
  if (G) {
    float g[VW], y[VW];
    ldv<VW>(G + r * ldg + VW * q, g);
    ldv<VW>(Ypre + r * ldy + VW * q, y);
#pragma unroll
    for (int k = 0; k < VW; ++k) acc[k] = g[k] * act_grad_from_pre(act, ap, y[k]);
  }
  for (int s = 0; s < srcs.n; ++s) {
    float z[VW];
    ldv<VW>(reinterpret_cast<const TZ*>(srcs.z[s]) + r * ldz + VW * q, z);
#pragma unroll
    for (int k = 0; k < VW; ++k) acc[k] += z[k];
  }
  stv<VW>(OUT + r * ldo + VW * q, acc);
}

// OUT[r, :] = bf16([X1[i1(r), 0:K1] || X2[i2(r), 0:K2] || 0 ...])  -- the A operand of the tensor-core
// linear layers: torch.cat([V[src], E], 1) (mixins.py:9) and torch.cat((V, M), 1) (base.py:180)
template <typename T1, typename T2, typename TO = __nv_bfloat16>
__global__ void k_concat_bf16(const T1* __restrict__ X1, int64_t ld1, const int32_t* __restrict__ idx1, int K1,
                              const T2* __restrict__ X2, int64_t ld2, const int32_t* __restrict__ idx2, int K2,
                              TO* __restrict__ OUT, int64_t ldo, int width, int64_t R) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  if (r >= R) return;
  const int64_t r1 = idx1 ? (int64_t)idx1[r] : r;
  const int64_t r2 = (K2 > 0 && idx2) ? (int64_t)idx2[r] : r;
  for (int c = threadIdx.x; c < width; c += blockDim.x) {
    float v = 0.f;
    if (c < K1) v = ld_as_float(X1 + r1 * ld1 + c);
    else if (c < K1 + K2) v = ld_as_float(X2 + r2 * ld2 + (c - K1));
    st_from_float(OUT + r * ldo + c, v);
  }
}

// 4-wide variant: thread = one column quad of `rows_par` interleaved rows, so every lane of a warp stores
// (rows are 18..24 quads wide: a lane-per-quad-of-one-row mapping would idle a quarter of the warp)
template <typename T1, typename T2, typename TO = __nv_bfloat16>
__global__ void __launch_bounds__(256)
k_concat_bf16_v4(const T1* __restrict__ X1, int64_t ld1, const int32_t* __restrict__ idx1, int K1,
                 const T2* __restrict__ X2, int64_t ld2, const int32_t* __restrict__ idx2, int K2,
                 TO* __restrict__ OUT, int64_t ldo, int QW, int64_t R, int rows_per_block, int x1_vec) {
  const int rows_par = blockDim.x / QW;
  const int rsub = threadIdx.x / QW;
  const int c4 = 4 * (threadIdx.x - rsub * QW);
  if (rsub >= rows_par) return;
  const int64_t r_end = min(R, ((int64_t)blockIdx.x + 1) * rows_per_block);
  for (int64_t r = (int64_t)blockIdx.x * rows_per_block + rsub; r < r_end; r += rows_par) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (c4 + 4 <= K1) {               // quad inside the first operand: one row index, no per-element tests
      const int64_t r1 = idx1 ? (int64_t)__ldg(idx1 + r) : r;
      const T1* s1 = X1 + r1 * ld1 + c4;
      if (x1_vec) ld4(s1, v);
      else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = ld_as_float(s1 + i);
      }
    } else if (c4 < K1 + K2) {
      const int64_t r1 = idx1 ? (int64_t)idx1[r] : r;
      const int64_t r2 = (K2 > 0 && idx2) ? (int64_t)idx2[r] : r;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c4 + i;
        if (c < K1) v[i] = ld_as_float(X1 + r1 * ld1 + c);
        else if (c < K1 + K2) v[i] = ld_as_float(X2 + r2 * ld2 + (c - K1));
      }
    }
    st4(OUT + r * ldo + c4, v);
  }
}

}  // namespace dmpnn

using namespace dmpnn;

extern "C" int dmpnn_concat_bf16(const void* X1, int x1_dtype, int64_t ld1, const int32_t* idx1, int64_t K1,
                                 const void* X2, int x2_dtype, int64_t ld2, const int32_t* idx2, int64_t K2,
                                 void* OUT, int64_t ldo, int64_t width, int64_t R, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DMPNN_CHECK_ARG(R >= 0 && K1 > 0 && K2 >= 0 && width >= K1 + K2 && ldo >= width, "concat_bf16: bad sizes");
  if (R == 0) return 0;
  DMPNN_CHECK_ARG(X1 && OUT && (K2 == 0 || X2), "concat_bf16: null pointer");
  if (K2 == 0) x2_dtype = x1_dtype;
  dim3 block(32, 8), grid(ceil_div_i64(R, 8));
  const bool vec = width % 4 == 0 && width / 4 <= 256 && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(OUT) & 7) == 0;
  DMPNN_DISPATCH_DTYPE(x1_dtype, T1,
    DMPNN_DISPATCH_DTYPE(x2_dtype, T2,
      if (vec)
        k_concat_bf16_v4<T1, T2><<<ceil_div_i64(R, 64), 256, 0, st>>>(
            (const T1*)X1, ld1, idx1, (int)K1, (const T2*)X2, ld2, idx2, (int)K2, (__nv_bfloat16*)OUT, ldo, (int)(width / 4), R,
            64, vec4_ok<T1>(X1, ld1) ? 1 : 0);
      else
        k_concat_bf16<T1, T2><<<grid, block, 0, st>>>((const T1*)X1, ld1, idx1, (int)K1, (const T2*)X2, ld2, idx2, (int)K2,
                                                     (__nv_bfloat16*)OUT, ldo, (int)width, R);
    ))
  DMPNN_CHECK_LAUNCH("concat_bf16", 1);
  return 0;
}

// f32 output: the [V[src] || E] / [N || sum E] operands of the fp32 tier's tensor-core GEMMs (dmpnn_linear_x3)
extern "C" int dmpnn_concat_f32(const void* X1, int x1_dtype, int64_t ld1, const int32_t* idx1, int64_t K1,
                                const void* X2, int x2_dtype, int64_t ld2, const int32_t* idx2, int64_t K2,
                                float* OUT, int64_t ldo, int64_t width, int64_t R, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DMPNN_CHECK_ARG(R >= 0 && K1 > 0 && K2 >= 0 && width >= K1 + K2 && ldo >= width, "concat_f32: bad sizes");
  if (R == 0) return 0;
  DMPNN_CHECK_ARG(X1 && OUT && (K2 == 0 || X2), "concat_f32: null pointer");
  if (K2 == 0) x2_dtype = x1_dtype;
  dim3 block(32, 8), grid(ceil_div_i64(R, 8));
  const bool vec = width % 4 == 0 && width / 4 <= 256 && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(OUT) & 15) == 0;
  DMPNN_DISPATCH_DTYPE(x1_dtype, T1,
    DMPNN_DISPATCH_DTYPE(x2_dtype, T2,
      if (vec)
        k_concat_bf16_v4<T1, T2, float><<<ceil_div_i64(R, 64), 256, 0, st>>>(
            (const T1*)X1, ld1, idx1, (int)K1, (const T2*)X2, ld2, idx2, (int)K2, OUT, ldo, (int)(width / 4), R,
            64, vec4_ok<T1>(X1, ld1) ? 1 : 0);
      else
        k_concat_bf16<T1, T2, float><<<grid, block, 0, st>>>((const T1*)X1, ld1, idx1, (int)K1, (const T2*)X2, ld2, idx2, (int)K2,
                                                            OUT, ldo, (int)width, R);
    ))
  DMPNN_CHECK_LAUNCH("concat_f32", 1);
  return 0;
}

extern "C" int dmpnn_segment_sum(const void* X, int x_dtype, int64_t ldx, const int32_t* idx, const int32_t* ptr,
                                 int64_t n_seg, int64_t C, int act, float act_param, int scale_mode, float scale,
                                 void* Y, int y_dtype, int64_t ldy, int64_t ldy_pad, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DMPNN_CHECK_ARG(n_seg >= 0 && C > 0 && ptr && Y, "segment_sum: bad args");
  DMPNN_CHECK_ARG(ldy >= C && ldy_pad <= ldy, "segment_sum: ldy too small");
  if (n_seg == 0) return 0;
  const int warps = 8;
  DMPNN_DISPATCH_DTYPE(x_dtype, TX,
    DMPNN_DISPATCH_DTYPE(y_dtype, TY,
      const int64_t padc = ldy_pad > C ? ldy_pad : C;
      const int64_t c8 = (C + 7) / 8 * 8;
      if (padc % 8 == 0 && padc >= c8 && c8 <= ldx && padc / 8 <= 256 && vec8_ok<TX>(X, ldx) && vec8_ok<TY>(Y, ldy)) {
        const int spb = 64;
        if (act == DMPNN_ACT_NONE)
          k_segment_sum_flat8<false, TX, TY><<<ceil_div_i64(n_seg, spb), 256, 0, st>>>(
              (const TX*)X, ldx, idx, ptr, n_seg, (int)C, (int)(padc / 8), act, act_param, scale_mode, scale, (TY*)Y, ldy, spb);
        else
          k_segment_sum_flat8<true, TX, TY><<<ceil_div_i64(n_seg, spb), 256, 0, st>>>(
              (const TX*)X, ldx, idx, ptr, n_seg, (int)C, (int)(padc / 8), act, act_param, scale_mode, scale, (TY*)Y, ldy, spb);
      } else if (C % 4 == 0 && padc % 4 == 0 && vec4_ok<TX>(X, ldx) && vec4_ok<TY>(Y, ldy))
        k_segment_sum_v4<TX, TY><<<ceil_div_i64(n_seg, warps), warps * 32, 0, st>>>(
            (const TX*)X, ldx, idx, ptr, n_seg, (int)(C / 4), act, act_param, scale_mode, scale, (TY*)Y, ldy, (int)(padc / 4));
      else
        k_segment_sum<TX, TY><<<ceil_div_i64(n_seg, warps), warps * 32, 0, st>>>(
            (const TX*)X, ldx, idx, ptr, n_seg, (int)C, act, act_param, scale_mode, scale, (TY*)Y, ldy, (int)ldy_pad);
    ))
  DMPNN_CHECK_LAUNCH("segment_sum", 1);
  return 0;
}

extern "C" int dmpnn_segment_bcast(const void* G, int g_dtype, int64_t ldg, const int32_t* seg_of_row,
                                   const int32_t* ptr, int64_t n_seg, int64_t R, int64_t C, int scale_mode, float scale,
                                   void* Y, int y_dtype, int64_t ldy, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DMPNN_CHECK_ARG(R >= 0 && C > 0 && G && seg_of_row && Y, "segment_bcast: bad args");
  DMPNN_CHECK_ARG(scale_mode != DMPNN_SCALE_INV_COUNT || ptr, "segment_bcast: ptr needed for mean");
  if (R == 0) return 0;
  dim3 block(64, 4), grid(ceil_div_i64(R, 4), ceil_div_i64(C, 64));
  DMPNN_DISPATCH_DTYPE(g_dtype, TG,
    DMPNN_DISPATCH_DTYPE(y_dtype, TY,
      if (n_seg > 0 && ptr && C % 4 == 0 && C / 4 <= 256 && vec4_ok<TG>(G, ldg) && vec4_ok<TY>(Y, ldy))
        k_segment_bcast_seg_v4<TG, TY><<<n_seg, 256, 0, st>>>((const TG*)G, ldg, ptr, (int)(C / 4), scale_mode, scale,
                                                            (TY*)Y, ldy);
      else if (C % 4 == 0 && C / 4 <= 256 && vec4_ok<TG>(G, ldg) && vec4_ok<TY>(Y, ldy))
        k_segment_bcast_v4<TG, TY><<<ceil_div_i64(R, 64), 256, 0, st>>>((const TG*)G, ldg, seg_of_row, ptr, R,
                                                                       (int)(C / 4), scale_mode, scale, (TY*)Y, ldy, 64);
      else
        k_segment_bcast<TG, TY><<<grid, block, 0, st>>>((const TG*)G, ldg, seg_of_row, ptr, R, (int)C, scale_mode,
                                                        scale, (TY*)Y, ldy);
    ))
  DMPNN_CHECK_LAUNCH("segment_bcast", 1);
  return 0;
}

extern "C" int dmpnn_bond_message(const void* X, int x_dtype, int64_t ldx, const int32_t* rowptr,
                                  const int32_t* rev_row, int64_t V, int64_t C, int act, float act_param,
                                  int permute_on_read, void* OUT, int out_dtype, int64_t ldo, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DMPNN_CHECK_ARG(V >= 0 && C > 0 && X && rowptr && rev_row && OUT, "bond_message: bad args");
  DMPNN_CHECK_ARG(X != OUT, "bond_message: in-place not supported");
  if (V == 0) return 0;
  const int warps = 8;
  DMPNN_DISPATCH_DTYPE(x_dtype, TX,
    DMPNN_DISPATCH_DTYPE(out_dtype, TO,
      if (C % 8 == 0 && vec8_ok<TX>(X, ldx) && vec8_ok<TO>(OUT, ldo))
        k_bond_message_v4<8, TX, TO><<<ceil_div_i64(V, warps), warps * 32, 0, st>>>(
            (const TX*)X, ldx, rowptr, rev_row, V, (int)(C / 8), act, act_param, permute_on_read, (TO*)OUT, ldo,
            (const TO*)nullptr, 0, 0);
      else if (C % 4 == 0 && vec4_ok<TX>(X, ldx) && vec4_ok<TO>(OUT, ldo))
        k_bond_message_v4<4, TX, TO><<<ceil_div_i64(V, warps), warps * 32, 0, st>>>(
            (const TX*)X, ldx, rowptr, rev_row, V, (int)(C / 4), act, act_param, permute_on_read, (TO*)OUT, ldo,
            (const TO*)nullptr, 0, 0);
      else
        k_bond_message<TX, TO><<<ceil_div_i64(V, warps), warps * 32, 0, st>>>(
            (const TX*)X, ldx, rowptr, rev_row, V, (int)C, act, act_param, permute_on_read, (TO*)OUT, ldo);
    ))
  DMPNN_CHECK_LAUNCH("bond_message", 1);
  return 0;
}

extern "C" int dmpnn_rev_average(const void* X, int x_dtype, int64_t ldx, const int32_t* rev_row, int64_t R,
                                 int64_t C, int act, float act_param, void* OUT, int out_dtype, int64_t ldo,
                                 void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DMPNN_CHECK_ARG(R >= 0 && C > 0 && X && rev_row && OUT, "rev_average: bad args");
  DMPNN_CHECK_ARG(X != OUT, "rev_average: in-place not supported");
  if (R == 0) return 0;
  dim3 block(64, 4), grid(ceil_div_i64(R, 4), ceil_div_i64(C, 64));
  DMPNN_DISPATCH_DTYPE(x_dtype, TX,
    DMPNN_DISPATCH_DTYPE(out_dtype, TO,
      if (C % 4 == 0 && vec4_ok<TX>(X, ldx) && vec4_ok<TO>(OUT, ldo))
        k_rev_average_v4<TX, TO><<<ceil_div_i64(R * (C / 4), 256), 256, 0, st>>>((const TX*)X, ldx, rev_row, R, (int)(C / 4),
                                                                                 act, act_param, (TO*)OUT, ldo);
      else
        k_rev_average<TX, TO><<<grid, block, 0, st>>>((const TX*)X, ldx, rev_row, R, (int)C, act, act_param, (TO*)OUT, ldo);
    ))
  DMPNN_CHECK_LAUNCH("rev_average", 1);
  return 0;
}

extern "C" int dmpnn_act_bwd(const void* G, int g_dtype, int64_t ldg, const int32_t* gidx, const void* Yact,
                             int y_dtype, int64_t ldy, int from_preact, int act, float act_param, void* dZ,
                             int dz_dtype, int64_t lddz, void* ACC, int acc_dtype, int64_t ldacc, int64_t R,
                             int64_t C, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DMPNN_CHECK_ARG(R >= 0 && C > 0 && G && Yact, "act_bwd: bad args");
  DMPNN_CHECK_ARG(dZ || ACC, "act_bwd: nothing to write");
  if (R == 0) return 0;
  if (!dZ) dz_dtype = g_dtype;
  if (!ACC) acc_dtype = g_dtype;
  if (act == DMPNN_ACT_RELU && !ACC && g_dtype == DMPNN_BF16 && y_dtype == DMPNN_BF16 && dz_dtype == DMPNN_BF16 && C % 4 == 0 &&
      C / 4 <= 256 && vec4_ok<__nv_bfloat16>(G, ldg) && vec4_ok<__nv_bfloat16>(Yact, ldy) && vec4_ok<__nv_bfloat16>(dZ, lddz)) {
    constexpr int U = 8;
    const bool wide = C % 8 == 0 && vec8_ok<__nv_bfloat16>(G, ldg) && vec8_ok<__nv_bfloat16>(Yact, ldy) && vec8_ok<__nv_bfloat16>(dZ, lddz);
    const int Q = (int)(wide ? C / 8 : C / 4);
    const int rpb = 2 * U * (256 / Q);           // two full passes of U interleaved rows per thread
    if (wide)
      k_act_bwd_relu_bf16<16, U><<<ceil_div_i64(R, rpb), 256, 0, st>>>((const __nv_bfloat16*)G, ldg, gidx, (const __nv_bfloat16*)Yact,
                                                                        ldy, (__nv_bfloat16*)dZ, lddz, R, Q, rpb);
    else
      k_act_bwd_relu_bf16<8, U><<<ceil_div_i64(R, rpb), 256, 0, st>>>((const __nv_bfloat16*)G, ldg, gidx, (const __nv_bfloat16*)Yact,
                                                                       ldy, (__nv_bfloat16*)dZ, lddz, R, Q, rpb);
    DMPNN_CHECK_LAUNCH("act_bwd", 1);
    return 0;
  }
  dim3 block(64, 4), grid(ceil_div_i64(R, 4), ceil_div_i64(C, 64));
  DMPNN_DISPATCH_DTYPE(g_dtype, TG,
    DMPNN_DISPATCH_DTYPE(y_dtype, TYA,
      DMPNN_DISPATCH_DTYPE(dz_dtype, TZ,
        DMPNN_DISPATCH_DTYPE(acc_dtype, TA,
          constexpr int kActBwdRows = 96;     // rows per block of the row-interleaved kernels
          if (C % 4 == 0 && C / 4 <= 256 && vec4_ok<TG>(G, ldg) && vec4_ok<TYA>(Yact, ldy) && vec4_ok<TZ>(dZ, dZ ? lddz : 4) &&
              vec4_ok<TA>(ACC, ACC ? ldacc : 4)) {
            if (C % 8 == 0 && vec8_ok<TG>(G, ldg) && vec8_ok<TYA>(Yact, ldy) && vec8_ok<TZ>(dZ, dZ ? lddz : 8) &&
                vec8_ok<TA>(ACC, ACC ? ldacc : 8))
              k_act_bwd_v4<8, TG, TYA, TZ, TA><<<ceil_div_i64(R, kActBwdRows), 256, 0, st>>>(
                  (const TG*)G, ldg, gidx, (const TYA*)Yact, ldy, from_preact, act, act_param, (TZ*)dZ, lddz, (TA*)ACC, ldacc,
                  R, (int)(C / 8), kActBwdRows);
            else
              k_act_bwd_v4<4, TG, TYA, TZ, TA><<<ceil_div_i64(R, kActBwdRows), 256, 0, st>>>(
                  (const TG*)G, ldg, gidx, (const TYA*)Yact, ldy, from_preact, act, act_param, (TZ*)dZ, lddz, (TA*)ACC, ldacc,
                  R, (int)(C / 4), kActBwdRows);
          } else
            k_act_bwd<TG, TYA, TZ, TA><<<grid, block, 0, st>>>((const TG*)G, ldg, gidx, (const TYA*)Yact, ldy,
                                                              from_preact, act, act_param, (TZ*)dZ, lddz, (TA*)ACC,
                                                              ldacc, R, (int)C);
        ))))
  DMPNN_CHECK_LAUNCH("act_bwd", 1);
  return 0;
}

// Masked backward message: OUT[r] = (sum_{r' in seg(dst r)} X[rev r'] - X[rev r]) * tau'(Y[r])   (4-wide path only)
extern "C" int dmpnn_bond_message_bwd_masked(const void* X, int dtype, int64_t ldx, const int32_t* rowptr,
                                             const int32_t* rev_row, int64_t V, int64_t C, const void* Yact, int64_t ldy,
                                             int act, float act_param, void* OUT, int64_t ldo, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DMPNN_CHECK_ARG(V >= 0 && C > 0 && C % 4 == 0 && X && rowptr && rev_row && Yact && OUT && X != OUT,
                  "bond_message_bwd_masked: bad args (C must be a multiple of 4)");
  if (V == 0) return 0;
  const int warps = 8;
  DMPNN_DISPATCH_DTYPE(dtype, T,
    DMPNN_CHECK_ARG(vec4_ok<T>(X, ldx) && vec4_ok<T>(OUT, ldo) && vec4_ok<T>(Yact, ldy), "bond_message_bwd_masked: unaligned");
    if (C % 8 == 0 && vec8_ok<T>(X, ldx) && vec8_ok<T>(OUT, ldo) && vec8_ok<T>(Yact, ldy))
      k_bond_message_v4<8, T, T><<<ceil_div_i64(V, warps), warps * 32, 0, st>>>(
          (const T*)X, ldx, rowptr, rev_row, V, (int)(C / 8), DMPNN_ACT_NONE, act_param, 1, (T*)OUT, ldo, (const T*)Yact, ldy, act);
    else
      k_bond_message_v4<4, T, T><<<ceil_div_i64(V, warps), warps * 32, 0, st>>>(
          (const T*)X, ldx, rowptr, rev_row, V, (int)(C / 4), DMPNN_ACT_NONE, act_param, 1, (T*)OUT, ldo, (const T*)Yact, ldy, act);
  )
  DMPNN_CHECK_LAUNCH("bond_message_bwd_masked", 1);
  return 0;
}

extern "C" int dmpnn_sum_act_bwd(const void* const* Z, int n_z, int64_t ldz, const void* G, int64_t ldg, const void* Ypre,
                                 int64_t ldy, int dtype, int act, float act_param, void* OUT, int out_dtype, int64_t ldo,
                                 int64_t R, int64_t C, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DMPNN_CHECK_ARG(R >= 0 && C > 0 && C % 4 == 0 && n_z >= 0 && n_z <= 8 && OUT && (n_z == 0 || Z) && (!G || Ypre),
                  "sum_act_bwd: bad args (C %% 4 == 0, at most 8 addends)");
  if (R == 0) return 0;
  SumSrcs srcs;
  srcs.n = n_z;
  for (int i = 0; i < 8; ++i) srcs.z[i] = i < n_z ? Z[i] : nullptr;
  DMPNN_DISPATCH_DTYPE(dtype, TZ,
    DMPNN_DISPATCH_DTYPE(out_dtype, TO,
      bool ok = vec4_ok<TZ>(G, ldg) && vec4_ok<TZ>(Ypre, ldy) && vec4_ok<TO>(OUT, ldo) && ldz % 4 == 0;
      for (int i = 0; i < n_z; ++i) ok = ok && vec4_ok<TZ>(Z[i], ldz);
      DMPNN_CHECK_ARG(ok, "sum_act_bwd: unaligned operands");
      bool ok8 = C % 8 == 0 && vec8_ok<TZ>(G, ldg) && vec8_ok<TZ>(Ypre, ldy) && vec8_ok<TO>(OUT, ldo) && ldz % 8 == 0;
      for (int i = 0; i < n_z; ++i) ok8 = ok8 && vec8_ok<TZ>(Z[i], ldz);
      if (ok8)
        k_sum_act_bwd_v4<8, TZ, TO><<<ceil_div_i64(R * (C / 8), 256), 256, 0, st>>>(srcs, ldz, (const TZ*)G, ldg, (const TZ*)Ypre, ldy,
                                                                                   act, act_param, (TO*)OUT, ldo, R, (int)(C / 8));
      else
        k_sum_act_bwd_v4<4, TZ, TO><<<ceil_div_i64(R * (C / 4), 256), 256, 0, st>>>(srcs, ldz, (const TZ*)G, ldg, (const TZ*)Ypre, ldy,
                                                                                   act, act_param, (TO*)OUT, ldo, R, (int)(C / 4));
    ))
  DMPNN_CHECK_LAUNCH("sum_act_bwd", 1);
  return 0;
}
