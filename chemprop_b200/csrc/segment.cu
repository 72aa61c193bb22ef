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

}  // namespace dmpnn

using namespace dmpnn;

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
      k_segment_sum<TX, TY><<<ceil_div_i64(n_seg, warps), warps * 32, 0, st>>>(
          (const TX*)X, ldx, idx, ptr, n_seg, (int)C, act, act_param, scale_mode, scale, (TY*)Y, ldy, (int)ldy_pad);
    ))
  DMPNN_CHECK_LAUNCH("segment_sum", 1);
  return 0;
}

extern "C" int dmpnn_segment_bcast(const void* G, int g_dtype, int64_t ldg, const int32_t* seg_of_row,
                                   const int32_t* ptr, int64_t R, int64_t C, int scale_mode, float scale, void* Y,
                                   int y_dtype, int64_t ldy, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  DMPNN_CHECK_ARG(R >= 0 && C > 0 && G && seg_of_row && Y, "segment_bcast: bad args");
  DMPNN_CHECK_ARG(scale_mode != DMPNN_SCALE_INV_COUNT || ptr, "segment_bcast: ptr needed for mean");
  if (R == 0) return 0;
  dim3 block(64, 4), grid(ceil_div_i64(R, 4), ceil_div_i64(C, 64));
  DMPNN_DISPATCH_DTYPE(g_dtype, TG,
    DMPNN_DISPATCH_DTYPE(y_dtype, TY,
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
  dim3 block(64, 4), grid(ceil_div_i64(R, 4), ceil_div_i64(C, 64));
  DMPNN_DISPATCH_DTYPE(g_dtype, TG,
    DMPNN_DISPATCH_DTYPE(y_dtype, TYA,
      DMPNN_DISPATCH_DTYPE(dz_dtype, TZ,
        DMPNN_DISPATCH_DTYPE(acc_dtype, TA,
          k_act_bwd<TG, TYA, TZ, TA><<<grid, block, 0, st>>>((const TG*)G, ldg, gidx, (const TYA*)Yact, ldy,
                                                            from_preact, act, act_param, (TZ*)dZ, lddz, (TA*)ACC,
                                                            ldacc, R, (int)C);
        ))))
  DMPNN_CHECK_LAUNCH("act_bwd", 1);
  return 0;
}
