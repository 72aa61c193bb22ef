// Version / error plumbing and the host-side collate of libdmpnn_sm100.so.
#include <stdarg.h>
#include <atomic>
#include <string.h>

#include "common.cuh"

namespace dmpnn {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace dmpnn

extern "C" {

long long dmpnn_launch_count(void) { return dmpnn::g_launches.load(std::memory_order_relaxed); }

int dmpnn_version(void) { return DMPNN_VERSION; }

const char* dmpnn_last_error(void) { return dmpnn::g_err; }

int dmpnn_device_ok(void) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return major == 10 ? 1 : 0;
}

// Replaces the per-molecule Python loop of BatchMolGraph.__post_init__
// (chemprop/data/collate.py:37-62): concatenate V/E, offset edge_index by the running atom
// count (:51), rev_edge_index by the running edge count (:52), batch[i-th mol's atoms] = i (:53).
int dmpnn_collate_host(int64_t n_mols, const int64_t* n_atoms, const int64_t* n_edges,
                       const float* const* V_ptrs, const float* const* E_ptrs,
                       const int64_t* const* edge_index_ptrs, const int64_t* const* rev_ptrs,
                       int64_t d_v, int64_t d_e, float* V_out, float* E_out,
                       int64_t* edge_index_out, int64_t* rev_out, int64_t* batch_out) {
  DMPNN_CHECK_ARG(n_mols >= 0 && d_v >= 0 && d_e >= 0, "collate_host: negative size");
  int64_t E_tot = 0;
  for (int64_t i = 0; i < n_mols; ++i) {
    DMPNN_CHECK_ARG(n_atoms[i] >= 0 && n_edges[i] >= 0, "collate_host: negative molecule size");
    E_tot += n_edges[i];
  }
  int64_t a0 = 0, e0 = 0;
  for (int64_t i = 0; i < n_mols; ++i) {
    const int64_t na = n_atoms[i], ne = n_edges[i];
    if (na > 0 && d_v > 0) memcpy(V_out + a0 * d_v, V_ptrs[i], sizeof(float) * na * d_v);
    if (ne > 0 && d_e > 0) memcpy(E_out + e0 * d_e, E_ptrs[i], sizeof(float) * ne * d_e);
    const int64_t* ei = edge_index_ptrs[i];
    const int64_t* rv = rev_ptrs[i];
    for (int64_t j = 0; j < ne; ++j) {
      edge_index_out[e0 + j] = ei[j] + a0;
      edge_index_out[E_tot + e0 + j] = ei[ne + j] + a0;
      rev_out[e0 + j] = rv[j] + e0;
    }
    for (int64_t j = 0; j < na; ++j) batch_out[a0 + j] = i;
    a0 += na;
    e0 += ne;
  }
  return 0;
}

static inline uint16_t f32_to_bf16_rne(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  if ((x & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((x >> 16) | 0x40u);   // NaN stays NaN (quiet)
  x += 0x7fffu + ((x >> 16) & 1u);                                             // round to nearest, ties to even
  return (uint16_t)(x >> 16);
}

// Compact transfer copy of the same batch (bf16 features, int32 indices): what BatchMolGraph(transfer_dtype=bfloat16)
// ships over PCIe.  One pass over the per-molecule arrays, same offsets as dmpnn_collate_host.
int dmpnn_collate_host_compact(int64_t n_mols, const int64_t* n_atoms, const int64_t* n_edges,
                               const float* const* V_ptrs, const float* const* E_ptrs,
                               const int64_t* const* edge_index_ptrs, const int64_t* const* rev_ptrs,
                               int64_t d_v, int64_t d_e, uint16_t* V_out, uint16_t* E_out,
                               int32_t* edge_index_out, int32_t* rev_out, int32_t* batch_out) {
  DMPNN_CHECK_ARG(n_mols >= 0 && d_v >= 0 && d_e >= 0, "collate_host_compact: negative size");
  int64_t E_tot = 0, V_tot = 0;
  for (int64_t i = 0; i < n_mols; ++i) {
    DMPNN_CHECK_ARG(n_atoms[i] >= 0 && n_edges[i] >= 0, "collate_host_compact: negative molecule size");
    E_tot += n_edges[i];
    V_tot += n_atoms[i];
  }
  DMPNN_CHECK_ARG(E_tot < (1LL << 31) && V_tot < (1LL << 31) && n_mols < (1LL << 31),
                  "collate_host_compact: batch too large for int32 indices");
  int64_t a0 = 0, e0 = 0;
  for (int64_t i = 0; i < n_mols; ++i) {
    const int64_t na = n_atoms[i], ne = n_edges[i];
    const float* v = V_ptrs[i];
    const float* e = E_ptrs[i];
    for (int64_t j = 0; j < na * d_v; ++j) V_out[a0 * d_v + j] = f32_to_bf16_rne(v[j]);
    for (int64_t j = 0; j < ne * d_e; ++j) E_out[e0 * d_e + j] = f32_to_bf16_rne(e[j]);
    const int64_t* ei = edge_index_ptrs[i];
    const int64_t* rv = rev_ptrs[i];
    for (int64_t j = 0; j < ne; ++j) {
      edge_index_out[e0 + j] = (int32_t)(ei[j] + a0);
      edge_index_out[E_tot + e0 + j] = (int32_t)(ei[ne + j] + a0);
      rev_out[e0 + j] = (int32_t)(rv[j] + e0);
    }
    for (int64_t j = 0; j < na; ++j) batch_out[a0 + j] = (int32_t)i;
    a0 += na;
    e0 += ne;
  }
  return 0;
}

}  // extern "C"
