// placeholder: fused tcgen05 depth step (implemented next)
#include "common.cuh"
extern "C" int dmpnn_pack_weight_bf16_bytes(int64_t N, int64_t K, size_t* bytes) { dmpnn::set_error("not built"); return -3; }
extern "C" int dmpnn_pack_weight_bf16(const float* W, int64_t ldw, int64_t N, int64_t K, void* Wpk, void* stream) { dmpnn::set_error("not built"); return -3; }
extern "C" int dmpnn_bond_step_fused_bf16(const void* H_prev, const void* H_0, void* H_next, int64_t ld,
                               int64_t n_rows_alloc, int64_t h, const void* Wpk, const float* bias,
                               const int32_t* rowptr, const int32_t* rev_row,
                               const int32_t* mol_atom_ptr, const int32_t* mol_row_ptr,
                               const int32_t* tile_mol_ptr, int64_t n_tiles,
                               int act, float act_param, int first_step, void* stream) { dmpnn::set_error("not built"); return -3; }
