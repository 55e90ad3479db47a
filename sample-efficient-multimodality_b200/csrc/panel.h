// Fused projection + batch-reduction pass over one activation-gradient matrix (panel.cu).
#pragma once
#include "common.cuh"

namespace dmi {

// Shapes the fused kernel is compiled for; callers fall back to skinny_rows + outer_reduce otherwise.
bool panel_fused_supported(long long K, int R);

// out[M,R] = in[M,K] W[R,K]^T;  G[R,K] += scale * L[M,R]^T in;  colsum[K] += scale * 1^T in (colsum may be null);
// in_f32: `in` is fp32, its bf16 copy is written to `copy` (may be null).  All reductions see the bf16-rounded input.
int panel_fused(const void* in, long long ld_in, bool in_f32, const bf16* W, long long ldw, bf16* out, long long ld_out, bf16* copy,
                long long ld_copy, const bf16* L, long long ldl, float* G, long long ldg, float* colsum, float scale, long long M,
                long long K, int R, cudaStream_t s);

// tcgen05 form (panel_tc.cu): bf16 input only, R = 32, K = 1024 or 2048; `out` 16-byte aligned with ld_out % 8 == 0.
bool panel_fused_tc_supported(long long K, int R);
int panel_fused_tc(const bf16* in, long long ld_in, const bf16* W, long long ldw, bf16* out, long long ld_out, const bf16* L, long long ldl,
                   float* G, long long ldg, float* colsum, float scale, long long M, long long K, int R, cudaStream_t s);

int panel_fused_tc_mcs(const bf16* in, long long ld_in, const bf16* W, long long ldw, bf16* out, long long ld_out, const bf16* L, long long ldl,
                       float* G, long long ldg, float* colsum, float scale, long long M, long long K, int R, cudaStream_t s);

// fp32-input form (panel_tc32.cu): the dY pass, also writes the bf16 copy.  Not validated on a GPU yet (fused_panel bit 4).
int panel_fused_tc32(const float* in, long long ld_in, const bf16* W, long long ldw, bf16* out, long long ld_out, bf16* copy, long long ld_copy,
                     const bf16* L, long long ldl, float* G, long long ldg, float* colsum, float scale, long long M, long long K, int R,
                     cudaStream_t s);

// Single-mode launches of the same kernel (K = 768 / 1024 / 2048, R = 32): projection only, batch reduction only.
bool panel_tc_mode_supported(long long K, int R);
int panel_tc_project(const bf16* in, long long ld_in, const bf16* W, long long ldw, bf16* out, long long ld_out, long long M, long long K, int R,
                     cudaStream_t s);
int panel_tc_reduce(const bf16* in, long long ld_in, const bf16* L, long long ldl, float* G, long long ldg, int transpose_out, float* colsum,
                    float scale, long long M, long long K, int R, cudaStream_t s);

}  // namespace dmi
