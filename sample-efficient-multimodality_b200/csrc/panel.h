// tcgen05 row-panel kernels for the rank-r side products of the adapted-MLP step (panel_tc.cu, panel_tc32.cu).
#pragma once
#include "common.cuh"

namespace dmi {

// Fused projection + batch-reduction pass over one bf16 activation-gradient matrix (the dpre pass), R = 32, K = 1024 or 2048:
//   out[M,R] = in[M,K] W[R,K]^T;  G[R,K] += scale * L[M,R]^T in;  colsum[K] += scale * 1^T in (colsum may be null).
// `out` must be 16-byte aligned with ld_out % 8 == 0.  Callers fall back to skinny_rows + outer_reduce for other shapes.
bool panel_fused_tc_supported(long long K, int R);
int panel_fused_tc(const bf16* in, long long ld_in, const bf16* W, long long ldw, bf16* out, long long ld_out, const bf16* L, long long ldl,
                   float* G, long long ldg, float* colsum, float scale, long long M, long long K, int R, cudaStream_t s);

// fp32-input form (panel_tc32.cu): the dY pass; also writes the bf16 copy of `in` (copy may be null).  All reductions see the
// bf16-rounded input.  Same shapes as panel_fused_tc.
int panel_fused_tc32(const float* in, long long ld_in, const bf16* W, long long ldw, bf16* out, long long ld_out, bf16* copy, long long ld_copy,
                     const bf16* L, long long ldl, float* G, long long ldg, float* colsum, float scale, long long M, long long K, int R,
                     cudaStream_t s);

// Projection alone (v = h A1, u = x A0): K = 768 / 1024 / 2048, R = 32.
bool panel_tc_mode_supported(long long K, int R);
int panel_tc_project(const bf16* in, long long ld_in, const bf16* W, long long ldw, bf16* out, long long ld_out, long long M, long long K, int R,
                     cudaStream_t s);

}  // namespace dmi
