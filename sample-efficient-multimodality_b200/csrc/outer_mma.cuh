// Batch-reduction ("outer product") kernel for the adapter gradients:
//
//     G[P,Q] += scale * sum_b L[b,P] * R[b,Q]          (+ optionally  colsum[Q] += scale * sum_b R[b,Q])
//
// with P = adapter rank (8..64) and Q = H or D.  It produces dB1 = v^T dY (+ d beta1 = 1^T dY), dA1^T = dv^T h,
// dB0 = u^T dpre (+ d beta0), dA0^T = du^T x.  The contraction runs over the batch, so both operands are read
// "transposed"; the work is HBM/L2-bound (each launch streams one [B,Q] bf16 activation once, ~32 FLOP/byte), so it
// uses warp-level mma.sync m16n8k16 with ldmatrix.trans rather than tcgen05: the tensor pipe is nowhere near the limit.
// The column sum comes for free as an extra all-ones row of L^T.
#pragma once
#include "common.cuh"

namespace dmi {

void count_launch();

constexpr int OUTER_QC = 256;     // Q columns per CTA (8 warps x 32): 512 contiguous bytes per row per stage (see skinny.cuh)
constexpr int OUTER_KB = 64;      // batch rows per pipeline stage
constexpr int OUTER_THREADS = 256;

struct OuterParams {
  const bf16* L; long long ldl;   // [B, P]
  const bf16* R; long long ldr;   // [B, Q]
  int B, P, Q;
  float* G; long long ldg;        // fp32, atomically accumulated
  int transpose_out;              // 0: G[p*ldg+q]   1: G[q*ldg+p]
  float* colsum;                  // [Q] or nullptr
  float scale;
  int rows_per_split;             // multiple of OUTER_KB
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// MT = number of 16-row tiles covering P (+1 for the ones row when COLSUM).
template <int MT_P, bool COLSUM>
__global__ void __launch_bounds__(OUTER_THREADS)
outer_reduce_kernel(const OuterParams p) {
  pdl_prologue();
  constexpr int MT = MT_P + (COLSUM ? 1 : 0);
  constexpr int LW = MT * 16 + 8;           // smem row stride of the L tile (elements); +8 keeps ldmatrix conflict-free
  constexpr int RW = OUTER_QC + 8;
  extern __shared__ __align__(16) uint8_t osm[];
  bf16* sL = reinterpret_cast<bf16*>(osm);                         // [2][KB][LW]
  bf16* sR = sL + 2 * OUTER_KB * LW;                               // [2][KB][RW]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * OUTER_QC;
  const int b_begin = blockIdx.y * p.rows_per_split;
  const int b_end = min(p.B, b_begin + p.rows_per_split);
  if (b_begin >= b_end) return;
  const int n_chunks = (b_end - b_begin + OUTER_KB - 1) / OUTER_KB;

  // pad columns of sL (ones row + zero fill) are written once per buffer per chunk by plain stores
  auto load_chunk = [&](int chunk, int buf) {
    const int b0 = b_begin + chunk * OUTER_KB;
    bf16* dl = sL + buf * OUTER_KB * LW;
    bf16* dr = sR + buf * OUTER_KB * RW;
    // R: KB rows x QC cols, 16 x 16B per row
    for (int i = tid; i < OUTER_KB * (OUTER_QC / 8); i += OUTER_THREADS) {
      const int row = i / (OUTER_QC / 8), c8 = (i % (OUTER_QC / 8)) * 8;
      const int b = b0 + row;
      const bool ok = (b < b_end) && (q0 + c8 < p.Q);
      const bf16* src = p.R + static_cast<long long>(ok ? b : 0) * p.ldr + (ok ? q0 + c8 : 0);
      cp_async16(dr + row * RW + c8, src, ok);
    }
    // L: KB rows x P cols
    const int pch = MT_P * 2;      // 16B chunks per row covering MT_P*16 columns
    for (int i = tid; i < OUTER_KB * pch; i += OUTER_THREADS) {
      const int row = i / pch, c8 = (i % pch) * 8;
      const int b = b0 + row;
      const bool ok = (b < b_end) && (c8 < p.P);
      const bf16* src = p.L + static_cast<long long>(ok ? b : 0) * p.ldl + (ok ? c8 : 0);
      cp_async16(dl + row * LW + c8, src, ok);
    }
    if (COLSUM) {
      for (int i = tid; i < OUTER_KB * 2; i += OUTER_THREADS) {
        const int row = i >> 1, half = i & 1;
        const int b = b0 + row;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (half == 0 && b < b_end) q.x = 0x00003F80u;     // bf16(1.0) in element 0
        *reinterpret_cast<uint4*>(dl + row * LW + MT_P * 16 + half * 8) = q;
      }
    }
  };

  float acc[MT][4][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

  load_chunk(0, 0);
  cp_async_commit();
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < n_chunks) load_chunk(ch + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const bf16* cl = sL + buf * OUTER_KB * LW;
    const bf16* cr = sR + buf * OUTER_KB * RW;
#pragma unroll
    for (int ks = 0; ks < OUTER_KB / 16; ++ks) {
      // B fragments for this warp's 32 columns: two ldmatrix.x4.trans (n-tiles 0,1 and 2,3)
      uint32_t bfr[4][2];
      {
        const int j = lane >> 3, i = lane & 7;
        const int krow = ks * 16 + (j & 1) * 8 + i;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ncol = warp * 32 + h * 16 + (j >> 1) * 8;
          ldmatrix_x4_trans(smem_u32(cr + krow * RW + ncol), bfr[2 * h][0], bfr[2 * h][1], bfr[2 * h + 1][0], bfr[2 * h + 1][1]);
        }
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        uint32_t a0, a1, a2, a3;
        const int j = lane >> 3, i = lane & 7;
        const int krow = ks * 16 + (j >> 1) * 8 + i;
        const int mcol = mt * 16 + (j & 1) * 8;
        ldmatrix_x4_trans(smem_u32(cl + krow * LW + mcol), a0, a1, a2, a3);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[mt][nt], a0, a1, a2, a3, bfr[nt][0], bfr[nt][1]);
      }
    }
    __syncthreads();
  }

  // ---- epilogue: atomically accumulate the partial sums ----
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int prow = mt * 16 + g + (e >> 1) * 8;
        const int q = q0 + warp * 32 + nt * 8 + 2 * t + (e & 1);
        if (q >= p.Q) continue;
        const float val = acc[mt][nt][e] * p.scale;
        if (prow < p.P) {
          float* dst = p.transpose_out ? (p.G + static_cast<long long>(q) * p.ldg + prow)
                                       : (p.G + static_cast<long long>(prow) * p.ldg + q);
          atomicAdd(dst, val);
        } else if (COLSUM && prow == MT_P * 16 && p.colsum != nullptr) {
          atomicAdd(p.colsum + q, val);
        }
      }
    }
  }
}

template <int MT_P, bool COLSUM>
int launch_outer_inst(const OuterParams& p, int nsplit, cudaStream_t stream) {
  constexpr int MT = MT_P + (COLSUM ? 1 : 0);
  constexpr int LW = MT * 16 + 8, RW = OUTER_QC + 8;
  constexpr int smem = 2 * OUTER_KB * (LW + RW) * 2;
  auto kern = outer_reduce_kernel<MT_P, COLSUM>;
  static bool configured = false;
  if (!configured) {
    DMI_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid((p.Q + OUTER_QC - 1) / OUTER_QC, nsplit);
  DMI_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(OUTER_THREADS), smem, stream, p));
  DMI_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DMI_OK;
}

}  // namespace dmi
