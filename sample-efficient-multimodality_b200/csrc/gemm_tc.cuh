// Persistent warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T  (both operands K-major).
//
// This one kernel carries every dense contraction of the adapted-projector path:
//   * the low-rank adapter term is folded in as extra K columns:  [x | xA0] * [W1 | B0^T]^T   (one more K slab)
//   * fused epilogues: +bias, GELU(tanh) (writing both pre-activation and activation), GELU' multiply (backward)
//
// Structure (one CTA per SM, 192 threads):
//   warp 0      TMA producer   : cp.async.bulk.tensor 2-D tiles (128B swizzle) into a STAGES-deep smem ring
//   warp 1      MMA issuer     : lane 0 issues tcgen05.mma (M=128, N=BN, K=32 bytes) into a double-buffered TMEM accumulator
//   warps 2..5  epilogue       : tcgen05.ld 32 lanes x 32 columns per warp -> registers -> fused math -> 16-byte global stores
// Pipelines: smem full/empty mbarriers (TMA <-> MMA) and TMEM full/empty mbarriers (MMA <-> epilogue), so the epilogue of
// tile i overlaps the MMAs of tile i+1.
#pragma once
#include "common.cuh"

namespace dmi {

enum EpiMode : int { EPI_STORE = 0, EPI_GELU = 1, EPI_GELU_BWD = 2 };
enum GemmKind : int { KIND_BF16 = 0, KIND_TF32 = 1 };

struct GemmParams {
  int M, N, K;
  float alpha;            // acc is scaled by alpha before bias / activation
  const float* bias;      // [N] fp32 or nullptr
  void* out0;             // EPI_STORE: alpha*acc+bias | EPI_GELU: gelu(pre) | EPI_GELU_BWD: acc * gelu'(aux)
  long long ld0;          // row stride of out0 in elements
  int out0_f32;           // 1: out0 is float, 0: out0 is bf16
  bf16* out1;             // EPI_STORE: optional bf16 copy | EPI_GELU: optional pre-activation (bf16) | else unused
  long long ld1;
  const bf16* aux;        // EPI_GELU_BWD: stashed pre-activation
  long long ld_aux;
  const uint8_t* keep;    // optional dropout keep mask [M,N] (1 = keep): EPI_GELU scales the activation, EPI_GELU_BWD the gradient
  long long ld_keep;
  float keep_scale;       // 1/(1-p)
  int accumulate_out0;    // EPI_STORE with fp32 out0: out0 += result instead of out0 = result
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_SMEM_BUDGET = 200 * 1024;

template <int BN>
struct GemmCfg {
  static constexpr int STAGE_BYTES = (GEMM_BM + BN) * 128;
  static constexpr int STAGES = (GEMM_SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (GEMM_SMEM_BUDGET / STAGE_BYTES);
  static constexpr int TMEM_COLS = (2 * BN) < 32 ? 32 : (2 * BN);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// v[j] *= keep[j] ? scale : 0 for the 32 columns of one epilogue chunk (keep: bytes, 16-byte aligned rows)
__device__ __forceinline__ void apply_keep_mask(float (&v)[32], const uint8_t* keep, float scale, int ncols_left) {
#pragma unroll
  for (int j = 0; j < 32; j += 16) {
    if (j < ncols_left) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(keep + j));
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int t = 0; t < 16; ++t) v[j + t] *= ((w[t >> 2] >> (8 * (t & 3))) & 0xFFu) ? scale : 0.0f;
    }
  }
}

// AB_MN = false: both operands K-major (A [M,K], B [N,K] row-major).
// AB_MN = true : both operands MN-major (A [K,M], B [K,N] row-major, i.e. C = A^T B with the contraction over the ROWS of
//                both matrices) -- the weight-gradient GEMMs dW = dY^T h, whose K is the batch.  TMA then stages
//                [BK rows x 64 columns] boxes (one per 64 columns of M / N) and the descriptors use the MN-major
//                SWIZZLE_128B canonical layout: 64 MN-elements contiguous, K rows 128 B apart, 8-row groups 1024 B apart (SBO),
//                64-column chunks BK*128 B apart (LBO).
template <int BN, int MODE, int KIND, bool AB_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int A_BYTES = GEMM_BM * 128;
  constexpr int BK = (KIND == KIND_BF16) ? 64 : 32;     // elements per 128-byte K slab
  constexpr int UK = (KIND == KIND_BF16) ? 16 : 8;      // elements per tcgen05.mma (32 bytes of K)
  constexpr uint32_t IDESC = make_idesc(GEMM_BM, BN, KIND == KIND_BF16 ? 1 : 2) | (AB_MN ? ((1u << 15) | (1u << 16)) : 0u);
  static_assert(!AB_MN || (KIND == KIND_BF16 && BN % 64 == 0), "MN-major operands: bf16 and BN multiple of 64 only");
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN must be a multiple of 32 in [32,256]");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;     // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles_n = (p.N + BN - 1) / BN;
  const int n_tiles_m = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = n_tiles_m * n_tiles_n;
  const int nkb = (p.K + BK - 1) / BK;
  const int ksteps_last = ((p.K - (nkb - 1) * BK) + UK - 1) / UK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles_n) * GEMM_BM;
        const int n0 = (tile % n_tiles_n) * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          if (AB_MN) {
            // boxes of [BK rows (K)] x [64 columns (MN)]: coordinate 0 = column, coordinate 1 = row
#pragma unroll
            for (int c = 0; c < GEMM_BM / 64; ++c) tma_load_2d(sa + c * (BK * 128), &tmA, &full_bar[stage], m0 + c * 64, kb * BK);
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sa + A_BYTES + c * (BK * 128), &tmB, &full_bar[stage], n0 + c * 64, kb * BK);
          } else {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
            tma_load_2d(sa + A_BYTES, &tmB, &full_bar[stage], kb * BK, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t adesc = AB_MN ? make_mnmajor_sw128_desc(sa, BK * 128) : make_kmajor_sw128_desc(sa);
          const uint64_t bdesc = AB_MN ? make_mnmajor_sw128_desc(sa + A_BYTES, BK * 128) : make_kmajor_sw128_desc(sa + A_BYTES);
          const int ks = (kb == nkb - 1) ? ksteps_last : (BK / UK);
          // K-major: advancing K by 32 bytes inside the 128-byte swizzle span = +2 in the (addr >> 4) start-address field.
          // MN-major: advancing K by 16 rows of 128 bytes = +128.
          constexpr uint32_t KADV = AB_MN ? (16 * 128) >> 4 : 2;
          for (int k = 0; k < ks; ++k) {
            if (KIND == KIND_BF16) umma_f16(d_tmem, adesc + KADV * k, bdesc + KADV * k, IDESC, (kb | k) != 0);
            else                   umma_tf32(d_tmem, adesc + KADV * k, bdesc + KADV * k, IDESC, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);                 // smem slot is free once these MMAs have read it
          if (kb == nkb - 1) umma_commit(&tfull_bar[acc]); // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, 32*quarter+32) are accessible to this warp
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = (tile / n_tiles_n) * GEMM_BM;
      const int n0 = (tile % n_tiles_n) * BN;
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) break;                   // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32(t_addr + c * 32, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col0 + j < p.N) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
        }
        if (!row_ok) continue;
        if (MODE == EPI_GELU) {
          if (p.out1 != nullptr) {
            bf16* dst = p.out1 + static_cast<long long>(row) * p.ld1 + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (col0 + j < p.N) {
                uint4 q;
                q.x = pack_bf16x2(v[j], v[j + 1]); q.y = pack_bf16x2(v[j + 2], v[j + 3]);
                q.z = pack_bf16x2(v[j + 4], v[j + 5]); q.w = pack_bf16x2(v[j + 6], v[j + 7]);
                *reinterpret_cast<uint4*>(dst + j) = q;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_tanh(v[j]);
          if (p.keep != nullptr) apply_keep_mask(v, p.keep + static_cast<long long>(row) * p.ld_keep + col0, p.keep_scale, p.N - col0);
        } else if (MODE == EPI_GELU_BWD) {
          if (p.keep != nullptr) apply_keep_mask(v, p.keep + static_cast<long long>(row) * p.ld_keep + col0, p.keep_scale, p.N - col0);
          const bf16* src = p.aux + static_cast<long long>(row) * p.ld_aux + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (col0 + j < p.N) {
              const uint4 q = __ldg(reinterpret_cast<const uint4*>(src + j));
              const float2 a0 = unpack_bf16x2(q.x), a1 = unpack_bf16x2(q.y), a2 = unpack_bf16x2(q.z), a3 = unpack_bf16x2(q.w);
              v[j] *= gelu_tanh_grad(a0.x); v[j + 1] *= gelu_tanh_grad(a0.y);
              v[j + 2] *= gelu_tanh_grad(a1.x); v[j + 3] *= gelu_tanh_grad(a1.y);
              v[j + 4] *= gelu_tanh_grad(a2.x); v[j + 5] *= gelu_tanh_grad(a2.y);
              v[j + 6] *= gelu_tanh_grad(a3.x); v[j + 7] *= gelu_tanh_grad(a3.y);
            }
          }
        }
        // main output
        if (p.out0_f32) {
          float* dst = reinterpret_cast<float*>(p.out0) + static_cast<long long>(row) * p.ld0 + col0;
          if (MODE == EPI_STORE && p.accumulate_out0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (col0 + j < p.N) {
                const float4 o = *reinterpret_cast<const float4*>(dst + j);
                v[j] += o.x; v[j + 1] += o.y; v[j + 2] += o.z; v[j + 3] += o.w;
              }
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (col0 + j < p.N) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          bf16* dst = reinterpret_cast<bf16*>(p.out0) + static_cast<long long>(row) * p.ld0 + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (col0 + j < p.N) {
              uint4 q;
              q.x = pack_bf16x2(v[j], v[j + 1]); q.y = pack_bf16x2(v[j + 2], v[j + 3]);
              q.z = pack_bf16x2(v[j + 4], v[j + 5]); q.w = pack_bf16x2(v[j + 6], v[j + 7]);
              *reinterpret_cast<uint4*>(dst + j) = q;
            }
          }
        }
        if (MODE == EPI_STORE && p.out1 != nullptr) {
          bf16* dst = p.out1 + static_cast<long long>(row) * p.ld1 + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (col0 + j < p.N) {
              uint4 q;
              q.x = pack_bf16x2(v[j], v[j + 1]); q.y = pack_bf16x2(v[j + 2], v[j + 3]);
              q.z = pack_bf16x2(v[j + 4], v[j + 5]); q.w = pack_bf16x2(v[j + 6], v[j + 7]);
              *reinterpret_cast<uint4*>(dst + j) = q;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Row-major [rows, inner] matrix with row stride ld (elements) -> 2-D tensor map, 128-byte swizzle, box = [box_rows, 128 bytes].
int make_tmap_2d(CUtensorMap* tm, const void* ptr, int kind, long long inner, long long rows, long long ld, int box_rows);
// Same matrix, but boxes of [box_rows rows] x [64 columns] for MN-major operands (bf16).
int make_tmap_2d_mn(CUtensorMap* tm, const void* ptr, long long inner, long long rows, long long ld, int box_rows);

int num_sms();

void count_launch();

template <int BN, int MODE, int KIND, bool AB_MN = false>
int launch_gemm_inst(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  auto kern = gemm_tn_kernel<BN, MODE, KIND, AB_MN>;
  if (!configured) {
    DMI_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int n_tiles = ((p.M + GEMM_BM - 1) / GEMM_BM) * ((p.N + BN - 1) / BN);
  const int grid = n_tiles < num_sms() ? n_tiles : num_sms();
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, p);
  DMI_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DMI_OK;
}

}  // namespace dmi
