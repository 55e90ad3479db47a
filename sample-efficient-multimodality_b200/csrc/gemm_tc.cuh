// Persistent warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T  (both operands K-major).
//
// This one kernel carries every dense contraction of the adapted-projector path:
//   * the low-rank adapter term is folded in as extra K columns:  [x | xA0] * [W1 | B0^T]^T   (one more K slab)
//   * fused epilogues: +bias, GELU(tanh) (writing both pre-activation and activation), GELU' multiply (backward)
//
// Structure (one CTA per SM, 320 threads):
//   warp 0      TMA producer   : cp.async.bulk.tensor 2-D tiles (128B swizzle) into a STAGES-deep smem ring
//   warp 1      MMA issuer     : lane 0 issues tcgen05.mma (M=128, N=BN, K=32 bytes) into a double-buffered TMEM accumulator
//   warps 2..9  epilogue       : tcgen05.ld 32 lanes x 32 columns per warp -> registers -> fused math -> shared staging ->
//                                coalesced 16-byte global stores (two warps per TMEM lane quarter, one per column half)
// Pipelines: smem full/empty mbarriers (TMA <-> MMA) and TMEM full/empty mbarriers (MMA <-> epilogue), so the epilogue of
// tile i overlaps the MMAs of tile i+1.
#pragma once
#include <string.h>

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
  // EPI_STORE only: fp32 matrix added to alpha*acc (+bias) before the store, used by the exact adapter merge checks
  const float* addend; long long ld_add;
  int debug;              // measurement only (dmi_set_option "gemm_debug"): 1 = skip the epilogue, 2 = skip TMA loads / full waits
  // EPI_STORE with fp32 out0 only: rows >= split_row go to out0_b[(row - split_row) * ld0_b + col] instead (and get no out1 copy) --
  // two row groups that share the B operand as one launch (dmi_augment: batch rows and support rows times the same rotation)
  int split_row;          // 0 = off
  float* out0_b; long long ld0_b;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_THREADS = 64 + 8 * 32;   // TMA warp, MMA warp, 8 epilogue warps
constexpr int GEMM_SMEM_BUDGET = 193 * 1024;   // operand ring; + 32 KB epilogue staging + barriers <= 227 KB

template <int BN>
struct GemmCfg {
  static constexpr int STAGE_BYTES = (GEMM_BM + BN) * 128;
  static constexpr int STAGES = (GEMM_SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (GEMM_SMEM_BUDGET / STAGE_BYTES);
  static constexpr int TMEM_COLS = (2 * BN) < 32 ? 32 : (2 * BN);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 8 * 4096 /*epilogue staging*/;
};

// ---- epilogue staging: one 4 KB shared buffer per epilogue warp (32 rows x 32 fp32, 16-byte slots XOR-swizzled by the row) ----
constexpr int EPI_STAGE_BYTES = 4096;
constexpr int EPI_WARPS = 8;

// ---- explicit shared-state-space accessors (a generic pointer into dynamic smem makes the compiler emit generic LD/ST) ----
__device__ __forceinline__ void sts128(uint32_t addr, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// Drains one accumulator tile: this warp handles its 32 TMEM lanes (rows row0..row0+31) x its half of the BN columns.
// Per 32x32 chunk: tcgen05.ld (thread = row) -> raw fp32 to the warp's swizzled 4 KB staging buffer -> re-read TRANSPOSED
// (lane = 4 consecutive columns of rows i*4 + lane/8) -> bias / activation / gradient math -> coalesced global stores.
// Doing the math after the transpose means every lane works on FIXED columns: its 4 bias values live in registers, the
// stashed pre-activation and the dropout mask are read with coalesced loads, and each store instruction covers whole rows.
template <int BN, int MODE>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t stage_s, uint32_t t_addr, int row0, int n0, int half, int lane) {
  constexpr int CHUNKS = BN / 32;
  constexpr int CH_PER_WARP = (CHUNKS + 1) / 2;
  const int rows_valid = p.M - row0;          // rows >= rows_valid are outside the matrix
  const int ch = lane & 7;                    // this lane's 4-column group inside a chunk
  const int rsub = lane >> 3;                 // row inside each group of 4 rows
  const uint32_t put_base = stage_s + lane * 128;
  const int put_sw = lane & 7;
#pragma unroll 1
  for (int cc = 0; cc < CH_PER_WARP; ++cc) {
    const int c = half * CH_PER_WARP + cc;
    if (c >= CHUNKS) break;
    const int col0 = n0 + c * 32;
    if (col0 >= p.N) break;                   // warp-uniform
    const int colg = col0 + ch * 4;
    const bool col_ok = colg < p.N;
    uint32_t r[32];
    tmem_ld_32x32(t_addr + c * 32, r);
    // independent global loads go out while the TMEM load is in flight
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias != nullptr && col_ok) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + colg));
    uint2 aux[8];
    if (MODE == EPI_GELU_BWD) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + rsub;
        aux[i] = make_uint2(0u, 0u);
        if (row < rows_valid && col_ok) aux[i] = __ldg(reinterpret_cast<const uint2*>(p.aux + static_cast<long long>(row0 + row) * p.ld_aux + colg));
      }
    }
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sts128(put_base + ((j ^ put_sw) << 4), __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
    __syncwarp();
    // Fast path (warp-uniform): the whole 32x32 chunk is inside the matrix and none of the rarely used options is on.  Straight-line
    // code: 8 transposed reads, one block of independent FP32 work, then stores through row pointers that advance by 4 rows --
    // no per-store predicates, branches or 64-bit multiplies (ncu: those were ~30 % of the epilogue's instructions).
    const bool fast = rows_valid >= 32 && col0 + 32 <= p.N && !(MODE != EPI_STORE && p.keep != nullptr) && !(p.debug & 15) &&
                      !(MODE == EPI_STORE && ((p.accumulate_out0 && !p.out0_f32) || p.addend != nullptr || (p.out0_f32 && p.out1 != nullptr) || p.split_row > 0));
    if (fast) {
      float4 a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + rsub;
        a[i] = lds128(stage_s + row * 128 + ((ch ^ (row & 7)) << 4));
      }
      uint2 pre_pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v0 = fmaf(a[i].x, p.alpha, b4.x), v1 = fmaf(a[i].y, p.alpha, b4.y), v2 = fmaf(a[i].z, p.alpha, b4.z), v3 = fmaf(a[i].w, p.alpha, b4.w);
        if (MODE == EPI_GELU) {
          pre_pk[i] = make_uint2(pack_bf16x2(v0, v1), pack_bf16x2(v2, v3));
          v0 = gelu_tanh(v0); v1 = gelu_tanh(v1); v2 = gelu_tanh(v2); v3 = gelu_tanh(v3);
        } else if (MODE == EPI_GELU_BWD) {
          const float2 p01 = unpack_bf16x2(aux[i].x), p23 = unpack_bf16x2(aux[i].y);
          v0 *= gelu_tanh_grad(p01.x); v1 *= gelu_tanh_grad(p01.y); v2 *= gelu_tanh_grad(p23.x); v3 *= gelu_tanh_grad(p23.y);
        }
        a[i] = make_float4(v0, v1, v2, v3);
      }
      const long long first = static_cast<long long>(row0 + rsub);
      if (MODE == EPI_GELU && p.out1 != nullptr) {
        bf16* q = p.out1 + first * p.ld1 + colg;
        const long long step = 4 * p.ld1;
#pragma unroll
        for (int i = 0; i < 8; ++i, q += step) *reinterpret_cast<uint2*>(q) = pre_pk[i];
      }
      if (p.out0_f32) {
        float* q = reinterpret_cast<float*>(p.out0) + first * p.ld0 + colg;
        const long long step = 4 * p.ld0;
        if (MODE == EPI_STORE && p.accumulate_out0) {       // weight gradients: out0 += result (all 8 loads first, then the stores)
          float4 o[8];
          float* ql = q;
#pragma unroll
          for (int i = 0; i < 8; ++i, ql += step) o[i] = *reinterpret_cast<const float4*>(ql);
#pragma unroll
          for (int i = 0; i < 8; ++i) { a[i].x += o[i].x; a[i].y += o[i].y; a[i].z += o[i].z; a[i].w += o[i].w; }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i, q += step) *reinterpret_cast<float4*>(q) = a[i];
      } else {
        bf16* q = reinterpret_cast<bf16*>(p.out0) + first * p.ld0 + colg;
        const long long step = 4 * p.ld0;
#pragma unroll
        for (int i = 0; i < 8; ++i, q += step) *reinterpret_cast<uint2*>(q) = make_uint2(pack_bf16x2(a[i].x, a[i].y), pack_bf16x2(a[i].z, a[i].w));
      }
      __syncwarp();
      continue;
    }
    // three phases so that the FP32 work of the 32 elements is one block of independent instructions (the epilogue warps are few --
    // two per scheduler -- so instruction-level parallelism, not occupancy, has to hide the MUFU / FMA latencies):
    // (1) all transposed reads, (2) all math, (3) all global stores.
    float4 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = i * 4 + rsub;
      a[i] = lds128(stage_s + row * 128 + ((ch ^ (row & 7)) << 4));
    }
    uint32_t kw[8];
    const bool use_keep = MODE != EPI_STORE && p.keep != nullptr;
    if (use_keep) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + rsub;
        kw[i] = (row < rows_valid && col_ok) ? __ldg(reinterpret_cast<const uint32_t*>(p.keep + static_cast<long long>(row0 + row) * p.ld_keep + colg)) : 0u;
      }
    }
    uint2 pre_pk[8];                      // EPI_GELU: packed pre-activation
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v0 = fmaf(a[i].x, p.alpha, b4.x), v1 = fmaf(a[i].y, p.alpha, b4.y), v2 = fmaf(a[i].z, p.alpha, b4.z), v3 = fmaf(a[i].w, p.alpha, b4.w);
      if (MODE == EPI_GELU) {
        pre_pk[i] = make_uint2(pack_bf16x2(v0, v1), pack_bf16x2(v2, v3));
        if (!(p.debug & 4)) { v0 = gelu_tanh(v0); v1 = gelu_tanh(v1); v2 = gelu_tanh(v2); v3 = gelu_tanh(v3); }
      } else if (MODE == EPI_GELU_BWD) {
        const float2 p01 = unpack_bf16x2(aux[i].x), p23 = unpack_bf16x2(aux[i].y);
        v0 *= gelu_tanh_grad(p01.x); v1 *= gelu_tanh_grad(p01.y); v2 *= gelu_tanh_grad(p23.x); v3 *= gelu_tanh_grad(p23.y);
      }
      if (use_keep) {
        v0 *= (kw[i] & 0xFFu) ? p.keep_scale : 0.f; v1 *= (kw[i] & 0xFF00u) ? p.keep_scale : 0.f;
        v2 *= (kw[i] & 0xFF0000u) ? p.keep_scale : 0.f; v3 *= (kw[i] & 0xFF000000u) ? p.keep_scale : 0.f;
      }
      a[i] = make_float4(v0, v1, v2, v3);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = i * 4 + rsub;
      if (row >= rows_valid || !col_ok || (p.debug & 8)) continue;
      const long long grow = row0 + row;
      float v0 = a[i].x, v1 = a[i].y, v2 = a[i].z, v3 = a[i].w;
      if (MODE == EPI_STORE && p.addend != nullptr) {
        const float4 ad = __ldg(reinterpret_cast<const float4*>(p.addend + grow * p.ld_add + colg));
        v0 += ad.x; v1 += ad.y; v2 += ad.z; v3 += ad.w;
      }
      if (MODE == EPI_GELU && p.out1 != nullptr) *reinterpret_cast<uint2*>(p.out1 + grow * p.ld1 + colg) = pre_pk[i];      // pre-activation
      if (p.out0_f32) {
        const bool second = MODE == EPI_STORE && p.split_row > 0 && grow >= p.split_row;
        float4* g = second ? reinterpret_cast<float4*>(p.out0_b + (grow - p.split_row) * p.ld0_b + colg)
                           : reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out0) + grow * p.ld0 + colg);
        if (MODE == EPI_STORE && p.accumulate_out0) { const float4 o = *g; v0 += o.x; v1 += o.y; v2 += o.z; v3 += o.w; }
        *g = make_float4(v0, v1, v2, v3);
        if (MODE == EPI_STORE && p.out1 != nullptr && !second)
          *reinterpret_cast<uint2*>(p.out1 + grow * p.ld1 + colg) = make_uint2(pack_bf16x2(v0, v1), pack_bf16x2(v2, v3));
      } else {
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out0) + grow * p.ld0 + colg) = make_uint2(pack_bf16x2(v0, v1), pack_bf16x2(v2, v3));
      }
    }
    __syncwarp();
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
  const int tile0 = blockIdx.x, tile_stride = gridDim.x;
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
      mbar_init(&tempty_bar[a], 8);
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
  // programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) overlapped the predecessor's tail;
  // from here on global memory is read
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp runs the loop (warp-uniform control flow) and one elected lane issues: inside a `lane == 0` branch the compiler
    // cannot use the uniform datapath, so every UTMALDG was preceded by an ELECT / R2UR.BROADCAST loop (see the MMA issuer below).
    if (!(p.debug & 2)) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < n_tiles; tile += tile_stride) {
        const int m0 = (tile / n_tiles_n) * GEMM_BM;
        const int n0 = (tile % n_tiles_n) * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
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
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // All 32 lanes run the loop with warp-uniform control flow and ONE ELECTED lane issues the tcgen05 instructions.  Round 1 ran
    // the whole loop inside `if (lane == 0)`: in divergent code the compiler cannot use the uniform datapath, so it materialised
    // every descriptor in vector registers and wrapped each UTCHMMA / UTCBAR in an ELECT + 5x R2UR.BROADCAST loop -- ~17 dependent
    // instructions per MMA on a single thread, i.e. the issue path (not the operand feed) paced the tensor pipe at 165-190 cycles
    // per 128-cycle MMA (profiles/r1_gemm_duty_ncu.txt: 69-78 % pipe-active with loads AND epilogue switched off).  With
    // elect.sync as the guard the same source compiles to uniform-register descriptors and four back-to-back UTCHMMA per k block.
    {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = tile0; tile < n_tiles; tile += tile_stride, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          if (!(p.debug & 2)) mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
            const uint64_t adesc = AB_MN ? make_mnmajor_sw128_desc(sa, BK * 128) : make_kmajor_sw128_desc(sa);
            const uint64_t bdesc = AB_MN ? make_mnmajor_sw128_desc(sa + A_BYTES, BK * 128) : make_kmajor_sw128_desc(sa + A_BYTES);
            // K-major: advancing K by 32 bytes inside the 128-byte swizzle span = +2 in the (addr >> 4) start-address field.
            // MN-major: advancing K by 16 rows of 128 bytes = +128.
            constexpr uint32_t KADV = AB_MN ? (16 * 128) >> 4 : 2;
            if (kb != nkb - 1 || ksteps_last == BK / UK) {
#pragma unroll
              for (int k = 0; k < BK / UK; ++k) {
                if (KIND == KIND_BF16) umma_f16(d_tmem, adesc + KADV * k, bdesc + KADV * k, IDESC, (kb | k) != 0);
                else                   umma_tf32(d_tmem, adesc + KADV * k, bdesc + KADV * k, IDESC, (kb | k) != 0);
              }
            } else {
              for (int k = 0; k < ksteps_last; ++k) {
                if (KIND == KIND_BF16) umma_f16(d_tmem, adesc + KADV * k, bdesc + KADV * k, IDESC, (kb | k) != 0);
                else                   umma_tf32(d_tmem, adesc + KADV * k, bdesc + KADV * k, IDESC, (kb | k) != 0);
              }
            }
            umma_commit(&empty_bar[stage]);                   // smem slot is free once these MMAs have read it
            if (kb == nkb - 1) umma_commit(&tfull_bar[acc]); // accumulator complete -> epilogue
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // Two warps per TMEM lane quarter (warp & 3), each owning one half of the tile's columns, so every SM sub-partition has
    // two epilogue warps to interleave.  Every 32x32 chunk goes registers -> (swizzled) shared staging -> COALESCED global
    // stores: a thread owns one row of the accumulator, so storing straight from registers would touch 32 different
    // 128-byte lines per instruction (measured: the epilogue, not the MMA, bounded the kernel).
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, 32*quarter+32) are accessible to this warp
    const int half = (warp - 2) >> 2;             // 0: columns [0, BN/2)   1: columns [BN/2, BN)
    const uint32_t stage = smem_u32(smem + STAGES * Cfg::STAGE_BYTES + 256 + (warp - 2) * EPI_STAGE_BYTES);
    int it = 0;
    for (int tile = tile0; tile < n_tiles; tile += tile_stride, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = (tile / n_tiles_n) * GEMM_BM;
      const int n0 = (tile % n_tiles_n) * BN;
      const int row0 = m0 + quarter * 32;         // first row of this warp's 32-row slab
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
      if (!(p.debug & 1)) epilogue_tile<BN, MODE>(p, stage, t_addr, row0, n0, half, lane);
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
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1 + pdl_attribute(&attr[1]);
  DMI_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
  count_launch();
  return DMI_OK;
}

}  // namespace dmi
