// Row-panel "skinny" projection:  out[M, R] = in[M, K] * W[R, K]^T   with R = adapter rank (8..64), bf16 output.
//
// These are the rank-r products of the adapted MLP (u = x A0, v = h A1, dv = dY B1^T, du = dpre B0^T): 2*K*R FLOP per row
// against 2*K bytes read per row (R <= 64 -> <= 64 FLOP/byte, far below the machine balance), i.e. purely HBM-bound.  So the
// kernel is built for bytes in flight, not tensor throughput: one CTA owns a 64-row panel, streams it through a 4-deep
// cp.async ring in 128-column chunks and feeds warp-level mma.sync (m16n8k16).
//
// IN_F32 = true is the fused "convert + project" pass: the input is the caller's fp32 matrix (x or dY); the kernel converts
// it to bf16 on the fly, writes the bf16 copy that the big GEMMs consume (xext / dyext columns [0,K)) and computes the
// projection from the same registers -- one pass over the fp32 data instead of convert-then-reread.
#pragma once
#include "outer_mma.cuh"

namespace dmi {

// Tile shape (profiles/r1_side_kernels.txt, B200, [32768 x 2048] bf16): 64 rows x 4 stages 41 us, 128 x 2 33 us (4.0 TB/s), 256 x 2 35 us,
// 32 x 4 46 us; the fp32 path is best at 64 rows (5.5 TB/s).  The launcher uses 128 x 2 for bf16 input and 64 rows for fp32 input.
constexpr int SK_KC = 128;         // K columns per pipeline stage
constexpr int SK_THREADS = 512;    // 16 warps: (ROWS/16) row groups of 16 rows x (256/ROWS) K-slices of each 128-column chunk
constexpr int SK_AW = SK_KC + 8;   // padded smem row stride (elements): ldmatrix conflict-free

struct SkinnyParams {
  const void* in; long long ld_in;      // [M, K] bf16 (IN_F32 = false) or fp32 (IN_F32 = true)
  const bf16* W; long long ldw;         // [R, K] bf16, K-major
  bf16* out; long long ld_out;          // [M, R]
  bf16* copy; long long ld_copy;        // IN_F32: bf16 copy of `in` [M, K] (may alias the buffer that holds `out` in other columns)
  int M, K, R;
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

template <int R, bool IN_F32, int SK_ROWS = 64, int SK_STAGES = 4>
__global__ void __launch_bounds__(SK_THREADS)
skinny_rows_kernel(const SkinnyParams p) {
  pdl_prologue();
  constexpr int NT = R / 8;                       // n8 tiles
  constexpr int NSTG = IN_F32 ? 2 : SK_STAGES;    // the fp32 path prefetches through registers, 2 smem buffers suffice
  constexpr int SK_RGROUPS = SK_ROWS / 16;
  constexpr int SK_KSPLIT = (SK_THREADS / 32) / SK_RGROUPS;
  static_assert(SK_KC / 16 % SK_KSPLIT == 0, "K slices must divide the k16 steps of a chunk");
  extern __shared__ __align__(16) uint8_t ssm[];
  bf16* sA = reinterpret_cast<bf16*>(ssm);                            // [NSTG][SK_ROWS][SK_AW]
  bf16* sW = sA + NSTG * SK_ROWS * SK_AW;                             // [NSTG][R][SK_AW]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long row0 = static_cast<long long>(blockIdx.x) * SK_ROWS;
  const int n_chunks = (p.K + SK_KC - 1) / SK_KC;

  auto load_w = [&](int chunk, int buf) {
    bf16* dw = sW + buf * R * SK_AW;
    const int k0 = chunk * SK_KC;
    for (int i = tid; i < R * (SK_KC / 8); i += SK_THREADS) {
      const int n = i / (SK_KC / 8), c8 = (i % (SK_KC / 8)) * 8;
      const bool ok = k0 + c8 < p.K;
      cp_async16(dw + n * SK_AW + c8, p.W + static_cast<long long>(n) * p.ldw + (ok ? k0 + c8 : 0), ok);
    }
  };
  auto load_a_bf16 = [&](int chunk, int buf) {
    bf16* da = sA + buf * SK_ROWS * SK_AW;
    const int k0 = chunk * SK_KC;
    const bf16* src = reinterpret_cast<const bf16*>(p.in);
    for (int i = tid; i < SK_ROWS * (SK_KC / 8); i += SK_THREADS) {
      const int r = i / (SK_KC / 8), c8 = (i % (SK_KC / 8)) * 8;
      const bool ok = (row0 + r < p.M) && (k0 + c8 < p.K);
      cp_async16(da + r * SK_AW + c8, src + (ok ? (row0 + r) * p.ld_in + k0 + c8 : 0), ok);
    }
  };
  // fp32 path: each thread owns NPRE x (4 consecutive floats): chunk = 64 rows x 32 float4 = 2048 float4 / 512 threads
  constexpr int NPRE = SK_ROWS * (SK_KC / 4) / SK_THREADS;
  float4 pre[NPRE];
  auto fetch_a_f32 = [&](int chunk) {
    const int k0 = chunk * SK_KC;
    const float* src = reinterpret_cast<const float*>(p.in);
#pragma unroll
    for (int j = 0; j < NPRE; ++j) {
      const int i = tid + j * SK_THREADS;
      const int r = i >> 5, c4 = (i & 31) * 4;
      const bool ok = (row0 + r < p.M) && (k0 + c4 < p.K);
      pre[j] = ok ? __ldg(reinterpret_cast<const float4*>(src + (row0 + r) * p.ld_in + k0 + c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto commit_a_f32 = [&](int chunk, int buf) {
    bf16* da = sA + buf * SK_ROWS * SK_AW;
    const int k0 = chunk * SK_KC;
#pragma unroll
    for (int j = 0; j < NPRE; ++j) {
      const int i = tid + j * SK_THREADS;
      const int r = i >> 5, c4 = (i & 31) * 4;
      uint2 q;
      q.x = pack_bf16x2(pre[j].x, pre[j].y);
      q.y = pack_bf16x2(pre[j].z, pre[j].w);
      *reinterpret_cast<uint2*>(da + r * SK_AW + c4) = q;
      if (p.copy != nullptr && (row0 + r < p.M) && (k0 + c4 < p.K))
        *reinterpret_cast<uint2*>(p.copy + (row0 + r) * p.ld_copy + k0 + c4) = q;
    }
  };

  float acc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;

  const int rw = warp % SK_RGROUPS;  // row group: rows [16*rw, 16*rw+16) of the panel
  const int kq = warp / SK_RGROUPS;  // K slice of every chunk
  auto compute = [&](int buf) {
    const bf16* ca = sA + buf * SK_ROWS * SK_AW + rw * 16 * SK_AW;
    const bf16* cw = sW + buf * R * SK_AW;
#pragma unroll
    for (int ks = kq * (SK_KC / 16 / SK_KSPLIT); ks < (kq + 1) * (SK_KC / 16 / SK_KSPLIT); ++ks) {
      uint32_t a0, a1, a2, a3;
      ldmatrix_x4(smem_u32(ca + (lane & 15) * SK_AW + ks * 16 + (lane >> 4) * 8), a0, a1, a2, a3);
      if (NT >= 2) {
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
          uint32_t b0, b1, b2, b3;
          const int nrow = np * 16 + (lane >> 4) * 8 + (lane & 7);
          ldmatrix_x4(smem_u32(cw + nrow * SK_AW + ks * 16 + ((lane >> 3) & 1) * 8), b0, b1, b2, b3);
          mma_bf16_16816(acc[2 * np], a0, a1, a2, a3, b0, b1);
          mma_bf16_16816(acc[2 * np + 1], a0, a1, a2, a3, b2, b3);
        }
      } else {
        uint32_t b0, b1;
        ldmatrix_x2(smem_u32(cw + (lane & 7) * SK_AW + ks * 16 + ((lane >> 3) & 1) * 8), b0, b1);
        mma_bf16_16816(acc[0], a0, a1, a2, a3, b0, b1);
      }
    }
  };

  if (IN_F32) {
    fetch_a_f32(0);
    load_w(0, 0);
    cp_async_commit();
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int buf = ch & 1;
      commit_a_f32(ch, buf);                         // registers -> smem (+ global bf16 copy)
      if (ch + 1 < n_chunks) {
        fetch_a_f32(ch + 1);                         // next chunk's global loads fly during the MMAs below
        load_w(ch + 1, buf ^ 1);
      }
      cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();
      compute(buf);
      __syncthreads();
    }
  } else {
#pragma unroll
    for (int s = 0; s < SK_STAGES - 1; ++s) {
      if (s < n_chunks) { load_a_bf16(s, s); load_w(s, s); }
      cp_async_commit();
    }
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int nxt = ch + SK_STAGES - 1;
      if (nxt < n_chunks) { load_a_bf16(nxt, nxt % SK_STAGES); load_w(nxt, nxt % SK_STAGES); }
      cp_async_commit();
      cp_async_wait<SK_STAGES - 1>();
      __syncthreads();
      compute(ch % SK_STAGES);
      __syncthreads();
    }
  }

  // epilogue: sum the 4 K-quarter partials through shared memory (the operand ring is free now), then bf16 pairs
  float* red = reinterpret_cast<float*>(ssm);                     // [SK_KSPLIT][SK_ROWS][R]
  const int g = lane >> 2, t = lane & 3;
  __syncthreads();
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      float* d = red + (static_cast<long long>(kq) * SK_ROWS + rw * 16 + g + hrow * 8) * R + nt * 8 + 2 * t;
      d[0] = acc[nt][2 * hrow];
      d[1] = acc[nt][2 * hrow + 1];
    }
  __syncthreads();
  for (int i = tid; i < SK_ROWS * (R / 2); i += SK_THREADS) {
    const int r = i / (R / 2), c = (i % (R / 2)) * 2;
    const long long row = row0 + r;
    if (row < p.M) {
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int q = 0; q < SK_KSPLIT; ++q) {
        s0 += red[(static_cast<long long>(q) * SK_ROWS + r) * R + c];
        s1 += red[(static_cast<long long>(q) * SK_ROWS + r) * R + c + 1];
      }
      *reinterpret_cast<uint32_t*>(p.out + row * p.ld_out + c) = pack_bf16x2(s0, s1);
    }
  }
}

// A handful of rows (the B = 4 hypernet micro-step): the panel kernel above is ONE CTA walking K chunk by chunk (19.5 us for
// [4 x 2048] . [2048 x 32]).  Here one CTA per output column r takes the whole K at once -- every load of the call is in flight
// together -- and reduces across its 8 warps.  bf16 in, fp32 accumulation, bf16 out like the panel kernel.
constexpr int TINY_MAX_ROWS = 16;
template <bool IN_F32>
__global__ void __launch_bounds__(256)
tiny_rows_kernel(const SkinnyParams p) {
  pdl_prologue();
  __shared__ float red[8][TINY_MAX_ROWS];
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float acc[TINY_MAX_ROWS];
#pragma unroll
  for (int m = 0; m < TINY_MAX_ROWS; ++m) acc[m] = 0.f;
  for (int k = tid * 8; k < p.K; k += 256 * 8) {
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(p.W + static_cast<long long>(r) * p.ldw + k));
    const float2 w0 = unpack_bf16x2(wv.x), w1 = unpack_bf16x2(wv.y), w2 = unpack_bf16x2(wv.z), w3 = unpack_bf16x2(wv.w);
#pragma unroll
    for (int m = 0; m < TINY_MAX_ROWS; ++m) {
      if (m < p.M) {
        uint4 xv;
        if (IN_F32) {      // convert on the fly (round to nearest even, like the panel kernel); CTA 0 also writes the bf16 copy
          const float* src = reinterpret_cast<const float*>(p.in) + static_cast<long long>(m) * p.ld_in + k;
          const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
          xv = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
          if (r == 0 && p.copy != nullptr) *reinterpret_cast<uint4*>(p.copy + static_cast<long long>(m) * p.ld_copy + k) = xv;
        } else {
          xv = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.in) + static_cast<long long>(m) * p.ld_in + k);
        }
        const float2 x0 = unpack_bf16x2(xv.x), x1 = unpack_bf16x2(xv.y), x2 = unpack_bf16x2(xv.z), x3 = unpack_bf16x2(xv.w);
        float a = acc[m];
        a = fmaf(w0.x, x0.x, a); a = fmaf(w0.y, x0.y, a); a = fmaf(w1.x, x1.x, a); a = fmaf(w1.y, x1.y, a);
        a = fmaf(w2.x, x2.x, a); a = fmaf(w2.y, x2.y, a); a = fmaf(w3.x, x3.x, a); a = fmaf(w3.y, x3.y, a);
        acc[m] = a;
      }
    }
  }
#pragma unroll
  for (int m = 0; m < TINY_MAX_ROWS; ++m) {
    float v = acc[m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][m] = v;
  }
  __syncthreads();
  if (tid < p.M) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][tid];
    p.out[static_cast<long long>(tid) * p.ld_out + r] = __float2bfloat16(v);
  }
}

inline int launch_tiny_rows(const SkinnyParams& p, bool in_f32, cudaStream_t stream) {
  if (in_f32) DMI_CHECK_CUDA(launch_pdl(tiny_rows_kernel<true>, dim3(p.R), dim3(256), 0, stream, p));
  else        DMI_CHECK_CUDA(launch_pdl(tiny_rows_kernel<false>, dim3(p.R), dim3(256), 0, stream, p));
  count_launch();
  return DMI_OK;
}

template <int R, bool IN_F32, int SK_ROWS = 64, int SK_STAGES = 4>
int launch_skinny_inst(const SkinnyParams& p, cudaStream_t stream) {
  constexpr int NSTG = IN_F32 ? 2 : SK_STAGES;
  constexpr int SK_KSPLIT = (SK_THREADS / 32) / (SK_ROWS / 16);
  constexpr int ring = NSTG * (SK_ROWS + R) * SK_AW * 2;
  constexpr int redb = SK_KSPLIT * SK_ROWS * R * 4;
  constexpr int smem = ring > redb ? ring : redb;
  auto kern = skinny_rows_kernel<R, IN_F32, SK_ROWS, SK_STAGES>;
  static bool configured = false;
  if (!configured) {
    DMI_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  DMI_CHECK_CUDA(launch_pdl(kern, dim3((p.M + SK_ROWS - 1) / SK_ROWS), dim3(SK_THREADS), smem, stream, p));
  DMI_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DMI_OK;
}

}  // namespace dmi
