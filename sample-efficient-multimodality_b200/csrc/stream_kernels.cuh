// Register-streaming kernels for the HBM-bound side products of the adapted MLP:
//
//   project:  out[M, R]  = in[M, K] * W[R, K]^T          (u = x A0, v = h A1, dv = dY B1^T, du = dpre B0^T;  R = rank)
//   reduce :  G[P, Q]   += scale * L[B, P]^T * R[B, Q]    (dB = u^T dpre, dA^T = du^T x, ... ;  P = rank, + column sums)
//
// Both stream one large activation matrix exactly once.  They use NO shared memory and no block-level synchronisation:
// every warp loads its mma.sync fragments straight from global memory with 16-byte vector loads (the contraction index
// is permuted so that a thread's 8 consecutive elements ARE its fragment registers), keeps several loads in flight per
// thread and accumulates in registers.  That makes them (a) latency-tolerant without a pipeline to fill or drain and
// (b) small enough -- 128 threads, no smem -- to be co-resident with the persistent tcgen05 GEMM CTAs (which own all of
// an SM's shared memory but only ~60 % of its registers), so the library can run them on a side stream underneath the big
// GEMMs (api.cu, merged-weight schedule).
//
// The rank-sized operand of `reduce` (u, v, dv, du: [B, P]) is exchanged in a "pair-interleaved" layout LQ that `project`
// writes directly:  u32 LQ[b/2][g][jh] = { X[b&~1][j], X[b|1][j] }  with j = 8*jh + g,  jh < PJ/8,  PJ = max(P,16)
// -- i.e. two consecutive batch rows packed per 32-bit word (the K pair of an mma.sync A fragment), ordered so that the
// four words a thread needs for two 16-row M tiles are one 16-byte load.
#pragma once
#include "outer_mma.cuh"

namespace dmi {

constexpr int ST_THREADS = 128;    // 4 warps per CTA, no shared memory

struct ProjectParams {
  const void* in; long long ld_in;      // [M, K] bf16 or fp32
  const bf16* W; long long ldw;         // [R, K] bf16
  bf16* copy; long long ld_copy;        // IN_F32: bf16 copy of `in` (or nullptr)
  bf16* out; long long ld_out;          // plain [M, R] bf16 (or nullptr)
  uint32_t* out_lq;                     // pair-interleaved [ceil(M/2)][8][PJ/8] (or nullptr)
  int M, K;
};

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 ldg_nc_v2(const void* p) {
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ldg_v4(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// 8 consecutive elements of row `row` starting at column `col` as 4 packed bf16x2 words (zero when !ok).
template <bool IN_F32>
__device__ __forceinline__ uint4 load8(const void* base, long long ld, long long row, int col, bool ok) {
  if (!ok) return make_uint4(0u, 0u, 0u, 0u);
  if (IN_F32) {
    const float* p = reinterpret_cast<const float*>(base) + row * ld + col;
    const uint4 a = ldg_nc_v4(p), b = ldg_nc_v4(p + 4);
    return make_uint4(pack_bf16x2(__uint_as_float(a.x), __uint_as_float(a.y)), pack_bf16x2(__uint_as_float(a.z), __uint_as_float(a.w)),
                      pack_bf16x2(__uint_as_float(b.x), __uint_as_float(b.y)), pack_bf16x2(__uint_as_float(b.z), __uint_as_float(b.w)));
  }
  return ldg_nc_v4(reinterpret_cast<const bf16*>(base) + row * ld + col);
}

__device__ __forceinline__ uint32_t u4c(const uint4& v, int s) { return s == 0 ? v.x : (s == 1 ? v.y : (s == 2 ? v.z : v.w)); }

// ---------------------------------------------------------------------------------------------------------------------
// project: one warp = 16 rows x all of K.  Per 64-column chunk a thread (g = lane/4, t = lane%4) loads, for rows g and g+8,
// columns [8t, 8t+8) ("lo") and [32+8t, 32+8t+8) ("hi"); k16 step s uses word s of each: logical k = 2t+e <-> column 8t+2s+e,
// logical k = 2t+8+e <-> column 32+8t+2s+e.  W is loaded with the same column mapping (row 8*nt+g), so the products match.
// ---------------------------------------------------------------------------------------------------------------------
template <int R, bool IN_F32>
__global__ void __launch_bounds__(ST_THREADS)
stream_project_kernel(const ProjectParams p) {
  constexpr int NT = R / 8;
  constexpr int PJ = R < 16 ? 16 : R;
  constexpr int JH = PJ / 8;
  constexpr int PF = IN_F32 ? 2 : 3;                 // chunks of the streamed operand in flight per thread
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int warps_total = gridDim.x * (ST_THREADS / 32);
  const int n_chunks = (p.K + 63) / 64;
  const int n_mtiles = (p.M + 15) / 16;
  for (int mt = blockIdx.x * (ST_THREADS / 32) + (threadIdx.x >> 5); mt < n_mtiles; mt += warps_total) {
    const long long r_lo = static_cast<long long>(mt) * 16 + g, r_hi = r_lo + 8;
    const bool ok_lo = r_lo < p.M, ok_hi = r_hi < p.M;
    float acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;
    uint4 a[PF][4];                                   // [stage][row g lo, row g hi-cols, row g+8 lo, row g+8 hi-cols]
    auto fetch = [&](int ch, int st) {
      const int c_lo = ch * 64 + 8 * t, c_hi = c_lo + 32;
      const bool v_lo = c_lo < p.K, v_hi = c_hi < p.K;            // K % 8 == 0: an 8-element piece is entirely in or out
      a[st][0] = load8<IN_F32>(p.in, p.ld_in, r_lo, c_lo, ok_lo && v_lo);
      a[st][1] = load8<IN_F32>(p.in, p.ld_in, r_lo, c_hi, ok_lo && v_hi);
      a[st][2] = load8<IN_F32>(p.in, p.ld_in, r_hi, c_lo, ok_hi && v_lo);
      a[st][3] = load8<IN_F32>(p.in, p.ld_in, r_hi, c_hi, ok_hi && v_hi);
    };
#pragma unroll
    for (int s = 0; s < PF - 1; ++s)
      if (s < n_chunks) fetch(s, s);
#pragma unroll 1
    for (int ch0 = 0; ch0 < n_chunks; ch0 += PF) {
#pragma unroll
      for (int st = 0; st < PF; ++st) {
        const int ch = ch0 + st;
        if (ch >= n_chunks) break;
        if (ch + PF - 1 < n_chunks) fetch(ch + PF - 1, (st + PF - 1) % PF);
        const int c_lo = ch * 64 + 8 * t, c_hi = c_lo + 32;
        const bool v_lo = c_lo < p.K, v_hi = c_hi < p.K;
        if (IN_F32 && p.copy != nullptr) {
          if (ok_lo && v_lo) *reinterpret_cast<uint4*>(p.copy + r_lo * p.ld_copy + c_lo) = a[st][0];
          if (ok_lo && v_hi) *reinterpret_cast<uint4*>(p.copy + r_lo * p.ld_copy + c_hi) = a[st][1];
          if (ok_hi && v_lo) *reinterpret_cast<uint4*>(p.copy + r_hi * p.ld_copy + c_lo) = a[st][2];
          if (ok_hi && v_hi) *reinterpret_cast<uint4*>(p.copy + r_hi * p.ld_copy + c_hi) = a[st][3];
        }
        uint4 wl[NT], wh[NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const bf16* wr = p.W + static_cast<long long>(nt * 8 + g) * p.ldw;
          wl[nt] = v_lo ? ldg_v4(wr + c_lo) : make_uint4(0u, 0u, 0u, 0u);
          wh[nt] = v_hi ? ldg_v4(wr + c_hi) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const uint32_t a0 = u4c(a[st][0], s), a1 = u4c(a[st][2], s), a2 = u4c(a[st][1], s), a3 = u4c(a[st][3], s);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) mma_bf16_16816(acc[nt], a0, a1, a2, a3, u4c(wl[nt], s), u4c(wh[nt], s));
        }
      }
    }
    // ---- outputs ----
    if (p.out != nullptr) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        if (ok_lo) *reinterpret_cast<uint32_t*>(p.out + r_lo * p.ld_out + nt * 8 + 2 * t) = pack_bf16x2(acc[nt][0], acc[nt][1]);
        if (ok_hi) *reinterpret_cast<uint32_t*>(p.out + r_hi * p.ld_out + nt * 8 + 2 * t) = pack_bf16x2(acc[nt][2], acc[nt][3]);
      }
    }
    if (p.out_lq != nullptr) {
      // rows g and g^1 form a pair: the even lane keeps column 2t (takes the partner's c0/c2), the odd lane column 2t+1.
      const bool odd = g & 1;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float mine_keep = odd ? acc[nt][2 * h + 1] : acc[nt][2 * h];
          const float mine_send = odd ? acc[nt][2 * h] : acc[nt][2 * h + 1];
          const float got = __shfl_xor_sync(0xffffffffu, mine_send, 4);
          const long long row_even = static_cast<long long>(mt) * 16 + (g & ~1) + 8 * h;
          if (row_even < p.M) {
            const bool has_odd = row_even + 1 < p.M;
            const float ev = odd ? got : mine_keep, od = odd ? mine_keep : got;
            const int j = nt * 8 + 2 * t + (odd ? 1 : 0);
            p.out_lq[(row_even >> 1) * PJ + (j & 7) * JH + (j >> 3)] = pack_bf16x2(ev, has_odd ? od : 0.f);
          }
        }
      }
      if (R < 16) {     // zero the padding half (j in [8,16)) so that the reduce kernel's second M-tile half reads zeros
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const long long row_even = static_cast<long long>(mt) * 16 + (g & ~1) + 8 * h;
          if (row_even < p.M) p.out_lq[(row_even >> 1) * PJ + (2 * t + (odd ? 1 : 0)) * JH + 1] = 0u;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// reduce: D[j, q] = sum_b L[b, j] R[b, q]   as mma.sync with A = L^T (from the LQ layout: one 16-byte load gives the a0/a1
// words of two M tiles), B = R.  One warp = 32 columns of Q x a batch range.  Per k16 step (16 batch rows) a thread loads
// rows {2t, 2t+1, 2t+8, 2t+9} x columns [4g, 4g+4) (8 bytes each; lanes g = 0..7 cover 64 contiguous bytes of a row) and
// byte-permutes them into K pairs: n-tile jj (0..3), fragment column n = g  <->  matrix column c0 + 4g + jj.
// An extra all-ones M tile produces the column sums (the bias gradients).
// ---------------------------------------------------------------------------------------------------------------------
struct ReduceParams {
  const uint32_t* Lq;              // pair-interleaved [ceil(B/2)][8][PJ/8]
  const bf16* R; long long ldr;    // [B, Q]
  int B, Q;
  float* G; long long ldg;         // fp32, atomically accumulated: G[j*ldg + q]  (transpose_out: G[q*ldg + j])
  int transpose_out;
  float* colsum;                   // [Q] or nullptr
  float scale;
  int rows_per_split;              // multiple of 16
};

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int P, bool COLSUM>
__global__ void __launch_bounds__(ST_THREADS)
stream_reduce_kernel(const ReduceParams p) {
  constexpr int PJ = P < 16 ? 16 : P;
  constexpr int JH = PJ / 8;
  constexpr int MT = PJ / 16;                   // 16-row M tiles over j
  constexpr int MTT = MT + (COLSUM ? 1 : 0);
  constexpr int PF = 4;                         // k16 steps of R in flight per thread
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int warp = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * (ST_THREADS / 32) + warp) * 32;          // first column of this warp's 32-column block
  if (c0 >= p.Q) return;
  const int b_begin = blockIdx.y * p.rows_per_split;
  const int b_end = min(p.B, b_begin + p.rows_per_split);
  if (b_begin >= b_end) return;
  const int n_steps = (b_end - b_begin + 15) / 16;
  const int col = c0 + 4 * g;
  const bool col_ok = col < p.Q;                // Q % 4 == 0: a 4-column piece is entirely in or out

  float acc[MTT][4][4];
#pragma unroll
  for (int i = 0; i < MTT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

  uint2 rr[PF][4];                              // rows 2t, 2t+1, 2t+8, 2t+9
  auto fetch = [&](int step, int st) {
    const int b0 = b_begin + step * 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = b0 + 2 * t + (i & 1) + (i >> 1) * 8;
      rr[st][i] = (col_ok && b < b_end) ? ldg_nc_v2(p.R + static_cast<long long>(b) * p.ldr + col) : make_uint2(0u, 0u);
    }
  };
#pragma unroll
  for (int s = 0; s < PF - 1; ++s)
    if (s < n_steps) fetch(s, s);
  const uint32_t ones = (g == 0) ? 0x3F803F80u : 0u;
#pragma unroll 1
  for (int s0 = 0; s0 < n_steps; s0 += PF) {
#pragma unroll
    for (int st = 0; st < PF; ++st) {
      const int step = s0 + st;
      if (step >= n_steps) break;
      if (step + PF - 1 < n_steps) fetch(step + PF - 1, (st + PF - 1) % PF);
      const int b0 = b_begin + step * 16;
      // A fragments from LQ: pair (b0/2 + t) -> a0/a1 of every M tile, pair (b0/2 + t + 4) -> a2/a3
      uint32_t al[JH], ah[JH];
      {
        const int p_lo = (b0 >> 1) + t, p_hi = p_lo + 4;
        const bool v_lo = 2 * p_lo < b_end, v_hi = 2 * p_hi < b_end;
        const uint32_t* q_lo = p.Lq + static_cast<long long>(p_lo) * PJ + g * JH;
        const uint32_t* q_hi = p.Lq + static_cast<long long>(p_hi) * PJ + g * JH;
        if (JH == 2) {
          const uint2 x = v_lo ? __ldg(reinterpret_cast<const uint2*>(q_lo)) : make_uint2(0u, 0u);
          const uint2 y = v_hi ? __ldg(reinterpret_cast<const uint2*>(q_hi)) : make_uint2(0u, 0u);
          al[0] = x.x; al[1] = x.y; ah[0] = y.x; ah[1] = y.y;
        } else {
#pragma unroll
          for (int q4 = 0; q4 < JH / 4; ++q4) {
            const uint4 x = v_lo ? ldg_v4(q_lo + 4 * q4) : make_uint4(0u, 0u, 0u, 0u);
            const uint4 y = v_hi ? ldg_v4(q_hi + 4 * q4) : make_uint4(0u, 0u, 0u, 0u);
            al[4 * q4] = x.x; al[4 * q4 + 1] = x.y; al[4 * q4 + 2] = x.z; al[4 * q4 + 3] = x.w;
            ah[4 * q4] = y.x; ah[4 * q4 + 1] = y.y; ah[4 * q4 + 2] = y.z; ah[4 * q4 + 3] = y.w;
          }
        }
      }
      // B fragments: K pairs of the four columns this thread loaded
      uint32_t b_lo[4], b_hi[4];
      b_lo[0] = prmt(rr[st][0].x, rr[st][1].x, 0x5410u); b_lo[1] = prmt(rr[st][0].x, rr[st][1].x, 0x7632u);
      b_lo[2] = prmt(rr[st][0].y, rr[st][1].y, 0x5410u); b_lo[3] = prmt(rr[st][0].y, rr[st][1].y, 0x7632u);
      b_hi[0] = prmt(rr[st][2].x, rr[st][3].x, 0x5410u); b_hi[1] = prmt(rr[st][2].x, rr[st][3].x, 0x7632u);
      b_hi[2] = prmt(rr[st][2].y, rr[st][3].y, 0x5410u); b_hi[3] = prmt(rr[st][2].y, rr[st][3].y, 0x7632u);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) mma_bf16_16816(acc[mt][jj], al[2 * mt], al[2 * mt + 1], ah[2 * mt], ah[2 * mt + 1], b_lo[jj], b_hi[jj]);
      if (COLSUM) {
        // rows past b_end were loaded as zeros, so an unconditional ones row is exact
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) mma_bf16_16816(acc[MT][jj], ones, 0u, ones, 0u, b_lo[jj], b_hi[jj]);
      }
    }
  }
  // ---- accumulate the partial sums: thread holds, for M tile mt and n-tile jj,
  //      c0: (j = 16mt+g,   q = c0+8t+jj)   c1: (j = 16mt+g,   q = c0+8t+4+jj)
  //      c2: (j = 16mt+g+8, q = c0+8t+jj)   c3: (j = 16mt+g+8, q = c0+8t+4+jj)        -> 4 consecutive q over jj
#pragma unroll
  for (int mt = 0; mt < MTT; ++mt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = 16 * mt + g + (e >> 1) * 8;
      const int q = c0 + 8 * t + (e & 1) * 4;
      if (q >= p.Q) continue;
      const float v0 = acc[mt][0][e] * p.scale, v1 = acc[mt][1][e] * p.scale, v2 = acc[mt][2][e] * p.scale, v3 = acc[mt][3][e] * p.scale;
      if (COLSUM && mt == MT) {
        if (j == 16 * MT && p.colsum != nullptr) red_add_v4(p.colsum + q, v0, v1, v2, v3);
        continue;
      }
      if (j >= P) continue;
      if (!p.transpose_out) {
        red_add_v4(p.G + static_cast<long long>(j) * p.ldg + q, v0, v1, v2, v3);
      } else {
        atomicAdd(p.G + static_cast<long long>(q) * p.ldg + j, v0);
        atomicAdd(p.G + static_cast<long long>(q + 1) * p.ldg + j, v1);
        atomicAdd(p.G + static_cast<long long>(q + 2) * p.ldg + j, v2);
        atomicAdd(p.G + static_cast<long long>(q + 3) * p.ldg + j, v3);
      }
    }
  }
}

// plain [B, P] bf16 -> LQ layout (used when the rank-sized operand does not come from stream_project_kernel)
__global__ void lq_pack_kernel(const bf16* __restrict__ X, long long ldx, int B, int P, uint32_t* __restrict__ Lq) {
  const int PJ = P < 16 ? 16 : P, JH = PJ / 8;
  const long long total = static_cast<long long>((B + 1) / 2) * PJ;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pr = i / PJ;
    const int w = static_cast<int>(i % PJ), g = w / JH, jh = w % JH, j = jh * 8 + g;
    const long long b0 = 2 * pr;
    uint32_t lo = 0, hi = 0;
    if (j < P) {
      lo = __bfloat16_as_ushort(X[b0 * ldx + j]);
      if (b0 + 1 < B) hi = __bfloat16_as_ushort(X[(b0 + 1) * ldx + j]);
    }
    Lq[i] = lo | (hi << 16);
  }
}

}  // namespace dmi
