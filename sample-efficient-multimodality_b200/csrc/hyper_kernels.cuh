// HBM-bound fp32 kernels of the hypernetwork side of the path (k1, k3, k4, k6, k7, k12, k13).
// Everything here is exact fp32 arithmetic (no tensor cores): row normalisation, the 2-query attention pooling in its
// algebraically reduced form, the generator GEMV and its rank-1 gradient, the prefix splice.
#pragma once
#include "common.cuh"

namespace dmi {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum for up to 1024 threads; every thread gets the result.  `red` = 33 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : -INFINITY;
  if (w == 0) {
    t = warp_max(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// -------------------------------------------------------------------------------------------------------------------
// a1: row-wise L2 normalisation x / ||x||  (EmbeddingManager.get_embeddings, model_utils.py:54-59).  One warp per row.
// Also used as the augmentation prologue: optional column gather (perm), sign flip, and the 3xTF32 split
//   dst3[row] = [hi | hi | lo]  (hi = value with the low 13 mantissa bits cleared, lo = value - hi)
// so that the rotation GEMM on tf32 tensor cores is fp32-accurate:  x R = hi R_hi + hi R_lo + lo R_hi + O(2^-22).
// -------------------------------------------------------------------------------------------------------------------
struct RowPrepParams {
  const float* src; long long ld_src; int rows; int cols_in;      // source rows
  const int* perm;          // [cols_out] gather index into the source row, or nullptr (identity)
  const float* sign;        // [cols_out] +-1, or nullptr
  int cols_out;
  int normalize;            // divide by the L2 norm of the FULL source row (as the reference normalises before anything else)
  float* dst; long long ld_dst;        // optional fp32 output [rows, cols_out] (+ zero fill up to cols_pad)
  int cols_pad;                        // >= cols_out: columns [cols_out, cols_pad) of dst are zero-filled (pruned projector)
  float* dst3; long long ld_dst3;      // optional [rows, 3*cols_out] split output for the 3xTF32 GEMM
  bf16* dst_bf16; long long ld_bf16;   // optional bf16 copy [rows, cols_out]
};

__device__ __forceinline__ void row_prep_row(const RowPrepParams& p, int row, int lane) {
  const float* s = p.src + static_cast<long long>(row) * p.ld_src;
  float nrm = 1.0f;
  if (p.normalize) {
    float ss = 0.f;
    for (int c = lane; c < p.cols_in; c += 32) { const float v = s[c]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    nrm = sqrtf(ss);
  }
  for (int c = lane; c < p.cols_pad; c += 32) {
    float v = 0.f;
    if (c < p.cols_out) {
      const int sc = p.perm ? p.perm[c] : c;
      v = s[sc];
      if (p.normalize) v = v / nrm;            // x / ||x||, a true division as in the reference (model_utils.py:59)
      if (p.sign) v *= p.sign[c];
    }
    if (p.dst) p.dst[static_cast<long long>(row) * p.ld_dst + c] = v;
    if (c < p.cols_out) {
      if (p.dst3) {
        const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        const float lo = v - hi;
        float* d3 = p.dst3 + static_cast<long long>(row) * p.ld_dst3;
        d3[c] = hi; d3[p.cols_out + c] = hi; d3[2 * p.cols_out + c] = lo;
      }
      if (p.dst_bf16) p.dst_bf16[static_cast<long long>(row) * p.ld_bf16 + c] = __float2bfloat16(v);
    }
  }
}

// one warp per row
__global__ void row_prep_kernel(const RowPrepParams p) {
  pdl_prologue();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row < p.rows) row_prep_row(p, row, threadIdx.x & 31);
}

// Up to four independent row groups in ONE launch (the batch, support, text and instruction-prefix rows of dmi_augment: the hypernet
// micro-step is launch-bound, every launch removed is ~3 us of it).  Rows are numbered through the segments in order.
constexpr int ROW_PREP_MAX_SEGMENTS = 4;
struct RowPrepBatch {
  RowPrepParams seg[ROW_PREP_MAX_SEGMENTS];
  int n;
};

__global__ void row_prep_multi_kernel(const __grid_constant__ RowPrepBatch b) {
  pdl_prologue();
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
#pragma unroll
  for (int i = 0; i < ROW_PREP_MAX_SEGMENTS; ++i) {
    if (i >= b.n) return;
    if (row < b.seg[i].rows) { row_prep_row(b.seg[i], row, threadIdx.x & 31); return; }
    row -= b.seg[i].rows;
  }
}

// B operand of the 3xTF32 rotation GEMM: Rt3[n, :] = [R_hi[:, n] | R_lo[:, n] | R_hi[:, n]]  (K-major, K = 3*D_in)
__global__ void split_rotation_kernel(const float* __restrict__ R, int d_in, int d_out, float* __restrict__ Rt3) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int dk = threadIdx.y; dk < 32; dk += blockDim.y) {
    const int k = k0 + dk, n = n0 + threadIdx.x;
    tile[dk][threadIdx.x] = (k < d_in && n < d_out) ? R[static_cast<long long>(k) * d_out + n] : 0.f;
  }
  __syncthreads();
  for (int dn = threadIdx.y; dn < 32; dn += blockDim.y) {
    const int n = n0 + dn, k = k0 + threadIdx.x;
    if (n < d_out && k < d_in) {
      const float v = tile[threadIdx.x][dn];
      const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
      float* row = Rt3 + static_cast<long long>(n) * (3 * d_in);
      row[k] = hi; row[d_in + k] = v - hi; row[2 * d_in + k] = hi;
    }
  }
}

// -------------------------------------------------------------------------------------------------------------------
// GEMV (fp32, NV = number of right-hand vectors): fallback of the generators for widths the streaming kernels are not compiled for
//   gemv_rows: y[i, o] = out_scale * (sum_d W[o,d] x[i,d] + bias[o] * bias_scale[i])     one warp per output row o
// -------------------------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
gemv_rows_kernel(const float* __restrict__ W, long long ldw, int O, int D, const float* __restrict__ x, long long ldx,
                 const float* __restrict__ bias, const float* __restrict__ bias_scale, float out_scale, float* __restrict__ y, long long ldy) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const long long o = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= O) return;
  const float* w = W + o * ldw;
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;
  if ((D & 3) == 0 && (ldw & 3) == 0 && (ldx & 3) == 0) {
    for (int d = lane * 4; d < D; d += 128) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + d));
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + i * ldx + d));
        acc[i] = fmaf(wv.x, xv.x, acc[i]); acc[i] = fmaf(wv.y, xv.y, acc[i]);
        acc[i] = fmaf(wv.z, xv.z, acc[i]); acc[i] = fmaf(wv.w, xv.w, acc[i]);
      }
    }
  } else {
    for (int d = lane; d < D; d += 32) {
      const float wv = w[d];
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i] = fmaf(wv, x[i * ldx + d], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float s = warp_sum(acc[i]);
    if (lane == 0) {
      float b = 0.f;
      if (bias) b = bias[o] * (bias_scale ? bias_scale[i] : 1.0f);
      y[i * ldy + o] = out_scale * (s + b);
    }
  }
}

// -------------------------------------------------------------------------------------------------------------------
// k7 / k12: the generators.  Both directions are pure streams over G [O, D] fp32 (283 MB + 409 MB at the real widths; arithmetic
// intensity 0.5 FLOP/B), so they are written as PERSISTENT kernels: a fixed grid of 2 CTAs per SM, every warp walks rows
// o = warp, warp + n_warps, ... with TWO rows in flight (the loads of rows o + n_warps and o + 2 n_warps are issued before row o is
// reduced), streaming loads / stores (ld.global.cs / st.global.cs: the weights are touched once per call and must not evict the
// rest of the step from L2), the modality code e [D] held in registers.  Round 1 launched one short CTA per 8 (forward) / 32
// (backward) rows -- thousands of CTAs that each moved 24-190 KB -- and reached 0.53 / 0.54 of the HBM peak.
// NJ = float4 chunks per lane (D <= 128 NJ; D % 4 == 0).
// -------------------------------------------------------------------------------------------------------------------
template <int NJ>
__device__ __forceinline__ void gen_load_row(const float* __restrict__ row, int D, int lane, float4 (&w)[NJ]) {
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int d = lane * 4 + 128 * j;
    w[j] = (d < D) ? __ldcs(reinterpret_cast<const float4*>(row + d)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// forward: y[o] = out_scale * (G[o,:] . e + c[o])
template <int NJ>
__global__ void __launch_bounds__(256, 2)
generator_fwd_kernel(const float* __restrict__ Gw, long long ldw, long long O, int D, const float* __restrict__ e, const float* __restrict__ c,
                     float out_scale, float* __restrict__ y) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const long long nw = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  long long o = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  float4 ev[NJ], w0[NJ], w1[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int d = lane * 4 + 128 * j;
    ev[j] = (d < D) ? __ldg(reinterpret_cast<const float4*>(e + d)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  auto dot = [&](const float4 (&w)[NJ]) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      acc = fmaf(w[j].x, ev[j].x, acc); acc = fmaf(w[j].y, ev[j].y, acc); acc = fmaf(w[j].z, ev[j].z, acc); acc = fmaf(w[j].w, ev[j].w, acc);
    }
    return warp_sum(acc);
  };
  if (o < O) gen_load_row<NJ>(Gw + o * ldw, D, lane, w0);
  if (o + nw < O) gen_load_row<NJ>(Gw + (o + nw) * ldw, D, lane, w1);
  for (; o < O; o += 2 * nw) {
    float s0 = dot(w0);
    if (o + 2 * nw < O) gen_load_row<NJ>(Gw + (o + 2 * nw) * ldw, D, lane, w0);
    if (lane == 0) y[o] = out_scale * (s0 + (c ? c[o] : 0.f));
    if (o + nw < O) {
      float s1 = dot(w1);
      if (o + 3 * nw < O) gen_load_row<NJ>(Gw + (o + 3 * nw) * ldw, D, lane, w1);
      if (lane == 0) y[o + nw] = out_scale * (s1 + (c ? c[o + nw] : 0.f));
    }
  }
}

// backward in ONE pass over the weight rows: for row o, with g = dw_scale * dw[o]
//   de      +=  g * G[o,:]            (per-lane partial sums in registers, one shared-memory + atomic flush per CTA)
//   dc[o]   (+)= g                    (dc may be null)
//   dG[o,:] (+)= g * e                (WRITE_G; dense because torch.optim.AdamW wants a dense .grad -- with WRITE_G = false the rank-1
//                                      factors (dw, e) are kept instead and the kernel only READS G: half the traffic)
template <int NJ, bool WRITE_G>
__global__ void __launch_bounds__(256, WRITE_G ? 1 : 2)
generator_bwd_kernel(const float* __restrict__ Gw, long long ldw, long long O, int D, const float* __restrict__ dw, float dw_scale,
                     const float* __restrict__ e, float* __restrict__ dG, long long ldg, float* __restrict__ dc, float* __restrict__ de, int accumulate) {
  pdl_prologue();
  extern __shared__ float sde[];      // [D]
  for (int d = threadIdx.x; d < D; d += blockDim.x) sde[d] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long nw = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  long long o = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  float4 dacc[NJ], ev[NJ], w0[NJ], w1[NJ], g0[NJ], g1[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int d = lane * 4 + 128 * j;
    dacc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    ev[j] = (WRITE_G && d < D) ? __ldg(reinterpret_cast<const float4*>(e + d)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  auto load = [&](long long r, float4 (&w)[NJ], float4 (&gv)[NJ]) {
    gen_load_row<NJ>(Gw + r * ldw, D, lane, w);
    if (WRITE_G && accumulate) gen_load_row<NJ>(dG + r * ldg, D, lane, gv);
  };
  auto row = [&](long long r, const float4 (&w)[NJ], const float4 (&gv)[NJ]) {
    const float g = dw[r] * dw_scale;
    if (lane == 0 && dc != nullptr) dc[r] = accumulate ? dc[r] + g : g;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int d = lane * 4 + 128 * j;
      dacc[j].x = fmaf(g, w[j].x, dacc[j].x); dacc[j].y = fmaf(g, w[j].y, dacc[j].y);
      dacc[j].z = fmaf(g, w[j].z, dacc[j].z); dacc[j].w = fmaf(g, w[j].w, dacc[j].w);
      if (WRITE_G && d < D) {
        float4 acc = accumulate ? gv[j] : make_float4(0.f, 0.f, 0.f, 0.f);
        acc.x = fmaf(g, ev[j].x, acc.x); acc.y = fmaf(g, ev[j].y, acc.y); acc.z = fmaf(g, ev[j].z, acc.z); acc.w = fmaf(g, ev[j].w, acc.w);
        __stcs(reinterpret_cast<float4*>(dG + r * ldg + d), acc);
      }
    }
  };
  if (o < O) load(o, w0, g0);
  if (o + nw < O) load(o + nw, w1, g1);
  for (; o < O; o += 2 * nw) {
    row(o, w0, g0);
    if (o + 2 * nw < O) load(o + 2 * nw, w0, g0);
    if (o + nw < O) {
      row(o + nw, w1, g1);
      if (o + 3 * nw < O) load(o + 3 * nw, w1, g1);
    }
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int d = lane * 4 + 128 * j;
    if (d < D) {
      atomicAdd(&sde[d], dacc[j].x); atomicAdd(&sde[d + 1], dacc[j].y);
      atomicAdd(&sde[d + 2], dacc[j].z); atomicAdd(&sde[d + 3], dacc[j].w);
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) atomicAdd(de + d, sde[d]);
}

// -------------------------------------------------------------------------------------------------------------------
// k4 + k6: the support-set pooling.  The reference runs 1-head self-attention over S tokens and keeps rows 0..NQ-1
// (hypernet.py:46-82,175).  With q~_i = Wk^T q_i the scores are  (s_t . q~_i + q_i.bk) / sqrt(D)  and the context is
// e_i = Wv (sum_t P~[i,t] s_t) + bv * sum_t P~[i,t], so no K / V projection of the S tokens is ever materialised.
// The whole chain (and its backward) runs as one cooperative kernel per direction: pool_coop.cuh.  Shared definitions:
//   s_t = seq_t + PE_t; seq is given as two pieces: `prefix` rows [0, NQ) and `z` rows [NQ, S).
// -------------------------------------------------------------------------------------------------------------------
struct PoolParams {
  const float* prefix; const float* z; long long ldz; const float* pe; long long ldpe;   // pe may be nullptr
  int NQ, S, D;
  const float* qt;        // [NQ, D]  q~
  const float* qb;        // [NQ]     q_i . bk
  const float* keep;      // [NQ, S] dropout keep mask (0/1) or nullptr
  float keep_scale;       // 1/(1-p)
  float inv_sqrt_d;
  float* P;               // [NQ, S] raw scores (scratch, written by pool_scores_kernel)
  float* Pout;            // [NQ, S] softmax weights (before dropout), stash for backward
  float* c;               // [NQ, D]
  float* psum;            // [NQ] sum_t P~[i,t]
};

__device__ __forceinline__ float pool_token(const PoolParams& p, int t, int d) {
  const float base = (t < p.NQ) ? p.prefix[static_cast<long long>(t) * p.D + d] : p.z[static_cast<long long>(t - p.NQ) * p.ldz + d];
  return p.pe ? base + p.pe[static_cast<long long>(t) * p.ldpe + d] : base;
}

// -------------------------------------------------------------------------------------------------------------------
// k13: prefix splice (mmmodel.py:36-48).  out[b,0,:] = projected[b,:]; out[b,1+t,:] = table[ids[b,t],:];
// labels_out[b,:] = [-100, labels[b,:]]; mask_out[b,:] = [1, mask[b,:]].  One CTA per output row, 16-byte vectors.
// -------------------------------------------------------------------------------------------------------------------
struct SpliceParams {
  const float* proj_f32; const bf16* proj_bf16; long long ld_proj;   // exactly one of the two
  const void* table; int table_is_bf16; long long ld_table; long long vocab;
  const long long* ids; int B, T, H;
  void* out; int out_is_bf16;                       // [B, 1+T, H]
  const long long* labels; long long* labels_out;   // [B,T] -> [B,1+T] (may be nullptr)
  const void* mask; int mask_is_i64; float* mask_out;   // [B,T] (int64 or float) -> float [B,1+T] (may be nullptr)
  int* error_flag;                                  // set to 1 if an id is outside [0, vocab)
};

__device__ __forceinline__ float load_as_f32(const void* base, int is_bf16, long long idx) {
  return is_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(base)[idx]) : reinterpret_cast<const float*>(base)[idx];
}

__global__ void __launch_bounds__(256)
splice_kernel(const SpliceParams p) {
  pdl_prologue();
  const int row = blockIdx.x;                 // b * (1+T) + pos
  const int b = row / (1 + p.T), pos = row % (1 + p.T);
  const long long obase = static_cast<long long>(row) * p.H;
  if (threadIdx.x == 0) {
    if (p.labels_out) p.labels_out[row] = (pos == 0) ? -100LL : p.labels[static_cast<long long>(b) * p.T + pos - 1];
    if (p.mask_out) {
      float mv = 1.0f;
      if (pos > 0) {
        const long long mi = static_cast<long long>(b) * p.T + pos - 1;
        mv = p.mask_is_i64 ? static_cast<float>(reinterpret_cast<const long long*>(p.mask)[mi]) : reinterpret_cast<const float*>(p.mask)[mi];
      }
      p.mask_out[row] = mv;
    }
  }
  const void* src;
  int src_bf16;
  long long sbase;
  if (pos == 0) {
    src = p.proj_f32 ? static_cast<const void*>(p.proj_f32) : static_cast<const void*>(p.proj_bf16);
    src_bf16 = p.proj_f32 ? 0 : 1;
    sbase = static_cast<long long>(b) * p.ld_proj;
  } else {
    long long id = p.ids[static_cast<long long>(b) * p.T + pos - 1];
    if (id < 0 || id >= p.vocab) {
      if (threadIdx.x == 0 && p.error_flag) atomicExch(p.error_flag, 1);
      id = 0;
    }
    src = p.table; src_bf16 = p.table_is_bf16; sbase = id * p.ld_table;
  }
  // fast paths: same dtype -> 16-byte copies; bf16 -> fp32 widening (the reference's torch.cat promotion)
  if (src_bf16 && p.out_is_bf16 && (p.H & 7) == 0 && (sbase & 7) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(src) + sbase);
    uint4* d4 = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + obase);
    for (int i = threadIdx.x; i < p.H / 8; i += blockDim.x) d4[i] = __ldg(s4 + i);
  } else if (!src_bf16 && !p.out_is_bf16 && (p.H & 3) == 0 && (sbase & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + sbase);
    float4* d4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + obase);
    for (int i = threadIdx.x; i < p.H / 4; i += blockDim.x) d4[i] = __ldg(s4 + i);
  } else if (src_bf16 && !p.out_is_bf16 && (p.H & 7) == 0 && (sbase & 7) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(src) + sbase);
    float4* d4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + obase);
    for (int i = threadIdx.x; i < p.H / 8; i += blockDim.x) {
      const uint4 q = __ldg(s4 + i);
      const float2 a0 = unpack_bf16x2(q.x), a1 = unpack_bf16x2(q.y), a2 = unpack_bf16x2(q.z), a3 = unpack_bf16x2(q.w);
      d4[2 * i] = make_float4(a0.x, a0.y, a1.x, a1.y);
      d4[2 * i + 1] = make_float4(a2.x, a2.y, a3.x, a3.y);
    }
  } else {
    for (int i = threadIdx.x; i < p.H; i += blockDim.x) {
      const float v = load_as_f32(src, src_bf16, sbase + i);
      if (p.out_is_bf16) reinterpret_cast<bf16*>(p.out)[obase + i] = __float2bfloat16(v);
      else reinterpret_cast<float*>(p.out)[obase + i] = v;
    }
  }
}


// -------------------------------------------------------------------------------------------------------------------
// Embedding store gather: the embedding side of the reference collates + EmbeddingManager.get_embeddings
//   emb = FloatTensor(item['emb'])[selected_features]; embs = stack(...); embs -= emb_mean      (dmi/data/base.py:222-232)
//   embs = embs.to(device); embs /= embs.norm(dim=1, keepdim=True)                               (dmi/utils/model_utils.py:47-62)
// as ONE pass over a flat device-resident [N, D_store] table: row gather by sample index, optional column gather
// (InfFS feature selection), optional mean subtraction, optional L2 normalisation, fp32 and/or bf16 output.  One warp per row.
// -------------------------------------------------------------------------------------------------------------------
struct GatherParams {
  const void* store; int store_is_bf16; long long ld_store; long long n_rows;
  const long long* idx; int B;
  const int* sel; int d_store;         // column gather (values in [0, d_store)) or nullptr
  const float* mean;                   // [d_out] or nullptr
  int d_out; int normalize;
  float* out; long long ldo;
  bf16* out_bf16; long long ldo_bf16;
  int* error_flag;
  int vec_ok;                          // host-checked: d_out % 8 == 0, d_out <= 1024, every base pointer 16-byte aligned, every ld % 8 == 0
};

__global__ void __launch_bounds__(256)
gather_rows_kernel(const GatherParams p) {
  pdl_prologue();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.B) return;
  const long long src = p.idx != nullptr ? p.idx[row] : row;
  if (src < 0 || src >= p.n_rows) {
    // the reference raises an IndexError on the host here; on the device the row is poisoned with NaN (never left uninitialised)
    // and the error flag is raised for the caller's deferred check
    if (lane == 0 && p.error_flag != nullptr) atomicExch(p.error_flag, 1);
    const float qnan = __int_as_float(0x7fc00000);
    for (int j = lane; j < p.d_out; j += 32) {
      if (p.out != nullptr) p.out[static_cast<long long>(row) * p.ldo + j] = qnan;
      if (p.out_bf16 != nullptr) p.out_bf16[static_cast<long long>(row) * p.ldo_bf16 + j] = __float2bfloat16(qnan);
    }
    return;
  }
  const float* sf = reinterpret_cast<const float*>(p.store) + src * p.ld_store;
  const bf16* sb = reinterpret_cast<const bf16*>(p.store) + src * p.ld_store;
  // Fast path (no column gather, 16-byte aligned rows, d_out <= 1024): the row is read ONCE with 128-bit loads and lives in registers
  // (8 values per lane per 256-column block) through the norm and the stores.  Same arithmetic, same order of the per-lane partial sums
  // as the generic path is NOT required (fp32 sum of squares: 1e-6 parity, tested); the division stays a true division.
  if (p.sel == nullptr && p.vec_ok) {
    constexpr int NB = 4;                       // 256-column blocks
    float v[NB][8];
    float ssq = 0.f;
#pragma unroll
    for (int blk = 0; blk < NB; ++blk) {
      const int j0 = blk * 256 + lane * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[blk][e] = 0.f;
      if (j0 < p.d_out) {
        if (p.store_is_bf16) {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(sb + j0));
          const float2 a0 = unpack_bf16x2(q.x), a1 = unpack_bf16x2(q.y), a2 = unpack_bf16x2(q.z), a3 = unpack_bf16x2(q.w);
          v[blk][0] = a0.x; v[blk][1] = a0.y; v[blk][2] = a1.x; v[blk][3] = a1.y; v[blk][4] = a2.x; v[blk][5] = a2.y; v[blk][6] = a3.x; v[blk][7] = a3.y;
        } else {
          const float4 lo = __ldg(reinterpret_cast<const float4*>(sf + j0)), hi = __ldg(reinterpret_cast<const float4*>(sf + j0 + 4));
          v[blk][0] = lo.x; v[blk][1] = lo.y; v[blk][2] = lo.z; v[blk][3] = lo.w; v[blk][4] = hi.x; v[blk][5] = hi.y; v[blk][6] = hi.z; v[blk][7] = hi.w;
        }
        if (p.mean != nullptr) {
          const float4 m0 = __ldg(reinterpret_cast<const float4*>(p.mean + j0)), m1 = __ldg(reinterpret_cast<const float4*>(p.mean + j0 + 4));
          v[blk][0] -= m0.x; v[blk][1] -= m0.y; v[blk][2] -= m0.z; v[blk][3] -= m0.w; v[blk][4] -= m1.x; v[blk][5] -= m1.y; v[blk][6] -= m1.z; v[blk][7] -= m1.w;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) ssq = fmaf(v[blk][e], v[blk][e], ssq);
      }
    }
    float nrm1 = 1.0f;
    if (p.normalize) nrm1 = sqrtf(warp_sum(ssq));
#pragma unroll
    for (int blk = 0; blk < NB; ++blk) {
      const int j0 = blk * 256 + lane * 8;
      if (j0 >= p.d_out) continue;
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = p.normalize ? v[blk][e] / nrm1 : v[blk][e];
      if (p.out != nullptr) {
        float4* d = reinterpret_cast<float4*>(p.out + static_cast<long long>(row) * p.ldo + j0);
        d[0] = make_float4(o[0], o[1], o[2], o[3]);
        d[1] = make_float4(o[4], o[5], o[6], o[7]);
      }
      if (p.out_bf16 != nullptr)
        *reinterpret_cast<uint4*>(p.out_bf16 + static_cast<long long>(row) * p.ldo_bf16 + j0) =
            make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    }
    return;
  }
  float ss = 0.f;
  for (int j = lane; j < p.d_out; j += 32) {
    const int c = p.sel != nullptr ? p.sel[j] : j;
    float v = p.store_is_bf16 ? __bfloat162float(sb[c]) : sf[c];
    if (p.mean != nullptr) v -= p.mean[j];
    ss = fmaf(v, v, ss);
  }
  float nrm = 1.0f;
  if (p.normalize) {
    ss = warp_sum(ss);
    nrm = sqrtf(ss);
  }
  for (int j = lane; j < p.d_out; j += 32) {
    const int c = p.sel != nullptr ? p.sel[j] : j;
    float v = p.store_is_bf16 ? __bfloat162float(sb[c]) : sf[c];
    if (p.mean != nullptr) v -= p.mean[j];
    if (p.normalize) v = v / nrm;              // x / ||x||, a true division as in the reference (model_utils.py:59)
    if (p.out != nullptr) p.out[static_cast<long long>(row) * p.ldo + j] = v;
    if (p.out_bf16 != nullptr) p.out_bf16[static_cast<long long>(row) * p.ldo_bf16 + j] = __float2bfloat16(v);
  }
}

}  // namespace dmi
