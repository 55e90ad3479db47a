// Small HBM-bound helper kernels of the adapted-projector path: dtype packing, transposes, GELU' multiply.
#pragma once
#include "common.cuh"

namespace dmi {

// dst[row, 0:cols] (bf16, ld_dst) = scale * src[row, 0:cols] (fp32, ld_src); cols % 8 == 0, 8 elements per thread.
__global__ void cvt_rows_f32_bf16_kernel(const float* __restrict__ src, long long ld_src, bf16* __restrict__ dst,
                                         long long ld_dst, long long rows, int cols, float scale) {
  pdl_prologue();
  const int c8 = cols >> 3;
  const long long total = rows * c8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / c8;
    const int c = static_cast<int>(i % c8) * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + row * ld_src + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src + row * ld_src + c + 4));
    uint4 q;
    q.x = pack_bf16x2(a.x * scale, a.y * scale); q.y = pack_bf16x2(a.z * scale, a.w * scale);
    q.z = pack_bf16x2(b.x * scale, b.y * scale); q.w = pack_bf16x2(b.z * scale, b.w * scale);
    *reinterpret_cast<uint4*>(dst + row * ld_dst + c) = q;
  }
}

// dst[i*ld_dst + j] (bf16) = scale * src[j*ld_src + i] (fp32), i < n_i, j < n_j: tiled transpose through shared memory.
__global__ void transpose_f32_bf16_kernel(const float* __restrict__ src, long long ld_src, bf16* __restrict__ dst,
                                          long long ld_dst, int n_i, int n_j, float scale) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  for (int dj = threadIdx.y; dj < 32; dj += blockDim.y) {
    const int j = j0 + dj, i = i0 + threadIdx.x;
    tile[dj][threadIdx.x] = (i < n_i && j < n_j) ? src[static_cast<long long>(j) * ld_src + i] : 0.f;
  }
  __syncthreads();
  for (int di = threadIdx.y; di < 32; di += blockDim.y) {
    const int i = i0 + di, j = j0 + threadIdx.x;
    if (i < n_i && j < n_j) dst[static_cast<long long>(i) * ld_dst + j] = __float2bfloat16(tile[threadIdx.x][di] * scale);
  }
}

// out[i] = a[i] + (b ? b[i] : 0)
__global__ void add_vec_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + (b != nullptr ? b[i] : 0.f);
}

// H1 backward (reference lora_forward stops after the first GELU): dpre = dy * gelu'(pre)   (bf16 out)
__global__ void gelu_bwd_rows_kernel(const float* __restrict__ dy, long long lddy, const bf16* __restrict__ pre, long long ldpre,
                                     bf16* __restrict__ dpre, long long lddpre, long long rows, int cols) {
  pdl_prologue();
  const int c8 = cols >> 3;
  const long long total = rows * c8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / c8;
    const int c = static_cast<int>(i % c8) * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(dy + row * lddy + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(dy + row * lddy + c + 4));
    const uint4 pq = __ldg(reinterpret_cast<const uint4*>(pre + row * ldpre + c));
    const float2 p0 = unpack_bf16x2(pq.x), p1 = unpack_bf16x2(pq.y), p2 = unpack_bf16x2(pq.z), p3 = unpack_bf16x2(pq.w);
    uint4 q;
    q.x = pack_bf16x2(a.x * gelu_tanh_grad(p0.x), a.y * gelu_tanh_grad(p0.y));
    q.y = pack_bf16x2(a.z * gelu_tanh_grad(p1.x), a.w * gelu_tanh_grad(p1.y));
    q.z = pack_bf16x2(b.x * gelu_tanh_grad(p2.x), b.y * gelu_tanh_grad(p2.y));
    q.w = pack_bf16x2(b.z * gelu_tanh_grad(p3.x), b.w * gelu_tanh_grad(p3.y));
    *reinterpret_cast<uint4*>(dpre + row * lddpre + c) = q;
  }
}

// out[q] += scale * sum_b src[b, q]   (bf16 in, fp32 atomic accumulate).  CTA = 64 columns x one batch split.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const bf16* __restrict__ src, long long ld, long long rows, int cols, float* __restrict__ out, float scale,
                   long long rows_per_split) {
  pdl_prologue();
  __shared__ float red[32][65];
  const int cg = threadIdx.x & 7, rl = threadIdx.x >> 3;          // 8 column groups of 8, 32 row lanes
  const int c = blockIdx.x * 64 + cg * 8;
  const long long r0 = blockIdx.y * rows_per_split;
  const long long r1 = (r0 + rows_per_split < rows) ? (r0 + rows_per_split) : rows;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < cols) {
    for (long long r = r0 + rl; r < r1; r += 32) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(src + r * ld + c));
      const float2 a0 = unpack_bf16x2(q.x), a1 = unpack_bf16x2(q.y), a2 = unpack_bf16x2(q.z), a3 = unpack_bf16x2(q.w);
      acc[0] += a0.x; acc[1] += a0.y; acc[2] += a1.x; acc[3] += a1.y;
      acc[4] += a2.x; acc[5] += a2.y; acc[6] += a3.x; acc[7] += a3.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cg * 8 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) s += red[i][threadIdx.x];
    const int col = blockIdx.x * 64 + threadIdx.x;
    if (col < cols) atomicAdd(out + col, s * scale);
  }
}

// One launch that lays a generated adapter out for the GEMMs (all pieces are r x H or smaller; L2-resident):
//   w1ext[h, D+j] = s*B0[j,h]   a0t[j,d] = A0[d,j]   b0[j,h] = s*B0[j,h]   bias0[h] = b1[h]+beta0[h]
//   w2ext[h, H+j] = s*B1[j,h]   w2text[h, H+j] = A1[h,j]   a1t[j,h] = A1[h,j]   b1bf[j,h] = s*B1[j,h]   bias1[h] = b2[h]+beta1[h]
struct AdapterPackParams {
  int D, H, r;
  float scale;
  const float *A0, *B0, *beta0, *A1, *B1, *beta1, *b1, *b2;
  bf16 *w1ext, *w2ext, *w2text, *a0t, *a1t, *b0, *b1bf;
  float *bias0, *bias1;
};

__host__ __device__ inline long long adapter_pack_items(const AdapterPackParams& p) {
  const long long rH = static_cast<long long>(p.r) * p.H, rD = static_cast<long long>(p.r) * p.D;
  long long n = p.H;                         // bias0
  if (p.A0) n += 2 * rH + rD;                // w1ext cols, b0, a0t
  if (p.bias1) n += p.H;
  if (p.A1) n += 4 * rH;                     // w2ext cols, w2text cols, a1t, b1bf
  return n;
}

__global__ void adapter_pack_kernel(const AdapterPackParams p) {
  pdl_prologue();
  const long long rH = static_cast<long long>(p.r) * p.H, rD = static_cast<long long>(p.r) * p.D;
  const long long total = adapter_pack_items(p);
  const int D = p.D, H = p.H, r = p.r;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long k = i;
    if (k < H) { p.bias0[k] = p.b1[k] + (p.beta0 ? p.beta0[k] : 0.f); continue; }
    k -= H;
    if (p.A0) {
      if (k < rH) { const int h = k / r, j = k % r; p.w1ext[static_cast<long long>(h) * (D + r) + D + j] = __float2bfloat16(p.scale * p.B0[static_cast<long long>(j) * H + h]); continue; }
      k -= rH;
      if (k < rH) { p.b0[k] = __float2bfloat16(p.scale * p.B0[k]); continue; }
      k -= rH;
      if (k < rD) { const int j = k / D, d = k % D; p.a0t[k] = __float2bfloat16(p.A0[static_cast<long long>(d) * r + j]); continue; }
      k -= rD;
    }
    if (p.bias1) {
      if (k < H) { p.bias1[k] = p.b2[k] + (p.beta1 ? p.beta1[k] : 0.f); continue; }
      k -= H;
    }
    if (p.A1) {
      if (k < rH) { const int h = k / r, j = k % r; p.w2ext[static_cast<long long>(h) * (H + r) + H + j] = __float2bfloat16(p.scale * p.B1[static_cast<long long>(j) * H + h]); continue; }
      k -= rH;
      if (k < rH) { const int h = k / r, j = k % r; p.w2text[static_cast<long long>(h) * (H + r) + H + j] = __float2bfloat16(p.A1[k]); continue; }
      k -= rH;
      if (k < rH) { const int j = k / H, h = k % H; p.a1t[k] = __float2bfloat16(p.A1[static_cast<long long>(h) * r + j]); continue; }
      k -= rH;
      if (k < rH) { p.b1bf[k] = __float2bfloat16(p.scale * p.B1[k]); continue; }
    }
  }
}

// Adapter merge (reference Projector.combine_lora, projector.py:95-103), exact fp32:
//   Wm[o,i] = W[o,i] + scale * sum_j A[i,j] * B[j,o]      bm[o] = b[o] + beta[o]
// CTA = 32 (o) x 32 (i) outputs; A and B tiles staged in shared memory.  One-off per few-shot run, 2*in*H*r FLOP.
__global__ void __launch_bounds__(256)
merge_adapter_kernel(const float* __restrict__ W, long long ldw, const float* __restrict__ bias, const float* __restrict__ A,
                     const float* __restrict__ B, const float* __restrict__ beta, int in_dim, int H, int r, float scale,
                     float* __restrict__ Wm, long long ldwm, float* __restrict__ bm) {
  pdl_prologue();
  __shared__ float sA[32][65];     // [i][j]
  __shared__ float sB[64][33];     // [j][o]
  const int o0 = blockIdx.y * 32, i0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  for (int idx = threadIdx.x; idx < 32 * r; idx += 256) {
    const int i = idx / r, j = idx % r;
    sA[i][j] = (i0 + i < in_dim) ? A[static_cast<long long>(i0 + i) * r + j] : 0.f;
  }
  for (int idx = threadIdx.x; idx < 32 * r; idx += 256) {
    const int j = idx / 32, o = idx % 32;
    sB[j][o] = (o0 + o < H) ? B[static_cast<long long>(j) * H + o0 + o] : 0.f;
  }
  __syncthreads();
  for (int oo = ty; oo < 32; oo += 8) {
    const int o = o0 + oo, i = i0 + tx;
    if (o < H && i < in_dim) {
      float acc = 0.f;
      for (int j = 0; j < r; ++j) acc = fmaf(sA[tx][j], sB[j][oo], acc);
      Wm[static_cast<long long>(o) * ldwm + i] = W[static_cast<long long>(o) * ldw + i] + scale * acc;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    const int o = o0 + threadIdx.x;
    if (o < H) bm[o] = bias[o] + (beta != nullptr ? beta[o] : 0.f);
  }
}


inline int ew_grid(long long work_items, int threads) {
  long long g = (work_items + threads - 1) / threads;
  const long long cap = 148LL * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace dmi
