// Small HBM-bound helper kernels of the adapted-projector path: dtype packing, transposes, GELU' multiply.
#pragma once
#include "common.cuh"

namespace dmi {

// dst[row, 0:cols] (bf16, ld_dst) = scale * src[row, 0:cols] (fp32, ld_src); cols % 8 == 0, 8 elements per thread.
__global__ void cvt_rows_f32_bf16_kernel(const float* __restrict__ src, long long ld_src, bf16* __restrict__ dst,
                                         long long ld_dst, long long rows, int cols, float scale) {
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
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + (b != nullptr ? b[i] : 0.f);
}

// H1 backward (reference lora_forward stops after the first GELU): dpre = dy * gelu'(pre)   (bf16 out)
__global__ void gelu_bwd_rows_kernel(const float* __restrict__ dy, long long lddy, const bf16* __restrict__ pre, long long ldpre,
                                     bf16* __restrict__ dpre, long long lddpre, long long rows, int cols) {
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

inline int ew_grid(long long work_items, int threads) {
  long long g = (work_items + threads - 1) / threads;
  const long long cap = 148LL * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace dmi
