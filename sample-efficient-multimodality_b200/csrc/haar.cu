// C ABI, part 4: on-device Haar-random orthogonal matrix.
//
// The reference draws its isometry on the HOST for every training micro-step:
//   R = torch.FloatTensor(scipy.stats.ortho_group.rvs(mm_dim)).to(device)            (dmi/train_hypernet.py:56-57)
// i.e. LAPACK QR of an n x n Gaussian matrix plus the sign fix Q <- Q diag(sign(r_kk)) -- 72 ms at n = 768 and 0.7 s at n = 2048
// (SURVEY section 8a, row a2), two orders of magnitude more than the rest of the micro-step on a B200.
//
// Same distribution, no QR of a matrix: by the rotational invariance of the Gaussian, the k-th Householder reflector of that QR is
// determined by an independent Gaussian vector x_k of dimension n-k (Stewart 1980, "The efficient generation of random orthogonal
// matrices"; Mezzadri 2007).  With v_k = x_k + sign(x_k0) ||x_k|| e_0 (embedded in coordinates k..n-1), H_k = I - 2 v_k v_k^T / v_k^T v_k
// and d_k = -sign(x_k0) (= sign of the r_kk the QR would have produced):
//                                Q = H_0 H_1 ... H_{n-1} diag(d)
// All n reflector vectors are available up front, so the product is applied in blocks of 64 reflectors in compact WY form
// (H_b0 ... H_b63 = I - V_b T_b V_b^T,  T_b^{-1} = striu(V_b^T V_b) + diag(v_k^T v_k)/2), as LAPACK's dorgqr does:
//   for b = last .. first:   W = V_b^T Q;   Q -= V_b (T_b W)
// -- three small fp32 GEMMs per block (4 n^3 FLOP in total, 1.8 GFLOP at n = 768).  fp32 CUDA-core arithmetic throughout: the
// result is orthogonal to ~1e-6, which the 1e-5 parity budget of the rotation needs; tensor-core (tf32) products would not be.
#include <string.h>

#include "../../include/dmi_b200.h"
#include "common.cuh"

namespace dmi {

void count_launch();

constexpr int HB = 64;           // reflectors per block

// Row k of `gauss` (length n) holds x_k in its entries k..n-1 (entries 0..k-1 are ignored).
// Vt[k, i] = v_k[i] (0 for i < k),  vnorm2[k] = v_k^T v_k,  d[k] = -sign(x_k0).   One warp per reflector.
__global__ void __launch_bounds__(256)
haar_reflectors_kernel(const float* __restrict__ gauss, int n, float* __restrict__ Vt, float* __restrict__ vnorm2, float* __restrict__ d) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= n) return;
  const float* x = gauss + static_cast<long long>(k) * n;
  float ss = 0.f;
  for (int i = k + lane; i < n; i += 32) ss = fmaf(x[i], x[i], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float x0 = x[k];
  const float sgn = x0 >= 0.f ? 1.f : -1.f;
  const float nrm = sqrtf(ss);
  const float v0 = x0 + sgn * nrm;
  float* v = Vt + static_cast<long long>(k) * n;
  for (int i = lane; i < n; i += 32) v[i] = i < k ? 0.f : (i == k ? v0 : x[i]);
  if (lane == 0) {
    vnorm2[k] = ss - x0 * x0 + v0 * v0;
    d[k] = -sgn;
  }
}

// Generic fp32 GEMM on the CUDA cores with arbitrary element strides:  C[m,n] = alpha * sum_k A(m,k) B(k,n) + beta * C[m,n],
// A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn], C row-major with leading dimension ldc.  64x64 tile, 16-deep slabs,
// 256 threads x (4x4) outputs.  Optional column scale of the result (the diag(d) of the last update).
struct SgemmParams {
  const float* A; long long sam, sak;
  const float* B; long long sbk, sbn;
  float* C; long long ldc;
  int M, N, K;
  float alpha, beta;
  const float* colscale;           // [N] or nullptr: C[m,n] = (alpha*acc + beta*C[m,n]) * colscale[n]
  int kchunk;                      // K range per blockIdx.z; gridDim.z > 1: partial sums are atomically added into a ZEROED C
};

__global__ void __launch_bounds__(256)
sgemm_strided_kernel(const SgemmParams p) {
  __shared__ float sA[16][64 + 4];
  __shared__ float sB[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int k_begin = blockIdx.z * p.kchunk;
  const int k_end = min(p.K, k_begin + p.kchunk);
  for (int k0 = k_begin; k0 < k_end; k0 += 16) {
    for (int e = threadIdx.x; e < 16 * 64; e += 256) {
      // pick the faster-varying index along the unit-stride dimension of each operand
      int kk, mm;
      if (p.sak == 1) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
      const int m = m0 + mm, k = k0 + kk;
      sA[kk][mm] = (m < p.M && k < k_end) ? p.A[m * p.sam + k * p.sak] : 0.f;
      int kb, nn;
      if (p.sbk == 1) { kb = e & 15; nn = e >> 4; } else { nn = e & 63; kb = e >> 6; }
      const int n = n0 + nn, k2 = k0 + kb;
      sB[kb][nn] = (n < p.N && k2 < k_end) ? p.B[k2 * p.sbk + n * p.sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float* c = p.C + static_cast<long long>(m) * p.ldc + n;
      float v = p.alpha * acc[i][j];
      if (gridDim.z > 1) { atomicAdd(c, v); continue; }
      if (p.beta != 0.f) v = fmaf(p.beta, *c, v);
      if (p.colscale != nullptr) v *= p.colscale[n];
      *c = v;
    }
  }
}

// T = U^{-1} with U = striu(S) + diag(vnorm2)/2 (upper triangular nb x nb).  One CTA; thread j owns column j of T and solves
// U t = e_j by back substitution entirely in shared memory (U[i][l] is a broadcast read, Tsh[l][j] is conflict-free across j).
__global__ void __launch_bounds__(HB)
wy_tfactor_kernel(const float* __restrict__ S, int lds, const float* __restrict__ vnorm2, int nb, float* __restrict__ T, int ldt) {
  __shared__ float U[HB][HB + 1];
  __shared__ float Tsh[HB][HB + 1];
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
    const int i = e / nb, j = e % nb;
    U[i][j] = (j > i) ? S[i * lds + j] : (j == i ? 0.5f * vnorm2[i] : 0.f);
  }
  __syncthreads();
  const int j = threadIdx.x;
  if (j < nb) {
#pragma unroll 1
    for (int i = nb - 1; i >= 0; --i) {
      float t = 0.f;
      if (i <= j) {
        float s = (i == j) ? 1.f : 0.f;
        for (int l = i + 1; l <= j; ++l) s = fmaf(-U[i][l], Tsh[l][j], s);      // the inverse is upper triangular: T[l][j] = 0 for l > j
        t = s / U[i][i];
      }
      Tsh[i][j] = t;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) T[(e / nb) * ldt + (e % nb)] = Tsh[e / nb][e % nb];
}

__global__ void set_identity_kernel(float* Q, int n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < static_cast<long long>(n) * n) Q[i] = (i / n == i % n) ? 1.f : 0.f;
}

int num_sms();

// split_k: when the output has too few 64x64 tiles to fill the GPU (the [64, n] products of the WY update), the contraction is
// split over blockIdx.z and accumulated with atomics into C, which is zeroed here first (beta must be 0, no column scale).
static int sgemm(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* C, long long ldc, int M, int N, int K,
                 float alpha, float beta, const float* colscale, cudaStream_t s, bool split_k = false) {
  SgemmParams p;
  p.A = A; p.sam = sam; p.sak = sak; p.B = B; p.sbk = sbk; p.sbn = sbn; p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
  p.alpha = alpha; p.beta = beta; p.colscale = colscale;
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  p.kchunk = K > 0 ? K : 1;
  if (split_k && beta == 0.f && colscale == nullptr) {
    const int tiles = static_cast<int>(grid.x * grid.y);
    int z = (2 * num_sms() + tiles - 1) / tiles;
    const int max_z = (K + 63) / 64;
    if (z > max_z) z = max_z;
    if (z > 1) {
      p.kchunk = (((K + z - 1) / z) + 15) / 16 * 16;
      grid.z = (K + p.kchunk - 1) / p.kchunk;
      if (grid.z > 1) DMI_CHECK_CUDA(cudaMemset2DAsync(C, ldc * sizeof(float), 0, N * sizeof(float), M, s));
    }
  }
  sgemm_strided_kernel<<<grid, 256, 0, s>>>(p);
  DMI_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DMI_OK;
}

}  // namespace dmi

using namespace dmi;

extern "C" {

// floats of workspace: Vt[n*n] + vnorm2[n] + d[n] + S[HB*HB] + T[HB*HB] + W[HB*n] + W2[HB*n]
int64_t dmi_haar_workspace_bytes(int64_t n) {
  return static_cast<int64_t>(sizeof(float)) * (n * n + 2 * n + 2 * HB * HB + 2 * HB * n + 64);
}

int dmi_haar_orthogonal(const float* gauss, int64_t n64, float* Q, void* workspace, uint64_t workspace_bytes, void* stream) {
  DMI_REQUIRE(gauss != nullptr && Q != nullptr && workspace != nullptr && n64 >= 1 && n64 <= 16384, "haar_orthogonal: bad arguments (n=%lld)", (long long)n64);
  DMI_REQUIRE(workspace_bytes >= static_cast<uint64_t>(dmi_haar_workspace_bytes(n64)), "haar_orthogonal: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n = static_cast<int>(n64);
  float* Vt = static_cast<float*>(workspace);
  float* vnorm2 = Vt + static_cast<long long>(n) * n;
  float* d = vnorm2 + n;
  float* S = d + n;
  float* T = S + HB * HB;
  float* W = T + HB * HB;
  float* W2 = W + static_cast<long long>(HB) * n;
  haar_reflectors_kernel<<<(n + 7) / 8, 256, 0, s>>>(gauss, n, Vt, vnorm2, d);
  DMI_CHECK_CUDA(cudaGetLastError());
  count_launch();
  set_identity_kernel<<<static_cast<unsigned>((static_cast<long long>(n) * n + 255) / 256), 256, 0, s>>>(Q, n);
  DMI_CHECK_CUDA(cudaGetLastError());
  count_launch();
  const int nblocks = (n + HB - 1) / HB;
  for (int b = nblocks - 1; b >= 0; --b) {
    const int k0 = b * HB;
    const int nb = (n - k0 < HB) ? (n - k0) : HB;
    const float* Vb = Vt + static_cast<long long>(k0) * n;            // [nb, n] row-major: row j = v_{k0+j}; columns < k0 are zero
    // S = V_b V_b^T over columns k0..n-1
    int rc = sgemm(Vb + k0, n, 1, Vb + k0, 1, n, S, HB, nb, nb, n - k0, 1.f, 0.f, nullptr, s, true);
    if (rc != DMI_OK) return rc;
    wy_tfactor_kernel<<<1, HB, 0, s>>>(S, HB, vnorm2 + k0, nb, T, HB);
    DMI_CHECK_CUDA(cudaGetLastError());
    count_launch();
    // W = V_b Q  (rows of Q below k0 only: V_b is zero in columns < k0)       [nb, n]
    rc = sgemm(Vb + k0, n, 1, Q + static_cast<long long>(k0) * n, n, 1, W, n, nb, n, n - k0, 1.f, 0.f, nullptr, s, true);
    if (rc != DMI_OK) return rc;
    // W2 = T W                                                                   [nb, n]
    rc = sgemm(T, HB, 1, W, n, 1, W2, n, nb, n, nb, 1.f, 0.f, nullptr, s);
    if (rc != DMI_OK) return rc;
    // Q[k0:, :] -= V_b^T W2 ; the last update (b == 0) also applies the column signs diag(d)
    rc = sgemm(Vb + k0, 1, n, W2, n, 1, Q + static_cast<long long>(k0) * n, n, n - k0, n, nb, -1.f, 1.f, b == 0 ? d : nullptr, s);
    if (rc != DMI_OK) return rc;
  }
  return DMI_OK;
}

}  // extern "C"
