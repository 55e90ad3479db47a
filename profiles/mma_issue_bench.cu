// Microbenchmark: cycles per tcgen05.mma (bf16, K=16) on every SM at once, for the mainloop variants the GEMM could use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../sample-efficient-multimodality_b200/csrc mma_issue_bench.cu -o mma_issue_bench
// Variants: cta_group 1/2, N, commit cadence (none / every 4 MMAs, local or multicast), rotating operand stages,
// concurrent bulk-copy traffic into the operand ring (shared-memory write bandwidth interference).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

using namespace dmi;

__device__ __forceinline__ void umma_f16_cg2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void commit_cg2_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void commit_cg2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}

struct Variant {
  int cg;          // cta_group 1 or 2
  int n;           // UMMA N
  int commit;      // 0: only at the end, 1: every 4 MMAs (local), 2: every 4 MMAs multicast to both CTAs (cg 2 only)
  int rotate;      // 1: operand descriptors rotate over the ring stages, 0: always stage 0
  int traffic;     // bytes per 4 MMAs bulk-copied into the ring by a second warp (0 = none)
  int kstep;       // MMAs per "k block" (commit / traffic cadence)
  int data;        // operand ring contents: 0 = zeros, 1 = random bf16 ~ N(0,1), 2 = left as found
  int nacc;        // 1: single accumulator, 2: alternate between two accumulators every 128 MMAs
};

constexpr int STAGE_BYTES = 48 * 1024;
constexpr int STAGES = 4;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;

template <int CG>
__global__ void __launch_bounds__(128, 1) bench_kernel(Variant v, int n_mma, const uint8_t* gsrc, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* done_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* dummy_bar = done_bar + 1;          // [STAGES] arrive-only
  uint64_t* copy_bar = dummy_bar + STAGES;     // [STAGES]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(copy_bar + STAGES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    mbar_init(done_bar, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&dummy_bar[s], 1); mbar_init(&copy_bar[s], 1); }
    fence_barrier_init();
  }
  if (warp == 0) {
    if (CG == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (v.data != 2) {
    uint32_t* w = reinterpret_cast<uint32_t*>(smem);
    for (int i = threadIdx.x; i < STAGES * STAGE_BYTES / 4; i += blockDim.x) {
      uint32_t h = (i + 1) * 2654435761u + blockIdx.x * 40503u;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      // two bf16 values with sign random, exponent in [124,127] (|x| in [0.125, 2)), random mantissa
      const uint32_t lo = ((h & 0x8000u)) | ((124u + ((h >> 7) & 3u)) << 7) | (h & 0x7Fu);
      const uint32_t hi = (((h >> 16) & 0x8000u)) | ((124u + ((h >> 23) & 3u)) << 7) | ((h >> 16) & 0x7Fu);
      w[i] = v.data == 1 ? (lo | (hi << 16)) : 0u;
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
  }
  const uint32_t idesc = make_idesc(CG == 2 ? 256 : 128, v.n, 1);
  long long t0 = 0, t1 = 0;
  if (warp == 0 && rank == 0) {
    if (lane == 0) {
      t0 = clock64();
      int stage = 0;
      for (int i = 0; i < n_mma; i += v.kstep) {
        const uint32_t sa = smem_u32(smem + (v.rotate ? stage : 0) * STAGE_BYTES);
        const uint64_t adesc = make_kmajor_sw128_desc(sa);
        const uint64_t bdesc = make_kmajor_sw128_desc(sa + 16384);
        const uint32_t d = tmem_base + ((v.nacc == 2) ? (((i >> 7) & 1) * 256) : 0);
        for (int k = 0; k < v.kstep; ++k) {
          if (CG == 1) umma_f16(d, adesc + 2 * (k & 3), bdesc + 2 * (k & 3), idesc, 1u);
          else         umma_f16_cg2(d, adesc + 2 * (k & 3), bdesc + 2 * (k & 3), idesc, 1u);
        }
        if (v.commit == 1) { if (CG == 1) umma_commit(&dummy_bar[stage]); else commit_cg2(&dummy_bar[stage]); }
        if (v.commit == 2 && CG == 2) commit_cg2_mc(&dummy_bar[stage], 0x3);
        if (++stage == STAGES) stage = 0;
      }
      if (CG == 1) umma_commit(done_bar); else commit_cg2(done_bar);
      mbar_wait(done_bar, 0);
      t1 = clock64();
      out_cycles[blockIdx.x] = t1 - t0;
    }
  } else if (warp == 1 && v.traffic > 0) {
    // free-running bulk copies into the ring: one batch of v.traffic bytes per k block's worth of MMA time is NOT enforced;
    // the copy warp simply keeps STAGES copies in flight for the same number of k blocks
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int nkb = n_mma / v.kstep;
      for (int i = 0; i < nkb; ++i) {
        if (i >= STAGES) mbar_wait(&copy_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&copy_bar[stage], v.traffic);
        const uint8_t* src = gsrc + (static_cast<size_t>(blockIdx.x) * 64 + (i & 63)) * STAGE_BYTES;
        bulk_g2s(smem + stage * STAGE_BYTES, src, v.traffic, &copy_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      // drain
      for (int s = 0; s < STAGES && s < nkb; ++s) {
        const int idx = (nkb - 1 - s);
        mbar_wait(&copy_bar[idx % STAGES], (idx / STAGES) & 1);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (CG == 1) tmem_dealloc(tmem_base, 512);
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

namespace dmi { void set_error(const char*, ...) {} }

int main(int argc, char** argv) {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int n_mma = argc > 1 ? atoi(argv[1]) : 4096;
  const bool few = argc > 2;
  uint8_t* gsrc;
  const size_t gbytes = static_cast<size_t>(sms) * 64 * STAGE_BYTES;
  CK(cudaMalloc(&gsrc, gbytes));
  CK(cudaMemset(gsrc, 0x3c, gbytes));
  long long* d_cyc;
  CK(cudaMalloc(&d_cyc, sizeof(long long) * sms));
  CK(cudaFuncSetAttribute(bench_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  CK(cudaFuncSetAttribute(bench_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  std::vector<Variant> vs = {
      {1, 256, 0, 0, 0, 4, 2, 1}, {1, 256, 1, 0, 0, 4, 2, 1}, {1, 256, 1, 1, 0, 4, 2, 1}, {1, 256, 0, 1, 0, 4, 2, 1}, {1, 128, 1, 1, 0, 4, 2, 1}, {1, 64, 1, 1, 0, 4, 2, 1},
      {1, 256, 1, 1, 0, 8, 2, 1}, {1, 256, 1, 1, 0, 16, 2, 1},
      {1, 256, 1, 1, 24576, 4, 2, 1}, {1, 256, 1, 1, 49152, 4, 2, 1}, {1, 128, 1, 1, 32768, 4, 2, 1},
      {2, 256, 0, 0, 0, 4, 2, 1}, {2, 256, 1, 1, 0, 4, 2, 1}, {2, 256, 2, 1, 0, 4, 2, 1}, {2, 256, 2, 1, 0, 8, 2, 1}, {2, 128, 2, 1, 0, 4, 2, 1},
      {2, 256, 2, 1, 16384, 4, 2, 1}, {2, 256, 2, 1, 32768, 4, 2, 1}, {2, 256, 2, 1, 49152, 4, 2, 1},
  };
  if (few) vs = {{1, 256, 1, 1, 0, 4, 0, 1}, {1, 256, 1, 1, 0, 4, 1, 1}, {1, 256, 1, 1, 0, 4, 1, 2}, {2, 256, 2, 1, 0, 4, 0, 1}, {2, 256, 2, 1, 0, 4, 1, 1},
              {2, 256, 2, 1, 0, 4, 1, 2}, {1, 256, 1, 1, 0, 4, 1, 2}, {2, 256, 2, 1, 0, 4, 1, 2}};
  std::vector<long long> h(sms);
  for (const Variant& v : vs) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaMemset(d_cyc, 0, sizeof(long long) * sms));
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(sms);
      cfg.blockDim = dim3(128);
      cfg.dynamicSmemBytes = SMEM_BYTES;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = v.cg;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      CK(cudaEventRecord(e0));
      if (v.cg == 1) CK(cudaLaunchKernelEx(&cfg, bench_kernel<1>, v, n_mma, (const uint8_t*)gsrc, d_cyc));
      else           CK(cudaLaunchKernelEx(&cfg, bench_kernel<2>, v, n_mma, (const uint8_t*)gsrc, d_cyc));
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      CK(cudaMemcpy(h.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
      long long mx = 0, sum = 0;
      int cnt = 0;
      for (int i = 0; i < sms; ++i) if (h[i] > 0) { mx = h[i] > mx ? h[i] : mx; sum += h[i]; ++cnt; }
      if (rep == 1) {
        const double cyc = cnt ? static_cast<double>(sum) / cnt / n_mma : 0;
        const double flops = 2.0 * (v.cg == 2 ? 256 : 128) * v.n * 16 * n_mma * (v.cg == 2 ? sms / 2 : sms);
        printf("cg=%d N=%3d commit=%d rotate=%d traffic=%5d kstep=%2d data=%d nacc=%d : %7.1f cycles/MMA (max %7.1f)  %8.1f us  %7.0f TFLOP/s  issuers=%d\n", v.cg, v.n,
               v.commit, v.rotate, v.traffic, v.kstep, v.data, v.nacc, cyc, static_cast<double>(mx) / n_mma, ms * 1e3, flops / (ms * 1e-3) / 1e12, cnt);
      }
    }
  }
  return 0;
}
