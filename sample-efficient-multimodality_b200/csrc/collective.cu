// One-shot all-reduce of the adapter gradients over NVLink 5 / NVSwitch peer memory (data-parallel step).
//
// The per-step payload of the adapted-projector path is small (dA/dB/dbeta of both layers: 0.9 MB), so an all-reduce is pure latency.
// Issued through NCCL on a side stream it also collides with the step's persistent one-CTA-per-SM GEMMs: whichever SM the NCCL kernel
// occupies delays that SM's share of the next GEMM by the whole collective (measured at N = 2: 86 us of exposed time for a 0.9 MB
// all-reduce, profiles/r2_bench_n2b.json).  This kernel uses small CTAs (128 threads, no shared memory) that co-reside with a persistent
// GEMM CTA instead of waiting for -- or blocking -- an SM, so it can run on a side stream underneath the next step (the waiting for the
// slowest rank at the barrier is then hidden too), or in the step's own stream right after the last gradient kernel:
//   barrier-in   every CTA tells every peer (release, system scope) that this GPU's gradients are complete and waits for theirs
//   reduce       each element is read from all `world` copies -- one multimem.ld_reduce through the NVSwitch (in-switch reduction,
//                NVLS) when the buffer has a multicast mapping, otherwise `world` peer loads over NVLink -- and the sum is written to
//                the LOCAL output buffer only (no peer is written to, so nothing is in flight when the kernel ends)
//   barrier-out  nobody leaves before every peer has finished reading this GPU's copy (the buffer is zeroed again next step)
// The flags live in a second peer-mapped buffer: slot [cta][peer] is incremented once per barrier and never reset, so barrier k of
// the process waits for the value k (wrap-safe comparison); `epoch` (1, 2, 3, ...) is the call count, identical on all ranks.
// Every spin is bounded (~2 s) and traps, so that a lost rank surfaces as a CUDA error instead of a hung box.
#include <string.h>

#include "../../include/dmi_b200.h"
#include "common.cuh"

namespace dmi {

void count_launch();

namespace {

constexpr int AR_MAX_WORLD = 8;
constexpr int AR_MAX_CTAS = 64;
constexpr int AR_THREADS = 128;      // small CTAs (128 threads, < 64 registers, no shared memory) fit NEXT TO a resident persistent GEMM CTA

struct ArParams {
  const float* in[AR_MAX_WORLD];
  unsigned int* flags[AR_MAX_WORLD];
  const float* mc;          // multicast address of the input buffer, or null
  float* out;
  long long n;              // floats, multiple of 4
  int rank, world;
  unsigned int epoch;
  float scale;
};

__device__ __forceinline__ void ar_barrier(const ArParams& p, unsigned int target) {
  __syncthreads();                                   // every thread of the CTA has finished the phase before
  if (threadIdx.x < p.world) {
    const int peer = threadIdx.x;
    unsigned int* remote = p.flags[peer] + blockIdx.x * AR_MAX_WORLD + p.rank;
    asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(remote) : "memory");
    const unsigned int* mine = p.flags[p.rank] + blockIdx.x * AR_MAX_WORLD + peer;
    const long long t0 = clock64();
    for (;;) {
      unsigned int v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (static_cast<int>(v - target) >= 0) break;
      if (clock64() - t0 > 4000000000LL) {
        printf("dmi_b200: all-reduce barrier timed out (rank %d waits for rank %d, cta %d, flag %u < %u)\n", p.rank, peer, blockIdx.x, v, target);
        __trap();
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(AR_THREADS, 8) allreduce_oneshot_kernel(const ArParams p) {
  ar_barrier(p, 2 * p.epoch - 1);
  // four independent 16-byte reductions in flight per thread (the loads are issued back to back, then consumed): the kernel is pure
  // NVLink round-trip latency, so memory-level parallelism is what shortens it
  constexpr int U = 4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  const long long first = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  for (long long i0 = first; i0 < p.n; i0 += U * stride) {
    float4 acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i >= p.n) continue;
      if (p.mc != nullptr) {
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(acc[u].x), "=f"(acc[u].y), "=f"(acc[u].z), "=f"(acc[u].w) : "l"(p.mc + i) : "memory");
      } else {
#pragma unroll
        for (int k = 0; k < AR_MAX_WORLD; ++k) {
          if (k < p.world) {
            const int peer = (p.rank + k) % p.world;     // rank-rotated order spreads the reads over the links
            float4 v;          // volatile: peer lines cached by the previous step's reads must not be served from L1
            asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p.in[peer] + i) : "memory");
            acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i >= p.n) continue;
      *reinterpret_cast<float4*>(p.out + i) = make_float4(acc[u].x * p.scale, acc[u].y * p.scale, acc[u].z * p.scale, acc[u].w * p.scale);
    }
  }
  ar_barrier(p, 2 * p.epoch);
}

}  // namespace
}  // namespace dmi

using namespace dmi;

extern "C" {

int64_t dmi_allreduce_flag_words(void) { return static_cast<int64_t>(AR_MAX_CTAS) * AR_MAX_WORLD; }

int dmi_allreduce_oneshot(const void* const* peer_bufs, void* const* peer_flags, const void* multicast_ptr, int rank, int world, float* out,
                          int64_t n, float scale, uint32_t epoch, void* stream) {
  DMI_REQUIRE(peer_bufs && peer_flags && out && n > 0, "allreduce_oneshot: null argument");
  DMI_REQUIRE(world >= 1 && world <= AR_MAX_WORLD && rank >= 0 && rank < world, "allreduce_oneshot: bad rank %d / world %d (at most %d GPUs)", rank, world, AR_MAX_WORLD);
  DMI_REQUIRE(n % 4 == 0 && epoch >= 1, "allreduce_oneshot: n must be a multiple of 4 floats and epoch counts from 1");
  ArParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < world; ++i) {
    DMI_REQUIRE(peer_bufs[i] && peer_flags[i] && (reinterpret_cast<uintptr_t>(peer_bufs[i]) & 15) == 0, "allreduce_oneshot: peer %d pointer missing / misaligned", i);
    p.in[i] = static_cast<const float*>(peer_bufs[i]);
    p.flags[i] = static_cast<unsigned int*>(peer_flags[i]);
  }
  DMI_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && out != p.in[rank], "allreduce_oneshot: out must be 16-byte aligned and distinct from the input");
  p.mc = static_cast<const float*>(multicast_ptr);
  p.out = out; p.n = n; p.rank = rank; p.world = world; p.epoch = epoch; p.scale = scale;
  long long ctas = (n / 4 + AR_THREADS - 1) / AR_THREADS;
  if (ctas > AR_MAX_CTAS) ctas = AR_MAX_CTAS;          // all CTAs of all ranks must be co-resident: they wait for one another
  allreduce_oneshot_kernel<<<static_cast<unsigned>(ctas), AR_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DMI_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DMI_OK;
}

}  // extern "C"
