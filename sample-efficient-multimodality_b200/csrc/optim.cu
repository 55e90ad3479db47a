// C ABI, part 3: fused gradient-norm clipping + AdamW over a list of fp32 tensors.
//
// Replaces the optimizer step of the reference trainers -- torch.nn.utils.clip_grad_norm_ followed by optim.AdamW.step()
// (dmi/train_hypernet.py:148-149 with the optimizer built at :526-532, dmi/train_projector.py, dmi/train_lora.py) -- which
// in PyTorch is a sequence of multi-tensor passes (norm, scale, mul, lerp, addcmul, sqrt, div, addcdiv): ~60 bytes of HBM
// traffic per parameter.  Here: one read-only pass for the squared norm (4 B / parameter) and ONE pass that applies the clip
// factor, the decoupled weight decay and the Adam update: 16 B read + 12 B written per parameter, the HBM minimum for
// fp32 master weights with fp32 moments.  The clip factor is computed on the device from the accumulated squared norm, so
// there is no host synchronisation between backward and the update.
//
// Arithmetic follows torch.optim.AdamW (amsgrad=False, maximize=False) operation by operation:
//   p *= 1 - lr*wd;  m = m + (g - m)(1 - b1);  v = b2 v + (1 - b2) g g;  p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// and clip_grad_norm_:  coef = min(1, max_norm / (||g||_2 + 1e-6)),  g *= coef.
#include <math.h>
#include <string.h>

#include "../../include/dmi_b200.h"
#include "common.cuh"

namespace dmi {

void count_launch();
int num_sms();

constexpr int OPT_MAX_TENSORS = 24;      // descriptors per launch (kernel parameter space); longer lists take several launches
constexpr int OPT_CHUNK = 8192;          // elements per CTA work item
constexpr int OPT_THREADS = 256;

struct OptTensor {
  float* p;
  float* g;
  float* m;
  float* v;
  long long n;
  long long first_chunk;                 // index of this tensor's first chunk in the launch
};
struct OptBatch {
  OptTensor t[OPT_MAX_TENSORS];
  int count;
  long long total_chunks;
};

__device__ __forceinline__ int find_tensor(const OptBatch& b, long long chunk) {
  int i = 0;
#pragma unroll 1
  while (i + 1 < b.count && b.t[i + 1].first_chunk <= chunk) ++i;
  return i;
}

// sqnorm[0] += sum over all tensors of g^2, DETERMINISTICALLY: every thread accumulates its elements over all the chunks of its CTA
// (fixed grid -> fixed assignment), one block reduction per CTA, the per-CTA partials go to a fixed slot of a library-owned buffer and
// the LAST CTA to finish (threadfence + ticket) adds them up in index order.  Round 1 issued one fp32 atomic per 8192-element chunk
// (~21 k atomics in arrival order for the hypernet): the clip factor then depended on the scheduling, and the per-chunk
// barriers held the read-only stream at 0.66 of the HBM peak.  One call at a time per device (the library's single-stream contract).
constexpr int OPT_MAX_GRID = 4096;
__device__ float g_sq_partials[OPT_MAX_GRID];
__device__ unsigned int g_sq_ticket = 0;

__global__ void __launch_bounds__(OPT_THREADS)
grad_sqnorm_kernel(const __grid_constant__ OptBatch b, float* __restrict__ sqnorm) {
  __shared__ float red[OPT_THREADS / 32];
  __shared__ int last;
  float acc = 0.f;
  for (long long chunk = blockIdx.x; chunk < b.total_chunks; chunk += gridDim.x) {
    const int ti = find_tensor(b, chunk);
    const OptTensor& t = b.t[ti];
    const long long base = (chunk - t.first_chunk) * OPT_CHUNK;
    const long long end = (base + OPT_CHUNK < t.n) ? base + OPT_CHUNK : t.n;
    const bool vec = (reinterpret_cast<uintptr_t>(t.g) & 15) == 0;
    if (vec) {
      const long long e4 = base + ((end - base) & ~3LL);
      // a full chunk is 8 float4 per thread: all eight loads are issued before the first is consumed
      float4 g[OPT_CHUNK / (OPT_THREADS * 4)];
#pragma unroll
      for (int u = 0; u < OPT_CHUNK / (OPT_THREADS * 4); ++u) {
        const long long i = base + (u * OPT_THREADS + threadIdx.x) * 4LL;
        g[u] = (i < e4) ? __ldcs(reinterpret_cast<const float4*>(t.g + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < OPT_CHUNK / (OPT_THREADS * 4); ++u) {
        acc = fmaf(g[u].x, g[u].x, acc); acc = fmaf(g[u].y, g[u].y, acc); acc = fmaf(g[u].z, g[u].z, acc); acc = fmaf(g[u].w, g[u].w, acc);
      }
      for (long long i = e4 + threadIdx.x; i < end; i += OPT_THREADS) acc = fmaf(t.g[i], t.g[i], acc);
    } else {
      for (long long i = base + threadIdx.x; i < end; i += OPT_THREADS) acc = fmaf(t.g[i], t.g[i], acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < OPT_THREADS / 32; ++w) s += red[w];
    g_sq_partials[blockIdx.x] = s;
    __threadfence();
    last = atomicInc(&g_sq_ticket, gridDim.x - 1) == gridDim.x - 1;      // wraps to 0 for the next launch
  }
  __syncthreads();
  if (last) {
    __threadfence();
    // fixed-order sum of the partials: thread t takes slots t, t + 256, ...; then a fixed tree
    float s = 0.f;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += OPT_THREADS) s += *(reinterpret_cast<volatile float*>(&g_sq_partials[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int w = 0; w < OPT_THREADS / 32; ++w) tot += red[w];
      *sqnorm += tot;
    }
  }
}

struct AdamScalars {
  float decay;            // 1 - lr * weight_decay
  float one_minus_b1;
  float beta2;
  float one_minus_b2;
  float step_size;        // lr / (1 - beta1^t)
  float bc2_sqrt;         // sqrt(1 - beta2^t)
  float eps;
  float max_norm;         // <= 0: no clipping
  int write_grads;        // 1: store the clipped gradient back (clip_grad_norm_ semantics), 0: leave .grad untouched
};

__device__ __forceinline__ void adam_update(float& p, float& g, float& m, float& v, const AdamScalars& s, float coef) {
  g *= coef;
  p *= s.decay;
  m = fmaf(g - m, s.one_minus_b1, m);                        // lerp_(grad, 1 - beta1)
  v = fmaf(s.one_minus_b2 * g, g, v * s.beta2);              // mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
  const float denom = sqrtf(v) / s.bc2_sqrt + s.eps;
  p = fmaf(-s.step_size, m / denom, p);                      // addcdiv_(exp_avg, denom, value = -step_size)
}

__global__ void __launch_bounds__(OPT_THREADS)
adamw_step_kernel(const __grid_constant__ OptBatch b, const AdamScalars s, const float* __restrict__ sqnorm) {
  float coef = 1.0f;
  if (s.max_norm > 0.f && sqnorm != nullptr) {
    const float total = sqrtf(__ldg(sqnorm));
    coef = fminf(s.max_norm / (total + 1e-6f), 1.0f);
  }
  for (long long chunk = blockIdx.x; chunk < b.total_chunks; chunk += gridDim.x) {
    const int ti = find_tensor(b, chunk);
    const OptTensor& t = b.t[ti];
    const long long base = (chunk - t.first_chunk) * OPT_CHUNK;
    const long long end = (base + OPT_CHUNK < t.n) ? base + OPT_CHUNK : t.n;
    const bool vec = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) | reinterpret_cast<uintptr_t>(t.m) |
                       reinterpret_cast<uintptr_t>(t.v)) & 15) == 0;
    long long scalar_from = base;
    if (vec) {
      const long long e4 = base + ((end - base) & ~3LL);
      for (long long i = base + threadIdx.x * 4LL; i < e4; i += OPT_THREADS * 4LL) {
        float4 p = *reinterpret_cast<float4*>(t.p + i);
        float4 g = *reinterpret_cast<const float4*>(t.g + i);
        float4 m = *reinterpret_cast<float4*>(t.m + i);
        float4 v = *reinterpret_cast<float4*>(t.v + i);
        adam_update(p.x, g.x, m.x, v.x, s, coef);
        adam_update(p.y, g.y, m.y, v.y, s, coef);
        adam_update(p.z, g.z, m.z, v.z, s, coef);
        adam_update(p.w, g.w, m.w, v.w, s, coef);
        *reinterpret_cast<float4*>(t.p + i) = p;
        *reinterpret_cast<float4*>(t.m + i) = m;
        *reinterpret_cast<float4*>(t.v + i) = v;
        if (s.write_grads) *reinterpret_cast<float4*>(t.g + i) = g;
      }
      scalar_from = e4;
    }
    for (long long i = scalar_from + threadIdx.x; i < end; i += OPT_THREADS) {
      float p = t.p[i], g = t.g[i], m = t.m[i], v = t.v[i];
      adam_update(p, g, m, v, s, coef);
      t.p[i] = p; t.m[i] = m; t.v[i] = v;
      if (s.write_grads) t.g[i] = g;
    }
  }
}

// g *= min(1, max_norm / (sqrt(sqnorm) + 1e-6))   (clip_grad_norm_ alone, for callers that keep their own optimizer)
__global__ void __launch_bounds__(OPT_THREADS)
grad_scale_kernel(const __grid_constant__ OptBatch b, float max_norm, const float* __restrict__ sqnorm) {
  const float coef = fminf(max_norm / (sqrtf(__ldg(sqnorm)) + 1e-6f), 1.0f);
  for (long long chunk = blockIdx.x; chunk < b.total_chunks; chunk += gridDim.x) {
    const int ti = find_tensor(b, chunk);
    const OptTensor& t = b.t[ti];
    const long long base = (chunk - t.first_chunk) * OPT_CHUNK;
    const long long end = (base + OPT_CHUNK < t.n) ? base + OPT_CHUNK : t.n;
    for (long long i = base + threadIdx.x; i < end; i += OPT_THREADS) t.g[i] *= coef;
  }
}

static int make_batches(const dmi_opt_tensor* ts, int count, bool need_pmv, OptBatch* out, int max_batches, int* n_batches) {
  int nb = 0;
  OptBatch cur;
  memset(&cur, 0, sizeof(cur));
  for (int i = 0; i < count; ++i) {
    const dmi_opt_tensor& t = ts[i];
    DMI_REQUIRE(t.n >= 0 && (t.n == 0 || (t.g != nullptr && (!need_pmv || (t.p != nullptr && t.m != nullptr && t.v != nullptr)))),
                "optimizer tensor %d: null pointer", i);
    if (t.n == 0) continue;
    if (cur.count == OPT_MAX_TENSORS) {
      DMI_REQUIRE(nb < max_batches, "too many optimizer tensors");
      out[nb++] = cur;
      memset(&cur, 0, sizeof(cur));
    }
    OptTensor& o = cur.t[cur.count++];
    o.p = t.p; o.g = t.g; o.m = t.m; o.v = t.v; o.n = t.n; o.first_chunk = cur.total_chunks;
    cur.total_chunks += (t.n + OPT_CHUNK - 1) / OPT_CHUNK;
  }
  if (cur.count > 0) {
    DMI_REQUIRE(nb < max_batches, "too many optimizer tensors");
    out[nb++] = cur;
  }
  *n_batches = nb;
  return DMI_OK;
}

static unsigned opt_grid(long long chunks) {
  const long long cap = 8LL * num_sms();
  return static_cast<unsigned>(chunks < cap ? (chunks > 0 ? chunks : 1) : cap);
}

constexpr int OPT_MAX_BATCHES = 64;

}  // namespace dmi

using namespace dmi;

extern "C" {

int dmi_grad_sqnorm(const dmi_opt_tensor* tensors, int count, float* sqnorm_accum, void* stream) {
  DMI_REQUIRE(tensors != nullptr && count >= 0 && sqnorm_accum != nullptr, "grad_sqnorm: bad arguments");
  static thread_local OptBatch batches[OPT_MAX_BATCHES];
  int nb = 0;
  int rc = make_batches(tensors, count, false, batches, OPT_MAX_BATCHES, &nb);
  if (rc != DMI_OK) return rc;
  for (int i = 0; i < nb; ++i) {
    grad_sqnorm_kernel<<<opt_grid(batches[i].total_chunks), OPT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(batches[i], sqnorm_accum);
    DMI_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  return DMI_OK;
}

int dmi_grad_clip(const dmi_opt_tensor* tensors, int count, float max_norm, const float* sqnorm, void* stream) {
  DMI_REQUIRE(tensors != nullptr && count >= 0 && sqnorm != nullptr && max_norm > 0.f, "grad_clip: bad arguments");
  static thread_local OptBatch batches[OPT_MAX_BATCHES];
  int nb = 0;
  int rc = make_batches(tensors, count, false, batches, OPT_MAX_BATCHES, &nb);
  if (rc != DMI_OK) return rc;
  for (int i = 0; i < nb; ++i) {
    grad_scale_kernel<<<opt_grid(batches[i].total_chunks), OPT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(batches[i], max_norm, sqnorm);
    DMI_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  return DMI_OK;
}

int dmi_adamw_step(const dmi_opt_tensor* tensors, int count, double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                   float max_grad_norm, const float* sqnorm, int write_clipped_grads, void* stream) {
  DMI_REQUIRE(tensors != nullptr && count >= 0 && step >= 1, "adamw_step: bad arguments (step counts from 1)");
  DMI_REQUIRE(max_grad_norm <= 0.f || sqnorm != nullptr, "adamw_step: clipping needs the accumulated squared norm");
  static thread_local OptBatch batches[OPT_MAX_BATCHES];
  int nb = 0;
  int rc = make_batches(tensors, count, true, batches, OPT_MAX_BATCHES, &nb);
  if (rc != DMI_OK) return rc;
  AdamScalars s;
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
  s.decay = static_cast<float>(1.0 - lr * weight_decay);
  s.one_minus_b1 = static_cast<float>(1.0 - beta1);
  s.beta2 = static_cast<float>(beta2);
  s.one_minus_b2 = static_cast<float>(1.0 - beta2);
  s.step_size = static_cast<float>(lr / bc1);
  s.bc2_sqrt = static_cast<float>(sqrt(bc2));
  s.eps = static_cast<float>(eps);
  s.max_norm = max_grad_norm;
  s.write_grads = write_clipped_grads;
  for (int i = 0; i < nb; ++i) {
    adamw_step_kernel<<<opt_grid(batches[i].total_chunks), OPT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(batches[i], s, sqnorm);
    DMI_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  return DMI_OK;
}

}  // extern "C"
