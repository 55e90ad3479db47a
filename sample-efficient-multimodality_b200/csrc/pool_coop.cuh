// Support-set pooling of the hypernetwork (k4-k6 and its backward) as ONE cooperative kernel per direction.
//
// The pooling is a chain of tiny dependent steps -- five 768 x 768 GEMVs, two passes over the [S, D] support sequence, rank <= 2
// weight-gradient updates -- each of which needs the whole GPU for a microsecond or two and then a device-wide dependency.  As
// separate launches (8 forward, 12 backward) the chain cost ~55 + ~75 us inside a CUDA graph, almost all of it launch / drain latency
// (profiles/r2_hyper_microstep_launches.txt).  Here every step is a phase of a persistent grid (one 256-thread CTA per SM, cooperative
// launch) and the dependencies are grid barriers.
//
// Memory rule: everything one phase writes and a later phase reads goes through L2 (`__ldcg` loads / plain stores); only tensors that
// are constant for the whole kernel (weights, the support sequence, the keep mask) use the read-only path.
#pragma once
#include "hyper_kernels.cuh"

namespace dmi {

constexpr int PC_THREADS = 256;
constexpr int PC_TSPLIT = 8;         // token slices of the two passes over the support sequence

// Grid barrier for a cooperatively launched grid (all CTAs resident).  `bar` = {count, generation}, zero before the first use and left
// at {0, g} by every barrier, so consecutive launches reuse it without a reset.  One thread per CTA arrives (release) and spins on the
// generation (acquire); lighter than cooperative_groups' grid.sync() (measured: see DESIGN 3.4).
__device__ __forceinline__ void pc_grid_barrier(unsigned int* bar) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int gen;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(bar + 1) : "memory");
    unsigned int prev;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(bar) : "memory");
    if (prev == gridDim.x - 1) {
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(bar), "r"(0u) : "memory");
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 1), "r"(gen + 1) : "memory");
    } else {
      unsigned int g;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(bar + 1) : "memory");
      } while (g == gen);
    }
  }
  __syncthreads();
}

struct PoolCoopFwdParams {
  PoolParams pp;                                   // sequence, keep mask, scalars; P = raw scores, Pout, c, psum
  const float *wq, *bq, *wk, *bk, *wv, *bv;        // [D,D] / [D]
  float *sq, *q, *qt, *qb, *e;                     // stash: s_i = prefix_i + PE_i, q, q~, q.bk, e
  unsigned int* bar;                               // grid barrier state
};

struct PoolCoopBwdParams {
  PoolParams pp;                                   // as in the forward (P unused)
  const float *wq, *wk, *bk, *wv, *bv;
  const float *sq, *q, *e_unused;
  const float* de;                                 // [NQ, D] from the generator backward
  float *dc, *dqt, *dq, *dpsum, *dqb, *dP;         // scratch (dc and dqt zero-initialised by the caller)
  float *dprefix, *dwq, *dbq, *dwk, *dbk, *dwv, *dbv;   // gradients, accumulated (+=)
  unsigned int* bar;                               // grid barrier state
};

// y[i, o] = sum_d W[o, d] x[i, d] + bias[o] * bias_scale[i]      one warp per output row o (x, bias_scale: written earlier in this kernel)
template <int NV>
__device__ __forceinline__ void co_gemv_rows(const float* __restrict__ W, int O, int D, const float* x, const float* __restrict__ bias,
                                             const float* bias_scale, float* y, int gw, int GW, int lane) {
  for (int o = gw; o < O; o += GW) {
    const float* w = W + static_cast<long long>(o) * D;
    float acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.f;
    if ((D & 3) == 0) {
      for (int d = lane * 4; d < D; d += 128) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + d));
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float4 xv = __ldcg(reinterpret_cast<const float4*>(x + i * D + d));
          acc[i] = fmaf(wv.x, xv.x, acc[i]); acc[i] = fmaf(wv.y, xv.y, acc[i]);
          acc[i] = fmaf(wv.z, xv.z, acc[i]); acc[i] = fmaf(wv.w, xv.w, acc[i]);
        }
      }
    } else {
      for (int d = lane; d < D; d += 32) {
        const float wv = __ldg(w + d);
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = fmaf(wv, __ldcg(x + i * D + d), acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float s = warp_sum(acc[i]);
      if (lane == 0) y[i * O + o] = s + (bias ? __ldg(bias + o) * (bias_scale ? __ldcg(bias_scale + i) : 1.0f) : 0.f);
    }
  }
}

// y[i, d] += sum_o W[o, d] x[i, o]      (W^T x; half-CTAs of 128 threads over (128-column block, slice of o); atomics into y)
template <int NV>
__device__ __forceinline__ void co_gemv_cols(const float* __restrict__ W, int O, int D, const float* x, float* y) {
  const int dblocks = (D + 127) / 128;
  const int nvb = gridDim.x * 2;
  int nsplit = nvb / dblocks;
  if (nsplit < 1) nsplit = 1;
  const int rps = (O + nsplit - 1) / nsplit;
  for (int v = blockIdx.x * 2 + (threadIdx.x >> 7); v < dblocks * nsplit; v += nvb) {
    const int d = (v % dblocks) * 128 + (threadIdx.x & 127);
    const int o0 = (v / dblocks) * rps, o1 = min(O, o0 + rps);
    if (d >= D) continue;
    float acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.f;
#pragma unroll 4
    for (int o = o0; o < o1; ++o) {
      const float wv = __ldg(W + static_cast<long long>(o) * D + d);
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i] = fmaf(wv, __ldcg(x + i * O + o), acc[i]);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) atomicAdd(y + i * D + d, acc[i]);
  }
}

// G[o, d] += sum_i a[i, o] b[i, d]      one warp per row o (each row has exactly one owner: plain read-modify-write)
template <int NV>
__device__ __forceinline__ void co_rank_update(float* G, int O, int D, const float* a, const float* b, int gw, int GW, int lane) {
  for (int o = gw; o < O; o += GW) {
    float av[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) av[i] = __ldcg(a + i * O + o);
    float* g = G + static_cast<long long>(o) * D;
    if ((D & 3) == 0) {
      for (int d = lane * 4; d < D; d += 128) {
        float4 acc = *reinterpret_cast<const float4*>(g + d);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float4 bv = __ldcg(reinterpret_cast<const float4*>(b + i * D + d));
          acc.x = fmaf(av[i], bv.x, acc.x); acc.y = fmaf(av[i], bv.y, acc.y);
          acc.z = fmaf(av[i], bv.z, acc.z); acc.w = fmaf(av[i], bv.w, acc.w);
        }
        *reinterpret_cast<float4*>(g + d) = acc;
      }
    } else {
      for (int d = lane; d < D; d += 32) {
        float acc = g[d];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc = fmaf(av[i], __ldcg(b + i * D + d), acc);
        g[d] = acc;
      }
    }
  }
}

// y[d] += sum_i s[i] a[i, d]      (bias gradients; s == nullptr: weights 1)
template <int NV>
__device__ __forceinline__ void co_weighted_rowsum(const float* a, const float* s, int D, float* y) {
  for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < D; d += gridDim.x * blockDim.x) {
    float acc = y[d];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc = fmaf(s ? __ldcg(s + i) : 1.0f, __ldcg(a + i * D + d), acc);
    y[d] = acc;
  }
}

// y[i] = a_i . b for i < NV, by ONE warp (a: written earlier in this kernel, b: constant)
template <int NV>
__device__ __forceinline__ void co_dot_rows(const float* a, const float* __restrict__ b, int D, float* y, int lane) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc = fmaf(__ldcg(a + i * D + d), __ldg(b + d), acc);
    acc = warp_sum(acc);
    if (lane == 0) y[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// forward:  s_i -> q_i = Wq s_i + bq -> q~_i = Wk^T q_i, qb_i = q_i.bk -> scores -> softmax (+ dropout) -> c_i -> e_i = Wv c_i + bv psum_i
// dynamic shared memory: (S + 64 + 256) floats
// ---------------------------------------------------------------------------------------------------------------------------------
template <int NQ>
__global__ void __launch_bounds__(PC_THREADS)
pool_fwd_coop_kernel(const PoolCoopFwdParams a) {
  extern __shared__ float pc_sm[];
  const PoolParams& p = a.pp;
  const int D = p.D, S = p.S;
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * (PC_THREADS / 32) + (threadIdx.x >> 5), GW = gridDim.x * (PC_THREADS / 32);
  const int gtid = blockIdx.x * PC_THREADS + threadIdx.x, GT = gridDim.x * PC_THREADS;

  // ---- phase 1: s_i = prefix_i + PE_i (kept for the backward), q~ = 0 and c = 0 for the atomics of phases 3 and 5 ----
  for (int idx = gtid; idx < NQ * D; idx += GT) {
    const int i = idx / D, d = idx % D;
    a.sq[idx] = __ldg(p.prefix + idx) + (p.pe ? __ldg(p.pe + static_cast<long long>(i) * p.ldpe + d) : 0.f);
    a.qt[idx] = 0.f;
    p.c[idx] = 0.f;
  }
  pc_grid_barrier(a.bar);
  // ---- phase 2: q = Wq s + bq ----
  co_gemv_rows<NQ>(a.wq, D, D, a.sq, a.bq, nullptr, a.q, gw, GW, lane);
  pc_grid_barrier(a.bar);
  // ---- phase 3: q~ = Wk^T q ; qb = q . bk ----
  co_gemv_cols<NQ>(a.wk, D, D, a.q, a.qt);
  if (gw == GW - 1) co_dot_rows<NQ>(a.q, a.bk, D, a.qb, lane);
  pc_grid_barrier(a.bar);
  // ---- phase 4: raw scores, one warp per (query, token) ----
  for (int idx = gw; idx < NQ * S; idx += GW) {
    const int i = idx / S, t = idx % S;
    const float* qt = a.qt + static_cast<long long>(i) * D;
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc = fmaf(pool_token(p, t, d), __ldcg(qt + d), acc);
    acc = warp_sum(acc);
    if (lane == 0) p.P[idx] = (acc + __ldcg(a.qb + i)) * p.inv_sqrt_d;
  }
  pc_grid_barrier(a.bar);
  // ---- phase 5: softmax over the S valid tokens (+ dropout keep mask) and the context.  One CTA per (query, 128-column slice, token
  // slice): every CTA recomputes the (tiny) softmax, then 128 columns x 2 token groups walk the CTA's PC_TSPLIT-th of the tokens and
  // the partial contexts are added atomically (12 CTAs walking all S tokens were the longest phase of the kernel) ----
  {
    float* w = pc_sm;
    float* red = pc_sm + S;
    float* part = red + 64;
    const int dblocks = (D + 127) / 128;
    const int tchunk = (S + PC_TSPLIT - 1) / PC_TSPLIT;
    for (int v = blockIdx.x; v < dblocks * NQ * PC_TSPLIT; v += gridDim.x) {
      const int ts = v % PC_TSPLIT, db = (v / PC_TSPLIT) % dblocks, i = v / (PC_TSPLIT * dblocks);
      const bool first = db == 0 && ts == 0;
      const float* Prow = p.P + static_cast<long long>(i) * S;
      float m = -INFINITY;
      for (int t = threadIdx.x; t < S; t += PC_THREADS) { const float x = __ldcg(Prow + t); w[t] = x; m = fmaxf(m, x); }
      m = block_max(m, red);
      float s = 0.f;
      for (int t = threadIdx.x; t < S; t += PC_THREADS) { const float ev = expf(w[t] - m); w[t] = ev; s += ev; }
      s = block_sum(s, red);
      const float inv = 1.0f / s;
      float ps = 0.f;
      for (int t = threadIdx.x; t < S; t += PC_THREADS) {
        const float pr = w[t] * inv;
        const float pt = p.keep ? pr * __ldg(p.keep + static_cast<long long>(i) * S + t) * p.keep_scale : pr;
        w[t] = pt;
        ps += pt;
        if (first) p.Pout[static_cast<long long>(i) * S + t] = pr;
      }
      ps = block_sum(ps, red);
      if (first && threadIdx.x == 0) p.psum[i] = ps;
      __syncthreads();
      const int dx = threadIdx.x & 127, tg = threadIdx.x >> 7;
      const int d = db * 128 + dx;
      const int t1 = min(S, (ts + 1) * tchunk);
      float acc = 0.f;
      if (d < D) {
#pragma unroll 4
        for (int t = ts * tchunk + tg; t < t1; t += 2) acc = fmaf(w[t], pool_token(p, t, d), acc);
      }
      part[tg * 128 + dx] = acc;
      __syncthreads();
      if (tg == 0 && d < D) atomicAdd(p.c + static_cast<long long>(i) * D + d, part[dx] + part[128 + dx]);
      __syncthreads();
    }
  }
  pc_grid_barrier(a.bar);
  // ---- phase 6: e_i = Wv c_i + bv * psum_i ----
  co_gemv_rows<NQ>(a.wv, D, D, p.c, a.bv, p.psum, a.e, gw, GW, lane);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// backward, given de_i from the generator backward:
//   dbv += sum_i psum_i de_i ; dWv += sum_i de_i (x) c_i ; dc_i = Wv^T de_i ; dpsum_i = bv . de_i
//   dP~[t] = dc_i . s_t + dpsum_i ; dP = dP~ keep scale ; dsig[t] = P[t] (dP[t] - sum_t' dP[t'] P[t'])
//   dq~_i = sum_t dsig[t] s_t / sqrt(D) ; dqb_i = sum_t dsig[t] / sqrt(D) ; dprefix_t += P~[i,t] dc_i + dsig[t] q~_i / sqrt(D)   (t < NQ)
//   dWk += sum_i q_i (x) dq~_i ; dbk += sum_i dqb_i q_i ; dq_i = Wk dq~_i + dqb_i bk
//   dWq += sum_i dq_i (x) s_i ; dbq += sum_i dq_i ; dprefix_i += Wq^T dq_i
// dynamic shared memory: (S + 64 + 256) floats
// ---------------------------------------------------------------------------------------------------------------------------------
template <int NQ>
__global__ void __launch_bounds__(PC_THREADS)
pool_bwd_coop_kernel(const PoolCoopBwdParams b) {
  extern __shared__ float pc_sm[];
  const PoolParams& p = b.pp;
  const int D = p.D, S = p.S;
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * (PC_THREADS / 32) + (threadIdx.x >> 5), GW = gridDim.x * (PC_THREADS / 32);

  // ---- phase 1: value path ----
  co_weighted_rowsum<NQ>(b.de, p.psum, D, b.dbv);
  co_rank_update<NQ>(b.dwv, D, D, b.de, p.c, gw, GW, lane);
  co_gemv_cols<NQ>(b.wv, D, D, b.de, b.dc);
  if (gw == GW - 1) co_dot_rows<NQ>(b.de, b.bv, D, b.dpsum, lane);
  pc_grid_barrier(b.bar);
  // ---- phase 2: dP[i, t], one warp per (query, token) ----
  for (int idx = gw; idx < NQ * S; idx += GW) {
    const int i = idx / S, t = idx % S;
    const float* dc = b.dc + static_cast<long long>(i) * D;
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc = fmaf(pool_token(p, t, d), __ldcg(dc + d), acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float ks = p.keep ? __ldg(p.keep + idx) * p.keep_scale : 1.0f;
      b.dP[idx] = (acc + __ldcg(b.dpsum + i)) * ks;
    }
  }
  pc_grid_barrier(b.bar);
  // ---- phase 3: softmax backward and the D-sliced reductions, one CTA per (query, 128-column slice, token slice); dq~ is accumulated
  // atomically (zero-initialised scratch), the token-independent terms are added by the CTA of token slice 0 ----
  {
    float* dsig = pc_sm;
    float* red = pc_sm + S;
    float* part = red + 64;
    const int dblocks = (D + 127) / 128;
    const int tchunk = (S + PC_TSPLIT - 1) / PC_TSPLIT;
    for (int v = blockIdx.x; v < dblocks * NQ * PC_TSPLIT; v += gridDim.x) {
      const int ts = v % PC_TSPLIT, db = (v / PC_TSPLIT) % dblocks, i = v / (PC_TSPLIT * dblocks);
      const float* Prow = p.Pout + static_cast<long long>(i) * S;
      const float* dProw = b.dP + static_cast<long long>(i) * S;
      float dot = 0.f;
      for (int t = threadIdx.x; t < S; t += PC_THREADS) dot += __ldcg(dProw + t) * __ldg(Prow + t);
      dot = block_sum(dot, red);
      float sumsig = 0.f;
      for (int t = threadIdx.x; t < S; t += PC_THREADS) {
        const float x = __ldg(Prow + t) * (__ldcg(dProw + t) - dot);
        dsig[t] = x;
        sumsig += x;
      }
      sumsig = block_sum(sumsig, red);
      if (db == 0 && ts == 0 && threadIdx.x == 0) b.dqb[i] = sumsig * p.inv_sqrt_d;
      __syncthreads();
      const int dx = threadIdx.x & 127, tg = threadIdx.x >> 7;
      const int d = db * 128 + dx;
      const int t1 = min(S, (ts + 1) * tchunk);
      float acc = 0.f;
      if (d < D) {
#pragma unroll 4
        for (int t = ts * tchunk + tg; t < t1; t += 2) acc = fmaf(dsig[t], pool_token(p, t, d), acc);
      }
      part[tg * 128 + dx] = acc;
      __syncthreads();
      if (tg == 0 && d < D) {
        atomicAdd(b.dqt + static_cast<long long>(i) * D + d, (part[dx] + part[128 + dx]) * p.inv_sqrt_d);
        if (ts == 0) {
          const float dcd = __ldcg(b.dc + static_cast<long long>(i) * D + d);
          const float qtd = __ldg(p.qt + static_cast<long long>(i) * D + d);
          for (int t = 0; t < NQ; ++t) {
            const float ks = p.keep ? __ldg(p.keep + static_cast<long long>(i) * S + t) * p.keep_scale : 1.0f;
            atomicAdd(b.dprefix + static_cast<long long>(t) * D + d, __ldg(Prow + t) * ks * dcd + dsig[t] * qtd * p.inv_sqrt_d);
          }
        }
      }
      __syncthreads();
    }
  }
  pc_grid_barrier(b.bar);
  // ---- phase 4: key path ----
  co_rank_update<NQ>(b.dwk, D, D, b.q, b.dqt, gw, GW, lane);
  co_weighted_rowsum<NQ>(b.q, b.dqb, D, b.dbk);
  co_gemv_rows<NQ>(b.wk, D, D, b.dqt, b.bk, b.dqb, b.dq, gw, GW, lane);
  pc_grid_barrier(b.bar);
  // ---- phase 5: query path ----
  co_rank_update<NQ>(b.dwq, D, D, b.dq, b.sq, gw, GW, lane);
  co_weighted_rowsum<NQ>(b.dq, nullptr, D, b.dbq);
  co_gemv_cols<NQ>(b.wq, D, D, b.dq, b.dprefix);
}

}  // namespace dmi
