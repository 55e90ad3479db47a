// CTA-pair variant of the tcgen05 GEMM:  C[M,N] = A[M,K] * B[N,K]^T, bf16, K-major operands, 256 x 256 output tile per
// 2-CTA cluster, `tcgen05.mma.cta_group::2` (UMMA M = 256).
//
// Why: the 1-CTA kernel is bound by shared-memory bandwidth (128 B/cycle/SM): per 128-cycle 128x256x16 MMA the tensor core reads
// 4 KB of A + 8 KB of B while TMA writes the next 12 KB -- 192 B/cycle, i.e. at most 67 % of the MMA rate.
// In a CTA pair each SM holds its own 128 rows of A and only HALF of the B tile and the hardware feeds both tensor cores from the
// two halves: per SM 12 KB are still read (its half of B is also served to the peer) but only 8 KB are written -- 160 B/cycle,
// 80 % -- and each SM loads 32 KB instead of 48 KB per K block from L2.  Measured: 5-9 % faster than the 1-CTA kernel for the
// store / GELU epilogues on the step's shapes (profiles/r1_gemm_headroom.txt), selected automatically for them (api.cu, use_pair).
//
// Roles per CTA (320 threads): warp 0 TMA producer (own A rows + own half of B; completion is signalled on the LEADER's
// full barrier), warp 1 of the LEADER issues the MMAs for both CTAs and multicasts the commits (smem-slot release and
// accumulator-ready) to both CTAs, warps 2..9 drain the CTA's own half (128 rows) of the accumulator from its own TMEM.
#pragma once
#include "gemm_tc.cuh"

namespace dmi {

constexpr int G2_BN = 256;
constexpr int G2_STAGE_BYTES = (GEMM_BM + G2_BN / 2) * 128;      // 128 rows of A + 128 rows of B, 64 bf16 each = 32 KB
constexpr int G2_STAGES = 5;
constexpr int G2_EPI_BYTES = 8192;                                // per epilogue warp: two 4 KB TMA boxes ([32 rows x 128 bytes], SWIZZLE_128B)
constexpr int G2_OFF_EPI = G2_STAGES * G2_STAGE_BYTES;            // 1024-byte aligned (stage size is a multiple of 1024)
constexpr int G2_OFF_BAR = G2_OFF_EPI + EPI_WARPS * G2_EPI_BYTES;
constexpr int G2_SMEM_BYTES = G2_OFF_BAR + 256 + 1024 /*alignment slack*/;
static_assert(G2_SMEM_BYTES <= 227 * 1024, "pair GEMM: shared memory budget");

__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  // Default (CTA-scope release) semantics, as CUTLASS's ClusterBarrier::arrive(cta_id) issues it.  The explicit .release.cluster form
  // compiles to MEMBAR.ALL.GPU + ERRBAR, i.e. every epilogue warp waited for all of its global stores to drain before it could hand
  // the accumulator back (ncu source view: 8 % of all samples on those two instructions).  The TMEM reads that must be ordered
  // before the MMA's overwrite are ordered by tcgen05.wait::ld + tcgen05.fence::before_thread_sync, not by this arrive.
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of the pair; the mbarrier operand is a shared::cluster address (the leader's barrier).
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tm, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}


// ---- TMA store epilogue -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_saddr(uint32_t smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// same store with an L2 evict-first policy: the outputs are written once and must not push the operand tiles (re-read by the other N
// tiles of the wave) out of L2
__device__ __forceinline__ void tma_store_2d_ef(const CUtensorMap* tm, uint32_t smem_src, int c0, int c1, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_src), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// One 64-column slab of this warp's 32 accumulator rows: tcgen05.ld (thread = row, 2 x 32 columns) -> fused math in registers ->
// the row's 128 bytes (bf16) or 2 x 128 bytes (fp32) into SWIZZLE_128B staging boxes -> one elected lane issues the TMA store.
// No transposing re-read of the staging buffer, no per-thread global stores or address arithmetic: the old epilogue made the
// K = 800 GELU GEMM epilogue-bound (106 us against 65 us with the epilogue switched off, profiles/r2_headroom1.txt).
// EPI_GELU_BWD streams the stashed pre-activation slab in through TMA as well (box 1 of the staging pair, prefetched one slab ahead).
template <int MODE>
__device__ __forceinline__ void pair_epilogue_slab(const GemmParams& p, const CUtensorMap* tmO0, const CUtensorMap* tmO1, uint32_t buf, uint32_t t_addr,
                                                   int row0, int col0, int lane, bool release_tmem, uint32_t tempty_remote, uint64_t* aux_bar,
                                                   uint32_t& aux_phase, bool have_next, int next_row0, int next_col0) {
  uint32_t ra[32], rb[32];
  tmem_ld_32x32(t_addr, ra);
  tmem_ld_32x32(t_addr + 32, rb);
  uint4 pre[8];
  if (MODE == EPI_GELU_BWD) {
    mbar_wait(aux_bar, aux_phase);
    aux_phase ^= 1;
#pragma unroll
    for (int j = 0; j < 8; ++j) pre[j] = lds128u(buf + 4096 + lane * 128 + ((j ^ (lane & 7)) << 4));
    __syncwarp();                                   // every lane holds its row: the box may be refilled
    if (have_next && elect_one()) {
      mbar_arrive_expect_tx(aux_bar, 4096);
      tma_load_2d_saddr(buf + 4096, tmO1, aux_bar, next_col0, next_row0);
    }
    __syncwarp();
  }
  tmem_ld_wait();
  if (release_tmem) {                               // last read of this accumulator buffer: hand it back before the math and the stores
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_remote(tempty_remote);
  }
  const bool f32out = MODE == EPI_STORE && p.out0_f32;
  const uint32_t rowb = buf + lane * 128;
  const int sw = lane & 7;
  if (f32out) {
    // fp32 output: alpha * acc + bias in place, two [32 x 32 fp32] boxes
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * j));
      uint32_t* v = j < 8 ? ra + 4 * j : rb + 4 * (j - 8);
      v[0] = __float_as_uint(fmaf(__uint_as_float(v[0]), p.alpha, b4.x)); v[1] = __float_as_uint(fmaf(__uint_as_float(v[1]), p.alpha, b4.y));
      v[2] = __float_as_uint(fmaf(__uint_as_float(v[2]), p.alpha, b4.z)); v[3] = __float_as_uint(fmaf(__uint_as_float(v[3]), p.alpha, b4.w));
    }
    if (elect_one()) bulk_wait_read0();      // the staging boxes are free once the previous slab's TMA stores have READ them
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sts128u(rowb + ((j ^ sw) << 4), ra[4 * j], ra[4 * j + 1], ra[4 * j + 2], ra[4 * j + 3]);
      sts128u(rowb + 4096 + ((j ^ sw) << 4), rb[4 * j], rb[4 * j + 1], rb[4 * j + 2], rb[4 * j + 3]);
    }
  } else {
    uint32_t o0[32], o1[32];                        // packed bf16 pairs (o1: pre-activation of EPI_GELU)
    // dropout keep mask of the projector-training path (train_projector.py: Dropout(0.1) active): this thread's 64 mask bytes of its row
    const bool use_keep = MODE != EPI_STORE && p.keep != nullptr;
    const bool keep_row_ok = row0 + lane < p.M;
    const uint4* keep_row = use_keep ? reinterpret_cast<const uint4*>(p.keep + static_cast<long long>(row0 + lane) * p.ld_keep + col0) : nullptr;
    uint4 g4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE != EPI_GELU_BWD && p.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * j));
      const uint32_t* src = j < 8 ? ra + 4 * j : rb + 4 * (j - 8);
      float v0 = fmaf(__uint_as_float(src[0]), p.alpha, b4.x), v1 = fmaf(__uint_as_float(src[1]), p.alpha, b4.y);
      float v2 = fmaf(__uint_as_float(src[2]), p.alpha, b4.z), v3 = fmaf(__uint_as_float(src[3]), p.alpha, b4.w);
      if (MODE == EPI_GELU) {
        o1[2 * j] = pack_bf16x2(v0, v1); o1[2 * j + 1] = pack_bf16x2(v2, v3);
        v0 = gelu_tanh(v0); v1 = gelu_tanh(v1); v2 = gelu_tanh(v2); v3 = gelu_tanh(v3);
      } else if (MODE == EPI_GELU_BWD) {
        const uint32_t w0 = (j & 1) ? pre[j >> 1].z : pre[j >> 1].x, w1 = (j & 1) ? pre[j >> 1].w : pre[j >> 1].y;
        const float2 p01 = unpack_bf16x2(w0), p23 = unpack_bf16x2(w1);
        v0 *= gelu_tanh_grad(p01.x); v1 *= gelu_tanh_grad(p01.y); v2 *= gelu_tanh_grad(p23.x); v3 *= gelu_tanh_grad(p23.y);
      }
      if (use_keep) {                               // 4 mask bytes = word (j & 3) of the 16-byte group j >> 2
        if ((j & 3) == 0 && keep_row_ok) g4 = __ldg(keep_row + (j >> 2));
        const uint32_t kw = (j & 3) == 0 ? g4.x : ((j & 3) == 1 ? g4.y : ((j & 3) == 2 ? g4.z : g4.w));
        v0 *= (kw & 0xFFu) ? p.keep_scale : 0.f; v1 *= (kw & 0xFF00u) ? p.keep_scale : 0.f;
        v2 *= (kw & 0xFF0000u) ? p.keep_scale : 0.f; v3 *= (kw & 0xFF000000u) ? p.keep_scale : 0.f;
      }
      o0[2 * j] = pack_bf16x2(v0, v1); o0[2 * j + 1] = pack_bf16x2(v2, v3);
    }
    if (elect_one()) bulk_wait_read0();
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sts128u(rowb + ((j ^ sw) << 4), o0[4 * j], o0[4 * j + 1], o0[4 * j + 2], o0[4 * j + 3]);
      if (MODE == EPI_GELU) sts128u(rowb + 4096 + ((j ^ sw) << 4), o1[4 * j], o1[4 * j + 1], o1[4 * j + 2], o1[4 * j + 3]);
    }
  }
  fence_proxy_async_smem();                         // generic-proxy writes -> visible to the TMA engine
  __syncwarp();
  if (elect_one() && !(p.debug & 16)) {             // debug 16 (measurement only): everything but the TMA stores
    if (p.debug & 64) {                             // debug 64 (measurement only): stores with an L2 evict-first hint
      const uint64_t pol = l2_policy_evict_first();
      tma_store_2d_ef(tmO0, buf, col0, row0, pol);
      if (f32out) tma_store_2d_ef(tmO0, buf + 4096, col0 + 32, row0, pol);
      if (MODE == EPI_GELU) tma_store_2d_ef(tmO1, buf + 4096, col0, row0, pol);
    } else {
      tma_store_2d(tmO0, buf, col0, row0);
      if (f32out) tma_store_2d(tmO0, buf + 4096, col0 + 32, row0);
      if (MODE == EPI_GELU) tma_store_2d(tmO1, buf + 4096, col0, row0);
    }
    bulk_commit();
  }
  __syncwarp();
}

// tmO0: output 0 ([32 x 128 B] boxes); tmO1: EPI_GELU pre-activation output / EPI_GELU_BWD stashed pre-activation input (else = tmO0)
template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO0,
                 const __grid_constant__ CUtensorMap tmO1, const GemmParams p) {
  constexpr int BN = G2_BN, STAGES = G2_STAGES, BK = 64, UK = 16;
  constexpr int A_BYTES = GEMM_BM * 128;
  constexpr uint32_t IDESC = make_idesc(2 * GEMM_BM, BN, 1);          // M = 256 across the pair, N = 256, bf16 -> fp32

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G2_OFF_BAR);   // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;       // per CTA
  uint64_t* tfull_bar = empty_bar + STAGES;      // [2] per CTA
  uint64_t* tempty_bar = tfull_bar + 2;          // [2] used in the leader only (16 arrivals: 8 epilogue warps x 2 CTAs)
  uint64_t* aux_bar = tempty_bar + 2;            // [EPI_WARPS] EPI_GELU_BWD: pre-activation slab landed in the warp's staging box 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int n_tiles_n = (p.N + BN - 1) / BN;
  const int n_tiles_m2 = (p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const int n_tiles = n_tiles_m2 * n_tiles_n;
  const int nkb = (p.K + BK - 1) / BK;
  const int ksteps_last = ((p.K - (nkb - 1) * BK) + UK - 1) / UK;
  const int tile0 = blockIdx.x >> 1, tile_stride = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO0);
    tma_prefetch_desc(&tmO1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 2 * EPI_WARPS);
    }
    for (int w = 0; w < EPI_WARPS; ++w) mbar_init(&aux_bar[w], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  // programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) overlapped the predecessor's tail;
  // from here on global memory is read
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    // whole warp loops, one elected lane issues (uniform-datapath descriptors, see gemm_tc.cuh)
    if (!(p.debug & 2)) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < n_tiles; tile += tile_stride) {
        const int m0 = ((tile / n_tiles_n) * 2 + cta_rank) * GEMM_BM;
        const int n0 = (tile % n_tiles_n) * BN + cta_rank * (BN / 2);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);                       // my slot was released by the leader's commit
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);   // bytes of BOTH CTAs land on this barrier
            const uint32_t bar = mapa_rank(smem_u32(&full_bar[stage]), 0);
            uint8_t* sa = smem + stage * G2_STAGE_BYTES;
            tma_load_2d_pair(sa, &tmA, bar, kb * BK, m0);
            tma_load_2d_pair(sa + A_BYTES, &tmB, bar, kb * BK, n0);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {          // warp-uniform loop, one elected lane issues (see gemm_tc.cuh)
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = tile0; tile < n_tiles; tile += tile_stride, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          if (!(p.debug & 2)) mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + stage * G2_STAGE_BYTES);
            const uint64_t adesc = make_kmajor_sw128_desc(sa);
            const uint64_t bdesc = make_kmajor_sw128_desc(sa + A_BYTES);
            if (kb != nkb - 1 || ksteps_last == BK / UK) {
#pragma unroll
              for (int k = 0; k < BK / UK; ++k) umma_f16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb | k) != 0);
            } else {
              for (int k = 0; k < ksteps_last; ++k) umma_f16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb | k) != 0);
            }
            umma_commit_pair(&empty_bar[stage], 0x3);                      // both CTAs may refill this slot
            if (kb == nkb - 1) umma_commit_pair(&tfull_bar[acc], 0x3);    // both CTAs' epilogues may drain
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9, both CTAs, own 128 rows) =====================
    // warp -> TMEM lane quarter (warp & 3) x column half; per tile two 64-column slabs, each leaving through TMA stores
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t buf = smem_u32(smem + G2_OFF_EPI + (warp - 2) * G2_EPI_BYTES);
    uint64_t* my_aux = &aux_bar[warp - 2];
    uint32_t aux_phase = 0;
    const uint32_t tempty0 = mapa_rank(smem_u32(&tempty_bar[0]), 0);
    auto rows_of = [&](int tile) { return ((tile / n_tiles_n) * 2 + static_cast<int>(cta_rank)) * GEMM_BM + quarter * 32; };
    auto cols_of = [&](int tile) { return (tile % n_tiles_n) * BN + half * (BN / 2); };
    if (MODE == EPI_GELU_BWD && tile0 < n_tiles && !(p.debug & 1)) {
      if (elect_one()) {
        mbar_arrive_expect_tx(my_aux, 4096);
        tma_load_2d(smem + G2_OFF_EPI + (warp - 2) * G2_EPI_BYTES + 4096, &tmO1, my_aux, cols_of(tile0), rows_of(tile0));
      }
      __syncwarp();
    }
    int it = 0;
    for (int tile = tile0; tile < n_tiles; tile += tile_stride, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row0 = rows_of(tile), col0 = cols_of(tile);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + half * (BN / 2);
      const uint32_t tempty_remote = tempty0 + acc * 8;
      if (!(p.debug & 1)) {
        const int nt = tile + tile_stride;
        pair_epilogue_slab<MODE>(p, &tmO0, &tmO1, buf, t_addr, row0, col0, lane, false, tempty_remote, my_aux, aux_phase, true, row0, col0 + 64);
        pair_epilogue_slab<MODE>(p, &tmO0, &tmO1, buf, t_addr + 64, row0, col0 + 64, lane, true, tempty_remote, my_aux, aux_phase, nt < n_tiles,
                                 rows_of(nt), cols_of(nt));
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(tempty_remote);
      }
    }
    if (elect_one()) bulk_wait_all();      // the staging boxes must outlive the TMA engine's reads; the writes complete with the kernel
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

template <int MODE>
int launch_gemm_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to0, const CUtensorMap& to1, const GemmParams& p, cudaStream_t stream) {
  static bool configured = false;
  auto kern = gemm_pair_kernel<MODE>;
  if (!configured) {
    DMI_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES));
    configured = true;
  }
  const int n_super = ((p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM)) * ((p.N + G2_BN - 1) / G2_BN);
  const int max_pairs = num_sms() / 2;
  const int grid = (n_super < max_pairs ? n_super : max_pairs) * 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = G2_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1 + pdl_attribute(&attr[1]);
  DMI_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, to0, to1, p));
  count_launch();
  return DMI_OK;
}

}  // namespace dmi
