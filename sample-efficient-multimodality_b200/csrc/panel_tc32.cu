// fp32-input form of the tcgen05 fused panel pass (panel_tc.cu): the dY sweep of the adapted-MLP backward.
//
//     copy[M, K] = bf16(in)                    the A operand of the following dpre GEMM (written once, coalesced)
//     out[M, R]  = bf16(in) * W[R, K]^T        dv  = dY B1^T
//     G[R, K]   += scale * L^T * bf16(in)      dB1 = v^T dY
//     colsum[K] += scale * 1^T bf16(in)        dbeta1
//
// Everything downstream of the bf16 tile (descriptors, TMEM map, epilogue, exchange) is the panel_tc.cu design; the front end differs:
// TMA stages fp32 quarter tiles ([128 rows x 32 floats], SWIZZLE_128B) in a 4-deep ring, four converter warps turn each into bf16 --
// written into the MMA tile in the 128B-swizzled layout the UMMA descriptors expect, and to global memory as the bf16 copy -- and hand
// the tile to the MMA thread through the async-proxy fence.  Default dY pass of the backward for >= 8192 rows (api.cu: use_panel_tc).
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer (one thread), 2..5 = epilogue (TMEM lane quarters), 6..9 = converters, 10 = TMA producer of
// the W blocks.
#include "gemm_tc.cuh"
#include "panel.h"

namespace dmi {
namespace {

constexpr int PF_R = 32;
constexpr int PF_ROWS = 128;
constexpr int PF_THREADS = 352;                          // TMA (fp32 quarters), MMA, 4 epilogue, 4 converter warps, TMA (W blocks)
constexpr int PF_A_BYTES = 2 * PF_ROWS * 128;              // bf16 tile: two [128 x 64] chunks
constexpr int PF_W_BYTES = 2 * PF_R * 128;
constexpr int PF_TILE_BYTES = PF_A_BYTES + PF_W_BYTES;     // 40 KB, same layout as a panel_tc.cu stage
constexpr int PF_NT = 2;                                   // bf16 tile buffers
constexpr int PF_Q_BYTES = PF_ROWS * 128;                  // fp32 quarter tile: [128 rows x 32 floats]
constexpr int PF_NF = 5;                                   // fp32 staging ring depth
constexpr int PF_L_BYTES = PF_ROWS * 128;
constexpr int PF_OWN = PF_ROWS / 2;
constexpr int PF_OFF_F = PF_NT * PF_TILE_BYTES;
constexpr int PF_OFF_L = PF_OFF_F + PF_NF * PF_Q_BYTES;
constexpr int PF_OFF_ONES = PF_OFF_L + 2 * PF_L_BYTES;
constexpr int PF_OFF_X = PF_OFF_ONES + PF_L_BYTES;
constexpr int PF_OFF_BAR = PF_OFF_X + 2 * PF_OWN * PF_R * 4;
constexpr int PF_SMEM = PF_OFF_BAR + 256 + 1024;
static_assert(PF_SMEM <= 227 * 1024, "panel_tc32: shared memory budget");
static_assert(PF_OFF_F % 1024 == 0 && PF_OFF_L % 1024 == 0 && PF_OFF_ONES % 1024 == 0, "swizzled buffers must be 1024-byte aligned");

struct PanelTc32Params {
  bf16* out; long long ld_out;
  bf16* copy; long long ld_copy;
  float* G; long long ldg;
  float* colsum;
  float scale;
  int M;
  int n_panels, n_clusters;
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_f4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// NJ = column tiles (128 columns) per CTA = K / 256
template <int NJ>
__global__ void __launch_bounds__(PF_THREADS, 1)
panel_tc32_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmL,
                  const PanelTc32Params p) {
  constexpr int R = PF_R;
  constexpr uint32_t IDESC_P = make_idesc(PF_ROWS, R, 1);
  constexpr uint32_t IDESC_R = make_idesc(128, R, 1) | (1u << 15) | (1u << 16);
  constexpr uint32_t IDESC_C = make_idesc(128, 16, 1) | (1u << 15) | (1u << 16);
  constexpr uint32_t COL_RED = 0, COL_CS = NJ * R, COL_PROJ = NJ * R + NJ * 16;
  static_assert(COL_PROJ + 2 * R <= 512, "TMEM budget");

  extern __shared__ uint8_t pf_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(pf_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* ffull_bar = reinterpret_cast<uint64_t*>(smem + PF_OFF_BAR);    // [NF] fp32 quarter landed (TMA -> converters)
  uint64_t* fempty_bar = ffull_bar + PF_NF;                                // [NF] fp32 quarter consumed (4 converter warps)
  uint64_t* tlfull_bar = fempty_bar + PF_NF;                               // [NT] bf16 tile complete (4 warps x 4 quarters)
  uint64_t* tlempty_bar = tlfull_bar + PF_NT;                              // [NT] bf16 tile + W block consumed (MMA commit)
  uint64_t* wfull_bar = tlempty_bar + PF_NT;                               // [NT] W block of the tile landed
  uint64_t* lfull_bar = wfull_bar + PF_NT;                                 // [2]
  uint64_t* lempty_bar = lfull_bar + 2;                                    // [2]
  uint64_t* tfull_bar = lempty_bar + 2;                                    // [2] projection accumulator complete
  uint64_t* tempty_bar = tfull_bar + 2;                                    // [2]
  uint64_t* xfull_bar = tempty_bar + 2;                                    // [2]
  uint64_t* xempty_bar = xfull_bar + 2;                                    // [2]
  uint64_t* rfull_bar = xempty_bar + 2;                                    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rfull_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int col0 = static_cast<int>(crank) * (NJ * 128);
  const bool do_colsum = p.colsum != nullptr;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmIn);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmL);
    for (int s = 0; s < PF_NF; ++s) { mbar_init(&ffull_bar[s], 1); mbar_init(&fempty_bar[s], 4); }
    for (int s = 0; s < PF_NT; ++s) { mbar_init(&tlfull_bar[s], 16); mbar_init(&tlempty_bar[s], 1); mbar_init(&wfull_bar[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&lfull_bar[b], 1); mbar_init(&lempty_bar[b], 1);
      mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4);
      mbar_init(&xfull_bar[b], 1); mbar_init(&xempty_bar[b], 2);
    }
    mbar_init(rfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (warp >= 2 && warp < 6) {
    const uint32_t ones = smem_u32(smem + PF_OFF_ONES) + (threadIdx.x - 64) * 128;
    const float one2 = __uint_as_float(0x3F803F80u);
#pragma unroll
    for (int c = 0; c < 8; ++c) sts128(ones + c * 16, one2, one2, one2, one2);
    fence_proxy_async_smem();
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
    // ===================== TMA producer (warp-uniform loops, one elected lane issues; see panel_tc.cu) =====================
    // fp32 quarter tiles only: they need nothing but a free staging slot, so this warp runs ahead of everything else (the W blocks,
    // which have to wait for the tile buffers, are loaded by warp 10).  Measured at 32768 x 2048: ring of 4 slots 82.7 us, 5 slots
    // (all the shared memory there is) 80.6 us -- 74.9 us with a second converter group, removed for the phase hazard described at
    // the converters; the conflict-free row mapping and the separate W producer changed nothing measurable.  An L2 prefetch cursor
    // (cp.async.bulk.prefetch.tensor) ahead of the loads did not help either (75.9 / 78.4 us at distance 1 / 2) and was removed.
    {
      int fs = 0, it = 0;
      uint32_t fph = 0;
      for (int pi = cluster_id; pi < p.n_panels; pi += p.n_clusters, ++it) {
        const int b = it & 1;
        const int row0 = pi * PF_ROWS;
        mbar_wait(&lempty_bar[b], ((it >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&lfull_bar[b], PF_L_BYTES);
          tma_load_2d(smem + PF_OFF_L + b * PF_L_BYTES, &tmL, &lfull_bar[b], 0, row0);
        }
        __syncwarp();
        for (int j = 0; j < NJ; ++j) {
          const int col = col0 + j * 128;
          for (int q = 0; q < 4; ++q) {
            mbar_wait(&fempty_bar[fs], fph ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&ffull_bar[fs], PF_Q_BYTES);
              tma_load_2d(smem + PF_OFF_F + fs * PF_Q_BYTES, &tmIn, &ffull_bar[fs], col + q * 32, row0);
            }
            __syncwarp();
            if (++fs == PF_NF) { fs = 0; fph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 10) {
    // ===================== W-block producer =====================
    // the W block lives in the tile buffer, which is free once the MMAs that read it two tiles ago have completed
    int tb = 0;
    uint32_t tph = 0;
    for (int pi = cluster_id; pi < p.n_panels; pi += p.n_clusters) {
      for (int j = 0; j < NJ; ++j) {
        const int col = col0 + j * 128;
        mbar_wait(&tlempty_bar[tb], tph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&wfull_bar[tb], PF_W_BYTES);
          uint8_t* st = smem + tb * PF_TILE_BYTES;
          tma_load_2d(st + PF_A_BYTES, &tmW, &wfull_bar[tb], col, 0);
          tma_load_2d(st + PF_A_BYTES + PF_W_BYTES / 2, &tmW, &wfull_bar[tb], col + 64, 0);
        }
        __syncwarp();
        if (++tb == PF_NT) { tb = 0; tph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    {
      int tb = 0, it = 0;
      uint32_t tph = 0;
      for (int pi = cluster_id; pi < p.n_panels; pi += p.n_clusters, ++it) {
        const int b = it & 1;
        mbar_wait(&tempty_bar[b], ((it >> 1) & 1) ^ 1);
        mbar_wait(&lfull_bar[b], (it >> 1) & 1);
        tc_fence_after();
        for (int j = 0; j < NJ; ++j) {
          mbar_wait(&wfull_bar[tb], tph);
          mbar_wait(&tlfull_bar[tb], tph);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t odesc = make_mnmajor_sw128_desc(smem_u32(smem + PF_OFF_ONES), PF_L_BYTES);
            const uint32_t d_proj = tmem_base + COL_PROJ + b * R;
            const uint64_t ldesc = make_mnmajor_sw128_desc(smem_u32(smem + PF_OFF_L + b * PF_L_BYTES), PF_L_BYTES);
            const uint32_t sa = smem_u32(smem + tb * PF_TILE_BYTES);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const uint64_t ak = make_kmajor_sw128_desc(sa + c * (PF_A_BYTES / 2));
              const uint64_t wk = make_kmajor_sw128_desc(sa + PF_A_BYTES + c * (PF_W_BYTES / 2));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(d_proj, ak + 2 * k, wk + 2 * k, IDESC_P, (j | c | k) != 0);
            }
            const uint64_t am = make_mnmajor_sw128_desc(sa, PF_A_BYTES / 2);
#pragma unroll
            for (int k = 0; k < 8; ++k) umma_f16(tmem_base + COL_RED + j * R, am + 128 * k, ldesc + 128 * k, IDESC_R, (it | k) != 0);
            if (do_colsum) {
#pragma unroll
              for (int k = 0; k < 8; ++k) umma_f16(tmem_base + COL_CS + j * 16, am + 128 * k, odesc + 128 * k, IDESC_C, (it | k) != 0);
            }
            umma_commit(&tlempty_bar[tb]);
            if (j == NJ - 1) {
              umma_commit(&tfull_bar[b]);
              umma_commit(&lempty_bar[b]);
            }
          }
          __syncwarp();
          if (++tb == PF_NT) { tb = 0; tph ^= 1; }
        }
      }
      if (elect_one()) umma_commit(rfull_bar);
      __syncwarp();
    }
  } else if (warp >= 6 && warp < 10) {
    // ===================== converters (warps 6..9) =====================
    // Quarter tile = [128 rows x 32 floats]; warp cw owns rows [32 cw, +32) of EVERY quarter.  (A second group of converter warps taking
    // alternate quarters was tried and removed: it bought nothing, and a warp that skips a phase of ffull_bar can be fooled by the
    // parity of the phase before it when TMA completions arrive out of order -- a rare wrong tile at N = 1 and, with H2D and NCCL
    // traffic perturbing the timing, launch failures in the multi-rank end-to-end runs.  Every waiter now sees every phase.)
    // Per instruction a lane handles one 16-byte bf16 chunk (8 columns = two fp32 chunks): lane = (g = lane >> 2, c4 = lane & 3), 4
    // instructions cover the warp's 32 rows; the global store of an instruction covers 8 rows x 64 contiguous bytes.  Row of group g
    // inside the 8-row block: (g >> 1) ^ (g & 1 ? 5 : 0), so that the two rows of a quarter-warp differ in swizzle bits 0 and 2 -- both
    // the fp32 reads (chunks 2 c4 ^ sw) and the bf16 writes (chunks cc ^ sw) then touch 8 distinct 16-byte bank groups (rows r, r + 1
    // gave a 2-way conflict on every STS.128).
    const int cw = warp - 6;
    const int g8 = lane >> 2, c4 = lane & 3;
    const int row_sub = (g8 >> 1) ^ ((g8 & 1) ? 5 : 0);
    int fs = 0, tb = 0;
    uint32_t fph = 0, tph = 0;
    for (int pi = cluster_id; pi < p.n_panels; pi += p.n_clusters) {
      const long long row0 = static_cast<long long>(pi) * PF_ROWS;
      for (int j = 0; j < NJ; ++j) {
        mbar_wait(&tlempty_bar[tb], tph ^ 1);
        const uint32_t tile = smem_u32(smem + tb * PF_TILE_BYTES);
        for (int q = 0; q < 4; ++q) {
          mbar_wait(&ffull_bar[fs], fph);
          const uint32_t fq = smem_u32(smem + PF_OFF_F + fs * PF_Q_BYTES);
          const int cc = (q & 1) * 4 + c4;                   // bf16 chunk inside the 64-column half h = q >> 1
          const uint32_t thalf = tile + (q >> 1) * (PF_A_BYTES / 2);
          const int gcol = col0 + j * 128 + q * 32 + c4 * 8;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = cw * 32 + i * 8 + row_sub;
            const int sw = row & 7;
            const float4 a = lds128(fq + row * 128 + (((2 * c4) ^ sw) << 4));
            const float4 b4 = lds128(fq + row * 128 + (((2 * c4 + 1) ^ sw) << 4));
            const uint32_t w0 = pack_bf16x2(a.x, a.y), w1 = pack_bf16x2(a.z, a.w), w2 = pack_bf16x2(b4.x, b4.y), w3 = pack_bf16x2(b4.z, b4.w);
            sts128u(thalf + row * 128 + ((cc ^ sw) << 4), w0, w1, w2, w3);
            if (p.copy != nullptr && row0 + row < p.M)
              *reinterpret_cast<uint4*>(p.copy + (row0 + row) * p.ld_copy + gcol) = make_uint4(w0, w1, w2, w3);
          }
          fence_proxy_async_smem();            // generic-proxy tile writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&tlfull_bar[tb]);
            mbar_arrive(&fempty_bar[fs]);
          }
          if (++fs == PF_NF) { fs = 0; fph ^= 1; }
        }
        if (++tb == PF_NT) { tb = 0; tph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5), identical to panel_tc.cu =====================
    const int quarter = warp & 3;
    const int row_p = quarter * 32 + lane;
    const bool is_owner = static_cast<uint32_t>(row_p >> 6) == crank;
    const uint32_t peer = crank ^ 1u;
    const int row_l = row_p & (PF_OWN - 1);
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    int it = 0;
    for (int pi = cluster_id; pi < p.n_panels; pi += p.n_clusters, ++it) {
      const int b = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(&tfull_bar[b], ph);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32(lane_base + COL_PROJ + b * R, r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[b]);
      const uint32_t xrow = smem_u32(smem + PF_OFF_X) + static_cast<uint32_t>((b * PF_OWN + row_l) * R * 4);
      if (!is_owner) {
        mbar_wait(&xempty_bar[b], ph ^ 1);
        const uint32_t bar = mapa_u32(smem_u32(&xfull_bar[b]), peer);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          st_async_f4(mapa_u32(xrow + ((c ^ (row_l & 7)) << 4), peer), r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3], bar);
      } else {
        if (lane == 0 && (quarter & 1) == 0) mbar_arrive_expect_tx(&xfull_bar[b], PF_OWN * R * 4);
        mbar_wait(&xfull_bar[b], ph);
        uint4 o[4];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = lds128(xrow + ((c ^ (row_l & 7)) << 4));
          const uint32_t lo = pack_bf16x2(__uint_as_float(r[4 * c]) + v.x, __uint_as_float(r[4 * c + 1]) + v.y);
          const uint32_t hi = pack_bf16x2(__uint_as_float(r[4 * c + 2]) + v.z, __uint_as_float(r[4 * c + 3]) + v.w);
          if (c & 1) { o[c >> 1].z = lo; o[c >> 1].w = hi; } else { o[c >> 1].x = lo; o[c >> 1].y = hi; }
        }
        const long long row = static_cast<long long>(pi) * PF_ROWS + row_p;
        if (row < p.M) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + row * p.ld_out);
#pragma unroll
          for (int c = 0; c < 4; ++c) dst[c] = o[c];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&xempty_bar[b]), peer));
      }
    }
    mbar_wait(rfull_bar, 0);
    tc_fence_after();
    for (int j = 0; j < NJ; ++j) {
      uint32_t r[32], cs = 0;
      tmem_ld_32x32(lane_base + COL_RED + j * R, r);
      if (do_colsum) tmem_ld_32x1(lane_base + COL_CS + j * 16, cs);
      tmem_ld_wait();
      const int q = col0 + j * 128 + quarter * 32 + lane;
#pragma unroll
      for (int i = 0; i < R; ++i) atomicAdd(p.G + static_cast<long long>(i) * p.ldg + q, __uint_as_float(r[i]) * p.scale);
      if (do_colsum) atomicAdd(p.colsum + q, __uint_as_float(cs) * p.scale);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int NJ>
int launch_panel_tc32(const CUtensorMap& tIn, const CUtensorMap& tW, const CUtensorMap& tL, const PanelTc32Params& p0, cudaStream_t stream) {
  auto kern = panel_tc32_kernel<NJ>;
  static int max_clusters = 0;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(PF_THREADS);
  cfg.dynamicSmemBytes = PF_SMEM;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = 1 + pdl_attribute(&attr[1]);
  if (max_clusters == 0) {
    DMI_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_SMEM));
    cfg.gridDim = dim3(2 * (num_sms() / 2));
    int n = 0;
    DMI_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    if (n < 1) {
      set_error("panel_fused_tc32: no 2-CTA cluster with %d B of shared memory can be resident", PF_SMEM);
      return DMI_ERR_UNSUPPORTED;
    }
    max_clusters = n;
  }
  PanelTc32Params p = p0;
  p.n_clusters = max_clusters < p.n_panels ? max_clusters : p.n_panels;
  cfg.gridDim = dim3(2 * p.n_clusters);
  DMI_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tIn, tW, tL, p));
  count_launch();
  return DMI_OK;
}

}  // namespace

int panel_fused_tc32(const float* in, long long ld_in, const bf16* W, long long ldw, bf16* out, long long ld_out, bf16* copy, long long ld_copy,
                     const bf16* L, long long ldl, float* G, long long ldg, float* colsum, float scale, long long M, long long K, int R,
                     cudaStream_t s) {
  DMI_REQUIRE(in && W && out && L && G && M > 0, "panel_fused_tc32: bad arguments");
  DMI_REQUIRE(panel_fused_tc_supported(K, R), "panel_fused_tc32: K=%lld R=%d outside the compiled shapes (K 1024/2048, R 32)", K, R);
  DMI_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && ld_out % 8 == 0 &&
                  (copy == nullptr || ((reinterpret_cast<uintptr_t>(copy) & 15) == 0 && ld_copy % 8 == 0)),
              "panel_fused_tc32: out / copy must be 16-byte aligned with ld %% 8 == 0 (ld_out=%lld ld_copy=%lld)", ld_out, ld_copy);
  CUtensorMap tIn, tW, tL;
  int rc = make_tmap_2d(&tIn, in, KIND_TF32, K, M, ld_in, PF_ROWS);      // fp32: boxes of [128 rows x 32 floats]
  if (rc != DMI_OK) return rc;
  rc = make_tmap_2d(&tW, W, KIND_BF16, K, R, ldw, R);
  if (rc != DMI_OK) return rc;
  rc = make_tmap_2d(&tL, L, KIND_BF16, R, M, ldl, PF_ROWS);
  if (rc != DMI_OK) return rc;
  PanelTc32Params p;
  p.out = out; p.ld_out = ld_out; p.copy = copy; p.ld_copy = ld_copy; p.G = G; p.ldg = ldg; p.colsum = colsum; p.scale = scale;
  p.M = static_cast<int>(M);
  p.n_panels = static_cast<int>((M + PF_ROWS - 1) / PF_ROWS);
  p.n_clusters = 0;
  return K == 2048 ? launch_panel_tc32<8>(tIn, tW, tL, p, s) : launch_panel_tc32<4>(tIn, tW, tL, p, s);
}

}  // namespace dmi
