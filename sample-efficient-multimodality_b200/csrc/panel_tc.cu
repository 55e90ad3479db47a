// Fused projection + batch-reduction pass on the tensor cores: one sweep over a bf16 activation gradient
// `in` [M, K] computes
//
//     out[M, R]  = in * W[R, K]^T             D_proj[128 rows x R]      += tile   * W_block^T     (tile as K-major  A operand)
//     G[R, K]   += scale * L[M, R]^T * in     D_red [128 cols x R]      += tile^T * L_panel       (tile as MN-major A operand)
//     colsum[K] += scale * 1^T in             column R of D_red: a warp writes 1.0 into column R of the staged L panel (N = R + 16)
//
// A TMA box of [128 rows x 64 columns] bf16 with SWIZZLE_128B is, physically, both the canonical K-major layout (rows = M, 64
// K-elements per 128-byte span) and the canonical MN-major layout (64 MN-elements contiguous, K rows 128 B apart), so the tensor
// core reads every staged tile twice under two descriptors and nothing is re-read from HBM or moved by threads.
//
// A 2-CTA cluster shares each 128-row panel: CTA c owns columns [c K/2, (c+1) K/2).  Its batch-reduction accumulators (K/256 column
// tiles x (R + 16) TMEM columns) stay in TMEM for the whole persistent kernel; the projection accumulator is double-buffered and its
// two partial sums are exchanged through distributed shared memory with complete_tx-signalling stores (each CTA finishes 64 rows).
// Warp roles: 0 = TMA producer, 1 = MMA issuer (one thread), 2..5 = epilogue (one TMEM lane quarter each).
#include "gemm_tc.cuh"
#include "panel.h"

namespace dmi {
namespace {

constexpr int PT_R = 32;                         // adapter rank this kernel is compiled for
constexpr int PT_ROWS = 128;
constexpr int PT_STAGES = 4;
constexpr int PT_THREADS = 192;
constexpr int PT_A_BYTES = 2 * PT_ROWS * 128;    // two [128 x 64] bf16 boxes
constexpr int PT_W_BYTES = 2 * PT_R * 128;       // two [R x 64] bf16 boxes of the projection weights
constexpr int PT_STAGE_BYTES = PT_A_BYTES + PT_W_BYTES;
constexpr int PT_L_BYTES = PT_ROWS * 128;        // [128 rows x 64-wide chunk], columns >= R zero-filled by TMA
constexpr int PT_OWN = PT_ROWS / 2;              // rows of a panel finished by each CTA of the pair
constexpr int PT_OFF_L = PT_STAGES * PT_STAGE_BYTES;
constexpr int PT_OFF_X = PT_OFF_L + 2 * PT_L_BYTES;            // [2][PT_OWN][R] fp32, 16-byte chunks XOR-swizzled by the row
constexpr int PT_OFF_BAR = PT_OFF_X + 2 * PT_OWN * PT_R * 4;
constexpr int PT_SMEM = PT_OFF_BAR + 256 + 1024 /*alignment slack*/;
static_assert(PT_SMEM <= 227 * 1024, "panel_tc: shared memory budget");

struct PanelTcParams {
  bf16* out; long long ld_out;
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

// NJ = column tiles (128 columns) per CTA = K / 256.  PROJ / RED select the projection and the batch reduction (+ column sum);
// the fused pass has both, `v = h A1` / `u = x A0` are PROJ only, `dA1 = h^T dv` / `dA0 = x^T du` are RED only.
// MCS (merged column sum, used whenever a column sum is requested): an extra warp writes a 1.0 into column R of every row of the L
// panel after it lands, and the batch reduction runs with N = R + 16, so the bias gradient costs no MMAs of its own (16 UMMAs per
// stage; the first version spent 8 more against an all-ones B tile: 35.4 vs 33.3 us).
template <int NJ, bool PROJ = true, bool RED = true, bool MCS = false>
__global__ void __launch_bounds__(PT_THREADS + (MCS ? 32 : 0), 1)
panel_tc_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmL,
                const PanelTcParams p) {
  constexpr int R = PT_R;
  constexpr uint32_t IDESC_P = make_idesc(PT_ROWS, R, 1);                                  // projection: both operands K-major
  constexpr int NRED = MCS ? R + 16 : R;                                                   // N (and TMEM column stride) of the batch reduction
  constexpr uint32_t IDESC_R = make_idesc(128, NRED, 1) | (1u << 15) | (1u << 16);        // batch reduction: both MN-major
  constexpr uint32_t COL_RED = 0, COL_PROJ = NJ * (R + 16);                                  // TMEM column map: reduction tile j at j * NRED
  static_assert(!MCS || (PROJ && RED), "merged column sum is a variant of the fused pass");
  static_assert(COL_PROJ + 2 * R <= 512, "TMEM budget");

  extern __shared__ uint8_t pt_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(pt_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + PT_OFF_BAR);     // [STAGES] TMA -> MMA
  uint64_t* empty_bar = full_bar + PT_STAGES;                              // [STAGES] MMA -> TMA
  uint64_t* lfull_bar = empty_bar + PT_STAGES;                             // [2] L panel landed
  uint64_t* lempty_bar = lfull_bar + 2;                                    // [2] L panel consumed
  uint64_t* tfull_bar = lempty_bar + 2;                                    // [2] projection accumulator complete
  uint64_t* tempty_bar = tfull_bar + 2;                                    // [2] projection accumulator drained
  uint64_t* xfull_bar = tempty_bar + 2;                                    // [2] peer's partial rows landed (complete_tx)
  uint64_t* xempty_bar = xfull_bar + 2;                                    // [2] peer consumed what this CTA sent
  uint64_t* rfull_bar = xempty_bar + 2;                                    // [1] all batch-reduction MMAs complete
  uint64_t* lready_bar = rfull_bar + 1;                                    // [2] MCS: ones column written into the L panel
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lready_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int col0 = static_cast<int>(crank) * (NJ * 128);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmIn);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmL);
    for (int s = 0; s < PT_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&lfull_bar[b], 1); mbar_init(&lempty_bar[b], 1);
      mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4);
      mbar_init(&xfull_bar[b], 1); mbar_init(&xempty_bar[b], 2);
    }
    mbar_init(rfull_bar, 1);
    mbar_init(&lready_bar[0], 1);
    mbar_init(&lready_bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // barrier inits visible to the peer before it stores / arrives remotely
  tc_fence_after();
  // programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) overlapped the predecessor's tail;
  // from here on global memory is read
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp runs the loops (warp-uniform control flow) and one elected lane issues: inside an `if (lane == 0)` region the
    // compiler cannot use the uniform datapath and wraps every UTMALDG / UTCHMMA / UTCBAR in an ELECT + R2UR.BROADCAST sequence
    // (~17 dependent instructions per MMA on one thread -- the "96 cycles per small-N UMMA" of round 1 was this issue path).
    {
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int pi = cluster_id; pi < p.n_panels; pi += p.n_clusters, ++it) {
        const int b = it & 1;
        const int row0 = pi * PT_ROWS;
        if (RED) {
          mbar_wait(&lempty_bar[b], ((it >> 1) & 1) ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&lfull_bar[b], PT_L_BYTES);
            tma_load_2d(smem + PT_OFF_L + b * PT_L_BYTES, &tmL, &lfull_bar[b], 0, row0);
          }
          __syncwarp();
        }
        for (int j = 0; j < NJ; ++j) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[stage], PROJ ? PT_STAGE_BYTES : PT_A_BYTES);
            uint8_t* sa = smem + stage * PT_STAGE_BYTES;
            const int col = col0 + j * 128;
            tma_load_2d(sa, &tmIn, &full_bar[stage], col, row0);
            tma_load_2d(sa + PT_A_BYTES / 2, &tmIn, &full_bar[stage], col + 64, row0);
            if (PROJ) {
              tma_load_2d(sa + PT_A_BYTES, &tmW, &full_bar[stage], col, 0);
              tma_load_2d(sa + PT_A_BYTES + PT_W_BYTES / 2, &tmW, &full_bar[stage], col + 64, 0);
            }
          }
          __syncwarp();
          if (++stage == PT_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    {
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int pi = cluster_id; pi < p.n_panels; pi += p.n_clusters, ++it) {
        const int b = it & 1;
        if (PROJ) mbar_wait(&tempty_bar[b], ((it >> 1) & 1) ^ 1);
        if (RED) mbar_wait(MCS ? &lready_bar[b] : &lfull_bar[b], (it >> 1) & 1);
        tc_fence_after();
        for (int j = 0; j < NJ; ++j) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d_proj = tmem_base + COL_PROJ + b * R;
            const uint64_t ldesc = make_mnmajor_sw128_desc(smem_u32(smem + PT_OFF_L + b * PT_L_BYTES), PT_L_BYTES);
            const uint32_t sa = smem_u32(smem + stage * PT_STAGE_BYTES);
            // projection: 2 column chunks x 4 k16 steps, K advances by 32 bytes inside the swizzle span (+2 in the address field)
#pragma unroll
            for (int c = 0; c < (PROJ ? 2 : 0); ++c) {
              const uint64_t ak = make_kmajor_sw128_desc(sa + c * (PT_A_BYTES / 2));
              const uint64_t wk = make_kmajor_sw128_desc(sa + PT_A_BYTES + c * (PT_W_BYTES / 2));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(d_proj, ak + 2 * k, wk + 2 * k, IDESC_P, (j | c | k) != 0);
            }
            // batch reduction and column sum: M = the 128 columns of the stage (two 64-column chunks 16 KB apart), K = the 128 rows,
            // advancing by 16 rows of 128 bytes (+128 in the address field)
            const uint64_t am = make_mnmajor_sw128_desc(sa, PT_A_BYTES / 2);
#pragma unroll
            for (int k = 0; k < (RED ? 8 : 0); ++k) umma_f16(tmem_base + COL_RED + j * NRED, am + 128 * k, ldesc + 128 * k, IDESC_R, (it | k) != 0);
            umma_commit(&empty_bar[stage]);
            if (j == NJ - 1) {
              if (PROJ) umma_commit(&tfull_bar[b]);
              if (RED) umma_commit(&lempty_bar[b]);
            }
          }
          __syncwarp();
          if (++stage == PT_STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (RED) {
        if (elect_one()) umma_commit(rfull_bar);
        __syncwarp();
      }
    }
  } else if (MCS && warp == 6) {
    // ===================== ones column of the L panel (merged column sum) =====================
    // Row t of the [128 x 64] SWIZZLE_128B panel: columns R..R+7 are 16-byte chunk R/8, stored at chunk (R/8) ^ (t & 7).  TMA zero-filled
    // columns >= R; rows beyond M carry zero tile rows, so their 1.0 contributes nothing and needs no guard.
    int it = 0;
    for (int pi = cluster_id; pi < p.n_panels; pi += p.n_clusters, ++it) {
      const int b = it & 1;
      mbar_wait(&lfull_bar[b], (it >> 1) & 1);
      const uint32_t sl = smem_u32(smem + PT_OFF_L + b * PT_L_BYTES);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = i * 32 + lane;
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(sl + t * 128 + (((R / 8) ^ (t & 7)) << 4)), "h"(static_cast<unsigned short>(0x3F80)) : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&lready_bar[b]);
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;                          // TMEM lanes [32 quarter, +32)
    const int row_p = quarter * 32 + lane;                 // row of the panel held by this thread
    const uint32_t owner = static_cast<uint32_t>(row_p >> 6);
    const bool is_owner = owner == crank;
    const uint32_t peer = crank ^ 1u;
    const int row_l = row_p & (PT_OWN - 1);
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    int it = 0;
    for (int pi = cluster_id; PROJ && pi < p.n_panels; pi += p.n_clusters, ++it) {
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
      const uint32_t xrow = smem_u32(smem + PT_OFF_X) + static_cast<uint32_t>((b * PT_OWN + row_l) * R * 4);
      if (!is_owner) {
        // this row is finished by the peer: wait until it has consumed what was sent into the slot two panels ago, then send
        mbar_wait(&xempty_bar[b], ph ^ 1);
        const uint32_t bar = mapa_u32(smem_u32(&xfull_bar[b]), peer);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          st_async_f4(mapa_u32(xrow + ((c ^ (row_l & 7)) << 4), peer), r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3], bar);
      } else {
        if (lane == 0 && (quarter & 1) == 0) mbar_arrive_expect_tx(&xfull_bar[b], PT_OWN * R * 4);
        mbar_wait(&xfull_bar[b], ph);
        uint4 o[4];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = lds128(xrow + ((c ^ (row_l & 7)) << 4));
          const uint32_t lo = pack_bf16x2(__uint_as_float(r[4 * c]) + v.x, __uint_as_float(r[4 * c + 1]) + v.y);
          const uint32_t hi = pack_bf16x2(__uint_as_float(r[4 * c + 2]) + v.z, __uint_as_float(r[4 * c + 3]) + v.w);
          if (c & 1) { o[c >> 1].z = lo; o[c >> 1].w = hi; } else { o[c >> 1].x = lo; o[c >> 1].y = hi; }
        }
        const long long row = static_cast<long long>(pi) * PT_ROWS + row_p;
        if (row < p.M) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + row * p.ld_out);
#pragma unroll
          for (int c = 0; c < 4; ++c) dst[c] = o[c];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&xempty_bar[b]), peer));
      }
    }
    // ---- final flush: lane = column, registers = rank index -> coalesced atomics into G[R, K] and colsum[K] ----
    if (RED) {
      mbar_wait(rfull_bar, 0);
      tc_fence_after();
    }
    for (int j = 0; RED && j < NJ; ++j) {
      uint32_t r[32], cs = 0;
      tmem_ld_32x32(lane_base + COL_RED + j * NRED, r);
      if (MCS) tmem_ld_32x1(lane_base + COL_RED + j * NRED + R, cs);
      tmem_ld_wait();
      const int q = col0 + j * 128 + quarter * 32 + lane;
#pragma unroll
      for (int i = 0; i < R; ++i) atomicAdd(p.G + static_cast<long long>(i) * p.ldg + q, __uint_as_float(r[i]) * p.scale);
      if (MCS) atomicAdd(p.colsum + q, __uint_as_float(cs) * p.scale);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // nobody exits while the peer can still store into / arrive on this CTA's shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int NJ, bool PROJ = true, bool RED = true, bool MCS = false>
int launch_panel_tc(const CUtensorMap& tIn, const CUtensorMap& tW, const CUtensorMap& tL, const PanelTcParams& p0, cudaStream_t stream) {
  auto kern = panel_tc_kernel<NJ, PROJ, RED, MCS>;
  static int max_clusters = 0;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(PT_THREADS + (MCS ? 32 : 0));
  cfg.dynamicSmemBytes = PT_SMEM;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = 1 + pdl_attribute(&attr[1]);
  if (max_clusters == 0) {
    DMI_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PT_SMEM));
    cfg.gridDim = dim3(2 * (num_sms() / 2));
    int n = 0;
    DMI_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    if (n < 1) {
      set_error("panel_fused_tc: no 2-CTA cluster with %d B of shared memory can be resident", PT_SMEM);
      return DMI_ERR_UNSUPPORTED;
    }
    max_clusters = n;
  }
  PanelTcParams p = p0;
  p.n_clusters = max_clusters < p.n_panels ? max_clusters : p.n_panels;
  cfg.gridDim = dim3(2 * p.n_clusters);
  DMI_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tIn, tW, tL, p));
  count_launch();
  return DMI_OK;
}

}  // namespace

bool panel_fused_tc_supported(long long K, int R) { return R == PT_R && (K == 1024 || K == 2048); }
bool panel_tc_mode_supported(long long K, int R) { return R == PT_R && (K == 768 || K == 1024 || K == 2048); }

// W == nullptr: no projection (out unused).  L == nullptr: no batch reduction (G, colsum unused).
static int panel_tc_run(const bf16* in, long long ld_in, const bf16* W, long long ldw, bf16* out, long long ld_out, const bf16* L, long long ldl,
                        float* G, long long ldg, float* colsum, float scale, long long M, long long K, int R, cudaStream_t s) {
  const bool proj = W != nullptr, red = L != nullptr;
  DMI_REQUIRE(in && (proj || red) && M > 0, "panel_tc: bad arguments");
  DMI_REQUIRE(!proj || ((reinterpret_cast<uintptr_t>(out) & 15) == 0 && out != nullptr && ld_out % 8 == 0),
              "panel_tc: out must be 16-byte aligned with ld_out %% 8 == 0 (ld_out=%lld)", ld_out);
  DMI_REQUIRE(!red || G != nullptr, "panel_tc: missing G");
  CUtensorMap tIn, tW, tL;
  int rc = make_tmap_2d(&tIn, in, KIND_BF16, K, M, ld_in, PT_ROWS);
  if (rc != DMI_OK) return rc;
  tW = tIn;
  tL = tIn;
  if (proj) {
    rc = make_tmap_2d(&tW, W, KIND_BF16, K, R, ldw, R);
    if (rc != DMI_OK) return rc;
  }
  if (red) {
    rc = make_tmap_2d(&tL, L, KIND_BF16, R, M, ldl, PT_ROWS);   // inner extent R < the 64-element box: the rest is zero-filled
    if (rc != DMI_OK) return rc;
  }
  PanelTcParams p;
  p.out = out; p.ld_out = ld_out; p.G = G; p.ldg = ldg; p.colsum = red ? colsum : nullptr; p.scale = scale; p.M = static_cast<int>(M);
  p.n_panels = static_cast<int>((M + PT_ROWS - 1) / PT_ROWS);
  p.n_clusters = 0;
  if (proj && red) {
    DMI_REQUIRE(panel_fused_tc_supported(K, R), "panel_fused_tc: K=%lld R=%d outside the compiled shapes (K 1024/2048, R 32)", K, R);
    if (colsum != nullptr)      // column sum merged into the batch-reduction MMAs (16 UMMAs per stage; 33.3 vs 35.4 us with a separate ones tile)
      return K == 2048 ? launch_panel_tc<8, true, true, true>(tIn, tW, tL, p, s) : launch_panel_tc<4, true, true, true>(tIn, tW, tL, p, s);
    return K == 2048 ? launch_panel_tc<8>(tIn, tW, tL, p, s) : launch_panel_tc<4>(tIn, tW, tL, p, s);
  }
  DMI_REQUIRE(proj && panel_tc_mode_supported(K, R), "panel_tc_project: K=%lld R=%d outside the compiled shapes (K 768/1024/2048, R 32)", K, R);
  if (K == 2048) return launch_panel_tc<8, true, false>(tIn, tW, tL, p, s);
  if (K == 1024) return launch_panel_tc<4, true, false>(tIn, tW, tL, p, s);
  return launch_panel_tc<3, true, false>(tIn, tW, tL, p, s);
}

int panel_fused_tc(const bf16* in, long long ld_in, const bf16* W, long long ldw, bf16* out, long long ld_out, const bf16* L, long long ldl,
                   float* G, long long ldg, float* colsum, float scale, long long M, long long K, int R, cudaStream_t s) {
  DMI_REQUIRE(in && W && out && L && G && M > 0, "panel_fused_tc: bad arguments");
  return panel_tc_run(in, ld_in, W, ldw, out, ld_out, L, ldl, G, ldg, colsum, scale, M, K, R, s);
}

// out[M,R] = in W^T only (v = h A1, u = x A0)
int panel_tc_project(const bf16* in, long long ld_in, const bf16* W, long long ldw, bf16* out, long long ld_out, long long M, long long K, int R,
                     cudaStream_t s) {
  DMI_REQUIRE(in && W && out, "panel_tc_project: bad arguments");
  return panel_tc_run(in, ld_in, W, ldw, out, ld_out, nullptr, 0, nullptr, 0, nullptr, 1.0f, M, K, R, s);
}

}  // namespace dmi
