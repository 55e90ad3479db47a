// Fused row-panel pass of the adapted-MLP backward: ONE sweep over a [M, K] activation gradient computes
//
//     out[M, R]  = in[M, K] * W[R, K]^T                      (rank-r projection:  dv = dY B1^T,   du = dpre B0^T)
//     G[R, K]   += scale * L[M, R]^T * in[M, K]              (batch reduction:    dB1 = v^T dY,   dB0 = u^T dpre)
//     colsum[K] += scale * 1^T in[M, K]                      (bias gradient:      dbeta1,         dbeta0)
//     copy[M, K] = bf16(in)                                  (IN_F32 only: the bf16 operand of the following tcgen05 GEMM)
//
// It replaces a skinny_rows launch followed by an outer_reduce launch over the same matrix (skinny.cuh, outer_mma.cuh), i.e. one
// full HBM pass per use.  The two contractions pull in opposite directions -- the projection reduces over the columns of a row,
// the batch reduction over the rows of a column -- and the batch-reduction accumulators of a full row ((R+1) x K fp32 = 270 KB at
// K = 2048) exceed one SM's register file.  So a 4-CTA thread-block cluster shares each 64-row panel: CTA c streams columns
// [c K/4, (c+1) K/4), keeps its (R+1) x K/4 accumulators in registers for the whole kernel (persistent over panels, one atomic
// flush at the end) and contributes a partial projection; the partials are reduce-scattered through distributed shared memory
// (each CTA finishes 16 of the panel's 64 rows) with complete_tx-signalling remote stores (st.async) on the owner's mbarrier.
//
// Work is HBM-bound by construction (<= 66 FLOP/byte) and uses warp-level mma.sync m16n8k16 from a cp.async ring; per 32 KB stage
// the tensor pipe sees 640 MMAs and shared memory ~5x the HBM bytes (operand fragments are re-read by the warps that share them),
// which is why the batch-reduction's left operand is held in registers per panel and the all-ones row of the column sum is a
// constant fragment rather than a shared-memory tile.
#include "skinny.cuh"
#include "panel.h"

namespace dmi {

int num_sms();

namespace {

constexpr int PN_CLUSTER = 4;
constexpr int PN_ROWS = 64;                 // rows per panel
constexpr int PN_KC = 256;                  // columns per pipeline stage
constexpr int PN_THREADS = 256;             // 8 warps
constexpr int PN_TW = PN_KC + 8;            // bf16 tile row stride (elements); 528 B = odd multiple of 16 B: ldmatrix conflict-free
constexpr int PN_FW = PN_KC + 4;            // fp32 staging row stride (floats)
constexpr int PN_OWN = PN_ROWS / PN_CLUSTER;  // rows of a panel finished by each CTA of the cluster

struct PanelParams {
  const void* in; long long ld_in;
  const bf16* W; long long ldw;
  bf16* out; long long ld_out;
  bf16* copy; long long ld_copy;
  const bf16* L; long long ldl;
  float* G; long long ldg;
  float* colsum;
  float scale;
  int M, K;
  int n_panels, n_clusters;
};

__device__ __forceinline__ uint32_t mapa_cluster(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// complete_tx-signalling remote store: the 16 bytes land in the peer's shared memory and are counted on the peer's mbarrier --
// no cluster-scope fence (MEMBAR.ALL.GPU, which also waits for this thread's cp.async traffic) on the sender's side.
__device__ __forceinline__ void st_async_v4(uint32_t addr, float a, float b, float c, float d, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d), "r"(mbar) : "memory");
}
template <int R>
__device__ __forceinline__ int red_swizzle(int row) { return R >= 32 ? (row & 3) : ((row >> 1) & 1); }

template <int R, int NCH, bool IN_F32>
struct PanelSmem {
  static constexpr int LW = R + 8;                                   // L tile row stride (elements); odd multiple of 16 B (static_assert below)
  static constexpr int KQ = NCH * PN_KC;                             // columns owned by one CTA
  static constexpr int WW = KQ + 8;
  static constexpr int NST = IN_F32 ? 2 : 4;                         // ring depth (fp32 staging stages / bf16 tile stages)
  static constexpr int NLS = IN_F32 ? 3 : NST;                       // L tile slots (fp32 mode: panels pi, pi+1, pi+2 can be in flight)
  static constexpr int tile_bytes = PN_ROWS * PN_TW * 2;
  static constexpr int f32_bytes = PN_ROWS * PN_FW * 4;
  static constexpr int l_bytes = PN_ROWS * LW * 2;
  static constexpr int off_tiles = 0;                                // bf16 tiles: NST (bf16 mode) or 1 (fp32 mode)
  static constexpr int off_f32 = off_tiles + (IN_F32 ? 1 : NST) * tile_bytes;
  static constexpr int off_l = off_f32 + (IN_F32 ? NST * f32_bytes : 0);
  static constexpr int off_w = off_l + NLS * l_bytes;
  static constexpr int off_xch = off_w + R * WW * 2;                 // [2][PN_CLUSTER][PN_OWN][R] fp32
  static constexpr int off_bar = off_xch + 2 * PN_CLUSTER * PN_OWN * R * 4;   // 2 receive mbarriers
  static constexpr int total = off_bar + 16;
  static_assert(((LW * 2 / 16) & 1) == 1, "L tile stride must be an odd multiple of 16 bytes");
  static_assert((PN_THREADS / 32) * 32 * R * 4 <= tile_bytes, "projection partials must fit in one tile slot");
};

template <int R, int NCH, bool IN_F32>
__global__ void __launch_bounds__(PN_THREADS, 1)
panel_fused_kernel(const PanelParams p) {
  using S = PanelSmem<R, NCH, IN_F32>;
  constexpr int NT = R / 8;              // n8 tiles of the projection
  constexpr int MT = R / 16;             // m16 tiles of the batch-reduction's left operand (R = 16, 32, 64)
  constexpr int LW = S::LW, WW = S::WW, KQ = S::KQ, NST = S::NST;
  static_assert(R % 16 == 0, "rank must be a multiple of 16");
  extern __shared__ __align__(128) uint8_t psm[];
  bf16* sT = reinterpret_cast<bf16*>(psm + S::off_tiles);
  float* sF = reinterpret_cast<float*>(psm + S::off_f32);
  bf16* sL = reinterpret_cast<bf16*>(psm + S::off_l);
  bf16* sW = reinterpret_cast<bf16*>(psm + S::off_w);
  float* sX = reinterpret_cast<float*>(psm + S::off_xch);
  uint64_t* full = reinterpret_cast<uint64_t*>(psm + S::off_bar);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t crank = cluster_ctarank();
  const int cluster_id = blockIdx.x / PN_CLUSTER;
  const int col0 = static_cast<int>(crank) * KQ;                     // first column owned by this CTA
  const int my_panels = (p.n_panels > cluster_id) ? (p.n_panels - cluster_id + p.n_clusters - 1) / p.n_clusters : 0;
  const int n_stages = my_panels * NCH;

  // ---- loaders --------------------------------------------------------------------------------------------------
  // stage s = (panel s / NCH of this cluster, column chunk s % NCH)
  auto issue_stage = [&](int s) {
    const int pi = s / NCH, ch = s % NCH;
    const long long row0 = static_cast<long long>(cluster_id + pi * p.n_clusters) * PN_ROWS;
    const int c0 = col0 + ch * PN_KC;
    if (IN_F32) {
      float* df = sF + (s % NST) * (PN_ROWS * PN_FW);
      const float* src = reinterpret_cast<const float*>(p.in);
      for (int i = tid; i < PN_ROWS * (PN_KC / 4); i += PN_THREADS) {
        const int r = i / (PN_KC / 4), c4 = (i % (PN_KC / 4)) * 4;
        const bool ok = row0 + r < p.M;
        cp_async16(df + r * PN_FW + c4, src + (ok ? (row0 + r) * p.ld_in + c0 + c4 : 0), ok);
      }
    } else {
      bf16* dt = sT + (s % NST) * (PN_ROWS * PN_TW);
      const bf16* src = reinterpret_cast<const bf16*>(p.in);
      for (int i = tid; i < PN_ROWS * (PN_KC / 8); i += PN_THREADS) {
        const int r = i / (PN_KC / 8), c8 = (i % (PN_KC / 8)) * 8;
        const bool ok = row0 + r < p.M;
        cp_async16(dt + r * PN_TW + c8, src + (ok ? (row0 + r) * p.ld_in + c0 + c8 : 0), ok);
      }
    }
    if (ch == 0) {                       // the panel's L rows travel with its first chunk
      bf16* dl = sL + (IN_F32 ? (pi % 3) : (s % NST)) * (PN_ROWS * LW);
      for (int i = tid; i < PN_ROWS * (R / 8); i += PN_THREADS) {
        const int r = i / (R / 8), c8 = (i % (R / 8)) * 8;
        const bool ok = row0 + r < p.M;
        cp_async16(dl + r * LW + c8, p.L + (ok ? (row0 + r) * p.ldl + c8 : 0), ok);
      }
    }
  };

  // ---- accumulators ---------------------------------------------------------------------------------------------
  float acc_red[NCH][MT + 1][4][4];      // G^T partial: [chunk][m16 tile of L^T (+ the all-ones tile)][n8 tile of this warp's 32 columns]
  float acc_proj[2][NT][4];              // projection partial: [m16 tile of this warp's 32 rows][n8 tile]
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int i = 0; i < MT + 1; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc_red[c][i][j][k] = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc_proj[i][j][k] = 0.f;
  const bool do_colsum = p.colsum != nullptr;
  const uint32_t ones_frag = (g == 0) ? 0x3F803F80u : 0u;           // bf16x2(1, 1) in row 0 of the m16 tile

  const int rh = warp & 1;               // projection: rows [32 rh, 32 rh + 32) of the panel
  const int kq = warp >> 1;              // projection: columns [64 kq, 64 kq + 64) of every chunk

  // ---- prologue: projection weights of this CTA's columns + the first stages ---------------------------------------
  for (int i = tid; i < R * (KQ / 8); i += PN_THREADS) {
    const int n = i / (KQ / 8), c8 = (i % (KQ / 8)) * 8;
    cp_async16(sW + n * WW + c8, p.W + static_cast<long long>(n) * p.ldw + col0 + c8, true);
  }
  constexpr int AHEAD = IN_F32 ? 2 : NST - 1;     // stages in flight ahead of the one being consumed
#pragma unroll
  for (int s = 0; s < AHEAD; ++s) {
    if (s < n_stages) issue_stage(s);
    cp_async_commit();
  }
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_barrier_init();
  }
  cluster_sync_all();                    // barriers initialised and every CTA resident before any distributed-shared-memory store

  // Finish a panel whose partials were sent one panel ago: wait for all 4 contributions to this CTA's 16 rows, sum, store.
  // Deferred by one panel so that nobody waits on the exchange; the two parity slots make that safe (the peers cannot send
  // panel q + 2 before they have received this CTA's panel q + 1, which it sends only after finalize(q)).
  auto finalize = [&](int q) {
    const int par = q & 1;
    mbar_wait(&full[par], static_cast<uint32_t>((q >> 1) & 1));
    const long long qrow0 = static_cast<long long>(cluster_id + q * p.n_clusters) * PN_ROWS + static_cast<long long>(crank) * PN_OWN;
    for (int i = tid; i < PN_OWN * (R / 2); i += PN_THREADS) {
      const int r = i / (R / 2), c = (i % (R / 2)) * 2;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int src = 0; src < PN_CLUSTER; ++src) {
        const float2 v = *reinterpret_cast<const float2*>(sX + ((par * PN_CLUSTER + src) * PN_OWN + r) * R + c);
        s0 += v.x; s1 += v.y;
      }
      if (qrow0 + r < p.M) *reinterpret_cast<uint32_t*>(p.out + (qrow0 + r) * p.ld_out + c) = pack_bf16x2(s0, s1);
    }
  };

  uint32_t lfrag[PN_ROWS / 16][MT][4];

  for (int pi = 0; pi < my_panels; ++pi) {
    const long long row0 = static_cast<long long>(cluster_id + pi * p.n_clusters) * PN_ROWS;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int s = pi * NCH + ch;
      const bf16* ct;
      if (IN_F32) {
        // stage s has landed in the fp32 staging buffer once at most one younger group is pending
        cp_async_wait<AHEAD - 1>();
        __syncthreads();
        const float* cf = sF + (s % NST) * (PN_ROWS * PN_FW);
        const int c0 = col0 + ch * PN_KC;
#pragma unroll 4
        for (int i = tid; i < PN_ROWS * (PN_KC / 4); i += PN_THREADS) {
          const int r = i / (PN_KC / 4), c4 = (i % (PN_KC / 4)) * 4;
          const float4 v = *reinterpret_cast<const float4*>(cf + r * PN_FW + c4);
          uint2 q;
          q.x = pack_bf16x2(v.x, v.y);
          q.y = pack_bf16x2(v.z, v.w);
          *reinterpret_cast<uint2*>(sT + r * PN_TW + c4) = q;
          if (p.copy != nullptr && row0 + r < p.M) *reinterpret_cast<uint2*>(p.copy + (row0 + r) * p.ld_copy + c0 + c4) = q;
        }
        __syncthreads();                 // bf16 tile complete, staging buffer s % NST free
        if (s + AHEAD < n_stages) issue_stage(s + AHEAD);
        cp_async_commit();
        ct = sT;
      } else {
        // stage s has landed once at most AHEAD - 1 younger groups are pending; the barrier also retires every thread's reads of
        // the slot consumed one stage ago, which is the one refilled now (one barrier per stage)
        cp_async_wait<AHEAD - 1>();
        __syncthreads();
        if (s + AHEAD < n_stages) issue_stage(s + AHEAD);
        cp_async_commit();
        ct = sT + (s % NST) * (PN_ROWS * PN_TW);
      }

      if (ch == 0) {
        // left operand of the batch reduction (L^T: [R x rows]) for the whole panel -> registers
        const bf16* cl = sL + (IN_F32 ? (pi % 3) : (s % NST)) * (PN_ROWS * LW);
        const int j = lane >> 3, i = lane & 7;
#pragma unroll
        for (int ks = 0; ks < PN_ROWS / 16; ++ks)
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const int krow = ks * 16 + (j >> 1) * 8 + i;
            const int mcol = mt * 16 + (j & 1) * 8;
            ldmatrix_x4_trans(smem_u32(cl + krow * LW + mcol), lfrag[ks][mt][0], lfrag[ks][mt][1], lfrag[ks][mt][2], lfrag[ks][mt][3]);
          }
      }

      // ---- projection: rows [32 rh, +32) x this chunk's columns [64 kq, +64) ------------------------------------------
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int kcol = kq * 64 + ks * 16;
        uint32_t a[2][4];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
          ldmatrix_x4(smem_u32(ct + (rh * 32 + mi * 16 + (lane & 15)) * PN_TW + kcol + (lane >> 4) * 8), a[mi][0], a[mi][1], a[mi][2], a[mi][3]);
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
          uint32_t b0, b1, b2, b3;
          const int nrow = np * 16 + (lane >> 4) * 8 + (lane & 7);
          ldmatrix_x4(smem_u32(sW + nrow * WW + ch * PN_KC + kcol + ((lane >> 3) & 1) * 8), b0, b1, b2, b3);
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            mma_bf16_16816(acc_proj[mi][2 * np], a[mi][0], a[mi][1], a[mi][2], a[mi][3], b0, b1);
            mma_bf16_16816(acc_proj[mi][2 * np + 1], a[mi][0], a[mi][1], a[mi][2], a[mi][3], b2, b3);
          }
        }
      }

      // ---- batch reduction: all 64 rows x this warp's 32 columns of the chunk ------------------------------------------
#pragma unroll
      for (int ks = 0; ks < PN_ROWS / 16; ++ks) {
        uint32_t bfr[4][2];
        {
          const int j = lane >> 3, i = lane & 7;
          const int krow = ks * 16 + (j & 1) * 8 + i;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int ncol = warp * 32 + h * 16 + (j >> 1) * 8;
            ldmatrix_x4_trans(smem_u32(ct + krow * PN_TW + ncol), bfr[2 * h][0], bfr[2 * h][1], bfr[2 * h + 1][0], bfr[2 * h + 1][1]);
          }
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
            mma_bf16_16816(acc_red[ch][mt][nt], lfrag[ks][mt][0], lfrag[ks][mt][1], lfrag[ks][mt][2], lfrag[ks][mt][3], bfr[nt][0], bfr[nt][1]);
        if (do_colsum) {
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc_red[ch][MT][nt], ones_frag, 0u, ones_frag, 0u, bfr[nt][0], bfr[nt][1]);
        }
      }
      if (ch == 0 && pi > 0) finalize(pi - 1);
      // no trailing barrier: the next stage's leading barrier orders these reads before the slot is refilled / reconverted
    }

    // ---- panel end: projection partials -> CTA-local sum over the 4 column slices -> reduce-scatter over the cluster -----
    // The tile slot consumed last is free until the next issue_stage (bf16 mode) or conversion (fp32 mode): it holds the partials.
    // Rows are XOR-swizzled in 8-float groups so that the float2 fragment stores of a half-warp hit 32 distinct banks.
    {
      const int s_last = pi * NCH + NCH - 1;
      float* red = reinterpret_cast<float*>(IN_F32 ? sT : sT + (s_last % NST) * (PN_ROWS * PN_TW));   // [4 kq][64 rows][R]
      __syncthreads();                   // every warp is done with the tile that `red` overlays
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int hrow = 0; hrow < 2; ++hrow) {
            const int row = rh * 32 + mi * 16 + g + hrow * 8;
            float* d = red + (kq * PN_ROWS + row) * R + ((nt * 8 + 2 * t) ^ (red_swizzle<R>(row) << 3));
            *reinterpret_cast<float2*>(d) = make_float2(acc_proj[mi][nt][2 * hrow], acc_proj[mi][nt][2 * hrow + 1]);
            acc_proj[mi][nt][2 * hrow] = 0.f;
            acc_proj[mi][nt][2 * hrow + 1] = 0.f;
          }
      __syncthreads();
      const int par = pi & 1;
      // This CTA's receive barrier for the panel: 4 sources x 16 rows x R floats.  Its previous phase (panel pi - 2) completed
      // before finalize(pi - 2) returned, and early bytes of faster peers only drive the transaction count negative.
      if (tid == 0) mbar_arrive_expect_tx(&full[par], PN_CLUSTER * PN_OWN * R * 4);
      // PN_ROWS * R / 4 float4 groups, each sent to the CTA that owns the row; the store itself signals the owner's barrier
      for (int i = tid; i < PN_ROWS * (R / 4); i += PN_THREADS) {
        const int r = i / (R / 4), c4 = ((i % (R / 4)) * 4);
        const int cs = c4 ^ (red_swizzle<R>(r) << 3);
        float4 sum = *reinterpret_cast<const float4*>(red + r * R + cs);
#pragma unroll
        for (int q = 1; q < 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(red + (q * PN_ROWS + r) * R + cs);
          sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
        }
        const uint32_t owner = static_cast<uint32_t>(r / PN_OWN);
        float* slot = sX + ((par * PN_CLUSTER + static_cast<int>(crank)) * PN_OWN + (r % PN_OWN)) * R + c4;
        st_async_v4(mapa_cluster(smem_u32(slot), owner), sum.x, sum.y, sum.z, sum.w, mapa_cluster(smem_u32(&full[par]), owner));
      }
      // `red` is next overwritten (issue_stage / conversion) only behind the leading barrier of the next stage
    }
  }
  if (my_panels > 0) finalize(my_panels - 1);
  cp_async_wait<0>();

  // ---- final flush of the batch-reduction accumulators -----------------------------------------------------------------
  if (my_panels > 0) {
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int mt = 0; mt < MT + 1; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int prow = mt * 16 + g + (e >> 1) * 8;
            const int q = col0 + ch * PN_KC + warp * 32 + nt * 8 + 2 * t + (e & 1);
            const float val = acc_red[ch][mt][nt][e] * p.scale;
            if (mt < MT) {
              atomicAdd(p.G + static_cast<long long>(prow) * p.ldg + q, val);
            } else if (do_colsum && prow == MT * 16) {
              atomicAdd(p.colsum + q, val);
            }
          }
  }
  cluster_sync_all();                    // no CTA leaves while a peer could still address its shared memory
}

template <int R, int NCH, bool IN_F32>
int launch_panel(const PanelParams& p0, cudaStream_t stream) {
  using S = PanelSmem<R, NCH, IN_F32>;
  auto kern = panel_fused_kernel<R, NCH, IN_F32>;
  static int max_clusters = 0;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PN_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(PN_THREADS);
  cfg.dynamicSmemBytes = S::total;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (max_clusters == 0) {
    DMI_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total));
    cfg.gridDim = dim3(PN_CLUSTER * (num_sms() / PN_CLUSTER));
    int n = 0;
    DMI_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    if (n < 1) {
      set_error("panel_fused: no %d-CTA cluster with %d B of shared memory can be resident", PN_CLUSTER, S::total);
      return DMI_ERR_UNSUPPORTED;
    }
    max_clusters = n;
  }
  PanelParams p = p0;
  p.n_clusters = max_clusters < p.n_panels ? max_clusters : p.n_panels;
  cfg.gridDim = dim3(PN_CLUSTER * p.n_clusters);
  DMI_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  count_launch();
  return DMI_OK;
}

}  // namespace

bool panel_fused_supported(long long K, int R) { return (R == 16 || R == 32) && (K == 1024 || K == 2048); }

int panel_fused(const void* in, long long ld_in, bool in_f32, const bf16* W, long long ldw, bf16* out, long long ld_out, bf16* copy,
                long long ld_copy, const bf16* L, long long ldl, float* G, long long ldg, float* colsum, float scale, long long M,
                long long K, int R, cudaStream_t s) {
  DMI_REQUIRE(in && W && out && L && G && M > 0, "panel_fused: bad arguments");
  DMI_REQUIRE(panel_fused_supported(K, R), "panel_fused: K=%lld R=%d outside the compiled shapes (K 1024/2048, R 16/32)", K, R);
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  DMI_REQUIRE(al16(in) && al16(W) && al16(L) && (in_f32 ? ld_in % 4 == 0 : ld_in % 8 == 0) && ldw % 8 == 0 && ldl % 8 == 0 && ld_out % 2 == 0 &&
                  (reinterpret_cast<uintptr_t>(out) & 3) == 0 && (copy == nullptr || (ld_copy % 4 == 0 && (reinterpret_cast<uintptr_t>(copy) & 7) == 0)),
              "panel_fused: misaligned operands (ld_in=%lld ldw=%lld ldl=%lld ld_out=%lld)", ld_in, ldw, ldl, ld_out);
  PanelParams p;
  p.in = in; p.ld_in = ld_in; p.W = W; p.ldw = ldw; p.out = out; p.ld_out = ld_out; p.copy = in_f32 ? copy : nullptr; p.ld_copy = ld_copy;
  p.L = L; p.ldl = ldl; p.G = G; p.ldg = ldg; p.colsum = colsum; p.scale = scale;
  p.M = static_cast<int>(M); p.K = static_cast<int>(K);
  p.n_panels = static_cast<int>((M + PN_ROWS - 1) / PN_ROWS);
  p.n_clusters = 0;
  const int nch = static_cast<int>(K / (PN_CLUSTER * PN_KC));
#define DMI_PANEL(RR)                                                                              \
  case RR:                                                                                         \
    if (nch == 1) return in_f32 ? launch_panel<RR, 1, true>(p, s) : launch_panel<RR, 1, false>(p, s); \
    return in_f32 ? launch_panel<RR, 2, true>(p, s) : launch_panel<RR, 2, false>(p, s);
  switch (R) {
    DMI_PANEL(16) DMI_PANEL(32)
  }
#undef DMI_PANEL
  return DMI_ERR_UNSUPPORTED;
}

}  // namespace dmi
