// C ABI of libdmi_b200 (see include/dmi_b200.h): error plumbing, TMA descriptor creation, GEMM / outer-reduce dispatch
// and the adapted-MLP forward / backward schedules.
#include <stdarg.h>
#include <atomic>
#include <string.h>

#include "../../include/dmi_b200.h"
#include "common.cuh"
#include "elementwise.cuh"
#include "gemm_tc.cuh"
#include "gemm2_tc.cuh"
#include "outer_mma.cuh"
#include "skinny.cuh"
#include "panel.h"

namespace dmi {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};      // bumped from whichever host thread enqueues (no data race)
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      n = 0;
      return 148;
    }
  }
  return n;
}

// cuTensorMapEncodeTiled is a driver entry point; fetch it through the runtime so that the library has no link-time
// dependency on libcuda (it must load on a machine without a driver for the symbol-export test).
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* tm, const void* ptr, int kind, long long inner, long long rows, long long ld, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return DMI_ERR_CUDA;
  }
  const int esz = (kind == KIND_BF16) ? 2 : 4;
  DMI_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand base %p is not 16-byte aligned", ptr);
  DMI_REQUIRE((ld * esz) % 16 == 0, "TMA operand row stride %lld elements is not a multiple of 16 bytes", ld);
  DMI_REQUIRE(inner > 0 && rows > 0 && box_rows > 0 && box_rows <= 256, "bad TMA extents inner=%lld rows=%lld box_rows=%d", inner, rows, box_rows);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * esz};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, kind == KIND_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (ptr=%p inner=%lld rows=%lld ld=%lld box_rows=%d)", static_cast<int>(r), ptr, inner, rows, ld, box_rows);
    return DMI_ERR_CUDA;
  }
  return DMI_OK;
}

int make_tmap_2d_mn(CUtensorMap* tm, const void* ptr, long long inner, long long rows, long long ld, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return DMI_ERR_CUDA;
  }
  DMI_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * 2) % 16 == 0, "TMA (MN-major) operand misaligned: ptr=%p ld=%lld", ptr, ld);
  DMI_REQUIRE(inner > 0 && rows > 0 && box_rows > 0 && box_rows <= 256, "bad TMA extents inner=%lld rows=%lld", inner, rows);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (MN-major) failed with CUresult %d", static_cast<int>(r));
    return DMI_ERR_CUDA;
  }
  return DMI_OK;
}

// C[M,N] (+)= alpha * A[K,M]^T B[K,N]  (bf16, both operands MN-major; the weight-gradient GEMM, K = batch)
int gemm_mn(const void* A, long long lda, const void* B, long long ldb, const GemmParams& p, cudaStream_t s) {
  DMI_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0 && p.M % 8 == 0 && p.N % 8 == 0, "gemm_mn: bad extents M=%d N=%d K=%d", p.M, p.N, p.K);
  DMI_REQUIRE(p.out0 != nullptr && (reinterpret_cast<uintptr_t>(p.out0) & 15) == 0 && (p.ld0 % (p.out0_f32 ? 4 : 8)) == 0, "gemm_mn: out0 misaligned");
  CUtensorMap ta, tb;
  int rc = make_tmap_2d_mn(&ta, A, p.M, p.K, lda, 64);
  if (rc != DMI_OK) return rc;
  rc = make_tmap_2d_mn(&tb, B, p.N, p.K, ldb, 64);
  if (rc != DMI_OK) return rc;
  const long long mt = (p.M + GEMM_BM - 1) / GEMM_BM;
  const bool big = p.N > 128 && mt * ((p.N + 255) / 256) >= num_sms() / 2;
  if (big) return launch_gemm_inst<256, EPI_STORE, KIND_BF16, true>(ta, tb, p, s);
  if (p.N > 64) return launch_gemm_inst<128, EPI_STORE, KIND_BF16, true>(ta, tb, p, s);
  return launch_gemm_inst<64, EPI_STORE, KIND_BF16, true>(ta, tb, p, s);
}

static int colsum_bf16(const bf16* src, long long ld, long long rows, int cols, float* out, float scale, cudaStream_t s) {
  DMI_REQUIRE(cols % 8 == 0 && ld % 8 == 0 && rows > 0, "colsum: bad extents");
  const int cblocks = (cols + 63) / 64;
  long long nsplit = (2LL * num_sms() + cblocks - 1) / cblocks;
  const long long max_split = (rows + 31) / 32;
  if (nsplit > max_split) nsplit = max_split;
  if (nsplit < 1) nsplit = 1;
  const long long rps = (rows + nsplit - 1) / nsplit;
  nsplit = (rows + rps - 1) / rps;
  DMI_CHECK_CUDA(launch_pdl(colsum_bf16_kernel, dim3(cblocks, static_cast<unsigned>(nsplit)), dim3(256), 0, s, src, ld, rows, cols, out, scale, rps));
  DMI_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DMI_OK;
}

static int g_pair_mode = -1;         // -1 auto, 0 never, 1 always
static int g_gemm_debug = 0;
// The pair kernel's TMA-store epilogue covers the three hot forms (store to fp32 or bf16, GELU writing h and pre, GELU' reading pre);
// with an optional dropout keep mask on the two GELU forms; anything else (accumulate, addend, fp32 + bf16 dual output, ragged N) stays on
// the 1-CTA kernel.
static bool pair_epilogue_ok(const GemmParams& p, int mode) {
  auto ok = [](const void* q, long long ld, int esz) { return q != nullptr && (reinterpret_cast<uintptr_t>(q) & 15) == 0 && (ld * esz) % 16 == 0; };
  if (p.N % 128 != 0 || p.addend != nullptr || p.accumulate_out0) return false;
  if (p.keep != nullptr && ((reinterpret_cast<uintptr_t>(p.keep) & 15) != 0 || p.ld_keep % 16 != 0 || mode == EPI_STORE)) return false;
  if (p.bias != nullptr && (reinterpret_cast<uintptr_t>(p.bias) & 15) != 0) return false;
  if (!ok(p.out0, p.ld0, p.out0_f32 ? 4 : 2)) return false;
  if (mode == EPI_STORE) return p.out1 == nullptr;
  if (mode == EPI_GELU) return !p.out0_f32 && ok(p.out1, p.ld1, 2);
  if (mode == EPI_GELU_BWD) return !p.out0_f32 && ok(p.aux, p.ld_aux, 2);
  return false;
}
static bool use_pair(long long M, long long N, int mode = -1) {
  // Measured on B200 (profiles/r2_headroom1.txt, after the elected-lane issue loops): the CTA-pair kernel beats the 1-CTA kernel on
  // every epilogue of the step's shapes (204 vs 235 us store fp32, 106 vs 120 us GELU, 206 vs 257 us GELU').
  if (g_pair_mode >= 0) return g_pair_mode > 0;
  return mode >= 0 && M >= 8192 && N >= 1024;
}

template <int BN>
static int launch_gemm_bn(int kind, int mode, const void* A, long long lda, const void* B, long long ldb, const GemmParams& p, cudaStream_t s) {
  CUtensorMap ta, tb;
  int rc = make_tmap_2d(&ta, A, kind, p.K, p.M, lda, GEMM_BM);
  if (rc != DMI_OK) return rc;
  rc = make_tmap_2d(&tb, B, kind, p.K, p.N, ldb, BN);
  if (rc != DMI_OK) return rc;
  if (kind == KIND_TF32) {
    DMI_REQUIRE(mode == EPI_STORE, "tf32 GEMM supports only the store epilogue");
    return launch_gemm_inst<BN, EPI_STORE, KIND_TF32>(ta, tb, p, s);
  }
  if (BN == 256 && pair_epilogue_ok(p, mode) && use_pair(p.M, p.N, mode)) {
    // CTA-pair MMA (cta_group::2): 256x256 tile per 2-CTA cluster, each CTA stages 128 rows of A and 128 of the 256 B rows;
    // outputs (and the stashed pre-activation of the GELU' epilogue) move through TMA in [32 rows x 128 bytes] boxes
    rc = make_tmap_2d(&tb, B, kind, p.K, p.N, ldb, BN / 2);
    if (rc != DMI_OK) return rc;
    CUtensorMap to0, to1;
    rc = make_tmap_2d(&to0, p.out0, p.out0_f32 ? KIND_TF32 : KIND_BF16, p.N, p.M, p.ld0, 32);
    if (rc != DMI_OK) return rc;
    to1 = to0;
    if (mode == EPI_GELU) rc = make_tmap_2d(&to1, p.out1, KIND_BF16, p.N, p.M, p.ld1, 32);
    if (mode == EPI_GELU_BWD) rc = make_tmap_2d(&to1, p.aux, KIND_BF16, p.N, p.M, p.ld_aux, 32);
    if (rc != DMI_OK) return rc;
    switch (mode) {
      case EPI_STORE: return launch_gemm_pair<EPI_STORE>(ta, tb, to0, to1, p, s);
      case EPI_GELU: return launch_gemm_pair<EPI_GELU>(ta, tb, to0, to1, p, s);
      case EPI_GELU_BWD: return launch_gemm_pair<EPI_GELU_BWD>(ta, tb, to0, to1, p, s);
    }
  }
  switch (mode) {
    case EPI_STORE: return launch_gemm_inst<BN, EPI_STORE, KIND_BF16>(ta, tb, p, s);
    case EPI_GELU: return launch_gemm_inst<BN, EPI_GELU, KIND_BF16>(ta, tb, p, s);
    case EPI_GELU_BWD: return launch_gemm_inst<BN, EPI_GELU_BWD, KIND_BF16>(ta, tb, p, s);
  }
  set_error("unknown GEMM epilogue mode %d", mode);
  return DMI_ERR_INVALID;
}

static int pick_bn(long long M, long long N) {
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  const long long mt = (M + GEMM_BM - 1) / GEMM_BM;
  const long long tiles256 = mt * ((N + 255) / 256);
  return tiles256 >= num_sms() ? 256 : 128;
}

int gemm_tn(int kind, int mode, const void* A, long long lda, const void* B, long long ldb, const GemmParams& p, cudaStream_t s, int force_bn = 0) {
  DMI_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "GEMM with empty extent M=%d N=%d K=%d", p.M, p.N, p.K);
  DMI_REQUIRE(p.N % 8 == 0, "GEMM N=%d must be a multiple of 8", p.N);
  DMI_REQUIRE(A != nullptr && B != nullptr && p.out0 != nullptr, "GEMM null operand");
  DMI_REQUIRE((reinterpret_cast<uintptr_t>(p.out0) & 15) == 0 && (p.ld0 % (p.out0_f32 ? 4 : 8)) == 0, "GEMM out0 misaligned");
  DMI_REQUIRE(p.out1 == nullptr || ((reinterpret_cast<uintptr_t>(p.out1) & 15) == 0 && p.ld1 % 8 == 0), "GEMM out1 misaligned");
  DMI_REQUIRE(mode != EPI_GELU_BWD || (p.aux != nullptr && (reinterpret_cast<uintptr_t>(p.aux) & 15) == 0 && p.ld_aux % 8 == 0), "GEMM aux missing/misaligned");
  DMI_REQUIRE(p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0, "GEMM bias misaligned");
  const int bn = force_bn ? force_bn : pick_bn(p.M, p.N);
  GemmParams q = p;
  q.debug = g_gemm_debug;
  switch (bn) {
    case 32: return launch_gemm_bn<32>(kind, mode, A, lda, B, ldb, q, s);
    case 64: return launch_gemm_bn<64>(kind, mode, A, lda, B, ldb, q, s);
    case 128: return launch_gemm_bn<128>(kind, mode, A, lda, B, ldb, q, s);
    case 256: return launch_gemm_bn<256>(kind, mode, A, lda, B, ldb, q, s);
  }
  set_error("unsupported BN %d", bn);
  return DMI_ERR_UNSUPPORTED;
}

int outer_reduce(const bf16* L, long long ldl, const bf16* R, long long ldr, long long B, int P, int Q, float* G, long long ldg,
                 int transpose_out, float* colsum, float scale, cudaStream_t s) {
  DMI_REQUIRE(B > 0 && P > 0 && Q > 0, "outer_reduce with empty extent");
  DMI_REQUIRE(P % 8 == 0 && P <= 64, "outer_reduce: P=%d must be a multiple of 8 and <= 64", P);
  DMI_REQUIRE(Q % 8 == 0, "outer_reduce: Q=%d must be a multiple of 8", Q);
  DMI_REQUIRE(ldl % 8 == 0 && ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(L) & 15) == 0 && (reinterpret_cast<uintptr_t>(R) & 15) == 0,
              "outer_reduce: operands must be 16-byte aligned with ld %% 8 == 0");
  OuterParams p;
  p.L = L; p.ldl = ldl; p.R = R; p.ldr = ldr; p.B = static_cast<int>(B); p.P = P; p.Q = Q;
  p.G = G; p.ldg = ldg; p.transpose_out = transpose_out; p.colsum = colsum; p.scale = scale;
  const int qchunks = (Q + OUTER_QC - 1) / OUTER_QC;
  int nsplit = (2 * num_sms() + qchunks - 1) / qchunks;       // one wave of 2 resident CTAs (16 warps) per SM
  const long long max_split = (B + OUTER_KB - 1) / OUTER_KB;
  if (nsplit > max_split) nsplit = static_cast<int>(max_split);
  if (nsplit < 1) nsplit = 1;
  long long rps = (B + nsplit - 1) / nsplit;
  rps = ((rps + OUTER_KB - 1) / OUTER_KB) * OUTER_KB;
  nsplit = static_cast<int>((B + rps - 1) / rps);
  p.rows_per_split = static_cast<int>(rps);
  const int mt = (P + 15) / 16;
  const bool cs = colsum != nullptr;
  switch (mt) {
    case 1: return cs ? launch_outer_inst<1, true>(p, nsplit, s) : launch_outer_inst<1, false>(p, nsplit, s);
    case 2: return cs ? launch_outer_inst<2, true>(p, nsplit, s) : launch_outer_inst<2, false>(p, nsplit, s);
    case 3: return cs ? launch_outer_inst<3, true>(p, nsplit, s) : launch_outer_inst<3, false>(p, nsplit, s);
    case 4: return cs ? launch_outer_inst<4, true>(p, nsplit, s) : launch_outer_inst<4, false>(p, nsplit, s);
  }
  return DMI_ERR_UNSUPPORTED;
}


#define DMI_LAUNCHED()                  \
  do {                                  \
    DMI_CHECK_CUDA(cudaGetLastError()); \
    count_launch();                     \
  } while (0)

// Rank-r side passes of large batches run on the tcgen05 panel kernels (panel_tc.cu, panel_tc32.cu): the projections v = h A1 /
// u = x A0, the fp32 dY pass (bf16 copy + dv + dB1 + dbeta1 in one sweep) and the dpre pass (du + dB0 + dbeta0 in one sweep).
// Measured against the mma.sync kernels they replace (profiles/r2_panel_modes2.txt, 32768 rows): 28.7 vs 34.4 us, 85.8 vs 102.7 us,
// 33.3 vs 63 us.  Below PANEL_TC_MIN_ROWS (few panels per cluster) and for shapes the panel kernels are not compiled for, the
// mma.sync row-panel kernels (skinny.cuh, outer_mma.cuh) do the same work in separate passes.
// dmi_set_option("fused_panel", v): -1 = auto (by batch size), 0 = always the separate mma.sync passes, 1 = panel kernels at any size.
static int g_fused_panel = -1;
int g_pdl = 1;          // common.cuh: programmatic dependent launch of the small kernels
constexpr long long PANEL_TC_MIN_ROWS = 8192;
static bool use_panel_tc(long long rows) { return g_fused_panel < 0 ? rows >= PANEL_TC_MIN_ROWS : g_fused_panel > 0; }

// out[M,R] = in[M,K] W[R,K]^T  (R = rank); in_f32: fp32 input converted on the fly, bf16 copy written to `copy`
static int skinny_rows(const void* in, long long ld_in, bool in_f32, const bf16* W, long long ldw, bf16* out, long long ld_out, bf16* copy,
                       long long ld_copy, long long M, long long K, int R, cudaStream_t s) {
  DMI_REQUIRE(in && W && out && M > 0 && K > 0, "skinny_rows: bad arguments");
  DMI_REQUIRE(K % 8 == 0 && ldw % 8 == 0 && ld_out % 2 == 0 && (in_f32 ? (ld_in % 4 == 0 && (copy == nullptr || ld_copy % 4 == 0)) : ld_in % 8 == 0),
              "skinny_rows: misaligned operands (K=%lld ld_in=%lld)", K, ld_in);
  SkinnyParams p;
  p.in = in; p.ld_in = ld_in; p.W = W; p.ldw = ldw; p.out = out; p.ld_out = ld_out; p.copy = copy; p.ld_copy = ld_copy;
  p.M = static_cast<int>(M); p.K = static_cast<int>(K); p.R = R;
  const bool aligned16 = (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                         (!in_f32 || copy == nullptr || ((reinterpret_cast<uintptr_t>(copy) & 15) == 0 && ld_copy % 8 == 0));
  if (M <= TINY_MAX_ROWS && aligned16) return launch_tiny_rows(p, in_f32, s);
  switch (R) {
    case 8: return in_f32 ? launch_skinny_inst<8, true>(p, s) : launch_skinny_inst<8, false, 128, 2>(p, s);
    case 16: return in_f32 ? launch_skinny_inst<16, true>(p, s) : launch_skinny_inst<16, false, 128, 2>(p, s);
    case 32:
      return in_f32 ? launch_skinny_inst<32, true>(p, s) : launch_skinny_inst<32, false, 128, 2>(p, s);
    case 64: return in_f32 ? launch_skinny_inst<64, true>(p, s) : launch_skinny_inst<64, false, 128, 2>(p, s);
  }
  set_error("skinny_rows: rank %d unsupported", R);
  return DMI_ERR_UNSUPPORTED;
}

static GemmParams gp(long long M, long long N, long long K) {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
  p.alpha = 1.0f;
  return p;
}

static int check_mlp(const dmi_mlp_args* a, bool bwd) {
  DMI_REQUIRE(a != nullptr, "null dmi_mlp_args");
  DMI_REQUIRE(a->B > 0 && a->D > 0 && a->H > 0, "adapted_mlp: bad extents B=%lld D=%lld H=%lld", (long long)a->B, (long long)a->D, (long long)a->H);
  DMI_REQUIRE(a->D % 8 == 0 && a->H % 8 == 0, "adapted_mlp: D and H must be multiples of 8");
  const bool adapter = !(a->flags & DMI_MLP_NO_ADAPTER);
  DMI_REQUIRE(adapter ? (a->r == 8 || a->r == 16 || a->r == 32 || a->r == 64) : a->r == 0,
              "adapted_mlp: rank %lld unsupported (8/16/32/64, or 0 with DMI_MLP_NO_ADAPTER)", (long long)a->r);
  DMI_REQUIRE(a->w1ext && a->bias0 && a->xext && a->pre, "adapted_mlp: missing layer-0 buffers");
  if (!(a->flags & DMI_MLP_STOP_AFTER_FIRST_ACT)) DMI_REQUIRE(a->w2ext && a->bias1 && a->hext, "adapted_mlp: missing layer-1 buffers");
  if (adapter) DMI_REQUIRE(a->a0t != nullptr, "adapted_mlp: missing A0^T");
  if (!bwd && !(a->flags & DMI_MLP_X_PREPACKED)) DMI_REQUIRE(a->x != nullptr && a->ldx % 4 == 0, "adapted_mlp: missing x");
  if (bwd) DMI_REQUIRE(a->dy != nullptr && a->dpre != nullptr, "adapted_mlp_bwd: missing dy/dpre");
  if (a->flags & DMI_MLP_DROPOUT) DMI_REQUIRE(a->keep != nullptr && a->dropout_p >= 0.f && a->dropout_p < 1.f && a->H % 16 == 0, "adapted_mlp: dropout needs a keep mask, 0<=p<1, H %% 16 == 0");
  if (a->flags & DMI_MLP_BASE_GRADS) DMI_REQUIRE(!adapter, "adapted_mlp: BASE_GRADS is implemented for the plain / merged projector (DMI_MLP_NO_ADAPTER)");
  return DMI_OK;
}


int adapted_mlp_fwd(const dmi_mlp_args* a, cudaStream_t s) {
  int rc = check_mlp(a, false);
  if (rc != DMI_OK) return rc;
  const long long B = a->B, D = a->D, H = a->H, r = a->r;
  const bool adapter = !(a->flags & DMI_MLP_NO_ADAPTER);
  const bool h1 = a->flags & DMI_MLP_STOP_AFTER_FIRST_ACT;
  const long long KX = D + r, KH = H + r;
  bf16* xext = static_cast<bf16*>(a->xext);
  bf16* hext = static_cast<bf16*>(a->hext);
  // 1+2. x -> bf16 columns [0,D) of xext and u = x A0 -> columns [D, D+r), in ONE pass over the fp32 input
  const bool panels = use_panel_tc(B);
  if (adapter) {
    if (!(a->flags & DMI_MLP_X_PREPACKED)) rc = skinny_rows(a->x, a->ldx, true, static_cast<const bf16*>(a->a0t), D, xext + D, KX, xext, KX, B, D, static_cast<int>(r), s);
    else if (panels && panel_tc_mode_supported(D, static_cast<int>(r)))
      rc = panel_tc_project(xext, KX, static_cast<const bf16*>(a->a0t), D, xext + D, KX, B, D, static_cast<int>(r), s);
    else rc = skinny_rows(xext, KX, false, static_cast<const bf16*>(a->a0t), D, xext + D, KX, nullptr, 0, B, D, static_cast<int>(r), s);
    if (rc != DMI_OK) return rc;
  } else if (!(a->flags & DMI_MLP_X_PREPACKED)) {
    DMI_CHECK_CUDA(launch_pdl(cvt_rows_f32_bf16_kernel, dim3(ew_grid(B * (D / 8), 256)), dim3(256), 0, s, a->x, a->ldx, xext, KX, B, static_cast<int>(D), 1.0f));
    DMI_LAUNCHED();
  }
  // 3. pre = [x|u] [W1|B0^T]^T + (b1+beta0);  h = gelu(pre)
  {
    GemmParams p = gp(B, H, KX);
    p.bias = a->bias0;
    p.out1 = static_cast<bf16*>(a->pre); p.ld1 = H;
    if (a->flags & DMI_MLP_DROPOUT) { p.keep = a->keep; p.ld_keep = H; p.keep_scale = 1.0f / (1.0f - a->dropout_p); }
    if (h1) {
      // reference-as-written: the projector output IS h
      if (a->y != nullptr) { p.out0 = a->y; p.ld0 = a->ldy; p.out0_f32 = 1; }
      else { p.out0 = a->y_bf16; p.ld0 = a->ldy_bf16; p.out0_f32 = 0; }
      DMI_REQUIRE(p.out0 != nullptr, "adapted_mlp_fwd: no output buffer");
    } else {
      p.out0 = hext; p.ld0 = KH; p.out0_f32 = 0;
    }
    rc = gemm_tn(KIND_BF16, EPI_GELU, xext, KX, a->w1ext, KX, p, s);
    if (rc != DMI_OK) return rc;
    if (h1 && a->y != nullptr && a->y_bf16 != nullptr) {
      DMI_CHECK_CUDA(launch_pdl(cvt_rows_f32_bf16_kernel, dim3(ew_grid(B * (H / 8), 256)), dim3(256), 0, s, a->y, a->ldy, static_cast<bf16*>(a->y_bf16), a->ldy_bf16, B, static_cast<int>(H), 1.0f));
      DMI_LAUNCHED();
    }
  }
  if (h1) return DMI_OK;
  // 4. v = h A1 -> columns [H, H+r) of hext
  if (adapter) {
    DMI_REQUIRE(a->a1t != nullptr, "adapted_mlp_fwd: missing A1^T");
    if (panels && panel_tc_mode_supported(H, static_cast<int>(r)))
      rc = panel_tc_project(hext, KH, static_cast<const bf16*>(a->a1t), H, hext + H, KH, B, H, static_cast<int>(r), s);
    else
      rc = skinny_rows(hext, KH, false, static_cast<const bf16*>(a->a1t), H, hext + H, KH, nullptr, 0, B, H, static_cast<int>(r), s);
    if (rc != DMI_OK) return rc;
  }
  // 5. y = [h|v] [W2|B1^T]^T + (b2+beta1)
  {
    GemmParams p = gp(B, H, KH);
    p.bias = a->bias1;
    if (a->y != nullptr) {
      p.out0 = a->y; p.ld0 = a->ldy; p.out0_f32 = 1;
      p.out1 = static_cast<bf16*>(a->y_bf16); p.ld1 = a->ldy_bf16;
    } else {
      DMI_REQUIRE(a->y_bf16 != nullptr, "adapted_mlp_fwd: no output buffer");
      p.out0 = a->y_bf16; p.ld0 = a->ldy_bf16; p.out0_f32 = 0;
    }
    rc = gemm_tn(KIND_BF16, EPI_STORE, hext, KH, a->w2ext, KH, p, s);
    if (rc != DMI_OK) return rc;
  }
  return DMI_OK;
}

int adapted_mlp_bwd(const dmi_mlp_args* a, cudaStream_t s) {
  int rc = check_mlp(a, true);
  if (rc != DMI_OK) return rc;
  const long long B = a->B, D = a->D, H = a->H, r = a->r;
  const bool adapter = !(a->flags & DMI_MLP_NO_ADAPTER);
  const bool h1 = a->flags & DMI_MLP_STOP_AFTER_FIRST_ACT;
  const long long KX = D + r, KH = H + r;
  if (!adapter) {
    // ---- plain / merged MLP2: gradients to W1,b1,W2,b2 (train_projector.py:67, few-shot fine-tune train_hypernet.py:229-251) ----
    DMI_REQUIRE((a->flags & DMI_MLP_BASE_GRADS) && !(a->flags & DMI_MLP_STOP_AFTER_FIRST_ACT), "adapted_mlp_bwd: NO_ADAPTER needs BASE_GRADS on the full MLP2");
    DMI_REQUIRE(a->dyext && a->w2text && a->dW1 && a->db1 && a->dW2 && a->db2 && a->hext, "adapted_mlp_bwd: missing base-gradient buffers");
    const bf16* xb = static_cast<const bf16*>(a->xext);
    const bf16* hb = static_cast<const bf16*>(a->hext);
    bf16* dyb = static_cast<bf16*>(a->dyext);
    bf16* dpb = static_cast<bf16*>(a->dpre);
    const float g = a->grad_scale;
    DMI_CHECK_CUDA(launch_pdl(cvt_rows_f32_bf16_kernel, dim3(ew_grid(B * (H / 8), 256)), dim3(256), 0, s, a->dy, a->lddy, dyb, H, B, static_cast<int>(H), 1.0f));
    DMI_LAUNCHED();
    rc = colsum_bf16(dyb, H, B, static_cast<int>(H), a->db2, g, s);                       // db2 += 1^T dY
    if (rc != DMI_OK) return rc;
    {                                                                                     // dW2 += dY^T h
      GemmParams p = gp(H, H, B);
      p.alpha = g; p.out0 = a->dW2; p.ld0 = H; p.out0_f32 = 1; p.accumulate_out0 = 1;
      rc = gemm_mn(dyb, H, hb, H, p, s);
      if (rc != DMI_OK) return rc;
    }
    if (a->ev_layer1_grads != nullptr) DMI_CHECK_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(a->ev_layer1_grads), s));
    {                                                                                     // dpre = (dY W2) * gelu'(pre) [* keep/(1-p)]
      GemmParams p = gp(B, H, H);
      p.out0 = dpb; p.ld0 = H; p.out0_f32 = 0;
      p.aux = static_cast<const bf16*>(a->pre); p.ld_aux = H;
      if (a->flags & DMI_MLP_DROPOUT) { p.keep = a->keep; p.ld_keep = H; p.keep_scale = 1.0f / (1.0f - a->dropout_p); }
      rc = gemm_tn(KIND_BF16, EPI_GELU_BWD, dyb, H, a->w2text, H, p, s);
      if (rc != DMI_OK) return rc;
    }
    rc = colsum_bf16(dpb, H, B, static_cast<int>(H), a->db1, g, s);                       // db1 += 1^T dpre
    if (rc != DMI_OK) return rc;
    {                                                                                     // dW1 += dpre^T x
      GemmParams p = gp(H, D, B);
      p.alpha = g; p.out0 = a->dW1; p.ld0 = D; p.out0_f32 = 1; p.accumulate_out0 = 1;
      rc = gemm_mn(dpb, H, xb, D, p, s);
      if (rc != DMI_OK) return rc;
    }
    return DMI_OK;
  }
  const bf16* xext = static_cast<const bf16*>(a->xext);
  const bf16* hext = static_cast<const bf16*>(a->hext);
  bf16* dyext = static_cast<bf16*>(a->dyext);
  bf16* dpre = static_cast<bf16*>(a->dpre);
  bf16* du = static_cast<bf16*>(a->du);
  const float gs = a->grad_scale;
  if (h1) {
    // dpre = dy * gelu'(pre)
    DMI_CHECK_CUDA(launch_pdl(gelu_bwd_rows_kernel, dim3(ew_grid(B * (H / 8), 256)), dim3(256), 0, s, a->dy, a->lddy, static_cast<const bf16*>(a->pre), H, dpre, H, B, static_cast<int>(H)));
    DMI_LAUNCHED();
  } else {
    DMI_REQUIRE(dyext && a->w2text && a->b1 && a->dA1 && a->dB1, "adapted_mlp_bwd: missing layer-1 buffers");
    // 1+2+3a. dy -> bf16 columns [0,H) of dyext, dv = dy B1^T -> columns [H,H+r), dB1 += v^T dy, dbeta1 += 1^T dy
    if (use_panel_tc(B) && panel_fused_tc_supported(H, static_cast<int>(r))) {
      // ONE tcgen05 sweep over the fp32 gradient (panel_tc32.cu)
      rc = panel_fused_tc32(a->dy, a->lddy, static_cast<const bf16*>(a->b1), H, dyext + H, KH, dyext, KH, hext + H, KH, a->dB1, H, a->dbeta1, gs, B, H,
                            static_cast<int>(r), s);
      if (rc != DMI_OK) return rc;
    } else {
      rc = skinny_rows(a->dy, a->lddy, true, static_cast<const bf16*>(a->b1), H, dyext + H, KH, dyext, KH, B, H, static_cast<int>(r), s);
      if (rc != DMI_OK) return rc;
      rc = outer_reduce(hext + H, KH, dyext, KH, B, static_cast<int>(r), static_cast<int>(H), a->dB1, H, 0, a->dbeta1, gs, s);
      if (rc != DMI_OK) return rc;
    }
    // 3b. dA1^T += dv^T h
    rc = outer_reduce(dyext + H, KH, hext, KH, B, static_cast<int>(r), static_cast<int>(H), a->dA1, r, 1, nullptr, gs, s);
    if (rc != DMI_OK) return rc;
    if (a->ev_layer1_grads != nullptr) DMI_CHECK_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(a->ev_layer1_grads), s));
    // 4. dpre = ([dy|dv] [W2^T|A1]^T) * gelu'(pre)
    {
      GemmParams p = gp(B, H, KH);
      p.out0 = dpre; p.ld0 = H; p.out0_f32 = 0;
      p.aux = static_cast<const bf16*>(a->pre); p.ld_aux = H;
      rc = gemm_tn(KIND_BF16, EPI_GELU_BWD, dyext, KH, a->w2text, KH, p, s);
      if (rc != DMI_OK) return rc;
    }
  }
  DMI_REQUIRE(du && a->b0 && a->dA0 && a->dB0, "adapted_mlp_bwd: missing layer-0 buffers");
  // 5+6a. du = dpre B0^T, dB0 += u^T dpre, dbeta0 += 1^T dpre
  if (use_panel_tc(B) && panel_fused_tc_supported(H, static_cast<int>(r)) && (reinterpret_cast<uintptr_t>(du) & 15) == 0) {
    // ONE tcgen05 sweep over dpre (panel_tc.cu; the column sum rides in the batch-reduction MMAs)
    rc = panel_fused_tc(dpre, H, static_cast<const bf16*>(a->b0), H, du, r, xext + D, KX, a->dB0, H, a->dbeta0, gs, B, H, static_cast<int>(r), s);
    if (rc != DMI_OK) return rc;
  } else {
    rc = skinny_rows(dpre, H, false, static_cast<const bf16*>(a->b0), H, du, r, nullptr, 0, B, H, static_cast<int>(r), s);
    if (rc != DMI_OK) return rc;
    rc = outer_reduce(xext + D, KX, dpre, H, B, static_cast<int>(r), static_cast<int>(H), a->dB0, H, 0, a->dbeta0, gs, s);
    if (rc != DMI_OK) return rc;
  }
  // 6b. dA0^T += du^T x
  rc = outer_reduce(du, r, xext, KX, B, static_cast<int>(r), static_cast<int>(D), a->dA0, r, 1, nullptr, gs, s);
  return rc;
}

}  // namespace dmi

// =================================================================================================
// extern "C" surface
// =================================================================================================
using namespace dmi;

extern "C" {

int dmi_version(void) { return 100; }
const char* dmi_last_error(void) { return g_err; }
int dmi_num_sms(void) { return num_sms(); }
int64_t dmi_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int dmi_set_option(const char* name, int value) {
  if (name != nullptr && strcmp(name, "gemm_pair") == 0) { g_pair_mode = value; return DMI_OK; }
  if (name != nullptr && strcmp(name, "gemm_debug") == 0) { g_gemm_debug = value; return DMI_OK; }
  if (name != nullptr && strcmp(name, "fused_panel") == 0) { g_fused_panel = value; return DMI_OK; }
  if (name != nullptr && strcmp(name, "pdl") == 0) { g_pdl = value != 0; return DMI_OK; }
  set_error("dmi_set_option: unknown option %s", name ? name : "(null)");
  return DMI_ERR_INVALID;
}

int dmi_gemm_mn(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N, int64_t K, float alpha, float* out,
                int64_t ldo, int accumulate, void* stream) {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
  p.alpha = alpha; p.out0 = out; p.ld0 = ldo; p.out0_f32 = 1; p.accumulate_out0 = accumulate;
  return gemm_mn(A, lda, B, ldb, p, static_cast<cudaStream_t>(stream));
}

int dmi_gemm_tn(int kind, int mode, const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                float alpha, const float* bias, void* out0, int64_t ld0, int out0_is_f32, void* out1_bf16, int64_t ld1,
                const void* aux_bf16, int64_t ld_aux, void* stream) {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
  p.alpha = alpha; p.bias = bias; p.out0 = out0; p.ld0 = ld0; p.out0_f32 = out0_is_f32;
  p.out1 = static_cast<bf16*>(out1_bf16); p.ld1 = ld1; p.aux = static_cast<const bf16*>(aux_bf16); p.ld_aux = ld_aux;
  DMI_REQUIRE(kind == KIND_BF16 || kind == KIND_TF32, "dmi_gemm_tn: unknown kind %d", kind);
  return gemm_tn(kind, mode, A, lda, B, ldb, p, static_cast<cudaStream_t>(stream));
}

int dmi_skinny_rows(const void* in, int64_t ld_in, int in_is_f32, const void* W, int64_t ldw, void* out, int64_t ld_out, void* copy, int64_t ld_copy,
                    int64_t M, int64_t K, int64_t R, void* stream) {
  return skinny_rows(in, ld_in, in_is_f32 != 0, static_cast<const bf16*>(W), ldw, static_cast<bf16*>(out), ld_out, static_cast<bf16*>(copy), ld_copy, M, K,
                     static_cast<int>(R), static_cast<cudaStream_t>(stream));
}

int dmi_panel_fused_tc(const void* in, int64_t ld_in, const void* W, int64_t ldw, void* out, int64_t ld_out, const void* L, int64_t ldl, float* G,
                       int64_t ldg, float* colsum, float scale, int64_t M, int64_t K, int64_t R, void* stream) {
  return panel_fused_tc(static_cast<const bf16*>(in), ld_in, static_cast<const bf16*>(W), ldw, static_cast<bf16*>(out), ld_out,
                        static_cast<const bf16*>(L), ldl, G, ldg, colsum, scale, M, K, static_cast<int>(R), static_cast<cudaStream_t>(stream));
}

int dmi_panel_tc_project(const void* in, int64_t ld_in, const void* W, int64_t ldw, void* out, int64_t ld_out, int64_t M, int64_t K, int64_t R,
                         void* stream) {
  return panel_tc_project(static_cast<const bf16*>(in), ld_in, static_cast<const bf16*>(W), ldw, static_cast<bf16*>(out), ld_out, M, K,
                          static_cast<int>(R), static_cast<cudaStream_t>(stream));
}

int dmi_panel_fused_tc32(const float* in, int64_t ld_in, const void* W, int64_t ldw, void* out, int64_t ld_out, void* copy, int64_t ld_copy,
                         const void* L, int64_t ldl, float* G, int64_t ldg, float* colsum, float scale, int64_t M, int64_t K, int64_t R,
                         void* stream) {
  return panel_fused_tc32(in, ld_in, static_cast<const bf16*>(W), ldw, static_cast<bf16*>(out), ld_out, static_cast<bf16*>(copy), ld_copy,
                          static_cast<const bf16*>(L), ldl, G, ldg, colsum, scale, M, K, static_cast<int>(R), static_cast<cudaStream_t>(stream));
}

int dmi_outer_reduce(const void* L, int64_t ldl, const void* R, int64_t ldr, int64_t B, int64_t P, int64_t Q, float* G, int64_t ldg,
                     int transpose_out, float* colsum, float scale, void* stream) {
  return outer_reduce(static_cast<const bf16*>(L), ldl, static_cast<const bf16*>(R), ldr, B, static_cast<int>(P), static_cast<int>(Q), G, ldg,
                      transpose_out, colsum, scale, static_cast<cudaStream_t>(stream));
}

int dmi_projector_pack_base(const float* W1, int64_t ldw1, const float* W2, int64_t D, int64_t H, int64_t r, void* w1ext, void* w2ext,
                            void* w2text, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DMI_REQUIRE(W1 && w1ext && D % 8 == 0 && H % 8 == 0 && r % 8 == 0 && ldw1 % 4 == 0, "projector_pack_base: bad arguments");
  DMI_CHECK_CUDA(launch_pdl(cvt_rows_f32_bf16_kernel, dim3(ew_grid(H * (D / 8), 256)), dim3(256), 0, s, W1, ldw1, static_cast<bf16*>(w1ext), D + r, H, static_cast<int>(D), 1.0f));
  DMI_LAUNCHED();
  if (W2 != nullptr && w2ext != nullptr) {
    DMI_CHECK_CUDA(launch_pdl(cvt_rows_f32_bf16_kernel, dim3(ew_grid(H * (H / 8), 256)), dim3(256), 0, s, W2, H, static_cast<bf16*>(w2ext), H + r, H, static_cast<int>(H), 1.0f));
    DMI_LAUNCHED();
  }
  if (W2 != nullptr && w2text != nullptr) {
    dim3 grid((H + 31) / 32, (H + 31) / 32), block(32, 8);
    DMI_CHECK_CUDA(launch_pdl(transpose_f32_bf16_kernel, dim3(grid), dim3(block), 0, s, W2, H, static_cast<bf16*>(w2text), H + r, static_cast<int>(H), static_cast<int>(H), 1.0f));
    DMI_LAUNCHED();
  }
  return DMI_OK;
}

int dmi_adapter_pack(const float* A0, const float* B0, const float* beta0, const float* A1, const float* B1, const float* beta1,
                     const float* b1, const float* b2, int64_t D, int64_t H, int64_t r, float scale, void* w1ext, void* w2ext, void* w2text,
                     void* a0t, void* a1t, void* b0, void* b1_bf16, float* bias0, float* bias1, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DMI_REQUIRE(b1 && bias0 && D % 8 == 0 && H % 8 == 0, "adapter_pack: missing base bias / bad extents");
  AdapterPackParams p;
  memset(&p, 0, sizeof(p));
  p.D = static_cast<int>(D); p.H = static_cast<int>(H); p.r = static_cast<int>(r); p.scale = scale;
  p.b1 = b1; p.b2 = b2; p.beta0 = beta0; p.beta1 = beta1; p.bias0 = bias0; p.bias1 = (b2 != nullptr) ? bias1 : nullptr;
  if (A0 != nullptr) {
    DMI_REQUIRE(B0 && w1ext && a0t && b0, "adapter_pack: missing layer-0 arguments");
    DMI_REQUIRE(r == 8 || r == 16 || r == 32 || r == 64, "adapter_pack: rank %lld unsupported", (long long)r);
    p.A0 = A0; p.B0 = B0; p.w1ext = static_cast<bf16*>(w1ext); p.a0t = static_cast<bf16*>(a0t); p.b0 = static_cast<bf16*>(b0);
  }
  if (A1 != nullptr) {
    DMI_REQUIRE(A0 && B1 && b2 && w2ext && w2text && a1t && b1_bf16 && bias1, "adapter_pack: missing layer-1 arguments");
    p.A1 = A1; p.B1 = B1; p.w2ext = static_cast<bf16*>(w2ext); p.w2text = static_cast<bf16*>(w2text);
    p.a1t = static_cast<bf16*>(a1t); p.b1bf = static_cast<bf16*>(b1_bf16);
  }
  const long long total = adapter_pack_items(p);
  DMI_CHECK_CUDA(launch_pdl(adapter_pack_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, s, p));
  DMI_LAUNCHED();
  return DMI_OK;
}

int dmi_merge_adapter(const float* W, int64_t ldw, const float* bias, const float* A, const float* B, const float* beta, int64_t in_dim,
                      int64_t H, int64_t r, float scale, float* W_out, int64_t ldwo, float* bias_out, void* stream) {
  DMI_REQUIRE(W && bias && A && B && W_out && bias_out, "merge_adapter: null argument");
  DMI_REQUIRE(in_dim > 0 && H > 0 && r > 0 && r <= 64, "merge_adapter: bad extents in=%lld H=%lld r=%lld (r <= 64)", (long long)in_dim, (long long)H, (long long)r);
  dim3 grid(static_cast<unsigned>((in_dim + 31) / 32), static_cast<unsigned>((H + 31) / 32));
  DMI_CHECK_CUDA(launch_pdl(merge_adapter_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), W, ldw, bias, A, B, beta, static_cast<int>(in_dim), static_cast<int>(H),
                                                                            static_cast<int>(r), scale, W_out, ldwo, bias_out));
  DMI_LAUNCHED();
  return DMI_OK;
}

int dmi_adapted_mlp_fwd(const dmi_mlp_args* args, void* stream) { return adapted_mlp_fwd(args, static_cast<cudaStream_t>(stream)); }
int dmi_adapted_mlp_bwd(const dmi_mlp_args* args, void* stream) { return adapted_mlp_bwd(args, static_cast<cudaStream_t>(stream)); }

}  // extern "C"
