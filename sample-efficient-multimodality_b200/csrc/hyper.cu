// C ABI, part 2: augmentation (a1-a3), hypernetwork forward/backward (a4-a6 and their autograd), prefix splice (a12).
#include <string.h>

#include "../../include/dmi_b200.h"
#include "common.cuh"
#include "gemm_tc.cuh"
#include <map>
#include <mutex>
#include <utility>

#include "hyper_kernels.cuh"
#include "pool_coop.cuh"

namespace dmi {

int gemm_tn(int kind, int mode, const void* A, long long lda, const void* B, long long ldb, const GemmParams& p, cudaStream_t s, int force_bn);

#define HY_LAUNCHED()                   \
  do {                                  \
    DMI_CHECK_CUDA(cudaGetLastError()); \
    count_launch();                     \
  } while (0)

static int row_prep(const RowPrepParams& p, cudaStream_t s) {
  if (p.rows <= 0) return DMI_OK;
  DMI_CHECK_CUDA(launch_pdl(row_prep_kernel, dim3((p.rows + 7) / 8), dim3(256), 0, s, p));
  HY_LAUNCHED();
  return DMI_OK;
}

struct RowPrepQueue {          // collects the row groups of one dmi_augment call; flush() = one launch
  RowPrepBatch b;
  RowPrepQueue() { memset(&b, 0, sizeof(b)); }
  void add(const RowPrepParams& p) { if (p.rows > 0) b.seg[b.n++] = p; }
  int flush(cudaStream_t s) {
    long long rows = 0;
    for (int i = 0; i < b.n; ++i) rows += b.seg[i].rows;
    if (rows == 0) return DMI_OK;
    if (b.n == 1) return row_prep(b.seg[0], s);
    DMI_CHECK_CUDA(launch_pdl(row_prep_multi_kernel, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, s, b));
    HY_LAUNCHED();
    return DMI_OK;
  }
};

template <int NV>
static int gemv_rows(const float* W, long long ldw, long long O, int D, const float* x, long long ldx, const float* bias, const float* bias_scale,
                     float out_scale, float* y, long long ldy, cudaStream_t s) {
  const long long blocks = (O + 7) / 8;
  DMI_CHECK_CUDA(launch_pdl(gemv_rows_kernel<NV>, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, s, W, ldw, static_cast<int>(O), D, x, ldx, bias, bias_scale, out_scale, y, ldy));
  HY_LAUNCHED();
  return DMI_OK;
}

// stash layout (floats): sq[NQ*D] q[NQ*D] qt[NQ*D] c[NQ*D] e[NQ*D] P[NQ*S] qb[NQ] psum[NQ]  (NQ padded to 2 for the small vectors)
struct Stash {
  float *sq, *q, *qt, *c, *e, *P, *raw, *qb, *psum;
};
static long long stash_floats(long long NQ, long long S, long long D) { return 5 * NQ * D + 2 * NQ * S + 8; }
static Stash carve_stash(float* base, long long NQ, long long S, long long D) {
  Stash st;
  st.sq = base; st.q = st.sq + NQ * D; st.qt = st.q + NQ * D; st.c = st.qt + NQ * D; st.e = st.c + NQ * D;
  st.P = st.e + NQ * D; st.raw = st.P + NQ * S; st.qb = st.raw + NQ * S; st.psum = st.qb + 4;
  return st;
}
// scratch layout for backward (floats): de[NQ*D] dc[NQ*D] dqt[NQ*D] dq[NQ*D] dpsum[4] dqb[4]
static long long scratch_floats(long long NQ, long long S, long long D) { return 4 * NQ * D + 8 + NQ * S; }

static int check_hyper(const dmi_hypernet_args* a, bool bwd, bool need_gens = true, bool need_pool = true) {
  DMI_REQUIRE(a != nullptr, "null dmi_hypernet_args");
  DMI_REQUIRE(a->NQ == 1 || a->NQ == 2, "hypernet: %lld prefix tokens unsupported (the MLP2 projector has 2; 1 or 2 are built)", (long long)a->NQ);
  DMI_REQUIRE(a->D >= 8, "hypernet: bad hypnet_dim %lld", (long long)a->D);
  if (need_gens) {
    DMI_REQUIRE(a->n_layers >= 1 && a->n_layers <= a->NQ && a->n_layers <= DMI_MAX_GEN_LAYERS, "hypernet: bad generator count %lld", (long long)a->n_layers);
    for (int l = 0; l < a->n_layers; ++l) DMI_REQUIRE(a->gen_w[l] && a->gen_b[l] && a->gen_out[l] > 0 && a->w_out[l], "hypernet: generator %d incomplete", l);
  }
  if (need_pool) {
    DMI_REQUIRE(a->S_z >= 1, "hypernet: bad extent S_z=%lld", (long long)a->S_z);
    DMI_REQUIRE(a->z && a->prefix_tokens && a->wq && a->bq && a->wk && a->bk && a->wv && a->bv && a->stash, "hypernet: null parameter / stash");
    DMI_REQUIRE(a->ldz >= a->D, "hypernet: ldz < D");
  }
  if (bwd) DMI_REQUIRE(a->scratch && a->dprefix && a->dwq && a->dbq && a->dwk && a->dbk && a->dwv && a->dbv, "hypernet_bwd: missing gradient buffers");
  return DMI_OK;
}

// w_l = out_scale * (G_l e_l + c_l) for every layer, e = [n_layers, D] modality codes (one streaming pass over the generator weights)
static int generator_grid() { return 2 * num_sms(); }

static int hypernet_generate(const dmi_hypernet_args* a, const float* e, cudaStream_t s) {
  const int D = static_cast<int>(a->D);
  for (int l = 0; l < a->n_layers; ++l) {
    const float* el = e + static_cast<long long>(l) * D;
    const bool vec = D % 4 == 0 && D <= 1024 && (reinterpret_cast<uintptr_t>(a->gen_w[l]) & 15) == 0 && (reinterpret_cast<uintptr_t>(el) & 15) == 0;
    if (!vec) {
      int rc = gemv_rows<1>(a->gen_w[l], D, a->gen_out[l], D, el, D, a->gen_b[l], nullptr, a->out_scale, a->w_out[l], 0, s);
      if (rc != DMI_OK) return rc;
      continue;
    }
    if (D <= 768) DMI_CHECK_CUDA(launch_pdl(generator_fwd_kernel<6>, dim3(generator_grid()), dim3(256), 0, s, a->gen_w[l], D, a->gen_out[l], D, el, a->gen_b[l], a->out_scale, a->w_out[l]));
    else          DMI_CHECK_CUDA(launch_pdl(generator_fwd_kernel<8>, dim3(generator_grid()), dim3(256), 0, s, a->gen_w[l], D, a->gen_out[l], D, el, a->gen_b[l], a->out_scale, a->w_out[l]));
    HY_LAUNCHED();
  }
  return DMI_OK;
}

// barrier state of the cooperative pooling kernels: {count, generation}, zeroed once and left reusable by every barrier.  One state per
// (device, stream): kernels of one stream are ordered, kernels of different streams may run side by side and must not share a counter.
static unsigned int* pool_barrier_state(cudaStream_t s) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, unsigned int*> states;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  auto it = states.find({dev, s});
  if (it != states.end()) return it->second;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
    set_error("hypernet: first use of a stream inside a CUDA graph capture (run one eager warm-up call on it before capturing)");
    return nullptr;
  }
  unsigned int* ptr = nullptr;
  if (cudaMalloc(&ptr, 2 * sizeof(unsigned int)) != cudaSuccess) return nullptr;
  if (cudaMemset(ptr, 0, 2 * sizeof(unsigned int)) != cudaSuccess) return nullptr;
  states[{dev, s}] = ptr;
  return ptr;
}

// cooperative launch of a pooling kernel: one 256-thread CTA per SM (fewer if the device cannot hold that many at once)
template <typename P>
static int launch_pool_coop(void (*kern)(P), const P& prm, long long S, cudaStream_t s) {
  const size_t smem = (S + 64 + 256) * sizeof(float);
  DMI_REQUIRE(smem <= 48 * 1024, "hypernet: support sequence of %lld tokens is too long for the pooling kernel", S);
  int per_sm = 0;
  DMI_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PC_THREADS, smem));
  DMI_REQUIRE(per_sm >= 1, "hypernet: the pooling kernel does not fit an SM");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(num_sms()); cfg.blockDim = dim3(PC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  DMI_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, prm));
  HY_LAUNCHED();
  return DMI_OK;
}

template <int NQ>
static int hypernet_fwd_t(const dmi_hypernet_args* a, cudaStream_t s, bool pool_only = false) {
  const long long S = a->NQ + a->S_z;
  const int D = static_cast<int>(a->D);
  Stash st = carve_stash(a->stash, NQ, S, D);
  // pooling: query rows, q, q~, scores, softmax (+ dropout), context, e -- one cooperative kernel (pool_coop.cuh)
  PoolCoopFwdParams f;
  memset(&f, 0, sizeof(f));
  PoolParams& pp = f.pp;
  pp.prefix = a->prefix_tokens; pp.z = a->z; pp.ldz = a->ldz; pp.pe = a->pe; pp.ldpe = a->ldpe;
  pp.NQ = NQ; pp.S = static_cast<int>(S); pp.D = D; pp.qt = st.qt; pp.qb = st.qb;
  pp.keep = a->keep; pp.keep_scale = (a->keep != nullptr) ? 1.0f / (1.0f - a->dropout_p) : 1.0f;
  pp.inv_sqrt_d = 1.0f / sqrtf(static_cast<float>(D));
  pp.P = st.raw; pp.Pout = st.P; pp.c = st.c; pp.psum = st.psum;
  f.wq = a->wq; f.bq = a->bq; f.wk = a->wk; f.bk = a->bk; f.wv = a->wv; f.bv = a->bv;
  f.sq = st.sq; f.q = st.q; f.qt = st.qt; f.qb = st.qb; f.e = st.e;
  f.bar = pool_barrier_state(s);
  if (f.bar == nullptr) { if (dmi_last_error()[0] == 0) set_error("hypernet: cannot allocate the grid-barrier state"); return DMI_ERR_CUDA; }
  int rc = launch_pool_coop(pool_fwd_coop_kernel<NQ>, f, S, s);
  if (rc != DMI_OK) return rc;
  if (pool_only) return DMI_OK;
  // 6. generators: w_l = (alpha/r) (G_l e_l + c_l), streamed once from HBM
  return hypernet_generate(a, st.e, s);
}

template <int NQ>
static int hypernet_bwd_t(const dmi_hypernet_args* a, cudaStream_t s) {
  const long long S = a->NQ + a->S_z;
  const int D = static_cast<int>(a->D);
  Stash st = carve_stash(a->stash, NQ, S, D);
  float* de = a->scratch;
  float* dc = de + NQ * D;
  float* dqt = dc + NQ * D;
  float* dq = dqt + NQ * D;
  float* dpsum = dq + NQ * D;
  float* dqb = dpsum + 4;
  float* dP = dqb + 4;
  DMI_CHECK_CUDA(cudaMemsetAsync(de, 0, sizeof(float) * scratch_floats(NQ, S, D), s));
  // generators: dG += g (x) e, dc += g, de = G^T g   with g = (alpha/r) dw
  for (int l = 0; l < a->n_layers; ++l) {
    if (a->dw[l] == nullptr) continue;               // H1: generators.1 never receives a gradient
    // dgen_w[l] == NULL: the caller keeps the rank-1 factors (dw_l, e_l) instead of a dense gradient (Rank1FactorSync); only de is needed
    const long long O = a->gen_out[l];
    const float* el = st.e + static_cast<long long>(l) * D;
    float* del = de + static_cast<long long>(l) * D;
    const int acc = a->overwrite_gen_grads ? 0 : 1;
    DMI_REQUIRE(D % 4 == 0 && D <= 1024, "hypernet_bwd: hypnet_dim %d must be a multiple of 4 and <= 1024", D);
    const size_t sm = D * sizeof(float);
    if (a->dgen_w[l] != nullptr) {
      if (D <= 768) DMI_CHECK_CUDA(launch_pdl(generator_bwd_kernel<6, true>, dim3(num_sms()), dim3(256), sm, s, a->gen_w[l], D, O, D, a->dw[l], a->out_scale, el, a->dgen_w[l], D, a->dgen_b[l], del, acc));
      else          DMI_CHECK_CUDA(launch_pdl(generator_bwd_kernel<8, true>, dim3(num_sms()), dim3(256), sm, s, a->gen_w[l], D, O, D, a->dw[l], a->out_scale, el, a->dgen_w[l], D, a->dgen_b[l], del, acc));
    } else {
      if (D <= 768) DMI_CHECK_CUDA(launch_pdl(generator_bwd_kernel<6, false>, dim3(generator_grid()), dim3(256), sm, s, a->gen_w[l], D, O, D, a->dw[l], a->out_scale, el, nullptr, D, a->dgen_b[l], del, acc));
      else          DMI_CHECK_CUDA(launch_pdl(generator_bwd_kernel<8, false>, dim3(generator_grid()), dim3(256), sm, s, a->gen_w[l], D, O, D, a->dw[l], a->out_scale, el, nullptr, D, a->dgen_b[l], del, acc));
    }
    HY_LAUNCHED();
  }
  // pooling backward (value path, softmax / scores, key path, query path): one cooperative kernel (pool_coop.cuh)
  PoolCoopBwdParams pb;
  memset(&pb, 0, sizeof(pb));
  pb.pp.prefix = a->prefix_tokens; pb.pp.z = a->z; pb.pp.ldz = a->ldz; pb.pp.pe = a->pe; pb.pp.ldpe = a->ldpe;
  pb.pp.NQ = NQ; pb.pp.S = static_cast<int>(S); pb.pp.D = D; pb.pp.qt = st.qt; pb.pp.qb = st.qb;
  pb.pp.keep = a->keep; pb.pp.keep_scale = (a->keep != nullptr) ? 1.0f / (1.0f - a->dropout_p) : 1.0f;
  pb.pp.inv_sqrt_d = 1.0f / sqrtf(static_cast<float>(D));
  pb.pp.P = st.raw; pb.pp.Pout = st.P; pb.pp.c = st.c; pb.pp.psum = st.psum;
  pb.wq = a->wq; pb.wk = a->wk; pb.bk = a->bk; pb.wv = a->wv; pb.bv = a->bv;
  pb.sq = st.sq; pb.q = st.q;
  pb.de = de; pb.dc = dc; pb.dqt = dqt; pb.dq = dq; pb.dpsum = dpsum; pb.dqb = dqb; pb.dP = dP;
  pb.dprefix = a->dprefix; pb.dwq = a->dwq; pb.dbq = a->dbq; pb.dwk = a->dwk; pb.dbk = a->dbk; pb.dwv = a->dwv; pb.dbv = a->dbv;
  pb.bar = pool_barrier_state(s);
  if (pb.bar == nullptr) { if (dmi_last_error()[0] == 0) set_error("hypernet: cannot allocate the grid-barrier state"); return DMI_ERR_CUDA; }
  return launch_pool_coop(pool_bwd_coop_kernel<NQ>, pb, S, s);
}

}  // namespace dmi

using namespace dmi;

extern "C" {

int dmi_l2_normalize(const float* x, int64_t ldx, int64_t rows, int64_t cols, float* out, int64_t ldo, void* stream) {
  DMI_REQUIRE(x && out && rows >= 0 && cols > 0, "l2_normalize: bad arguments");
  RowPrepParams p;
  memset(&p, 0, sizeof(p));
  p.src = x; p.ld_src = ldx; p.rows = static_cast<int>(rows); p.cols_in = static_cast<int>(cols); p.cols_out = p.cols_pad = static_cast<int>(cols);
  p.normalize = 1; p.dst = out; p.ld_dst = ldo;
  return row_prep(p, static_cast<cudaStream_t>(stream));
}

int64_t dmi_augment_workspace_bytes(int64_t B, int64_t K, int64_t D) {
  // split operands of the 3xTF32 rotation: A3 [(B+K), 3D] + Rt3 [D, 3D] fp32 (+ alignment slack)
  return static_cast<int64_t>(sizeof(float)) * (3 * D * (B + K) + 3 * D * D) + 256;
}

int dmi_augment(const dmi_augment_args* a, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DMI_REQUIRE(a != nullptr, "null dmi_augment_args");
  DMI_REQUIRE(a->B >= 0 && a->K >= 0 && a->D > 0 && a->Dh >= a->D && a->D_src >= a->D, "augment: bad extents B=%lld K=%lld D=%lld Dh=%lld", (long long)a->B, (long long)a->K, (long long)a->D, (long long)a->Dh);
  DMI_REQUIRE(a->perm != nullptr || a->D_src == a->D, "augment: D_src != D needs a gather index");
  const int norm = (a->flags & DMI_AUG_NORMALIZE) ? 1 : 0;
  const bool rotate = a->R != nullptr;
  const int D = static_cast<int>(a->D), Dh = static_cast<int>(a->Dh);
  float* A3 = nullptr;
  float* Rt3 = nullptr;
  if (rotate) {
    DMI_REQUIRE(D % 8 == 0, "augment: rotation needs D %% 8 == 0");
    DMI_REQUIRE(a->workspace != nullptr && a->workspace_bytes >= static_cast<uint64_t>(dmi_augment_workspace_bytes(a->B, a->K, a->D)), "augment: workspace too small");
    A3 = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(a->workspace) + 255) & ~uintptr_t(255));
    Rt3 = A3 + 3LL * D * (a->B + a->K);
    dim3 grid((D + 31) / 32, (D + 31) / 32), block(32, 8);
    DMI_CHECK_CUDA(launch_pdl(split_rotation_kernel, dim3(grid), dim3(block), 0, s, a->R, D, D, Rt3));
    HY_LAUNCHED();
  }
  int rc;
  RowPrepQueue rows;
  // batch rows
  if (a->B > 0) {
    DMI_REQUIRE(a->mm && (a->mm_out || a->mm_out_bf16), "augment: missing batch embeddings / output");
    RowPrepParams p;
    memset(&p, 0, sizeof(p));
    p.src = a->mm; p.ld_src = a->ld_mm; p.rows = static_cast<int>(a->B); p.cols_in = static_cast<int>(a->D_src);
    p.perm = a->perm; p.sign = a->sign; p.cols_out = p.cols_pad = D; p.normalize = norm;
    if (rotate) { p.dst3 = A3; p.ld_dst3 = 3LL * D; }
    else { p.dst = a->mm_out; p.ld_dst = a->ld_mm_out; p.dst_bf16 = static_cast<bf16*>(a->mm_out_bf16); p.ld_bf16 = a->ld_mm_bf16; }
    rows.add(p);
  }
  // support rows: rotated mm rows go to z[1+2k], text rows to z[2+2k], the instruction-prefix row to z[0]
  if (a->K > 0 || a->prefix != nullptr) DMI_REQUIRE(a->z != nullptr, "augment: missing z");
  if (a->K > 0) {
    DMI_REQUIRE(a->sup && a->txt, "augment: missing support embeddings");
    RowPrepParams p;
    memset(&p, 0, sizeof(p));
    p.src = a->sup; p.ld_src = a->ld_sup; p.rows = static_cast<int>(a->K); p.cols_in = static_cast<int>(a->D_src);
    p.perm = a->perm; p.sign = a->sign; p.cols_out = D; p.normalize = norm;
    if (rotate) {
      p.dst3 = A3 + 3LL * D * a->B; p.ld_dst3 = 3LL * D; p.cols_pad = D;
      if (Dh > D) {   // zero the padding columns of the support rows (train_hypernet.py:99-100)
        DMI_CHECK_CUDA(cudaMemset2DAsync(a->z + Dh + D, sizeof(float) * 2 * Dh, 0, sizeof(float) * (Dh - D), a->K, s));
      }
    } else {
      p.dst = a->z + Dh; p.ld_dst = 2LL * Dh; p.cols_pad = Dh;
    }
    rows.add(p);
    RowPrepParams t;
    memset(&t, 0, sizeof(t));
    t.src = a->txt; t.ld_src = a->ld_txt; t.rows = static_cast<int>(a->K); t.cols_in = Dh; t.cols_out = t.cols_pad = Dh; t.normalize = norm;
    t.dst = a->z + 2LL * Dh; t.ld_dst = 2LL * Dh;
    rows.add(t);
  }
  if (a->prefix != nullptr) {
    RowPrepParams t;
    memset(&t, 0, sizeof(t));
    t.src = a->prefix; t.ld_src = Dh; t.rows = 1; t.cols_in = Dh; t.cols_out = t.cols_pad = Dh; t.normalize = norm;
    t.dst = a->z; t.ld_dst = Dh;
    rows.add(t);
  }
  rc = rows.flush(s);
  if (rc != DMI_OK) return rc;
  if (rotate) {
    // x' = x R as ONE tf32 tcgen05 GEMM with K = 3D: [hi|hi|lo] . [R_hi|R_lo|R_hi]^T  (fp32-accurate).  The batch rows and the support rows
    // are consecutive in A3 and share R: with an fp32 batch output they are one launch whose epilogue sends rows >= B to the z slots.
    const long long Mall = a->B + a->K;
    const bool one_launch = a->B > 0 && a->K > 0 && a->mm_out != nullptr && a->B <= 512;      // launch-bound sizes only: the split takes the epilogue's per-row path
    GemmParams g;
    memset(&g, 0, sizeof(g));
    g.N = D; g.K = 3 * D; g.alpha = 1.0f;
    if (one_launch) {
      g.M = static_cast<int>(Mall);
      g.out0 = a->mm_out; g.ld0 = a->ld_mm_out; g.out0_f32 = 1; g.out1 = static_cast<bf16*>(a->mm_out_bf16); g.ld1 = a->ld_mm_bf16;
      g.split_row = static_cast<int>(a->B); g.out0_b = a->z + Dh; g.ld0_b = 2LL * Dh;          // interleaved slots z[1+2k]
      rc = gemm_tn(KIND_TF32, EPI_STORE, A3, 3LL * D, Rt3, 3LL * D, g, s, Mall <= 512 ? 32 : 0);     // few rows: narrow tiles spread R over more SMs
      if (rc != DMI_OK) return rc;
    } else {
      if (a->B > 0) {
        g.M = static_cast<int>(a->B);
        if (a->mm_out != nullptr) { g.out0 = a->mm_out; g.ld0 = a->ld_mm_out; g.out0_f32 = 1; g.out1 = static_cast<bf16*>(a->mm_out_bf16); g.ld1 = a->ld_mm_bf16; }
        else { g.out0 = a->mm_out_bf16; g.ld0 = a->ld_mm_bf16; g.out0_f32 = 0; }
        rc = gemm_tn(KIND_TF32, EPI_STORE, A3, 3LL * D, Rt3, 3LL * D, g, s, a->B <= 512 ? 32 : 0);
        if (rc != DMI_OK) return rc;
      }
      if (a->K > 0) {
        memset(&g, 0, sizeof(g));
        g.M = static_cast<int>(a->K); g.N = D; g.K = 3 * D; g.alpha = 1.0f;
        g.out0 = a->z + Dh; g.ld0 = 2LL * Dh; g.out0_f32 = 1;
        rc = gemm_tn(KIND_TF32, EPI_STORE, A3 + 3LL * D * a->B, 3LL * D, Rt3, 3LL * D, g, s, a->K <= 512 ? 32 : 0);
        if (rc != DMI_OK) return rc;
      }
    }
  }
  return DMI_OK;
}

int64_t dmi_hypernet_stash_floats(int64_t NQ, int64_t S_z, int64_t D) { return stash_floats(NQ, NQ + S_z, D); }
int64_t dmi_hypernet_stash_code_offset(int64_t NQ, int64_t S_z, int64_t D) { (void)S_z; return 4 * NQ * D; }
int64_t dmi_hypernet_scratch_floats(int64_t NQ, int64_t S_z, int64_t D) { return scratch_floats(NQ, NQ + S_z, D); }

int dmi_hypernet_fwd(const dmi_hypernet_args* a, void* stream) {
  int rc = check_hyper(a, false);
  if (rc != DMI_OK) return rc;
  return a->NQ == 2 ? hypernet_fwd_t<2>(a, static_cast<cudaStream_t>(stream)) : hypernet_fwd_t<1>(a, static_cast<cudaStream_t>(stream));
}

__global__ void axpy_kernel(const float* __restrict__ x, float w, float* __restrict__ y, int n) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fmaf(w, x[i], y[i]);
}

int dmi_hypernet_pool(const dmi_hypernet_args* a, float* e_accum, float weight, void* stream) {
  int rc = check_hyper(a, false, /*need_gens=*/false);
  if (rc != DMI_OK) return rc;
  DMI_REQUIRE(e_accum != nullptr, "hypernet_pool: null output");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  rc = a->NQ == 2 ? hypernet_fwd_t<2>(a, s, true) : hypernet_fwd_t<1>(a, s, true);
  if (rc != DMI_OK) return rc;
  const int n = static_cast<int>(a->NQ * a->D);
  Stash st = carve_stash(a->stash, a->NQ, a->NQ + a->S_z, a->D);
  DMI_CHECK_CUDA(launch_pdl(axpy_kernel, dim3((n + 255) / 256), dim3(256), 0, s, st.e, weight, e_accum, n));
  HY_LAUNCHED();
  return DMI_OK;
}

int dmi_hypernet_generate(const dmi_hypernet_args* a, const float* e, void* stream) {
  int rc = check_hyper(a, false, /*need_gens=*/true, /*need_pool=*/false);
  if (rc != DMI_OK) return rc;
  DMI_REQUIRE(e != nullptr, "hypernet_generate: null modality codes");
  return hypernet_generate(a, e, static_cast<cudaStream_t>(stream));
}

int dmi_hypernet_bwd(const dmi_hypernet_args* a, void* stream) {
  int rc = check_hyper(a, true);
  if (rc != DMI_OK) return rc;
  return a->NQ == 2 ? hypernet_bwd_t<2>(a, static_cast<cudaStream_t>(stream)) : hypernet_bwd_t<1>(a, static_cast<cudaStream_t>(stream));
}

int dmi_gather_rows(const void* store, int store_is_bf16, int64_t ld_store, int64_t n_store_rows, int64_t d_store, const int64_t* idx, int64_t B,
                    int64_t d_out, const int32_t* selected_features, const float* mean, int flags, float* out, int64_t ldo, void* out_bf16,
                    int64_t ldo_bf16, int* error_flag, void* stream) {
  DMI_REQUIRE(store && n_store_rows > 0 && d_store > 0 && B >= 0 && d_out > 0 && (out || out_bf16), "gather_rows: bad arguments");
  DMI_REQUIRE(selected_features != nullptr || d_out <= d_store, "gather_rows: d_out=%lld exceeds the stored width %lld", (long long)d_out, (long long)d_store);
  DMI_REQUIRE(ld_store >= d_store && (out == nullptr || ldo >= d_out) && (out_bf16 == nullptr || ldo_bf16 >= d_out), "gather_rows: leading dimension too small");
  if (B == 0) return DMI_OK;
  GatherParams p;
  memset(&p, 0, sizeof(p));
  p.store = store; p.store_is_bf16 = store_is_bf16; p.ld_store = ld_store; p.n_rows = n_store_rows;
  p.idx = reinterpret_cast<const long long*>(idx); p.B = static_cast<int>(B);
  p.sel = selected_features; p.d_store = static_cast<int>(d_store); p.mean = mean; p.d_out = static_cast<int>(d_out);
  p.normalize = (flags & DMI_AUG_NORMALIZE) ? 1 : 0;
  p.out = out; p.ldo = ldo; p.out_bf16 = static_cast<bf16*>(out_bf16); p.ldo_bf16 = ldo_bf16; p.error_flag = error_flag;
  auto al16 = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  p.vec_ok = d_out % 8 == 0 && d_out <= 1024 && ld_store % 8 == 0 && al16(store) && al16(mean) && al16(out) && al16(out_bf16) &&
             (out == nullptr || ldo % 4 == 0) && (out_bf16 == nullptr || ldo_bf16 % 8 == 0);
  DMI_CHECK_CUDA(launch_pdl(gather_rows_kernel, dim3(static_cast<unsigned>((B + 7) / 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), p));
  HY_LAUNCHED();
  return DMI_OK;
}

int dmi_splice(const float* proj_f32, const void* proj_bf16, int64_t ld_proj, const void* table, int table_is_bf16, int64_t ld_table, int64_t vocab,
               const int64_t* ids, int64_t B, int64_t T, int64_t H, void* out, int out_is_bf16, const int64_t* labels, int64_t* labels_out,
               const void* mask, int mask_is_i64, float* mask_out, int* error_flag, void* stream) {
  DMI_REQUIRE((proj_f32 != nullptr) != (proj_bf16 != nullptr), "splice: give exactly one of proj_f32 / proj_bf16");
  DMI_REQUIRE(out && B >= 0 && T >= 0 && H > 0 && (T == 0 || (table && ids)), "splice: bad arguments");
  DMI_REQUIRE((labels_out == nullptr) || (labels != nullptr || T == 0), "splice: labels_out without labels");
  DMI_REQUIRE((mask_out == nullptr) || (mask != nullptr || T == 0), "splice: mask_out without mask");
  if (B == 0) return DMI_OK;
  SpliceParams p;
  memset(&p, 0, sizeof(p));
  p.proj_f32 = proj_f32; p.proj_bf16 = static_cast<const bf16*>(proj_bf16); p.ld_proj = ld_proj;
  p.table = table; p.table_is_bf16 = table_is_bf16; p.ld_table = ld_table; p.vocab = vocab;
  p.ids = reinterpret_cast<const long long*>(ids); p.B = static_cast<int>(B); p.T = static_cast<int>(T); p.H = static_cast<int>(H);
  p.out = out; p.out_is_bf16 = out_is_bf16;
  p.labels = reinterpret_cast<const long long*>(labels); p.labels_out = reinterpret_cast<long long*>(labels_out);
  p.mask = mask; p.mask_is_i64 = mask_is_i64; p.mask_out = mask_out; p.error_flag = error_flag;
  DMI_CHECK_CUDA(launch_pdl(splice_kernel, dim3(static_cast<unsigned>(B * (1 + T))), dim3(256), 0, static_cast<cudaStream_t>(stream), p));
  HY_LAUNCHED();
  return DMI_OK;
}

}  // extern "C"
