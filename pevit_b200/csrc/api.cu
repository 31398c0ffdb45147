// C ABI (include/pevit_b200.h) and the per-block forward / backward schedules.
#include "../../include/pevit_b200.h"

#include "common.cuh"
#include "kernels.h"

using namespace pevit;

namespace {

inline cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }

// Bump allocator over a caller-provided buffer (256-byte aligned pieces).
struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<uint8_t*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct Saved {
  bf16* xn1; float *mean1, *rstd1; bf16* qkv_hm; bf16* T; bf16* o_tok; float* lse;
  float* x1; float *mean2, *rstd2; bf16* z;
  // bottleneck
  float* m; float *mean_a, *rstd_a; bf16* a_n; bf16* zd; bf16* u;
  size_t bytes;
};

struct Work {
  // forward
  bf16* xn2; bf16* h;
  // backward
  bf16* dy_bf16; bf16* dz; float* dxn; float* dx1; bf16* dx1_bf16; bf16* do_tok; bf16* dqkv; bf16* ddelta;
  bf16* dzd; float* dm; bf16* dm_bf16;
  size_t bytes;
};

inline bool has_lowrank(const pevit_block_desc& d) { return d.method == PEVIT_KADAPTATION || d.method == PEVIT_LORA; }
inline bool has_bottleneck(const pevit_block_desc& d) { return d.method == PEVIT_ADAPTER || d.method == PEVIT_COMPACTER; }

Saved carve_saved(const pevit_block_desc& d, void* p) {
  const size_t M = static_cast<size_t>(d.L) * d.NB, D = d.D;
  Carver c(p);
  Saved s{};
  s.xn1 = c.take<bf16>(M * D);
  s.mean1 = c.take<float>(M); s.rstd1 = c.take<float>(M);
  s.qkv_hm = c.take<bf16>(3 * M * D);
  s.T = c.take<bf16>(M * 2 * (d.r > 0 ? d.r : 4));
  s.o_tok = c.take<bf16>(M * D);
  s.lse = c.take<float>(M * d.H);
  s.x1 = c.take<float>(M * D);
  s.mean2 = c.take<float>(M); s.rstd2 = c.take<float>(M);
  s.z = c.take<bf16>(M * 4 * D);
  if (has_bottleneck(d)) {
    s.m = c.take<float>(M * D);
    s.mean_a = c.take<float>(M); s.rstd_a = c.take<float>(M);
    s.a_n = c.take<bf16>(M * D);
    s.zd = c.take<bf16>(M * 64);
    s.u = c.take<bf16>(M * 64);
  }
  s.bytes = (c.off + 255) & ~size_t(255);
  return s;
}

Work carve_work(const pevit_block_desc& d, void* p) {
  const size_t M = static_cast<size_t>(d.L) * d.NB, D = d.D;
  const size_t W3 = 3 * D + 2 * d.r;
  Work w{};
  Carver f(p);
  w.xn2 = f.take<bf16>(M * D);
  w.h = f.take<bf16>(M * 4 * D);
  Carver b(p);  // backward reuses the same bytes
  w.dy_bf16 = b.take<bf16>(M * D);
  w.dz = b.take<bf16>(M * 4 * D);
  w.dxn = b.take<float>(M * D);
  w.dx1 = b.take<float>(M * D);
  w.dx1_bf16 = b.take<bf16>(M * D);
  w.do_tok = b.take<bf16>(M * D);
  w.dqkv = b.take<bf16>(M * W3);
  w.ddelta = b.take<bf16>(2 * M * D);
  if (has_bottleneck(d)) {
    w.dzd = b.take<bf16>(M * 64);
    w.dm = b.take<float>(M * D);
    w.dm_bf16 = b.take<bf16>(M * D);
  }
  const size_t mx = f.off > b.off ? f.off : b.off;
  w.bytes = (mx + 255) & ~size_t(255);
  return w;
}

int check_desc(const pevit_block_desc* d) {
  PEVIT_REQUIRE(d != nullptr, "null block descriptor");
  PEVIT_REQUIRE(d->L > 0 && d->NB > 0 && d->D > 0 && d->H > 0, "bad block shape L=%d NB=%d D=%d H=%d", d->L, d->NB,
                d->D, d->H);
  PEVIT_REQUIRE(d->D == 64 * d->H, "head_dim must be 64 (D=%d, H=%d)", d->D, d->H);
  PEVIT_REQUIRE(d->D % 128 == 0, "D=%d must be a multiple of 128", d->D);
  PEVIT_REQUIRE(d->method >= PEVIT_PLAIN && d->method <= PEVIT_COMPACTER, "unknown method %d", d->method);
  if (has_lowrank(*d))
    PEVIT_REQUIRE(d->r > 0 && d->r <= 32 && (2 * d->r) % 8 == 0, "low-rank width r=%d unsupported", d->r);
  else
    PEVIT_REQUIRE(d->r == 0, "r must be 0 for method %d", d->method);
  PEVIT_REQUIRE(d->out_rows >= 0 && d->out_rows <= d->L * d->NB && d->out_rows % d->NB == 0,
                "out_rows=%d must be a multiple of NB=%d within L*NB", d->out_rows, d->NB);
  PEVIT_REQUIRE(d->causal == 0 || (d->causal == 1 && d->method == PEVIT_PLAIN && d->save == 0 && d->L <= 128),
                "causal=%d: the masked attention is forward-only (save=0), method plain, L <= 128 (L=%d)", d->causal, d->L);
  return 0;
}

int attn_fwd_dispatch(cudaStream_t s, const AttnShape& a, int impl, const bf16* q, const bf16* k, const bf16* v,
                      const bf16* T, const float* qmat, const float* bias, bf16* o_tok, float* lse) {
  // impl 0: tcgen05 kernels (L <= 128: one tile per head or two heads per tile; longer: head-resident); impl 2: the
  // round-1 pair-streaming kernels for 128 < L <= 384 (kept as a second implementation and for shapes whose
  // operands do not fit in shared memory); impl 1: CUDA-core cross-check with the delta expanded in-kernel
  if (a.causal) {  // text tower (L = 77): only the L <= 128 kernel carries the mask
    PEVIT_REQUIRE(attn_tc_supported(a), "causal attention needs L <= 128 and no in-kernel delta (L=%d r=%d)", a.L, a.r);
    return attn_fwd_tc(s, a, q, k, v, o_tok, lse);
  }
  if (impl != 1 && attn_tc_supported(a)) return attn_fwd_tc(s, a, q, k, v, o_tok, lse);
  if (impl == 0 && attn_hr_supported(a)) return attn_fwd_hr(s, a, q, k, v, o_tok, lse);
  if (impl != 1 && attn_tc_long_supported(a)) return attn_fwd_tc_long(s, a, q, k, v, o_tok, lse);
  return attn_delta_fwd_ref(s, a, q, k, v, T, qmat, bias, o_tok, lse);
}

int attn_bwd_dispatch(cudaStream_t s, const AttnShape& a, int impl, const bf16* q, const bf16* k, const bf16* v,
                      const bf16* T, const float* qmat, const float* bias, const bf16* o_tok, const bf16* do_tok,
                      const float* lse, bf16* dqkv, int ld, bf16* ddelta) {
  PEVIT_REQUIRE(!a.causal, "causal attention is forward-only (the text tower is frozen)");
  if (impl != 1 && attn_tc_supported(a)) return attn_bwd_tc(s, a, q, k, v, do_tok, lse, dqkv, ld, ddelta);
  if (impl == 0 && attn_bwd_hr_supported(a)) return attn_bwd_hr(s, a, q, k, v, o_tok, do_tok, lse, dqkv, ld, ddelta);
  if (impl != 1 && attn_tc_long_supported(a)) return attn_bwd_tc_long(s, a, q, k, v, o_tok, do_tok, lse, dqkv, ld, ddelta);
  return attn_delta_bwd_ref(s, a, q, k, v, T, qmat, bias, o_tok, do_tok, lse, dqkv, ld, ddelta);
}

#define TRY(expr)            \
  do {                       \
    int rc__ = (expr);       \
    if (rc__ != 0) return rc__; \
  } while (0)

}  // namespace

extern "C" {

int pevit_abi_version(void) { return PEVIT_ABI_VERSION; }
const char* pevit_last_error(void) { return last_error(); }

int pevit_check_device(void) {
  int dev = 0;
  PEVIT_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  PEVIT_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  PEVIT_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  PEVIT_REQUIRE(major == 10, "pevit_b200 kernels are built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
  return 0;
}

int pevit_gemm_tn(const pevit_gemm_args* a, void* stream) {
  PEVIT_REQUIRE(a != nullptr, "null gemm args");
  GemmEpilogue ep;
  ep.bias = a->bias; ep.resid = a->resid; ep.out_f32 = a->out_f32;
  ep.out_bf16 = static_cast<bf16*>(a->out_bf16); ep.out2_bf16 = static_cast<bf16*>(a->out2_bf16);
  ep.aux_bf16 = static_cast<const bf16*>(a->aux_bf16); ep.ld_out = a->ld_out;
  ep.qkv_hm = static_cast<bf16*>(a->qkv_hm); ep.t_out = static_cast<bf16*>(a->t_out);
  ep.resid_bf16 = static_cast<const bf16*>(a->resid_bf16);
  ep.debug = a->force_bn >= 100000 ? a->force_bn / 100000 : 0;  // diagnostics (tools/gemm_bench.py)
  const int force_bn = a->force_bn >= 100000 ? a->force_bn % 100000 : a->force_bn;
  ep.L = a->L; ep.NB = a->NB; ep.H = a->H; ep.D = a->D; ep.r2 = a->r2;
  int epi = a->epilogue;
  // the public enum folds the activation kind into the epilogue id
  if (epi == PEVIT_EPI_QGELU) { epi = EPI_ACT; ep.act = ACT_QUICKGELU; }
  else if (epi == PEVIT_EPI_DQGELU) { epi = EPI_DACT; ep.act = ACT_QUICKGELU; }
  return gemm_tn(as_stream(stream), static_cast<const bf16*>(a->a), a->lda, static_cast<const bf16*>(a->b), a->ldb,
                 a->m, a->n, a->k, epi, ep, force_bn);
}

int pevit_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32, float* mean,
                        float* rstd, int32_t rows, int32_t d, void* stream) {
  return layernorm_fwd(as_stream(stream), x, gamma, beta, static_cast<bf16*>(y_bf16), y_f32, mean, rstd, rows, d);
}

int pevit_layernorm_bwd(const float* dyn, const float* x, const float* gamma, const float* mean, const float* rstd,
                        const float* dres, float* dx, void* dx_bf16, float* dgamma, float* dbeta, int32_t rows,
                        int32_t d, void* stream) {
  return layernorm_bwd(as_stream(stream), dyn, x, gamma, mean, rstd, dres, dx, static_cast<bf16*>(dx_bf16), dgamma,
                       dbeta, rows, d);
}

int pevit_attn_fwd(const pevit_attn_args* a, void* stream) {
  PEVIT_REQUIRE(a != nullptr, "null attention args");
  AttnShape sh{a->L, a->NB, a->H, a->D, a->r, a->alpha, a->causal};
  return attn_fwd_dispatch(as_stream(stream), sh, a->impl, static_cast<const bf16*>(a->q),
                           static_cast<const bf16*>(a->k), static_cast<const bf16*>(a->v),
                           static_cast<const bf16*>(a->t), a->qmat, a->delta_bias, static_cast<bf16*>(a->o_tok), a->lse);
}

int pevit_attn_bwd(const pevit_attn_args* a, void* stream) {
  PEVIT_REQUIRE(a != nullptr, "null attention args");
  AttnShape sh{a->L, a->NB, a->H, a->D, a->r, a->alpha, a->causal};
  return attn_bwd_dispatch(as_stream(stream), sh, a->impl, static_cast<const bf16*>(a->q),
                           static_cast<const bf16*>(a->k), static_cast<const bf16*>(a->v),
                           static_cast<const bf16*>(a->t), a->qmat, a->delta_bias,
                           static_cast<const bf16*>(a->o_tok), static_cast<const bf16*>(a->do_tok),
                           a->lse, static_cast<bf16*>(a->dqkv), a->ld_dqkv, static_cast<bf16*>(a->ddelta));
}

int pevit_kad_expand(const float* u1, const float* v1, const float* u2, const float* v2, const float* s, const float* t,
                     int32_t d, float alpha, void* w_ext, void* w_ext_t, float* qmat, void* qmat_t, void* delta_w,
                     void* stream) {
  return kad_expand(as_stream(stream), u1, v1, u2, v2, s, t, d, alpha, static_cast<bf16*>(w_ext),
                    static_cast<bf16*>(w_ext_t), qmat, static_cast<bf16*>(qmat_t), static_cast<bf16*>(delta_w));
}

int pevit_lora_expand(const float* aq, const float* av, const float* bq, const float* bv, int32_t d, int32_t r,
                      float alpha, void* w_ext, void* w_ext_t, float* qmat, void* qmat_t, void* delta_w,
                      void* stream) {
  return lora_expand(as_stream(stream), aq, av, bq, bv, d, r, alpha, static_cast<bf16*>(w_ext),
                     static_cast<bf16*>(w_ext_t), qmat, static_cast<bf16*>(qmat_t), static_cast<bf16*>(delta_w));
}

int pevit_atb_accumulate(const void* a, int32_t a_is_bf16, int32_t lda, const void* b, int32_t b_is_bf16, int32_t ldb,
                         int32_t m, int32_t kc, int32_t nc, float scale, float* c, void* stream) {
  return atb_accumulate(as_stream(stream), a, a_is_bf16, lda, b, b_is_bf16, ldb, m, kc, nc, scale, c);
}

int pevit_atb_tc(const void* a, int32_t lda, const void* b, int32_t ldb, int32_t nb_cols, int32_t m, int32_t kc,
                 int32_t n_lo, int32_t n_cnt, float scale, float* c, int32_t ldc, void* stream) {
  return atb_tc(as_stream(stream), static_cast<const bf16*>(a), lda, static_cast<const bf16*>(b), ldb, nb_cols, m, kc,
                n_lo, n_cnt, scale, c, ldc);
}

int pevit_colsum_bf16(const void* x, int32_t m, int32_t d, float* out, void* stream) {
  return colsum_bf16(as_stream(stream), static_cast<const bf16*>(x), nullptr, d, m, d, out);
}

int pevit_kad_factor_grads(const float* dP, const float* dQ, const float* u1, const float* v1, const float* u2,
                           const float* v2, const float* s, const float* t, int32_t d, float* du1, float* dv1,
                           float* du2, float* dv2, float* ds, float* dt, void* stream) {
  return kad_factor_grads(as_stream(stream), dP, dQ, u1, v1, u2, v2, s, t, d, du1, dv1, du2, dv2, ds, dt, false);
}

int pevit_kad_factor_grads_acc(const float* dP, const float* dQ, const float* u1, const float* v1, const float* u2,
                               const float* v2, const float* s, const float* t, int32_t d, float* du1, float* dv1,
                               float* du2, float* dv2, float* ds, float* dt, void* stream) {
  return kad_factor_grads(as_stream(stream), dP, dQ, u1, v1, u2, v2, s, t, d, du1, dv1, du2, dv2, ds, dt, true);
}

int pevit_kad_factor_grads_acc_batch(int32_t count, const float* const* dP, const float* const* dQ, const float* const* s,
                                     const float* const* t, float* const* ds, float* const* dt, const float* u1,
                                     const float* v1, const float* u2, const float* v2, int32_t d, float* du1, float* dv1,
                                     float* du2, float* dv2, void* stream) {
  PEVIT_REQUIRE(dP && dQ && s && t && ds && dt && u1 && v1 && u2 && v2 && du1 && dv1 && du2 && dv2,
                "pevit_kad_factor_grads_acc_batch: null pointer");
  return kad_factor_grads_batch(as_stream(stream), count, dP, dQ, s, t, ds, dt, u1, v1, u2, v2, d, du1, dv1, du2, dv2);
}

int pevit_phm_expand(const float* rule, int32_t n, const float* down_left, const float* down_right, const float* up_left,
                     const float* up_right, int32_t d, int32_t bottleneck, void* w_down, void* w_down_t, void* w_up,
                     void* w_up_t, void* stream) {
  PEVIT_REQUIRE(rule && down_left && down_right && up_left && up_right && w_down && w_down_t && w_up && w_up_t,
                "pevit_phm_expand: null pointer");
  return phm_expand(as_stream(stream), rule, n, down_left, down_right, up_left, up_right, d, bottleneck,
                    static_cast<bf16*>(w_down), static_cast<bf16*>(w_down_t), static_cast<bf16*>(w_up),
                    static_cast<bf16*>(w_up_t));
}

int pevit_phm_factor_grads(const float* d_w_down, const float* d_w_up, const float* rule, int32_t n, const float* down_left,
                           const float* down_right, const float* up_left, const float* up_right, int32_t d,
                           int32_t bottleneck, float* d_rule, float* d_down_left, float* d_down_right, float* d_up_left,
                           float* d_up_right, int32_t accumulate, void* stream) {
  PEVIT_REQUIRE(d_w_down && d_w_up && rule && down_left && down_right && up_left && up_right && d_down_left &&
                    d_down_right && d_up_left && d_up_right, "pevit_phm_factor_grads: null pointer");
  return phm_factor_grads(as_stream(stream), d_w_down, d_w_up, rule, n, down_left, down_right, up_left, up_right, d,
                          bottleneck, d_rule, d_down_left, d_down_right, d_up_left, d_up_right, accumulate != 0);
}

int pevit_bottleneck_pack(const float* w_down, const float* w_up, int32_t d, int32_t bottleneck, void* w_down_bf16,
                          void* w_down_t, void* w_up_bf16, void* w_up_t, void* stream) {
  PEVIT_REQUIRE(w_down && w_up && w_down_bf16 && w_down_t && w_up_bf16 && w_up_t, "pevit_bottleneck_pack: null pointer");
  return bottleneck_pack(as_stream(stream), w_down, w_up, d, bottleneck, static_cast<bf16*>(w_down_bf16),
                         static_cast<bf16*>(w_down_t), static_cast<bf16*>(w_up_bf16), static_cast<bf16*>(w_up_t));
}

int pevit_cast_bf16(const float* src, void* dst, size_t n, void* stream) {
  return cast_f32_to_bf16(as_stream(stream), src, static_cast<bf16*>(dst), n);
}

int pevit_transpose_bf16(const float* src, int32_t rows, int32_t cols, void* dst, int32_t ldd, void* stream) {
  return transpose_f32_to_bf16(as_stream(stream), src, rows, cols, static_cast<bf16*>(dst), ldd);
}

int pevit_head_ce_fwd(const float* feat, const float* w, const float* b, const int64_t* labels, int32_t n, int32_t e,
                      int32_t c, float* logits, float* dlogits, float* loss, void* stream) {
  PEVIT_REQUIRE(feat && w && labels && logits && dlogits && loss, "pevit_head_ce_fwd: null pointer");
  return head_ce_fwd(as_stream(stream), feat, w, b, reinterpret_cast<const long long*>(labels), n, e, c, logits, dlogits, loss);
}

int pevit_head_ce_bwd(const float* dlogits, const float* feat, const float* w, const float* gscale, int32_t n, int32_t e,
                      int32_t c, void* dfeat_bf16, float* dw, float* db, int32_t accumulate, void* stream) {
  PEVIT_REQUIRE(dlogits && feat && w, "pevit_head_ce_bwd: null pointer");
  return head_ce_bwd(as_stream(stream), dlogits, feat, w, gscale, n, e, c, static_cast<bf16*>(dfeat_bf16), dw, db, accumulate);
}

int pevit_sgd_momentum(float* p, const float* g, float* m, size_t n, float lr, float momentum, float weight_decay,
                       float grad_scale, void* stream) {
  return sgd_momentum(as_stream(stream), p, g, m, n, lr, momentum, weight_decay, grad_scale);
}

int pevit_prof_enable(int32_t on) { return prof_enable(on); }
int pevit_prof_reset(void) { return prof_reset(); }
int pevit_prof_num_classes(void) { return PC_COUNT; }
const char* pevit_prof_class_name(int32_t cls) { return prof_class_name(cls); }
int pevit_prof_read(double* ms, int64_t* launches, int32_t n) {
  static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");
  return prof_read(ms, reinterpret_cast<long long*>(launches), n);
}
int64_t pevit_launch_count(void) { return launch_count(); }

size_t pevit_peer_buffer_bytes(size_t n_floats) { return peer_buffer_bytes(n_floats); }
int pevit_peer_alloc(size_t n_floats, void** ptr, void* ipc_handle) { return peer_alloc(n_floats, ptr, ipc_handle); }
int pevit_peer_open(const void* ipc_handle, void** ptr) { return peer_open(ipc_handle, ptr); }
int pevit_peer_close(void* ptr) { return peer_close(ptr); }
int pevit_peer_free(void* ptr) { return peer_free(ptr); }
int pevit_peer_status(const void* own, size_t n_floats, int32_t* timed_out, void* stream) {
  int t = 0;
  int rc = peer_status(as_stream(stream), own, n_floats, &t);
  if (rc == 0) *timed_out = t;
  return rc;
}
int pevit_allreduce_sgd(void* const* peers, int32_t world, int32_t rank, size_t n, size_t n_decayed, float* p, float* m,
                        float lr, float momentum, float weight_decay, float gscale, void* stream) {
  return allreduce_sgd(as_stream(stream), peers, world, rank, n, n_decayed, p, m, lr, momentum, weight_decay, gscale);
}

size_t pevit_patch_embed_workspace_bytes(int32_t nb, int32_t resolution, int32_t patch, int32_t d) {
  return patch_embed_workspace_bytes(nb, resolution, patch, d);
}

int pevit_patch_embed(const float* images, const void* w_patch, const float* cls, const float* pos, const float* ln_g,
                      const float* ln_b, float* x, void* workspace, int32_t nb, int32_t resolution, int32_t patch,
                      int32_t d, int32_t pos_rows, void* stream) {
  PEVIT_REQUIRE(images && w_patch && cls && pos && ln_g && ln_b && x && workspace, "pevit_patch_embed: null pointer");
  PEVIT_REQUIRE(patch > 0 && resolution % patch == 0 && pos_rows == (resolution / patch) * (resolution / patch) + 1,
                "pevit_patch_embed: positional embedding has %d rows, resolution %d / patch %d needs %d", pos_rows,
                resolution, patch, patch > 0 ? (resolution / patch) * (resolution / patch) + 1 : 0);
  return patch_embed(as_stream(stream), images, PEVIT_PX_F32, nullptr, nullptr, static_cast<const bf16*>(w_patch), cls,
                     pos, ln_g, ln_b, x, workspace, nb, resolution, patch, d);
}

int pevit_patch_embed_px(const void* images, int32_t px_dtype, const float* mean, const float* std_, const void* w_patch,
                         const float* cls, const float* pos, const float* ln_g, const float* ln_b, float* x,
                         void* workspace, int32_t nb, int32_t resolution, int32_t patch, int32_t d, int32_t pos_rows,
                         void* stream) {
  PEVIT_REQUIRE(images && w_patch && cls && pos && ln_g && ln_b && x && workspace, "pevit_patch_embed_px: null pointer");
  PEVIT_REQUIRE(patch > 0 && resolution % patch == 0 && pos_rows == (resolution / patch) * (resolution / patch) + 1,
                "pevit_patch_embed_px: positional embedding has %d rows, resolution %d / patch %d needs %d", pos_rows,
                resolution, patch, patch > 0 ? (resolution / patch) * (resolution / patch) + 1 : 0);
  return patch_embed(as_stream(stream), images, px_dtype, mean, std_, static_cast<const bf16*>(w_patch), cls, pos, ln_g,
                     ln_b, x, workspace, nb, resolution, patch, d);
}

size_t pevit_block_saved_bytes(const pevit_block_desc* desc) {
  if (check_desc(desc) != 0) return 0;
  return carve_saved(*desc, nullptr).bytes;
}

size_t pevit_block_workspace_bytes(const pevit_block_desc* desc) {
  if (check_desc(desc) != 0) return 0;
  return carve_work(*desc, nullptr).bytes;
}

// Forward of one ResidualAttentionBlock:
//   x1 = x + out_proj(attn(ln_1(x)))            model.py:973
//   y  = x1 + mlp(ln_2(x1)) [+ bottleneck]      model.py:974, adapter_model.py:333, compacter_model.py:500
int pevit_block_fwd(const pevit_block_desc* desc, const pevit_block_weights* w, const float* x, float* y, void* saved,
                    void* workspace, void* stream) {
  TRY(check_desc(desc));
  PEVIT_REQUIRE(w && x && y && saved && workspace, "pevit_block_fwd: null pointer argument");
  const pevit_block_desc& d = *desc;
  cudaStream_t s = as_stream(stream);
  const int M = d.L * d.NB, D = d.D, r2 = 2 * d.r, W3 = 3 * D + r2;
  const int Mo = d.out_rows > 0 ? d.out_rows : M;  // rows whose output is needed (everything after attention)
  Saved sv = carve_saved(d, saved);
  Work wk = carve_work(d, workspace);
  const size_t plane = static_cast<size_t>(M) * D;

  // ln_1 -> bf16 A operand
  TRY(layernorm_fwd(s, x, w->ln1_g, w->ln1_b, sv.xn1, nullptr, sv.mean1, sv.rstd1, M, D));
  // in-projection (+ low-rank T columns), head split, q scale
  {
    GemmEpilogue ep;
    ep.bias = w->b_qkv; ep.qkv_hm = sv.qkv_hm; ep.t_out = sv.T;
    ep.L = d.L; ep.NB = d.NB; ep.H = d.H; ep.D = D; ep.r2 = r2;
    prof_set_tag(PC_GEMM_QKV);
    TRY(gemm_tn(s, sv.xn1, D, static_cast<const bf16*>(w->w_qkv_ext), D, M, W3, D, EPI_QKV, ep));
  }
  const bool fused_delta = has_lowrank(d) && d.attn_impl == 1;  // cross-check kernel expands the delta itself
  if (has_lowrank(d) && !fused_delta) {
    // q' = q + scr(alpha T_q Q_q^T + b), v' = v + scr(alpha T_v Q_v^T + b): F4 makes scr() the identity on the
    // flat head-major buffer viewed as [M][D], so the delta is a rank-2r GEMM accumulated in place (model.py:796-799)
    const bf16* dw = static_cast<const bf16*>(w->delta_w);
    if (M % 256 == 0) {
      // both deltas in ONE launch: batch 0 -> q plane, batch 1 -> v plane (two planes further down the same tensor);
      // T is read once per tile pair, delta_w[which] are consecutive row blocks of one [2D][2r] matrix
      GemmEpilogue ep;
      ep.bias = d.method == PEVIT_KADAPTATION ? w->delta_bias : nullptr;
      ep.out_bf16 = sv.qkv_hm; ep.resid_bf16 = sv.qkv_hm; ep.ld_out = D;
      ep.batch = 2; ep.b_batch_rows = D; ep.c_batch_rows = 2 * M;
      prof_set_tag(PC_GEMM_DELTA);
      TRY(gemm_tn(s, sv.T, r2, dw, r2, M, D, r2, EPI_BF16, ep));
    } else {
      for (int which = 0; which < 2; ++which) {
        bf16* dst = sv.qkv_hm + (which == 0 ? 0 : 2) * plane;
        GemmEpilogue ep;
        ep.bias = d.method == PEVIT_KADAPTATION ? w->delta_bias : nullptr;
        ep.out_bf16 = dst; ep.resid_bf16 = dst; ep.ld_out = D;
        prof_set_tag(PC_GEMM_DELTA);
        TRY(gemm_tn(s, sv.T, r2, dw + static_cast<size_t>(which) * D * r2, r2, M, D, r2, EPI_BF16, ep));
      }
    }
  }
  // attention core
  {
    AttnShape a{d.L, d.NB, d.H, D, fused_delta ? d.r : 0, d.alpha, d.causal};
    TRY(attn_fwd_dispatch(s, a, d.attn_impl, sv.qkv_hm, sv.qkv_hm + plane, sv.qkv_hm + 2 * plane,
                          fused_delta ? sv.T : nullptr, fused_delta ? w->qmat : nullptr,
                          fused_delta && d.method == PEVIT_KADAPTATION ? w->delta_bias : nullptr, sv.o_tok, sv.lse));
  }
  // out-projection + residual
  {
    GemmEpilogue ep;
    ep.bias = w->b_o; ep.resid = x; ep.out_f32 = sv.x1; ep.ld_out = D;
    prof_set_tag(PC_GEMM_OUT);
    TRY(gemm_tn(s, sv.o_tok, D, static_cast<const bf16*>(w->w_o), D, Mo, D, D, EPI_F32, ep));
  }
  // ln_2, c_fc + QuickGELU
  TRY(layernorm_fwd(s, sv.x1, w->ln2_g, w->ln2_b, wk.xn2, nullptr, sv.mean2, sv.rstd2, Mo, D));
  {
    GemmEpilogue ep;
    ep.bias = w->b_fc; ep.out_bf16 = wk.h; ep.out2_bf16 = d.save ? sv.z : nullptr; ep.ld_out = 4 * D;
    ep.act = ACT_QUICKGELU;
    prof_set_tag(PC_GEMM_FC);
    TRY(gemm_tn(s, wk.xn2, D, static_cast<const bf16*>(w->w_fc), D, Mo, 4 * D, D, EPI_ACT, ep));
  }
  if (!has_bottleneck(d)) {
    GemmEpilogue ep;
    ep.bias = w->b_proj; ep.resid = sv.x1; ep.out_f32 = y; ep.ld_out = D;
    prof_set_tag(PC_GEMM_PROJ);
    TRY(gemm_tn(s, wk.h, 4 * D, static_cast<const bf16*>(w->w_proj), 4 * D, Mo, D, 4 * D, EPI_F32, ep));
    return 0;
  }
  // bottleneck: y = x1 + m + up(act(down(LN_a(m)))), m = mlp output  (adapter_model.py:264-282, F8)
  {
    GemmEpilogue ep;
    ep.bias = w->b_proj; ep.out_f32 = sv.m; ep.ld_out = D;
    prof_set_tag(PC_GEMM_PROJ);
    TRY(gemm_tn(s, wk.h, 4 * D, static_cast<const bf16*>(w->w_proj), 4 * D, Mo, D, 4 * D, EPI_F32, ep));
  }
  TRY(layernorm_fwd(s, sv.m, w->lna_g, w->lna_b, sv.a_n, nullptr, sv.mean_a, sv.rstd_a, Mo, D));
  {
    GemmEpilogue ep;
    ep.bias = w->b_down; ep.out_bf16 = sv.u; ep.out2_bf16 = sv.zd; ep.ld_out = 64;
    ep.act = d.method == PEVIT_ADAPTER ? ACT_RELU : ACT_GELU_NEW;
    prof_set_tag(PC_GEMM_BOTTLENECK);
    TRY(gemm_tn(s, sv.a_n, D, static_cast<const bf16*>(w->w_down), D, Mo, 64, D, EPI_ACT, ep));
  }
  {
    GemmEpilogue ep;
    ep.bias = w->b_up; ep.resid = sv.x1; ep.resid2 = sv.m; ep.out_f32 = y; ep.ld_out = D;
    prof_set_tag(PC_GEMM_BOTTLENECK);
    TRY(gemm_tn(s, sv.u, 64, static_cast<const bf16*>(w->w_up), 64, Mo, D, 64, EPI_F32, ep));
  }
  return 0;
}

// Backward: activation gradients flow through every frozen GEMM (dgrad only);
// weight gradients exist only for the PEFT tensors.
int pevit_block_bwd(const pevit_block_desc* desc, const pevit_block_weights* w, const float* x, const float* dy,
                    const void* dy_bf16, float* dx, void* dx_bf16, const pevit_block_grads* g, const void* saved,
                    void* workspace, void* stream) {
  TRY(check_desc(desc));
  PEVIT_REQUIRE(w && x && dy && g && saved && workspace && (dx || !desc->need_dx),
                "pevit_block_bwd: null pointer argument");
  const pevit_block_desc& d = *desc;
  cudaStream_t s = as_stream(stream);
  const int M = d.L * d.NB, D = d.D, r = d.r, r2 = 2 * r, W3 = 3 * D + r2;
  const int Mo = d.out_rows > 0 ? d.out_rows : M;  // dy holds Mo rows; the rows beyond carry no gradient
  Saved sv = carve_saved(d, const_cast<void*>(saved));
  Work wk = carve_work(d, workspace);
  const size_t plane = static_cast<size_t>(M) * D;
  bf16* dxn16 = reinterpret_cast<bf16*>(wk.dxn);  // dgrad GEMM -> LayerNorm backward hand-over (bf16 view of dxn)

  // bf16 copy of dy (A operand of the first dgrad GEMM): handed over by the block above when it produced one
  const bf16* dyb = static_cast<const bf16*>(dy_bf16);
  if (dyb == nullptr) {
    TRY(cast_f32_to_bf16(s, dy, wk.dy_bf16, static_cast<size_t>(Mo) * D));
    dyb = wk.dy_bf16;
  }
  const bf16* dmlp_bf16 = dyb;  // gradient w.r.t. the MLP output m
  if (has_bottleneck(d)) {
    const int act = d.method == PEVIT_ADAPTER ? ACT_RELU : ACT_GELU_NEW;
    // up projection: dW_up = dy^T u, db_up = colsum(dy), du = dy W_up, dzd = du * act'(zd)
    // (the bias gradient's column sum reads the same dy rows: it rides in the A^T B launch as extra CTAs)
    {
      const AtbColsum cs{dyb, nullptr, D, Mo, D, g->d_b_up};
      const bool ride = g->d_w_up && g->d_b_up && atb_colsum_rider_supported(D, D);
      const AtbProblem q{dyb, D, sv.u, 64, 64, 0, 64, 1.f, g->d_w_up, 64};
      if (g->d_w_up) TRY(atb_tc_batch(s, &q, 1, Mo, D, ride ? &cs : nullptr));
      if (g->d_b_up && !ride) TRY(colsum_bf16(s, dyb, nullptr, D, Mo, D, g->d_b_up));
    }
    {
      GemmEpilogue ep;
      ep.out_bf16 = wk.dzd; ep.aux_bf16 = sv.zd; ep.ld_out = 64; ep.act = act;
      prof_set_tag(PC_GEMM_BOTTLENECK);
      TRY(gemm_tn(s, dyb, D, static_cast<const bf16*>(w->w_up_t), D, Mo, 64, D, EPI_DACT, ep));
    }
    // down projection: dW_down^T = a_n^T dzd, db_down = colsum(dzd), da_n = dzd W_down
    {
      const AtbColsum cs{wk.dzd, nullptr, 64, Mo, 64, g->d_b_down};
      const bool ride = g->d_w_down && g->d_b_down && atb_colsum_rider_supported(64, 64);
      const AtbProblem q{sv.a_n, D, wk.dzd, 64, 64, 0, 64, 1.f, g->d_w_down, 64};
      if (g->d_w_down) TRY(atb_tc_batch(s, &q, 1, Mo, D, ride ? &cs : nullptr));
      if (g->d_b_down && !ride) TRY(colsum_bf16(s, wk.dzd, nullptr, 64, Mo, 64, g->d_b_down));
    }
    {
      GemmEpilogue ep;
      ep.out_f32 = wk.dxn; ep.ld_out = D;
      prof_set_tag(PC_GEMM_BOTTLENECK);
      TRY(gemm_tn(s, wk.dzd, 64, static_cast<const bf16*>(w->w_down_t), 64, Mo, D, 64, EPI_F32, ep));
    }
    // adapter LayerNorm backward (+ the direct `+ m` path: dres = dy), with its affine grads
    TRY(layernorm_bwd(s, wk.dxn, sv.m, w->lna_g, sv.mean_a, sv.rstd_a, dy, nullptr, wk.dm_bf16, g->d_lna_g, g->d_lna_b,
                      Mo, D));
    dmlp_bf16 = wk.dm_bf16;
  }
  // c_proj dgrad fused with QuickGELU': dz = (dm W_proj) * g'(z)
  {
    GemmEpilogue ep;
    ep.out_bf16 = wk.dz; ep.aux_bf16 = sv.z; ep.ld_out = 4 * D; ep.act = ACT_QUICKGELU;
    prof_set_tag(PC_GEMM_DPROJ);
    TRY(gemm_tn(s, dmlp_bf16, D, static_cast<const bf16*>(w->w_proj_t), D, Mo, 4 * D, D, EPI_DACT, ep));
  }
  // c_fc dgrad -> d ln_2 output
  {
    GemmEpilogue ep;
    ep.out_bf16 = dxn16; ep.ld_out = D;  // handed to the LayerNorm backward in bf16 (fp32 accumulate, one rounding)
    prof_set_tag(PC_GEMM_DFC);
    TRY(gemm_tn(s, wk.dz, 4 * D, static_cast<const bf16*>(w->w_fc_t), 4 * D, Mo, D, 4 * D, EPI_BF16, ep));
  }
  // ln_2 backward + residual path
  TRY(layernorm_bwd(s, nullptr, sv.x1, w->ln2_g, sv.mean2, sv.rstd2, dy, wk.dx1, wk.dx1_bf16, nullptr, nullptr, Mo, D, -1,
                    dxn16));
  // out-proj dgrad -> dO (token rows)
  {
    GemmEpilogue ep;
    ep.out_bf16 = wk.do_tok; ep.ld_out = D;
    prof_set_tag(PC_GEMM_DOUT);
    TRY(gemm_tn(s, wk.dx1_bf16, D, static_cast<const bf16*>(w->w_o_t), D, Mo, D, D, EPI_BF16, ep));
    // query rows beyond out_rows received no gradient: their dO is exactly zero (they still act as keys)
    if (Mo < M)
      PEVIT_CHECK_CUDA(cudaMemsetAsync(wk.do_tok + static_cast<size_t>(Mo) * D, 0,
                                       static_cast<size_t>(M - Mo) * D * sizeof(bf16), s));
  }
  // attention backward
  {
    const bool fused_delta = has_lowrank(d) && d.attn_impl == 1;
    AttnShape a{d.L, d.NB, d.H, D, fused_delta ? r : 0, d.alpha};
    TRY(attn_bwd_dispatch(s, a, d.attn_impl, sv.qkv_hm, sv.qkv_hm + plane, sv.qkv_hm + 2 * plane,
                          fused_delta ? sv.T : nullptr, fused_delta ? w->qmat : nullptr,
                          fused_delta && d.method == PEVIT_KADAPTATION ? w->delta_bias : nullptr, sv.o_tok, wk.do_tok,
                          sv.lse, wk.dqkv, W3, has_lowrank(d) ? wk.ddelta : nullptr));
  }
  if (has_lowrank(d)) {
    const bf16* qmat_t = static_cast<const bf16*>(w->qmat_t);
    {
      // dT = alpha * dDelta * Q  -> bf16 straight into the extra K columns of the QKV dgrad operand; q and v in one
      // launch: batch b reads plane b of d(delta) ([M][D] in LND rows, F4) and rows [b r, +r) of qmat_t, and writes
      // columns 3D + b r of dqkv
      GemmEpilogue ep;
      ep.out_bf16 = wk.dqkv + 3 * D; ep.ld_out = W3;
      ep.batch = 2; ep.a_batch_rows = M; ep.b_batch_rows = r; ep.c_batch_elems = r;
      prof_set_tag(PC_GEMM_DT);
      TRY(gemm_tn(s, wk.ddelta, D, qmat_t, D, M, r, D, EPI_BF16, ep));
    }
    // the three weight-gradient shaped products of the layer in one launch:
    //   dQ_q = alpha dDelta_q^T T_q,  dQ_v = alpha dDelta_v^T T_v,  dP = X^T dT  ([D][2r], q | v)
    {
      AtbProblem probs[3];
      int n = 0;
      if (g->d_qmat) {
        for (int which = 0; which < 2; ++which)
          probs[n++] = AtbProblem{wk.ddelta + which * plane, D, sv.T, r2, r2, which * r, r, d.alpha,
                                  g->d_qmat + static_cast<size_t>(which) * D * r, r};
      }
      if (g->d_pmat) probs[n++] = AtbProblem{sv.xn1, D, wk.dqkv + 3 * D, W3, r2, 0, r2, 1.f, g->d_pmat, r2};
      // KAdaptation's shared bias gradient colsum(dDelta_q) + colsum(dDelta_v) rides in the same launch
      const bool want_bias = d.method == PEVIT_KADAPTATION && g->d_bias != nullptr;
      const AtbColsum cs{wk.ddelta, wk.ddelta + plane, D, M, D, g->d_bias};
      const bool ride = want_bias && n > 0 && atb_colsum_rider_supported(D, D);
      if (n > 0) TRY(atb_tc_batch(s, probs, n, M, D, ride ? &cs : nullptr));
      if (want_bias && !ride) TRY(colsum_bf16(s, wk.ddelta, wk.ddelta + plane, D, M, D, g->d_bias));
    }
  }
  if (!d.need_dx) return 0;  // first layer: nothing upstream of this block trains
  // in-projection dgrad (K = 3D + 2r: the low-rank columns ride along)
  {
    GemmEpilogue ep;
    ep.out_bf16 = dxn16; ep.ld_out = D;
    prof_set_tag(PC_GEMM_DQKV);
    TRY(gemm_tn(s, wk.dqkv, W3, static_cast<const bf16*>(w->w_qkv_ext_t), W3, M, D, W3, EPI_BF16, ep));
  }
  // ln_1 backward + residual path
  TRY(layernorm_bwd(s, nullptr, x, w->ln1_g, sv.mean1, sv.rstd1, wk.dx1, dx, static_cast<bf16*>(dx_bf16), nullptr, nullptr,
                    M, D, Mo, dxn16));
  return 0;
}

}  // extern "C"
