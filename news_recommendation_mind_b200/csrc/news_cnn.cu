// CNN news encoder, C-ABI entry points and the MR_F32 (SIMT, verification) implementation.
// Reference: models/Encoders/CNN.py:30-51 with models/Embeddings/BERT.py:39 fused in front and
// models/Modules/Attention.py:5-30,56-80 fused behind.  The MR_BF16 implementation (tcgen05) lives
// in news_cnn_tc.cu and is dispatched from here.
#include "gemm_simt.cuh"
#include "pool_kernels.cuh"
#include "news_cnn_tc.cuh"
#include "embed.cuh"

namespace mr {

// im2col view of the title tokens:  A(m, k=(tap,e)) = x[m+tap-1, e] inside the title, else 0
struct ConvTokenView {
  const void* ids; int is64; const float* table; const float* emb; int L, E; int64_t V;
  __device__ __forceinline__ float at(int64_t t, int e) const {
    if (ids) {
      int64_t row = load_index(ids, is64, t);
      row = row < 0 ? 0 : (row >= V ? V - 1 : row);
      return __ldg(table + row * E + e);
    }
    return __ldg(emb + t * E + e);
  }
  __device__ __forceinline__ float operator()(int64_t m, int64_t k) const {   // forward A(m,k)
    int tap = (int)(k / E), e = (int)(k - (int64_t)tap * E);
    int l = (int)(m % L) + tap - 1;
    if (l < 0 || l >= L) return 0.f;
    return at(m + tap - 1, e);
  }
};
struct ConvTokenViewT {      // wgrad A(m=(tap,e), k=t)
  ConvTokenView v;
  __device__ __forceinline__ float operator()(int64_t m, int64_t k) const { return v(k, m); }
};
// dgrad view: A(m=t, k=(tap,h)) = dconv[t-tap+1, h] inside the title, else 0
struct ConvGradView {
  const float* g; int L, H;
  __device__ __forceinline__ float operator()(int64_t m, int64_t k) const {
    int tap = (int)(k / H), h = (int)(k - (int64_t)tap * H);
    int l = (int)(m % L) - tap + 1;
    if (l < 0 || l >= L) return 0.f;
    return __ldg(g + (m - tap + 1) * H + h);
  }
};
struct ConvWStore {          // (m=(tap,e), n=h) -> d_conv_w[h, e, tap]
  float* out; int E;
  __device__ __forceinline__ void operator()(int64_t m, int64_t n, float v) const {
    int tap = (int)(m / E), e = (int)(m - (int64_t)tap * E);
    out[(n * E + e) * 3 + tap] = v;
  }
};
struct ReluGradEpi {         // dconv = (acc + dc_pool (+ d_c)) * (c > 0), written over dc_pool
  float* dc; const float* c; const float* extra; int64_t ld;
  __device__ __forceinline__ void operator()(int64_t m, int64_t n, float v) const {
    int64_t i = m * ld + n;
    float t = v + dc[i];
    if (extra) t += extra[i];
    dc[i] = c[i] > 0.f ? t : 0.f;
  }
};

// conv_w [H,E,3] -> wf [3E,H] (forward B operand) and wd [3H,E] (dgrad B operand)
__global__ void conv_w_permute_kernel(const float* __restrict__ w, float* __restrict__ wf, float* __restrict__ wd, int E, int H) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)H * E * 3;
  if (i >= total) return;
  int tap = (int)(i % 3);
  int e = (int)((i / 3) % E);
  int h = (int)(i / (3 * (int64_t)E));
  float v = w[i];
  if (wf) wf[((int64_t)tap * E + e) * H + h] = v;
  if (wd) wd[((int64_t)tap * H + h) * E + e] = v;
}

static int64_t ws_f32(const mr_cnn_shape* s, int backward) {
  const int64_t T = s->N * s->L, E = s->E, H = s->H;
  int64_t b = 0;
  if (!backward) {
    b += arena_bytes(3 * E * H, 4);
    return b;
  }
  b += arena_bytes(3 * H * E, 4);                         // wd
  b += 2 * arena_bytes(T * H, 4);                         // dkp, dc
  b += arena_bytes(s->N * H, 4);                          // dq partial
  int64_t sp = 64 * ((3 * E * H) > (H * H) ? (3 * E * H) : (H * H));
  b += arena_bytes(sp, 4);                                // split-K partials
  b += arena_bytes(colsum_chunks(T) * H, 4);              // colsum partials
  return b;
}

static int fwd_f32(const mr_cnn_shape* s, const void* ids, int ids_i64, const float* emb, const void* mask, int mask_i64,
                   const float* table, const float* conv_w, const float* conv_b, const float* proj_w, const float* proj_b,
                   const float* query, float* c, float* key, float* prob, float* news, void* ws, int64_t wsb,
                   cudaStream_t st) {
  const int64_t N = s->N, L = s->L, E = s->E, H = s->H, T = N * L;
  Arena ar(ws, wsb);
  float* wf = ar.take<float>(3 * E * H);
  MR_REQUIRE(ar.ok(), MR_ERR_WORKSPACE, "mr_news_cnn_fwd: workspace too small (%lld given)", (long long)wsb);
  conv_w_permute_kernel<<<(unsigned)ceil_div(3 * E * H, 256), 256, 0, st>>>(conv_w, wf, nullptr, (int)E, (int)H);
  MR_CHECK_LAUNCH("conv_w_permute_kernel");
  ConvTokenView av{ids, ids_i64, table, emb, (int)L, (int)E, s->V};
  cudaError_t e = gemm_simt<true, true>(T, H, 3 * E, av, RowMajor{wf, H}, BiasActEpi{c, H, conv_b, 1}, 1, nullptr, st);
  MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "conv gemm: %s", cudaGetErrorString(e));
  e = gemm_simt<true, false>(T, H, H, RowMajor{c, H}, Transposed{proj_w, H}, BiasActEpi{key, H, proj_b, 2}, 1, nullptr, st);
  MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "proj gemm: %s", cudaGetErrorString(e));
  cnn_pool_fwd_kernel<float><<<(unsigned)ceil_div(N, 8), 256, 0, st>>>(c, key, H, mask, mask_i64, query, prob, news, N, (int)L, (int)H);
  MR_CHECK_LAUNCH("cnn_pool_fwd_kernel");
  return MR_OK;
}

static int bwd_f32(const mr_cnn_shape* s, const void* ids, int ids_i64, const float* emb, const float* table,
                   const float* conv_w, const float* proj_w, const float* query, const float* c, const float* key,
                   const float* prob, const float* d_news, const float* d_c, float* d_conv_w, float* d_conv_b,
                   float* d_proj_w, float* d_proj_b, float* d_query, float* d_emb, void* ws, int64_t wsb, cudaStream_t st) {
  const int64_t N = s->N, L = s->L, E = s->E, H = s->H, T = N * L;
  Arena ar(ws, wsb);
  float* wd = ar.take<float>(3 * H * E);
  float* dkp = ar.take<float>(T * H);
  float* dc = ar.take<float>(T * H);
  float* dqp = ar.take<float>(N * H);
  int64_t sp_elems = 64 * ((3 * E * H) > (H * H) ? (3 * E * H) : (H * H));
  float* sp = ar.take<float>(sp_elems);
  float* cp = ar.take<float>(colsum_chunks(T) * H);
  MR_REQUIRE(ar.ok(), MR_ERR_WORKSPACE, "mr_news_cnn_bwd: workspace too small (%lld given)", (long long)wsb);
  cudaError_t e;
  conv_w_permute_kernel<<<(unsigned)ceil_div(3 * E * H, 256), 256, 0, st>>>(conv_w, nullptr, wd, (int)E, (int)H);
  MR_CHECK_LAUNCH("conv_w_permute_kernel");
  cnn_pool_bwd_kernel<float, float><<<(unsigned)ceil_div(N, 8), 256, 0, st>>>(c, key, H, prob, query, d_news, dkp, dc, H, dqp, nullptr, N, (int)L, (int)H);
  MR_CHECK_LAUNCH("cnn_pool_bwd_kernel");
  e = colsum(dqp, d_query, N, H, cp, st);
  MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "colsum dq: %s", cudaGetErrorString(e));
  e = colsum(dkp, d_proj_b, T, H, cp, st);
  MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "colsum dbq: %s", cudaGetErrorString(e));
  // d_proj_w[n,k] = sum_t dkp[t,n] c[t,k]
  e = gemm_simt<false, true>(H, H, T, Transposed{dkp, H}, RowMajor{c, H}, StoreEpi{d_proj_w, H}, pick_splits(H, H, T), sp, st);
  MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "dWq gemm: %s", cudaGetErrorString(e));
  // dconv = relu'(c) * (dc_pool + dkp Wq (+ d_c))
  e = gemm_simt<true, true>(T, H, H, RowMajor{dkp, H}, RowMajor{proj_w, H}, ReluGradEpi{dc, c, d_c, H}, 1, nullptr, st);
  MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "dc gemm: %s", cudaGetErrorString(e));
  e = colsum(dc, d_conv_b, T, H, cp, st);
  MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "colsum dbc: %s", cudaGetErrorString(e));
  ConvTokenViewT avt{ConvTokenView{ids, ids_i64, table, emb, (int)L, (int)E, s->V}};
  e = gemm_simt<false, true>(3 * E, H, T, avt, RowMajor{dc, H}, ConvWStore{d_conv_w, (int)E}, pick_splits(3 * E, H, T), sp, st);
  MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "dWc gemm: %s", cudaGetErrorString(e));
  if (d_emb) {
    e = gemm_simt<true, true>(T, E, 3 * H, ConvGradView{dc, (int)L, (int)H}, RowMajor{wd, E}, StoreEpi{d_emb, E}, 1, nullptr, st);
    MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "dX gemm: %s", cudaGetErrorString(e));
  }
  return MR_OK;
}

static int check_shape(const mr_cnn_shape* s, const char* who) {
  MR_REQUIRE(s != nullptr, MR_ERR_NULL, "%s: null shape", who);
  MR_REQUIRE(s->N >= 0 && s->L >= 1 && s->E >= 1 && s->H >= 1, MR_ERR_BAD_SHAPE, "%s: N=%lld L=%lld E=%lld H=%lld", who,
             (long long)s->N, (long long)s->L, (long long)s->E, (long long)s->H);
  MR_REQUIRE(s->L <= 32 * PL_MAXR, MR_ERR_UNSUPPORTED, "%s: signal_length %lld > %d", who, (long long)s->L, 32 * PL_MAXR);
  MR_REQUIRE(s->N * s->L < (1ll << 31), MR_ERR_UNSUPPORTED, "%s: too many tokens", who);
  MR_REQUIRE(s->precision == MR_F32 || s->precision == MR_BF16, MR_ERR_UNSUPPORTED, "%s: precision %d", who, s->precision);
  return MR_OK;
}

}  // namespace mr

extern "C" {

int64_t mr_news_cnn_workspace_bytes(const mr_cnn_shape* s, int backward) {
  if (!s) return -1;
  if (s->precision == MR_BF16) return mr::news_cnn_tc_workspace_bytes(s, backward);
  return mr::ws_f32(s, backward) + 256;
}

int mr_news_cnn_fwd(const mr_cnn_shape* s, const void* ids, int ids_i64, const float* emb, const void* mask, int mask_i64,
                    const void* table, const float* conv_w, const float* conv_b, const float* proj_w, const float* proj_b,
                    const float* query, void* c_save, void* key_save, float* prob, float* news, void* workspace,
                    int64_t workspace_bytes, void* stream) {
  using namespace mr;
  if (int rc = require_sm100()) return rc;
  if (int rc = check_shape(s, "mr_news_cnn_fwd")) return rc;
  MR_REQUIRE((ids != nullptr) != (emb != nullptr), MR_ERR_NULL, "mr_news_cnn_fwd: exactly one of ids / emb must be given");
  MR_REQUIRE(!ids || table, MR_ERR_NULL, "mr_news_cnn_fwd: ids given without a table");
  MR_REQUIRE(conv_w && conv_b && proj_w && proj_b && query && c_save && key_save && prob && news, MR_ERR_NULL,
             "mr_news_cnn_fwd: null pointer");
  if (s->N == 0) return MR_OK;
  if (s->precision == MR_BF16)
    return news_cnn_tc_fwd(s, ids, ids_i64, emb, mask, mask_i64, table, conv_w, conv_b, proj_w, proj_b, query, c_save,
                           key_save, prob, news, workspace, workspace_bytes, as_stream(stream));
  return fwd_f32(s, ids, ids_i64, emb, mask, mask_i64, static_cast<const float*>(table), conv_w, conv_b, proj_w, proj_b,
                 query, static_cast<float*>(c_save), static_cast<float*>(key_save), prob, news, workspace, workspace_bytes,
                 as_stream(stream));
}

int mr_news_cnn_bwd(const mr_cnn_shape* s, const void* ids, int ids_i64, const float* emb, const void* table,
                    const float* conv_w, const float* proj_w, const float* query, const void* c_save, const void* key_save,
                    const float* prob, const float* d_news, const float* d_c, float* d_conv_w, float* d_conv_b,
                    float* d_proj_w, float* d_proj_b, float* d_query, void* d_emb, void* workspace, int64_t workspace_bytes,
                    void* stream) {
  using namespace mr;
  if (int rc = require_sm100()) return rc;
  if (int rc = check_shape(s, "mr_news_cnn_bwd")) return rc;
  MR_REQUIRE((ids != nullptr) != (emb != nullptr), MR_ERR_NULL, "mr_news_cnn_bwd: exactly one of ids / emb must be given");
  MR_REQUIRE(conv_w && proj_w && query && c_save && key_save && prob && d_news && d_conv_w && d_conv_b && d_proj_w &&
                 d_proj_b && d_query, MR_ERR_NULL, "mr_news_cnn_bwd: null pointer");
  if (s->N == 0) {
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(d_conv_w, 0, sizeof(float) * s->H * s->E * 3, st);
    cudaMemsetAsync(d_conv_b, 0, sizeof(float) * s->H, st);
    cudaMemsetAsync(d_proj_w, 0, sizeof(float) * s->H * s->H, st);
    cudaMemsetAsync(d_proj_b, 0, sizeof(float) * s->H, st);
    cudaMemsetAsync(d_query, 0, sizeof(float) * s->H, st);
    return MR_OK;
  }
  if (s->precision == MR_BF16)
    return news_cnn_tc_bwd(s, ids, ids_i64, emb, table, conv_w, proj_w, query, c_save, key_save, prob, d_news, d_c,
                           d_conv_w, d_conv_b, d_proj_w, d_proj_b, d_query, d_emb, workspace, workspace_bytes,
                           as_stream(stream));
  return bwd_f32(s, ids, ids_i64, emb, static_cast<const float*>(table), conv_w, proj_w, query,
                 static_cast<const float*>(c_save), static_cast<const float*>(key_save), prob, d_news, d_c, d_conv_w,
                 d_conv_b, d_proj_w, d_proj_b, d_query, static_cast<float*>(d_emb), workspace, workspace_bytes,
                 as_stream(stream));
}

int64_t mr_token_group_plan_bytes(int64_t n_tokens, int64_t V) { return mr::token_group_plan_bytes(n_tokens, V); }

int mr_token_group_plan(const void* ids, int ids_i64, int64_t n_tokens, int64_t V, void* plan, int64_t plan_bytes, void* stream) {
  using namespace mr;
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(ids && plan, MR_ERR_NULL, "mr_token_group_plan: null pointer");
  MR_REQUIRE(n_tokens >= 0 && V >= 1 && n_tokens < (1ll << 31), MR_ERR_BAD_SHAPE, "mr_token_group_plan: n_tokens=%lld V=%lld",
             (long long)n_tokens, (long long)V);
  if (n_tokens == 0) return MR_OK;
  return token_group_plan_build(ids, ids_i64, n_tokens, V, plan, plan_bytes, as_stream(stream));
}

int64_t mr_news_cnn_bwd_table_workspace_bytes(const mr_cnn_shape* s) {
  if (!s || s->precision != MR_BF16) return -1;
  return mr::news_cnn_tc_workspace_bytes(s, 2);
}

int mr_news_cnn_bwd_table(const mr_cnn_shape* s, const void* ids, int ids_i64, const void* table_bf16, int64_t table_rows,
                          const float* conv_w, const float* proj_w, const float* query, const void* c_save,
                          const void* key_save, const float* prob, const float* d_news, float* d_conv_w, float* d_conv_b,
                          float* d_proj_w, float* d_proj_b, float* d_query, float* d_table, int64_t padding_idx,
                          const void* group_plan, int64_t group_plan_bytes, void* table_ready_event, void* workspace, int64_t workspace_bytes,
                          void* stream) {
  using namespace mr;
  if (int rc = require_sm100()) return rc;
  if (int rc = check_shape(s, "mr_news_cnn_bwd_table")) return rc;
  MR_REQUIRE(s->precision == MR_BF16, MR_ERR_UNSUPPORTED, "mr_news_cnn_bwd_table: bf16 path only");
  MR_REQUIRE(ids && table_bf16 && conv_w && proj_w && query && c_save && key_save && prob && d_news && d_conv_w && d_conv_b &&
                 d_proj_w && d_proj_b && d_query && d_table, MR_ERR_NULL, "mr_news_cnn_bwd_table: null pointer");
  MR_REQUIRE(group_plan == nullptr || group_plan_bytes >= token_group_plan_bytes(s->N * s->L, s->V), MR_ERR_WORKSPACE,
             "mr_news_cnn_bwd_table: grouping plan of %lld bytes is too small", (long long)group_plan_bytes);
  cudaStream_t st = as_stream(stream);
  if (s->N == 0) {
    cudaMemsetAsync(d_conv_w, 0, sizeof(float) * s->H * s->E * 3, st);
    cudaMemsetAsync(d_conv_b, 0, sizeof(float) * s->H, st);
    cudaMemsetAsync(d_proj_w, 0, sizeof(float) * s->H * s->H, st);
    cudaMemsetAsync(d_proj_b, 0, sizeof(float) * s->H, st);
    cudaMemsetAsync(d_query, 0, sizeof(float) * s->H, st);
    cudaMemsetAsync(d_table, 0, sizeof(float) * s->V * s->E, st);
    return MR_OK;
  }
  return news_cnn_tc_bwd(s, ids, ids_i64, nullptr, table_bf16, conv_w, proj_w, query, c_save, key_save, prob, d_news, nullptr,
                         d_conv_w, d_conv_b, d_proj_w, d_proj_b, d_query, nullptr, workspace, workspace_bytes, st, d_table,
                         table_rows, padding_idx, group_plan, table_ready_event);
}

}  // extern "C"
