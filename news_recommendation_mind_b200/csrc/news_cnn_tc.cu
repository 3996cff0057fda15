// MR_BF16 implementation of the CNN news encoder: bf16 operands on tcgen05 tensor cores, fp32
// accumulation in TMEM, every non-GEMM step in fp32.  Reference: models/Encoders/CNN.py:30-51 with the
// token gather of models/Embeddings/BERT.py:39 fused into the conv's A-operand producer.
//
// forward   ids --gather--> [conv3 tap GEMM + bias + ReLU] -> c_save (bf16 [T,Hp])
//           c_save --------> [proj tap GEMM + bias + tanh] -> key_save (bf16 [T,Hp])
//           (c_save, key_save, mask, query) -> masked-softmax pooling (one warp per title) -> prob, news
// backward  pooling bwd -> dkp, dc_pool (bf16)
//           dconv = relu'(c) * (dc_pool + dkp Wq)            [tap GEMM, RELUGRAD epilogue]
//           d_proj_w = dkp^T c,  d_conv_w[tap] = dconv^T x[.+tap-1]      [token-reduction GEMMs]
//           d_emb = sum_tap dconv[.-tap+1] Wc[tap]^T          [tap GEMM, dir = -1]
// Shapes: Hp = H rounded up to 16 (<= 256), Kp = E rounded up to 16; the bf16 token table has row
// pitch align_up(E, 64) (the shadow kept by the Python BERT_Embedding / written by mr_adam_step).
#include "news_cnn_tc.cuh"
#include "pool_kernels.cuh"
#include "tapgemm.cuh"
#include "gemm_simt.cuh"

namespace mr {

__global__ void cast_rows_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t rows,
                                      int64_t cols, int64_t ld) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ld) return;
  int64_t r = i / ld, c = i - r * ld;
  dst[i] = __float2bfloat16(c < cols ? src[r * cols + c] : 0.f);
}

static inline int64_t hp_of(const mr_cnn_shape* s) { return align_up(s->H, 16); }
static inline int64_t kp_of(const mr_cnn_shape* s) { return align_up(s->E, 16); }
static inline int64_t table_ld(const mr_cnn_shape* s) { return align_up(s->E, 64); }

int64_t news_cnn_tc_workspace_bytes(const mr_cnn_shape* s, int backward) {
  const int64_t T = s->N * s->L, Hp = hp_of(s), Kp = kp_of(s);
  int64_t b = 256;
  b += arena_bytes(tapgemm_pack_bytes(3, (int)Hp, (int)Kp), 1);     // conv weights
  b += arena_bytes(tapgemm_pack_bytes(1, (int)Hp, (int)Hp), 1);     // proj weights
  b += arena_bytes(T * Kp, 2);                                      // dense-embedding input as bf16
  if (backward) {
    b += 2 * arena_bytes(T * Hp, 2);                                // dkp, dc_pool / dconv
    b += arena_bytes(tapgemm_pack_bytes(3, (int)Kp, (int)Hp), 1);   // dgrad weights
    b += arena_bytes(s->N * s->H, 4);                               // dq partial
    b += arena_bytes(colsum_chunks(T) * Hp, 4);
  }
  return b;
}

static int check_tc(const mr_cnn_shape* s, const char* who) {
  MR_REQUIRE(hp_of(s) <= 256, MR_ERR_UNSUPPORTED, "%s: hidden_dim %lld > 256 is not supported by the bf16 path", who, (long long)s->H);
  MR_REQUIRE(s->L <= 128, MR_ERR_UNSUPPORTED, "%s: signal_length %lld > 128 is not supported by the bf16 path", who, (long long)s->L);
  return MR_OK;
}

int news_cnn_tc_fwd(const mr_cnn_shape* s, const void* ids, int ids_i64, const float* emb, const void* mask, int mask_i64,
                    const void* table, const float* conv_w, const float* conv_b, const float* proj_w, const float* proj_b,
                    const float* query, void* c_save, void* key_save, float* prob, float* news, void* ws, int64_t wsb,
                    cudaStream_t st) {
  if (int rc = check_tc(s, "mr_news_cnn_fwd")) return rc;
  const int64_t N = s->N, L = s->L, E = s->E, H = s->H, T = N * L, Hp = hp_of(s), Kp = kp_of(s);
  Arena ar(ws, wsb);
  uint8_t* wconv = ar.take<uint8_t>(tapgemm_pack_bytes(3, (int)Hp, (int)Kp));
  uint8_t* wproj = ar.take<uint8_t>(tapgemm_pack_bytes(1, (int)Hp, (int)Hp));
  __nv_bfloat16* xa = ids ? nullptr : ar.take<__nv_bfloat16>(T * Kp);
  MR_REQUIRE(ar.ok(), MR_ERR_WORKSPACE, "mr_news_cnn_fwd: workspace too small (%lld given)", (long long)wsb);
  __nv_bfloat16* c = static_cast<__nv_bfloat16*>(c_save);
  __nv_bfloat16* key = static_cast<__nv_bfloat16*>(key_save);

  if (int rc = tapgemm_pack(conv_w, wconv, 3, (int)Hp, (int)Kp, (int)H, (int)E, 3 * E, 3, 1, st)) return rc;
  if (int rc = tapgemm_pack(proj_w, wproj, 1, (int)Hp, (int)Hp, (int)H, (int)H, H, 1, 0, st)) return rc;
  if (!ids) {
    cast_rows_bf16_kernel<<<(unsigned)ceil_div(T * Kp, 256), 256, 0, st>>>(emb, xa, T, E, Kp);
    MR_CHECK_LAUNCH("cast_rows_bf16_kernel");
  }
  TapGemmArgs a{};
  TapGemmPlan plan;
  // conv: c = relu(conv3(x) + b)
  a.n_titles = N; a.L = (int)L; a.taps = 3; a.dir = 1; a.K = (int)Kp;
  a.n_sub = 1; a.nsz[0] = (int)Hp; a.nsz[1] = 0;
  if (ids) { a.ids = ids; a.ids_i64 = ids_i64; a.a = static_cast<const __nv_bfloat16*>(table); a.lda = table_ld(s); a.V = s->V; }
  else { a.ids = nullptr; a.a = xa; a.lda = Kp; a.V = 0; }
  a.wpack = wconv; a.epi = TG_EPI_BIAS_RELU; a.bias = conv_b; a.n_valid = (int)H;
  a.out = c; a.ldo = Hp;
  if (int rc = tapgemm_plan(a, &plan)) return rc;
  if (int rc = tapgemm_launch(plan, st)) return rc;
  // projection: key = tanh(c Wq^T + bq)
  TapGemmArgs b{};
  b.n_titles = N; b.L = (int)L; b.taps = 1; b.dir = 1; b.K = (int)Hp;
  b.n_sub = 1; b.nsz[0] = (int)Hp;
  b.ids = nullptr; b.a = c; b.lda = Hp;
  b.wpack = wproj; b.epi = TG_EPI_BIAS_TANH; b.bias = proj_b; b.n_valid = (int)H;
  b.out = key; b.ldo = Hp;
  if (int rc = tapgemm_plan(b, &plan)) return rc;
  if (int rc = tapgemm_launch(plan, st)) return rc;
  cnn_pool_fwd_kernel<__nv_bfloat16><<<(unsigned)ceil_div(N, 8), 256, 0, st>>>(c, key, Hp, mask, mask_i64, query, prob, news, N, (int)L, (int)H);
  MR_CHECK_LAUNCH("cnn_pool_fwd_kernel");
  return MR_OK;
}

int news_cnn_tc_bwd(const mr_cnn_shape*, const void*, int, const float*, const void*, const float*, const float*,
                    const float*, const void*, const void*, const float*, const float*, const float*, float*, float*,
                    float*, float*, float*, void*, void*, int64_t, cudaStream_t) {
  return set_err(MR_ERR_UNSUPPORTED, "MR_BF16 news encoder backward not built yet");
}
}  // namespace mr
