// Candidate scoring, pooling user encoders and ranking metrics (HBM / latency bound kernels).
//   scoring  : models/TwoTowerBaseModel.py:51-84   (+ NLLLoss, utils/Manager.py:381-382,641)
//   pooling  : models/Encoders/Pooling.py:5-43     (Attention_Pooling / Average_Pooling)
//   metrics  : utils/Manager.py:1205-1344,842-850  (auc, mrr, ndcg@5/10, ordinal rank)
#include "common.cuh"

namespace mr {

// ---------------------------------------------------------------------------------------------
// score + log-softmax: one warp per impression
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
score_logsoftmax_fwd_kernel(const float* __restrict__ cdd, const float* __restrict__ user, float* __restrict__ logp,
                            int64_t B, int C, int H) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float inv = rsqrtf((float)H);
  const float* u = user + b * H;
  float mx = -INFINITY;
  for (int c = 0; c < C; ++c) {
    const float* v = cdd + (b * C + c) * H;
    float d = 0.f;
    for (int h = lane; h < H; h += 32) d = fmaf(__ldg(v + h), __ldg(u + h), d);
    d = warp_sum(d) * inv;
    if (lane == 0) logp[b * C + c] = d;
    mx = fmaxf(mx, d);
  }
  __syncwarp();
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += expf(logp[b * C + c] - mx);
  sum = warp_sum(sum);
  const float lse = mx + logf(sum);
  for (int c = lane; c < C; c += 32) logp[b * C + c] -= lse;
}

__global__ void nll_mean_kernel(const float* __restrict__ logp, const void* __restrict__ label, int label_i64,
                                float* __restrict__ loss, int64_t B, int C) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sm[32];
  float s = 0.f;
  for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
    int64_t y = load_index(label, label_i64, b);
    if (y >= 0 && y < C) s -= logp[b * C + y];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) loss[0] = t / (float)B;
  }
}

__global__ void nll_bwd_kernel(const void* __restrict__ label, int label_i64, const float* __restrict__ d_loss,
                               float* __restrict__ d_logp, int64_t B, int C) {
  pdl_trigger();
  pdl_wait();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  int64_t b = i / C; int c = (int)(i - b * C);
  float g = d_loss ? d_loss[0] : 1.f;
  d_logp[i] = (load_index(label, label_i64, b) == c) ? -g / (float)B : 0.f;
}

__global__ void __launch_bounds__(256)
score_logsoftmax_bwd_kernel(const float* __restrict__ cdd, const float* __restrict__ user, const float* __restrict__ logp,
                            const float* __restrict__ d_logp, float* __restrict__ d_cdd, float* __restrict__ d_user,
                            int64_t B, int C, int H) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float inv = rsqrtf((float)H);
  float gs = 0.f;
  for (int c = lane; c < C; c += 32) gs += d_logp[b * C + c];
  gs = warp_sum(gs);
  const float* u = user + b * H;
  for (int h0 = 0; h0 < H; h0 += 32) {
    int h = h0 + lane;
    float uh = h < H ? __ldg(u + h) : 0.f;
    float du = 0.f;
    for (int c = 0; c < C; ++c) {
      float ds = (d_logp[b * C + c] - expf(logp[b * C + c]) * gs) * inv;     // d score
      if (h < H) {
        int64_t i = (b * C + c) * H + h;
        du = fmaf(ds, __ldg(cdd + i), du);
        d_cdd[i] = ds * uh;
      }
    }
    if (h < H) d_user[b * H + h] = du;
  }
}

// sigmoid(score): one warp per (impression, candidate)
__global__ void __launch_bounds__(256)
score_sigmoid_kernel(const float* __restrict__ cdd, const float* __restrict__ user, float* __restrict__ prob,
                     int64_t B, int C, int H, int apply_sigmoid) {
  const int lane = threadIdx.x & 31;
  const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= B * C) return;
  const float* v = cdd + j * H;
  const float* u = user + (j / C) * H;
  float d = 0.f;
  for (int h = lane; h < H; h += 32) d = fmaf(__ldg(v + h), __ldg(u + h), d);
  d = warp_sum(d) * rsqrtf((float)H);
  if (lane == 0) prob[j] = apply_sigmoid ? sigmoidf_(d) : d;
}

// fast eval: block per impression, user vector staged in smem, warps stride over candidates
__global__ void __launch_bounds__(128)
score_sigmoid_gather_kernel(const float* __restrict__ table, const void* __restrict__ cdd_id, int id_i64,
                            const int64_t* __restrict__ offsets, const float* __restrict__ user, float* __restrict__ prob,
                            int64_t n_impr, int64_t n_rows, int H) {
  extern __shared__ float u_s[];
  const int64_t i = blockIdx.x;
  if (i >= n_impr) return;
  for (int h = threadIdx.x; h < H; h += blockDim.x) u_s[h] = user[i * H + h];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float inv = rsqrtf((float)H);
  for (int64_t j = offsets[i] + w; j < offsets[i + 1]; j += nw) {
    int64_t row = load_index(cdd_id, id_i64, j);
    row = row < 0 ? 0 : (row >= n_rows ? n_rows - 1 : row);
    const float* v = table + row * H;
    float d = 0.f;
    for (int h = lane; h < H; h += 32) d = fmaf(__ldg(v + h), u_s[h], d);
    d = warp_sum(d) * inv;
    if (lane == 0) prob[j] = sigmoidf_(d);
  }
}

// ---------------------------------------------------------------------------------------------
// attention pooling over [B,S,H] with key = value (Pooling.py:12-25), warp per row b; S <= 256
// ---------------------------------------------------------------------------------------------
constexpr int AP_MAXR = 8;

__global__ void __launch_bounds__(256)
attnpool_fwd_kernel(const float* __restrict__ r, const float* __restrict__ mask, const float* __restrict__ q,
                    float* __restrict__ prob, float* __restrict__ out, int64_t B, int S, int H) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float inv = rsqrtf((float)H);
  const float* rb = r + b * S * H;
  float s[AP_MAXR]; bool keep[AP_MAXR];
#pragma unroll
  for (int i = 0; i < AP_MAXR; ++i) { s[i] = 0.f; keep[i] = false; }
  for (int t = 0; t < S; ++t) {
    float d = 0.f;
    for (int h = lane; h < H; h += 32) d = fmaf(__ldg(q + h), __ldg(rb + (int64_t)t * H + h), d);
    d = warp_sum(d) * inv;
    if ((t & 31) == lane) {
#pragma unroll
      for (int i = 0; i < AP_MAXR; ++i)
        if (i == (t >> 5)) { s[i] = d; keep[i] = mask ? (mask[b * S + t] != 0.f) : true; }
    }
  }
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < AP_MAXR; ++i) if (keep[i]) mx = fmaxf(mx, s[i]);
  mx = warp_max(mx);
  float e[AP_MAXR], sum = 0.f;
#pragma unroll
  for (int i = 0; i < AP_MAXR; ++i) { e[i] = keep[i] ? expf(s[i] - mx) : 0.f; sum += e[i]; }
  sum = warp_sum(sum);
  const float rs = sum > 0.f ? 1.f / sum : 0.f;
#pragma unroll
  for (int i = 0; i < AP_MAXR; ++i) {
    e[i] *= rs;
    int t = i * 32 + lane;
    if (t < S) prob[b * S + t] = e[i];
  }
  for (int h0 = 0; h0 < H; h0 += 32) {
    int h = h0 + lane;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < AP_MAXR; ++i) {
      if (i * 32 >= S) break;
      for (int j = 0; j < 32 && i * 32 + j < S; ++j) {
        float pj = __shfl_sync(0xffffffffu, e[i], j);
        if (h < H) acc = fmaf(pj, __ldg(rb + (int64_t)(i * 32 + j) * H + h), acc);
      }
    }
    if (h < H) out[b * H + h] = acc;
  }
}

__global__ void __launch_bounds__(256)
attnpool_bwd_kernel(const float* __restrict__ r, const float* __restrict__ q, const float* __restrict__ prob,
                    const float* __restrict__ d_out, float* __restrict__ d_r, float* __restrict__ dq_partial, int64_t B,
                    int S, int H) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float inv = rsqrtf((float)H);
  const float* rb = r + b * S * H;
  const float* go = d_out + b * H;
  float p[AP_MAXR], dp[AP_MAXR];
#pragma unroll
  for (int i = 0; i < AP_MAXR; ++i) {
    int t = i * 32 + lane;
    p[i] = t < S ? prob[b * S + t] : 0.f;
    dp[i] = 0.f;
  }
  for (int t = 0; t < S; ++t) {
    float d = 0.f;
    for (int h = lane; h < H; h += 32) d = fmaf(__ldg(go + h), __ldg(rb + (int64_t)t * H + h), d);
    d = warp_sum(d);
    if ((t & 31) == lane) {
#pragma unroll
      for (int i = 0; i < AP_MAXR; ++i) if (i == (t >> 5)) dp[i] = d;
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < AP_MAXR; ++i) dot = fmaf(p[i], dp[i], dot);
  dot = warp_sum(dot);
  float ds[AP_MAXR];
#pragma unroll
  for (int i = 0; i < AP_MAXR; ++i) ds[i] = p[i] * (dp[i] - dot) * inv;
  for (int h0 = 0; h0 < H; h0 += 32) {
    int h = h0 + lane;
    float qh = h < H ? __ldg(q + h) : 0.f, gh = h < H ? __ldg(go + h) : 0.f, dq = 0.f;
    // dq = sum_t ds[t] r[t,h] with sum_t ds[t] = 0: when the rows r[t,:] are nearly equal (the output of a self-attention
    // layer) the plain sum cancels catastrophically in fp32.  Summing ds[t] (r[t,h] - rbar[h]) with rbar = sum_t p[t] r[t,h]
    // is the same number with both factors small.
    float rbar = 0.f;
#pragma unroll
    for (int i = 0; i < AP_MAXR; ++i) {
      if (i * 32 >= S) break;
      for (int j = 0; j < 32 && i * 32 + j < S; ++j) {
        float pt = __shfl_sync(0xffffffffu, p[i], j);
        if (h < H) rbar = fmaf(pt, __ldg(r + (b * S + i * 32 + j) * H + h), rbar);
      }
    }
#pragma unroll
    for (int i = 0; i < AP_MAXR; ++i) {
      if (i * 32 >= S) break;
      for (int j = 0; j < 32 && i * 32 + j < S; ++j) {
        int t = i * 32 + j;
        float dst = __shfl_sync(0xffffffffu, ds[i], j);
        float pt = __shfl_sync(0xffffffffu, p[i], j);
        if (h < H) {
          int64_t o = (b * S + t) * H + h;
          dq = fmaf(dst, __ldg(r + o) - rbar, dq);
          d_r[o] = pt * gh + dst * qh;
        }
      }
    }
    if (h < H) dq_partial[b * H + h] = dq;
  }
}

__global__ void avgpool_fwd_kernel(const float* __restrict__ r, float* __restrict__ out, int64_t B, int S, int H) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  int64_t b = i / H; int h = (int)(i - b * H);
  float s = 0.f;
  for (int t = 0; t < S; ++t) s += r[(b * S + t) * H + h];
  out[i] = s / (float)S;
}
__global__ void avgpool_bwd_kernel(const float* __restrict__ d_out, float* __restrict__ d_r, int64_t B, int S, int H) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * S * H) return;
  int64_t b = i / ((int64_t)S * H); int h = (int)(i % H);
  d_r[i] = d_out[b * H + h] / (float)S;
}

// ---------------------------------------------------------------------------------------------
// ranking metrics: block per impression, O(n^2) rank counting in shared memory (n is tens to a
// few hundred in MIND).  Two tie rules, both the reference's: prediction.txt ranks are
// scipy.stats.rankdata(method="ordinal") = descending score, ties by ASCENDING position (Manager.py:846);
// MRR / nDCG order with np.argsort(score)[::-1] (Manager.py:1216,1269), i.e. a reversed ascending sort = ties by
// DESCENDING position wherever numpy's sort is stable (numpy 1.x insertion sort for n <= 16; beyond that, and in
// numpy 2.x with AVX-512, the reference's tie order is an implementation detail no rule can reproduce).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
rank_metrics_kernel(const float* __restrict__ prob, const float* __restrict__ label, const int64_t* __restrict__ offsets,
                    double* __restrict__ metrics, int32_t* __restrict__ rank_out, int64_t n_impr, int smem_cap) {
  extern __shared__ float sm[];
  __shared__ double red[4][4];
  const int64_t i = blockIdx.x;
  if (i >= n_impr) return;
  const int64_t beg = offsets[i];
  const int n = (int)(offsets[i + 1] - beg);
  const bool in_smem = n <= smem_cap;
  float* s_s = sm;
  float* l_s = sm + smem_cap;
  if (in_smem)
    for (int j = threadIdx.x; j < n; j += blockDim.x) { s_s[j] = prob[beg + j]; l_s[j] = label[beg + j]; }
  __syncthreads();
  const float* S_ = in_smem ? s_s : prob + beg;
  const float* L_ = in_smem ? l_s : label + beg;
  double a_num = 0.0, a_pos = 0.0, mrr_num = 0.0, dcg5 = 0.0, dcg10 = 0.0, idcg5 = 0.0, idcg10 = 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float sj = S_[j], lj = L_[j];
    int rank = 1, lrank = 1, orank = 1;
    double below = 0.0;
    for (int k = 0; k < n; ++k) {
      const float sk = S_[k], lk = L_[k];
      orank += (sk > sj) || (sk == sj && k < j);      // ordinal rank of prediction.txt
      rank += (sk > sj) || (sk == sj && k > j);       // position in argsort(score)[::-1]
      lrank += (lk > lj) || (lk == lj && k > j);      // ideal order (tie order irrelevant: equal labels, equal gains)
      if (lj == 1.f && lk != 1.f) below += (sj > sk) ? 1.0 : (sj == sk ? 0.5 : 0.0);
    }
    if (rank_out) rank_out[beg + j] = orank;
    const double gain = exp2((double)lj) - 1.0;
    if (lj == 1.f) { a_pos += 1.0; a_num += below; }
    mrr_num += (double)lj / (double)rank;
    if (rank <= 5) dcg5 += gain / log2((double)rank + 1.0);
    if (rank <= 10) dcg10 += gain / log2((double)rank + 1.0);
    if (lrank <= 5) idcg5 += gain / log2((double)lrank + 1.0);
    if (lrank <= 10) idcg10 += gain / log2((double)lrank + 1.0);
  }
  double v[8] = {a_num, a_pos, mrr_num, dcg5, dcg10, idcg5, idcg10, 0.0};
  // label sum for mrr denominator
  double lsum = 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) lsum += (double)L_[j];
  v[7] = lsum;
  __shared__ double acc[8][4];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    double w = warp_sum_d(v[t]);
    if ((threadIdx.x & 31) == 0) acc[t][threadIdx.x >> 5] = w;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[8];
    for (int k = 0; k < 8; ++k) t[k] = acc[k][0] + acc[k][1] + acc[k][2] + acc[k][3];
    const double nneg = (double)n - t[1];
    metrics[i * 4 + 0] = t[0] / (t[1] * nneg);
    metrics[i * 4 + 1] = t[2] / t[7];
    metrics[i * 4 + 2] = t[3] / t[5];
    metrics[i * 4 + 3] = t[4] / t[6];
  }
  (void)red;
}

}  // namespace mr

extern "C" {
using namespace mr;

int mr_score_logsoftmax_fwd(const float* cdd, const float* user, float* logp, const void* label, int label_i64,
                            float* loss_mean, int64_t B, int64_t C, int64_t H, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(cdd && user && logp, MR_ERR_NULL, "mr_score_logsoftmax_fwd: null pointer");
  MR_REQUIRE(B >= 0 && C >= 1 && H >= 1, MR_ERR_BAD_SHAPE, "mr_score_logsoftmax_fwd: B=%lld C=%lld H=%lld", (long long)B,
             (long long)C, (long long)H);
  if (B == 0) return MR_OK;
  cudaStream_t st = as_stream(stream);
  launch_pdl(score_logsoftmax_fwd_kernel, dim3((unsigned)ceil_div(B, 8)), dim3(256), 0, st, cdd, user, logp, B, (int)C, (int)H);
  MR_CHECK_LAUNCH("score_logsoftmax_fwd_kernel");
  if (loss_mean) {
    MR_REQUIRE(label != nullptr, MR_ERR_NULL, "mr_score_logsoftmax_fwd: loss requested without labels");
    launch_pdl(nll_mean_kernel, dim3(1), dim3(256), 0, st, logp, label, label_i64, loss_mean, B, (int)C);
    MR_CHECK_LAUNCH("nll_mean_kernel");
  }
  return MR_OK;
}

int mr_nll_loss_fwd(const float* logp, const void* label, int label_i64, float* loss, int64_t B, int64_t C, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(logp && label && loss, MR_ERR_NULL, "mr_nll_loss_fwd: null pointer");
  MR_REQUIRE(B >= 1 && C >= 1, MR_ERR_BAD_SHAPE, "mr_nll_loss_fwd: bad shape");
  launch_pdl(nll_mean_kernel, dim3(1), dim3(256), 0, as_stream(stream), logp, label, label_i64, loss, B, (int)C);
  MR_CHECK_LAUNCH("nll_mean_kernel");
  return MR_OK;
}

int mr_nll_loss_bwd(const void* label, int label_i64, const float* d_loss, float* d_logp, int64_t B, int64_t C, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(label && d_logp, MR_ERR_NULL, "mr_nll_loss_bwd: null pointer");
  MR_REQUIRE(B >= 1 && C >= 1, MR_ERR_BAD_SHAPE, "mr_nll_loss_bwd: bad shape");
  launch_pdl(nll_bwd_kernel, dim3((unsigned)ceil_div(B * C, 256)), dim3(256), 0, as_stream(stream), label, label_i64, d_loss, d_logp, B, (int)C);
  MR_CHECK_LAUNCH("nll_bwd_kernel");
  return MR_OK;
}

int mr_score_logsoftmax_bwd(const float* cdd, const float* user, const float* logp, const float* d_logp, float* d_cdd,
                            float* d_user, int64_t B, int64_t C, int64_t H, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(cdd && user && logp && d_logp && d_cdd && d_user, MR_ERR_NULL, "mr_score_logsoftmax_bwd: null pointer");
  MR_REQUIRE(B >= 0 && C >= 1 && H >= 1, MR_ERR_BAD_SHAPE, "mr_score_logsoftmax_bwd: bad shape");
  if (B == 0) return MR_OK;
  launch_pdl(score_logsoftmax_bwd_kernel, dim3((unsigned)ceil_div(B, 8)), dim3(256), 0, as_stream(stream), cdd, user, logp, d_logp, d_cdd,
                                                                                       d_user, B, (int)C, (int)H);
  MR_CHECK_LAUNCH("score_logsoftmax_bwd_kernel");
  return MR_OK;
}

int mr_score_sigmoid_fwd(const float* cdd, const float* user, float* prob, int64_t B, int64_t C, int64_t H,
                         int apply_sigmoid, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(cdd && user && prob, MR_ERR_NULL, "mr_score_sigmoid_fwd: null pointer");
  MR_REQUIRE(B >= 0 && C >= 1 && H >= 1, MR_ERR_BAD_SHAPE, "mr_score_sigmoid_fwd: bad shape");
  if (B == 0) return MR_OK;
  score_sigmoid_kernel<<<(unsigned)ceil_div(B * C, 8), 256, 0, as_stream(stream)>>>(cdd, user, prob, B, (int)C, (int)H,
                                                                                    apply_sigmoid);
  MR_CHECK_LAUNCH("score_sigmoid_kernel");
  return MR_OK;
}

int mr_score_sigmoid_gather_fwd(const float* news_table, const void* cdd_id, int id_i64, const int64_t* offsets,
                                const float* user, float* prob, int64_t n_impr, int64_t n_cand, int64_t n_news_rows,
                                int64_t H, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(news_table && cdd_id && offsets && user && prob, MR_ERR_NULL, "mr_score_sigmoid_gather_fwd: null pointer");
  MR_REQUIRE(n_impr >= 0 && n_cand >= 0 && n_news_rows >= 1 && H >= 1 && H <= 8192, MR_ERR_BAD_SHAPE,
             "mr_score_sigmoid_gather_fwd: bad shape");
  if (n_impr == 0) return MR_OK;
  score_sigmoid_gather_kernel<<<(unsigned)n_impr, 128, sizeof(float) * H, as_stream(stream)>>>(
      news_table, cdd_id, id_i64, offsets, user, prob, n_impr, n_news_rows, (int)H);
  MR_CHECK_LAUNCH("score_sigmoid_gather_kernel");
  return MR_OK;
}

int mr_attnpool_fwd(const float* r, const float* mask, const float* query, float* prob, float* out, int64_t B, int64_t S,
                    int64_t H, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(r && query && prob && out, MR_ERR_NULL, "mr_attnpool_fwd: null pointer");
  MR_REQUIRE(B >= 0 && S >= 1 && H >= 1, MR_ERR_BAD_SHAPE, "mr_attnpool_fwd: bad shape");
  MR_REQUIRE(S <= 32 * AP_MAXR, MR_ERR_UNSUPPORTED, "mr_attnpool_fwd: S=%lld > %d", (long long)S, 32 * AP_MAXR);
  if (B == 0) return MR_OK;
  attnpool_fwd_kernel<<<(unsigned)ceil_div(B, 8), 256, 0, as_stream(stream)>>>(r, mask, query, prob, out, B, (int)S, (int)H);
  MR_CHECK_LAUNCH("attnpool_fwd_kernel");
  return MR_OK;
}

int mr_attnpool_bwd(const float* r, const float* query, const float* prob, const float* d_out, float* d_r,
                    float* d_query_partial, int64_t B, int64_t S, int64_t H, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(r && query && prob && d_out && d_r && d_query_partial, MR_ERR_NULL, "mr_attnpool_bwd: null pointer");
  MR_REQUIRE(B >= 0 && S >= 1 && H >= 1 && S <= 32 * AP_MAXR, MR_ERR_BAD_SHAPE, "mr_attnpool_bwd: bad shape");
  if (B == 0) return MR_OK;
  attnpool_bwd_kernel<<<(unsigned)ceil_div(B, 8), 256, 0, as_stream(stream)>>>(r, query, prob, d_out, d_r, d_query_partial, B,
                                                                               (int)S, (int)H);
  MR_CHECK_LAUNCH("attnpool_bwd_kernel");
  return MR_OK;
}

int mr_avgpool_fwd(const float* r, float* out, int64_t B, int64_t S, int64_t H, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(r && out, MR_ERR_NULL, "mr_avgpool_fwd: null pointer");
  MR_REQUIRE(B >= 0 && S >= 1 && H >= 1, MR_ERR_BAD_SHAPE, "mr_avgpool_fwd: bad shape");
  if (B == 0) return MR_OK;
  avgpool_fwd_kernel<<<(unsigned)ceil_div(B * H, 256), 256, 0, as_stream(stream)>>>(r, out, B, (int)S, (int)H);
  MR_CHECK_LAUNCH("avgpool_fwd_kernel");
  return MR_OK;
}

int mr_avgpool_bwd(const float* d_out, float* d_r, int64_t B, int64_t S, int64_t H, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(d_out && d_r, MR_ERR_NULL, "mr_avgpool_bwd: null pointer");
  MR_REQUIRE(B >= 0 && S >= 1 && H >= 1, MR_ERR_BAD_SHAPE, "mr_avgpool_bwd: bad shape");
  if (B == 0) return MR_OK;
  avgpool_bwd_kernel<<<(unsigned)ceil_div(B * S * H, 256), 256, 0, as_stream(stream)>>>(d_out, d_r, B, (int)S, (int)H);
  MR_CHECK_LAUNCH("avgpool_bwd_kernel");
  return MR_OK;
}

int mr_rank_metrics(const float* prob, const float* label, const int64_t* offsets, double* metrics, int32_t* rank,
                    int64_t n_impr, int64_t n_cand, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(prob && label && offsets && metrics, MR_ERR_NULL, "mr_rank_metrics: null pointer");
  MR_REQUIRE(n_impr >= 0 && n_cand >= 0, MR_ERR_BAD_SHAPE, "mr_rank_metrics: bad shape");
  if (n_impr == 0) return MR_OK;
  const int cap = 2048;
  rank_metrics_kernel<<<(unsigned)n_impr, 128, sizeof(float) * 2 * cap, as_stream(stream)>>>(prob, label, offsets, metrics,
                                                                                             rank, n_impr, cap);
  MR_CHECK_LAUNCH("rank_metrics_kernel");
  return MR_OK;
}

}  // extern "C"
