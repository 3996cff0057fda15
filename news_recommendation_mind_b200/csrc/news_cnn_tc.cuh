// MR_BF16 (tcgen05) implementation of the CNN news encoder -- see news_cnn_tc.cu.
#pragma once
#include "common.cuh"

namespace mr {
int64_t news_cnn_tc_workspace_bytes(const mr_cnn_shape* s, int backward);     // backward: 0 fwd, 1 bwd, 2 bwd with d_table
int news_cnn_tc_fwd(const mr_cnn_shape* s, const void* ids, int ids_i64, const float* emb, const void* mask, int mask_i64,
                    const void* table, const float* conv_w, const float* conv_b, const float* proj_w, const float* proj_b,
                    const float* query, void* c_save, void* key_save, float* prob, float* news, void* ws, int64_t wsb,
                    cudaStream_t st);
int news_cnn_tc_bwd(const mr_cnn_shape* s, const void* ids, int ids_i64, const float* emb, const void* table,
                    const float* conv_w, const float* proj_w, const float* query, const void* c_save, const void* key_save,
                    const float* prob, const float* d_news, const float* d_c, float* d_conv_w, float* d_conv_b,
                    float* d_proj_w, float* d_proj_b, float* d_query, void* d_emb, void* ws, int64_t wsb, cudaStream_t st,
                    float* d_table = nullptr, int64_t table_rows = 0, int64_t padding_idx = -1,
                    const void* group_plan = nullptr, void* table_ready_event = nullptr);
}  // namespace mr
