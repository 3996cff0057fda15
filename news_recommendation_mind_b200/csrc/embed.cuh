// Token grouping for the conv backward of the bf16 news encoder (embed.cu): per vocabulary row sums of the
// conv-output gradient, S[v, tap, :] = sum_{t: ids[t] = v} dconv[t + 1 - tap, :] (rows of the same title only).
#pragma once
#include "common.cuh"

namespace mr {

// The PLAN (token positions sorted by id, segment bounds, <= 32-row chunks) depends on the ids only; it may be built ahead
// of time on another stream (mr_token_group_plan) and handed to token_group_taps, or left to it (plan == nullptr).
int64_t token_group_plan_bytes(int64_t T, int64_t V);
int token_group_plan_build(const void* ids, int ids_i64, int64_t T, int64_t V, void* plan, int64_t plan_bytes, cudaStream_t st);

int64_t token_group_workspace_bytes(int64_t T, int64_t Hp, int64_t V);
// dconv: bf16 [T, ld] (ld = Hp, a multiple of 8, 3*Hp <= 512); S: bf16 [>= V rows, 3*ld], every row v < V is written
// (rows of absent tokens as zeros).  Deterministic (sorted positions, fixed-order sums), no atomics, no host sync.
int token_group_taps(const void* ids, int ids_i64, const void* plan, const __nv_bfloat16* dconv, int64_t ld, int L, int64_t T,
                     int64_t V, __nv_bfloat16* S, void* workspace, int64_t workspace_bytes, cudaStream_t st);

}  // namespace mr
