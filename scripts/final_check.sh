#!/bin/bash
# Round-end verification on ONE B200 in one gpurun call (a few GPU-minutes): the GPU test-suite, the driver's bench line,
# smoke(), the per-kernel table, and the same tests + a quick bench with the opt-in tail-backward variant (MINDREC_TAIL_ROWDOT=1).
# Every step has its own time limit and its own log under gpurun_out/; steps are ordered by importance.
#     gpurun --timeout 260 -- 'bash scripts/final_check.sh'
mkdir -p gpurun_out
T0=$SECONDS
stamp() { echo "[$((SECONDS - T0)) s] $*"; }

# 1. correctness (the variant's run shares the GPU with the main one: no timing in either)
( timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_gputests.log 2>&1; echo "rc $?" >> gpurun_out/r2_final_gputests.log ) &
P1=$!
( MINDREC_TAIL_ROWDOT=1 timeout 110 python -m pytest tests -m gpu -x -q > gpurun_out/r2_rowdot_gputests.log 2>&1; echo "rc $?" >> gpurun_out/r2_rowdot_gputests.log ) &
P2=$!
wait $P1 $P2
stamp "gpu tests: default $(tail -n 2 gpurun_out/r2_final_gputests.log | tr '\n' ' ') | rowdot $(tail -n 2 gpurun_out/r2_rowdot_gputests.log | tr '\n' ' ')"

# 2. the bench line as the driver runs it
timeout 100 python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
stamp "bench rc $? $(grep -o '"value": [0-9.]*, "unit": "impressions/s", "n_gpus": 1' gpurun_out/r2_final_bench.json | head -n 1)"

# 3. the variant, headline + e2e + event timing of the three encoder kernels only
MINDREC_TAIL_ROWDOT=1 timeout 60 python bench.py --quick --steps 50 --warmup 5 > gpurun_out/r2_rowdot_bench.json 2> gpurun_out/r2_rowdot_bench.err
stamp "rowdot bench rc $? $(grep -o '"value": [0-9.]*, "unit": "impressions/s", "n_gpus": 1' gpurun_out/r2_rowdot_bench.json | head -n 1)"

# 4. per-kernel table of the final step (each kernel's own run time: no programmatic dependent launch)
MINDREC_PDL=0 BATCH=ids timeout 50 python scripts/step_kernels.py 8 > gpurun_out/r2_final_step_kernels.txt 2>&1
stamp "step_kernels rc $?"

# 5. the driver's smoke()
timeout 50 python __graft_entry__.py smoke > gpurun_out/r2_final_smoke.log 2>&1
stamp "smoke rc $? $(grep -c '^smoke' gpurun_out/r2_final_smoke.log) lines"
