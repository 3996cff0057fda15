#!/bin/bash
# Round-end verification on ONE B200 in one gpurun call (about three GPU-minutes): the GPU test-suite, the driver's bench line,
# the per-kernel table and smoke().  Every step has its own time limit and its own log under gpurun_out/; steps are ordered by
# importance.  (The run kept under profiles/r02_*_final.* also ran the suite and a quick bench a second time with the since
# removed MINDREC_TAIL_ROWDOT=1 variant of cnn_tail_bwd, see profiles/README.md.)
#     gpurun --timeout 260 -- 'bash scripts/final_check.sh'
mkdir -p gpurun_out
T0=$SECONDS
stamp() { echo "[$((SECONDS - T0)) s] $*"; }

# 1. correctness
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_gputests.log 2>&1
echo "rc $?" >> gpurun_out/r2_final_gputests.log
stamp "gpu tests: $(tail -n 2 gpurun_out/r2_final_gputests.log | tr '\n' ' ')"

# 2. the bench line as the driver runs it
timeout 100 python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
stamp "bench rc $? $(grep -o '"value": [0-9.]*, "unit": "impressions/s", "n_gpus": 1' gpurun_out/r2_final_bench.json | head -n 1)"

# 3. per-kernel table of the final step (each kernel's own run time: no programmatic dependent launch)
MINDREC_PDL=0 BATCH=ids timeout 50 python scripts/step_kernels.py 8 > gpurun_out/r2_final_step_kernels.txt 2>&1
stamp "step_kernels rc $?"

# 4. the driver's smoke()
timeout 50 python __graft_entry__.py smoke > gpurun_out/r2_final_smoke.log 2>&1
stamp "smoke rc $? $(grep -c '^smoke' gpurun_out/r2_final_smoke.log) lines"
