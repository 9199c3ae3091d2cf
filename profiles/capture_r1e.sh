#!/bin/bash
# r1e captures (run on the GPU box through gpurun; every ncu pass only after the same command exited 0 without ncu).
# Outputs go to gpurun_out/; the summaries kept under profiles/ are produced from them by profiles/summarise_r1e.py.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-learner > gpurun_out/r1e_bench_plain.json 2> gpurun_out/r1e_bench_plain.err || exit 1
# 1. launch list of the bench command (first 400 launches: warm-up, the timed steps, the host-array e2e leg)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1e_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-learner > gpurun_out/r1e_ncu_launches.log 2>&1
# 2. full capture of the two kernels of the timed step (T = 128 launches only: skip the 8-step chunks of the e2e leg)
ncu --set full --clock-control none --import-source on -k regex:'checkers_rollout_kernel|gae_tile_kernel' -s 6 -c 4 \
    -o gpurun_out/r1e_env_gae python bench.py --steps 2 --warmup 1 --no-learner > gpurun_out/r1e_ncu_env.log 2>&1
# 3. learner kernels new in r1e: GRU tile kernels, tensor-core mixer
python profiles/run_learner_kernels_r1e.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:'gru_fwd_tile|gru_bwd_tile|qmixer_tc' -s 3 -c 3 \
    -o gpurun_out/r1e_learner python profiles/run_learner_kernels_r1e.py > gpurun_out/r1e_ncu_learner.log 2>&1
ls -la gpurun_out
