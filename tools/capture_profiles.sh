#!/usr/bin/env bash
# Round-end evidence for profiles/: (1) the bench line, (2) the ncu launch list of the same command, (3) one ncu --set full
# capture each of the step kernel, the fused step + ring push, the policy kernel and the two transposing gathers.
# Run on the GPU box from the repo root: bash tools/capture_profiles.sh TAG
set -u
TAG="${1:-rX}"
OUT=gpurun_out
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err || { echo "bench failed"; tail -5 $OUT/${TAG}_bench.err; exit 1; }
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 || { echo "short bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_list.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
$NCU --kernel-name-base demangled -k 'regex:k_env_t<\(int\)0.*\(bool\)0, \(bool\)0>' --launch-skip 4 -c 1 -f -o $OUT/${TAG}_step \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-config2 --no-config3 --no-config4 --no-strong --no-rollout --no-obs > $OUT/${TAG}_ncu_step.log 2>&1
$NCU --kernel-name-base demangled -k 'regex:k_env_t<\(int\)0.*\(bool\)0, \(bool\)1>' --launch-skip 4 -c 2 -f -o $OUT/${TAG}_step_ring \
    python tools/profile_rollout.py 131072 8 > $OUT/${TAG}_ncu_step_ring.log 2>&1
$NCU -k 'regex:k_policy|k_window_gather|k_em_gather|k_transition_tail' --launch-skip 12 -c 6 -f -o $OUT/${TAG}_rollout \
    python tools/profile_rollout.py 131072 6 > $OUT/${TAG}_ncu_rollout.log 2>&1
ls -la $OUT/${TAG}_*
