#!/bin/bash
# Runs on the GPU box (gpurun -- 'bash tools/profile_round.sh'): everything profiles/ records for a round.
# Outputs land in gpurun_out/; tools/profile_collect.py (run in the build container) turns them into profiles/ files.
# bench.py semantics: one step = one rollout launch of --rollout (1024) env-steps of the whole batch.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/pytest_gpu.log
# the two arms exactly as the driver launches them
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
# device-resident variations: no info[original_state]; short launches
python bench.py --raw-obs 0 --skip-extras > $O/bench_noraw.json 2>> $O/bench.err
python bench.py --rollout 128 --skip-extras > $O/bench_T128.json 2>> $O/bench.err
python bench.py --rollout 20 --skip-extras > $O/bench_T20.json 2>> $O/bench.err
python tools/launch_length_probe.py > $O/launch_length.json 2> $O/launch_length.err
# launch list of the bench command (short run: ncu serialises and replays; 3 warm-up + 2 timed rollout launches)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --blocks 1 --skip-extras --skip-parity > $O/bench_under_ncu.json 2> $O/ncu_launches.err
# one full capture of the dominant kernel (4th launch: steady state)
ncu --set full --import-source on --clock-control none -k regex:atc_rollout_pipe -s 3 -c 1 -f -o $O/pipe_full \
    python bench.py --steps 2 --warmup 3 --blocks 1 --skip-extras --skip-parity > $O/ncu_full.log 2>&1
for cfg in 4096x1 16384x8_wind; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $O/launches_$cfg.csv \
      python bench.py --config $cfg --steps 2 --warmup 3 --blocks 1 --skip-extras --skip-parity > /dev/null 2>> $O/ncu_launches.err
done
# sanitizer: smoke (pair-per-CTA layout) and the one-CTA-per-SM layout forced on the same small case
timeout 180 compute-sanitizer --tool memcheck --print-limit 5 python __graft_entry__.py smoke > $O/sanitizer_memcheck.log 2>&1
ATC_B200_BIG_MIN_PAIRS=1 ATC_B200_BIG_MIN_STEPS=1 timeout 180 compute-sanitizer --tool memcheck --print-limit 5 \
    python __graft_entry__.py smoke > $O/sanitizer_memcheck_big.log 2>&1
timeout 180 compute-sanitizer --tool racecheck --print-limit 5 python __graft_entry__.py smoke > $O/sanitizer_racecheck.log 2>&1
ATC_B200_BIG_MIN_PAIRS=1 ATC_B200_BIG_MIN_STEPS=1 timeout 180 compute-sanitizer --tool racecheck --print-limit 5 \
    python __graft_entry__.py smoke > $O/sanitizer_racecheck_big.log 2>&1
tail -2 $O/pytest_gpu.log; tail -1 $O/sanitizer_memcheck.log; tail -1 $O/sanitizer_memcheck_big.log
tail -1 $O/sanitizer_racecheck.log; tail -1 $O/sanitizer_racecheck_big.log
cat $O/bench.json | head -c 1500
