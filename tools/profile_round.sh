#!/bin/bash
# Runs on the GPU box (gpurun -- 'bash tools/profile_round.sh'): everything profiles/ records for a round.
# Outputs land in gpurun_out/; tools/profile_collect.py (run in the build container) turns them into profiles/ files.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/pytest_gpu.log
python bench.py --impl reference --steps 64 --warmup 8 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py > $O/bench.json 2> $O/bench.err
python bench.py --raw-obs 0 --skip-extras > $O/bench_noraw.json 2>> $O/bench.err
python bench.py --rollout 128 --skip-extras > $O/bench_T128.json 2>> $O/bench.err
# launch list of the bench command (short run: ncu serialises and replays)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 4096 --warmup 1024 --skip-extras > $O/bench_under_ncu.json 2> $O/ncu_launches.err
# one full capture of the dominant kernel (4th launch: steady state)
ncu --set full --import-source on --clock-control none -k regex:atc_rollout_pipe -s 2 -c 1 -f -o $O/pipe_full \
    python bench.py --steps 4096 --warmup 1024 --skip-extras > $O/ncu_full.log 2>&1
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python __graft_entry__.py smoke > $O/sanitizer_memcheck.log 2>&1
timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python __graft_entry__.py smoke > $O/sanitizer_racecheck.log 2>&1
tail -2 $O/pytest_gpu.log; tail -1 $O/sanitizer_memcheck.log; tail -1 $O/sanitizer_racecheck.log
