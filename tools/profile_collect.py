#!/usr/bin/env python
"""Turns the outputs of tools/profile_round.sh (gpurun_out/) into the tracked files under profiles/:
   rN_bench*.json (the bench lines), rN_launches.csv (ncu launch list), rN_pipe_kernel_ncu_metrics.json (selected raw
   metrics of the `ncu --set full` capture), rN_pipe_kernel_roles.txt (per-role stall breakdown) and traffic.json
   (DRAM bytes per launch, read by bench.py into roofline.traffic).   Usage: python tools/profile_collect.py r1"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
KEEP = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'launch__block_size', 'launch__grid_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'lts__t_bytes.sum', 'dram__bytes_read.sum.per_second',
        'dram__bytes_write.sum.per_second']


def main(tag):
    for src, dst in (('bench.json', '%s_bench.json'), ('bench_reference.json', '%s_bench_reference.json'),
                     ('bench_noraw.json', '%s_bench_noraw.json'), ('bench_T128.json', '%s_bench_T128.json'),
                     ('bench_T20.json', '%s_bench_T20.json'), ('launch_length.json', '%s_launch_length.json'),
                     ('launches.csv', '%s_launches.csv'), ('launches_4096x1.csv', '%s_launches_4096x1.csv'),
                     ('launches_16384x8_wind.csv', '%s_launches_16384x8_wind.csv'),
                     ('sanitizer_memcheck.log', '%s_sanitizer_memcheck.log'),
                     ('sanitizer_memcheck_big.log', '%s_sanitizer_memcheck_big.log'),
                     ('pytest_gpu.log', '%s_pytest_gpu.log')):
        if os.path.exists(os.path.join(G, src)):
            shutil.copy(os.path.join(G, src), os.path.join(P, dst % tag))
    for name in ('sanitizer_racecheck', 'sanitizer_racecheck_big'):
        rc = os.path.join(G, name + '.log')
        if os.path.exists(rc):       # keep the summary, not the hundreds of "and Read access" lines
            keep = [l for l in open(rc) if not l.startswith('=========     ')]
            open(os.path.join(P, '%s_%s.log' % (tag, name)), 'w').writelines(keep)
    rep = os.path.join(G, 'pipe_full.ncu-rep')
    raw, src = os.path.join(G, 'pipe_full_raw.csv'), os.path.join(G, 'pipe_full_src.csv')
    subprocess.run('ncu -i %s --page raw --csv > %s 2>/dev/null' % (rep, raw), shell=True, check=True)
    subprocess.run('ncu -i %s --page source --csv > %s 2>/dev/null' % (rep, src), shell=True, check=True)
    rows = list(csv.reader(open(raw)))
    m = {}
    for h, u, v in zip(rows[0], rows[1], rows[2]):
        if h in KEEP or h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
            m[h] = {'value': v, 'unit': u}
    m['kernel'] = rows[2][rows[0].index('Kernel Name')]
    json.dump(m, open(os.path.join(P, '%s_pipe_kernel_ncu_metrics.json' % tag), 'w'), indent=1)

    def val(name):
        v, u = float(m[name]['value'].replace(',', '')), m[name]['unit']
        return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[u]
    bench = json.loads(open(os.path.join(G, 'bench_under_ncu.json')).read().strip().splitlines()[-1])
    T = bench['roofline']['launch']['n_steps']
    rd, wr = val('dram__bytes_read.sum'), val('dram__bytes_write.sum')
    cfg = bench['config']
    # bench.py copies dram_bytes_per_launch into roofline.traffic only when its own timed launch has exactly this shape
    json.dump({'kernel_name': m['kernel'],
               'launch': {'n_envs': cfg['envs_per_gpu'], 'n_aircraft': cfg['aircraft_per_env'], 'n_steps': T,
                          'raw_obs': int(bench['roofline']['launch']['raw_obs']),
                          'kernel': bench['roofline']['launch']['kernel']},
               'step_outputs': cfg['step_outputs'],
               'dram_bytes_read': int(rd), 'dram_bytes_written': int(wr), 'dram_bytes_per_launch': int(rd + wr),
               'algorithmic_bytes_per_launch': bench['roofline']['algorithmic_bytes_per_launch'],
               'source': 'profiles/%s_pipe_kernel_ncu_metrics.json (ncu --set full --clock-control none)' % tag},
              open(os.path.join(P, 'traffic.json'), 'w'), indent=1)
    n_iter = 2048.0 * T          # pair-steps of one launch: 65536 aircraft lanes / 32 x T
    m_inst = float(m['smsp__inst_executed.sum']['value'].replace(',', ''))
    print('warp instructions per pair-step: %.1f' % (m_inst / n_iter))
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_roles.py'), src, raw, str(n_iter)],
                         capture_output=True, text=True).stdout
    open(os.path.join(P, '%s_pipe_kernel_roles.txt' % tag), 'w').write(out)
    print(out[:3000])


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'r1')
