#!/bin/bash
mkdir -p gpurun_out/r02
T=${1:-v15a}
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra-legs --warps 1 > gpurun_out/r02/probe_${T}_w1.json 2> gpurun_out/r02/probe_${T}_w1.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra-legs --warps 2 --rpb 7 > gpurun_out/r02/probe_${T}_w2.json 2> gpurun_out/r02/probe_${T}_w2.err
M=sm__icc_request_hit_rate.pct,sm__icc_requests.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,gpu__time_duration.sum,launch__registers_per_thread
for w in 1 2; do
timeout 600 ncu --metrics $M --clock-control none -k regex:mc_sim_kernel --launch-skip 3 -c 1 --csv --log-file gpurun_out/r02/ncu_${T}_w$w.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-legs --warps $w --rpb 7 > /dev/null 2>&1
done
python - <<PY
import json
for w in (1,2):
    d=json.load(open('gpurun_out/r02/probe_${T}_w%d.json'%w))
    print(w, '%.4g'%d['value'], d['ms_per_step'], d['launch'])
PY
grep -h "icc_request_hit\|no_instruction\|issue_active\|inst_executed" gpurun_out/r02/ncu_${T}_w*.csv | cut -d, -f5,13- 
