#!/bin/bash
# Round profiles (run on the GPU box from the repo root): the bench line, ncu launch lists of the bench and of every kernel
# family, and one `ncu --set full` capture of the hot kernels.  Outputs land in gpurun_out/ (scratch); the summaries that are
# judged are condensed from them into profiles/ with tools/ncu_summary.py and tools/ncu_lines.py.
#   gpurun --timeout 1800 -- 'bash tools/collect_profiles.sh r02'
R=${1:-r02}
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/${R}_bench_reference_arm.json 2> /dev/null
python tools/probe_tma.py > $O/${R}_probe_tma.txt 2>&1
python tools/time_ranker.py 4096 > $O/${R}_time_ranker.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${R}_launches_bench.csv \
    python bench.py --steps 20 --warmup 5 --no-secondary > $O/${R}_ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${R}_launches_paths.csv \
    python tools/bench_paths.py --profile > $O/${R}_ncu_paths.log 2>&1
cap() {  # name, kernel regex, launches to skip, command...
    local name=$1 rx=$2 skip=$3; shift 3
    ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c 1 -o $O/${R}_$name -f "$@" > $O/${R}_$name.log 2>&1
}
cap ranker_fused ranker_fused_kernel 6 python tools/time_ranker.py 4096
cap lambda lambda_kernel 4 python tools/bench_paths.py --only lambda --profile
cap risk risk 4 python tools/bench_paths.py --only georisk --profile
cap approx_sym approx_ndcg_sym_kernel 30 python bench.py --steps 20 --warmup 5 --no-secondary
cap ndcg_key ndcg_key_kernel 4 python tools/bench_paths.py --only ndcg --profile
ls -la $O/${R}_*
