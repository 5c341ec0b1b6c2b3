#!/bin/bash
# One measurement pass on the GPU box (run under gpurun from the repo root): bench line, reference arm, ncu launch list of the bench
# command, one `ncu --set full` capture per workload. Everything lands in gpurun_out/ with the prefix given as $1.
# Numbers printed by a run under ncu are never bench values: the bench runs first, without a profiler.
P=${1:-r2f}
O=gpurun_out
python bench.py > $O/${P}_bench.json 2> $O/${P}_bench.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > $O/${P}_bench_ref.json 2>> $O/${P}_bench.err
BENCH_ARGS="--steps 2 --warmup 3 --spp 64 --side= --no-cpu-baseline --no-e2e"
python bench.py $BENCH_ARGS > $O/${P}_bench_short.json 2>> $O/${P}_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/${P}_launches_c3.csv python bench.py $BENCH_ARGS > $O/${P}_ncu_launches.log 2>&1
for wl in c3 c4 c5; do python scripts/profile_step.py $wl 8 > $O/${P}_plain_$wl.log 2>&1 || exit 1; done
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:k_bounce_small --launch-skip 3 -c 3 -o $O/${P}_prof_c3 python scripts/profile_step.py c3 8 > $O/${P}_ncu_c3.log 2>&1
$NCU -k 'regex:^k_primary$' --launch-skip 1 -c 1 -o $O/${P}_prof_c3_primary python scripts/profile_step.py c3 8 > $O/${P}_ncu_c3p.log 2>&1
$NCU -k 'regex:k_trace8|k_shade_surface|k_raygen' --launch-skip 10 -c 10 -o $O/${P}_prof_c4 python scripts/profile_step.py c4 8 > $O/${P}_ncu_c4.log 2>&1
$NCU -k regex:k_volume_paths --launch-skip 1 -c 1 -o $O/${P}_prof_c5 python scripts/profile_step.py c5 8 > $O/${P}_ncu_c5.log 2>&1
ls -la $O/${P}_*
