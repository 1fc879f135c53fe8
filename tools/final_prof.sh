set -x
SLU_STEP_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_step_launches_b256_v2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-gpu > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad -c 3 -o gpurun_out/wgrad_tma python tools/wgrad_only.py 0 1 > gpurun_out/wgrad_tma_ncu.log 2>&1
for c in 3 2 4 5; do timeout 400 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2b_bench_config$c.json 2> gpurun_out/r2b_bench_config$c.err; tail -c 300 gpurun_out/r2b_bench_config$c.json; echo; done
