mkdir -p gpurun_out
T=r1k
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 600 gpurun_out/bench_$T.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${T}_reference.json 2>> gpurun_out/bench_$T.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_bench_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:window_counts_kernel -s 2 -c 1 -f -o gpurun_out/prof_${T}_window python tools/prof_find.py > gpurun_out/ncu_w_$T.log 2>&1
TIDK_PROF_CFG=c5 ncu --set full --clock-control none --import-source on -k regex:explore_planes -s 2 -c 1 -f -o gpurun_out/prof_${T}_explore python tools/prof_explore.py > gpurun_out/ncu_e_$T.log 2>&1
ls -la gpurun_out/*$T*
