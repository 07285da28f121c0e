#!/bin/bash
# evidence pass: launch list of a short bench, full ncu capture of tools/prof_r02.py summarised ON the box (the .ncu-rep is larger
# than what gpurun copies back).  TIDK_FINAL_BENCH=1 also runs the bench line and the reference arm first.
mkdir -p gpurun_out
if [ "${TIDK_FINAL_BENCH:-0}" = "1" ]; then
  timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
  timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
fi
export TIDK_BENCH_PASSES=4
timeout 500 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
unset TIDK_BENCH_PASSES
timeout 300 python tools/prof_r02.py > gpurun_out/r02_prof_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none -k regex:"explore_vertical_kernel|window_counts_kernel|window_counts_flat_kernel|explore_planes_kernel|planes_build" -c 26 -o /tmp/r02_prof python tools/prof_r02.py > gpurun_out/r02_ncu_full.log 2>&1
echo "full capture rc=$?"
python tools/ncu_summary_all.py /tmp/r02_prof.ncu-rep gpurun_out/r02_ncu_full_prof_r02.json "python tools/prof_r02.py under ncu --set full --clock-control none (final build of round 2): launches 0-4 C5-shaped reads (K1 planes_build, explore_vertical<5,12> x3 with TMA staging, <0,0> with k = 6), 5 K1 on the C4 genome x0.25 (772 Mb), 6-8 flat window counter TTAGGG W = 10000 on resident planes, 9-10 W = 1000, 11-12 W = 2500, 13-14 FlatRuntime<12> GATTACAGATTA, 15-17 per-window kernel, ASCII source, W = 10000 (TIDK_PLANES_COUNT=0), 18-19 W = 1000, 20-21 W = 2500, 22-23 RuntimeMatcher<12> ASCII source, 24-25 fused Hymenoptera kernel (15 repeats, per-window form on the resident planes)"
tail -3 gpurun_out/r02_prof_plain.log
