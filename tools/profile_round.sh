#!/bin/bash
# One GPU call that refreshes the round's headline evidence under gpurun_out/ (copy what should be judged to profiles/):
# the bench line, the in-graph GEMM shape table, and the ncu launch list (durations + DRAM bytes) of one eager step.
mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/bench_n1.err
SPMM_BENCH_GEMM_TABLE=gpurun_out/r2_gemm_shape_table.txt python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_tab.json 2>/dev/null
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2400 --csv \
    --log-file gpurun_out/r2_launches_dram.csv python bench.py --profile --eager --steps 1 --warmup 0 > gpurun_out/ncu_run.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_dram.csv > gpurun_out/r2_launches_summary.txt
head -14 gpurun_out/r2_launches_summary.txt
cat gpurun_out/r2_bench_n1.json
