#!/bin/bash
# Runs on the GPU box (under gpurun): bench line, per-kernel tables, micro-benchmarks, ncu launch list and one
# ncu --set full capture of the dominant depthwise kernel.  Outputs land in gpurun_out/ (copied to profiles/ by hand).
set -u
R=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --kernel-table gpurun_out/${R}_kernel_table.json > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err
python tools/dw_microbench.py --json gpurun_out/${R}_dw_microbench.json > /dev/null 2>&1
python tools/pw_microbench.py --json gpurun_out/${R}_pw_microbench.json > /dev/null 2>&1
# launch list of the bench command (same command line ran above without ncu and exited 0)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
# full capture of the top depthwise kernels (stride-2 forward of layer1.0 and a stride-1 layer)
python tools/dw_microbench.py --layer l1.0 --iters 3 > gpurun_out/plain_dw.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dw3_ -s 9 -c 3 -o gpurun_out/${R}_dw_l1_0 -f \
    python tools/dw_microbench.py --layer l1.0 --iters 3 > gpurun_out/ncu_dw.log 2>&1
echo "ncu dw rc=$?"
# DRAM traffic of the depthwise C-ABI calls (tiled kernels dw3_*: fwd / dgrad / wgrad), eager steps
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:dw3_ \
    -c 400 --csv --log-file gpurun_out/${R}_dw_traffic.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic rc=$?"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${R}_smi.csv
