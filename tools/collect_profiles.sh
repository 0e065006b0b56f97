#!/bin/bash
# Runs on the GPU box (under gpurun): bench lines of every config, per-kernel tables, micro-benchmarks, the ncu launch list
# of the bench command, DRAM-traffic counters of the conv kernels and one ncu --set full capture of the dominant kernel
# of each family.  Outputs land in gpurun_out/; tools/summarize_profiles.py turns them into the tracked files under profiles/.
set -u
R=${1:-r02}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --kernel-table gpurun_out/${R}_kernel_table.json > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err
echo "reference rc=$?"
python bench.py --config multigrid --steps 5 --warmup 3 > gpurun_out/${R}_bench_multigrid.json 2>> gpurun_out/${R}_bench.err
python bench.py --config xl --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_xl.json 2>> gpurun_out/${R}_bench.err
python bench.py --config charades --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_charades.json 2>> gpurun_out/${R}_bench.err
python bench.py --version S --batch 16 --frames 13 --crop 160 --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/${R}_bench_s.json 2>> gpurun_out/${R}_bench.err
echo "configs done"
python tools/dw_microbench.py --json gpurun_out/${R}_dw_microbench.json > /dev/null 2>&1
python tools/pw_microbench.py --json gpurun_out/${R}_pw_microbench.json > /dev/null 2>&1
python tools/stem_microbench.py > gpurun_out/${R}_stem_microbench.jsonl 2>/dev/null
python tools/ew_microbench.py > gpurun_out/${R}_ew_microbench.jsonl 2>/dev/null
python tools/head_gemm_microbench.py > gpurun_out/${R}_head_gemm_microbench.txt 2>&1
python tools/head_gemm_microbench.py --rows 64 >> gpurun_out/${R}_head_gemm_microbench.txt 2>&1
python tools/head_gemm_microbench.py --rows 256 >> gpurun_out/${R}_head_gemm_microbench.txt 2>&1
python tools/marginal_cost.py --zero-input --json gpurun_out/${R}_marginal_cost.json > /dev/null 2>&1
python tools/split_concurrency_probe.py > gpurun_out/${R}_split_concurrency_probe.txt 2>&1
echo "microbenchmarks done"
# launch list of the bench command (same command line ran above without ncu and exited 0)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
# DRAM traffic of the conv C-ABI calls (tiled depthwise kernels dw3_*, tcgen05 GEMMs pw_*), one eager step
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'dw3_|pw_' \
    -c 600 --csv --log-file gpurun_out/${R}_conv_traffic.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-parity > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic rc=$?"
# full captures: pointwise wgrad (dominant class by time) on the 803 k-row stage-1 layer, depthwise wgrad / forward on a stride-1 layer
python tools/pw_microbench.py --layer l1.c3 --only wgrad --iters 3 > gpurun_out/plain_pw.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pw_wgrad_tc -s 3 -c 1 -o gpurun_out/${R}_pw_wgrad_l1c3 -f \
    python tools/pw_microbench.py --layer l1.c3 --only wgrad --iters 3 > gpurun_out/ncu_pw.log 2>&1
echo "ncu pw rc=$?"
python tools/dw_microbench.py --layer l2.x --iters 3 > gpurun_out/plain_dw.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dw3_ -s 9 -c 3 -o gpurun_out/${R}_dw_l2x -f \
    python tools/dw_microbench.py --layer l2.x --iters 3 > gpurun_out/ncu_dw.log 2>&1
echo "ncu dw rc=$?"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${R}_smi.csv
cuobjdump -sass x3d_multigrid_b200/libx3d_b200.so 2>/dev/null | grep -oE '^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+' | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn > gpurun_out/${R}_sass_mnemonics.txt
echo "done"
