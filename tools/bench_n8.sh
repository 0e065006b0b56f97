TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 170 $TR --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "train rc=$?"
timeout 170 $TR --master-port 29512 bench.py --gpus 8 --config xl --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_xl_n8.json 2> gpurun_out/r02_bench_xl_n8.err; echo "xl rc=$?"
timeout 200 $TR --master-port 29513 bench.py --gpus 8 --config multigrid --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_multigrid_n8.json 2> gpurun_out/r02_bench_multigrid_n8.err; echo "mg rc=$?"
tail -c 300 gpurun_out/r02_bench_n8.json
