#!/bin/bash
# Round-2 two-GPU session: multi-GPU parity tests, 1->1 bench (pull and push), NVLink byte counters under ncu.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multicast.py -x -q > gpurun_out/r02_pytest_gpu_multi_2gpu.log 2>&1; echo "pytest multi rc=$?"; tail -n 12 gpurun_out/r02_pytest_gpu_multi_2gpu.log
timeout 600 $TR bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02_bench_n2_pull.json 2> gpurun_out/r02_bench_n2_pull.err; echo "bench pull rc=$?"; cut -c1-2500 gpurun_out/r02_bench_n2_pull.json; tail -n 5 gpurun_out/r02_bench_n2_pull.err
timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 --direction push --quick --no-cpu-baseline > gpurun_out/r02_bench_n2_push.json 2> gpurun_out/r02_bench_n2_push.err; echo "bench push rc=$?"; cut -c1-900 gpurun_out/r02_bench_n2_push.json
for m in pull push; do
  timeout 200 ncu --metrics gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:paged_copy -c 6 --csv \
    --log-file gpurun_out/r02_ncu_nvlink_$m.csv benchmarks/copylab --peer $m --only lib --iters 2 > gpurun_out/r02_ncu_nvlink_$m.log 2>&1; echo "ncu nvlink $m rc=$?"; tail -n 4 gpurun_out/r02_ncu_nvlink_$m.csv | cut -c1-400
done
