#!/bin/bash
# Round-2 N-GPU session (N = 4 or 8): fan-out bench in both directions, replicate + NVLS, NVLS tuning sweep, ncclBcast baseline.
N=${1:-8}; mkdir -p gpurun_out; cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 420 $TR bench.py --gpus $N --steps 20 --warmup 4 > gpurun_out/r02_bench_n${N}_pull.json 2> gpurun_out/r02_bench_n${N}_pull.err; echo "bench pull rc=$?"; cut -c1-1200 gpurun_out/r02_bench_n${N}_pull.json; tail -n 4 gpurun_out/r02_bench_n${N}_pull.err
timeout 240 $TR bench.py --gpus $N --steps 20 --warmup 4 --direction push --quick --no-cpu-baseline > gpurun_out/r02_bench_n${N}_push.json 2> gpurun_out/r02_bench_n${N}_push.err; echo "bench push rc=$?"; cut -c1-600 gpurun_out/r02_bench_n${N}_push.json
timeout 200 python benchmarks/nvls_staged.py --steps 20 --out gpurun_out/r02_nvls_staged_n${N}.json 2>&1 | tail -n 2 | cut -c1-600
if [ "$2" != "short" ]; then
  timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 4 --replicate --no-cpu-baseline > gpurun_out/r02_bench_n${N}_rep.json 2> gpurun_out/r02_bench_n${N}_rep.err; echo "bench rep rc=$?"; cut -c1-600 gpurun_out/r02_bench_n${N}_rep.json
  timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 4 --replicate --nvls --quick --no-cpu-baseline > gpurun_out/r02_bench_n${N}_nvls.json 2> gpurun_out/r02_bench_n${N}_nvls.err; echo "bench nvls rc=$?"; cut -c1-600 gpurun_out/r02_bench_n${N}_nvls.json; tail -n 3 gpurun_out/r02_bench_n${N}_nvls.err
  timeout 120 benchmarks/nccl_bcast_baseline --iters 5 --warmup 1 > gpurun_out/r02_nccl_bcast_n${N}.json 2> gpurun_out/r02_nccl_bcast_n${N}.err; echo "nccl bcast rc=$?"; cat gpurun_out/r02_nccl_bcast_n${N}.json | cut -c1-500
  timeout 240 python benchmarks/mc_bench.py --sweep --iters 6 --out gpurun_out/r02_mc_sweep_n${N}.json > gpurun_out/r02_mc_sweep_n${N}.log 2>&1; echo "mc sweep rc=$?"; tail -n 12 gpurun_out/r02_mc_sweep_n${N}.log
fi
