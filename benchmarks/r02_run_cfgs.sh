#!/bin/bash
# Round-2 8-GPU session for BASELINE configs[2..4]: fp8 cast fan-out 1->4 (pull = fp8 on the wire, push = bf16 on the wire),
# Llama-3-70B TP4 shard pairs (8 KiB regions), Mixtral 64k-ctx 1->7.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
run() { # name, nproc, args...
  name=$1; np=$2; shift 2
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $np "$@" \
    > gpurun_out/r02_bench_$name.json 2> gpurun_out/r02_bench_$name.err; echo "$name rc=$?"; grep '^{' gpurun_out/r02_bench_$name.json | cut -c1-420; tail -n 2 gpurun_out/r02_bench_$name.err | cut -c1-300
}
run n5_fp8_pull 5 --steps 20 --warmup 4 --cast fp8 --quick --no-cpu-baseline
run n5_fp8_push 5 --steps 20 --warmup 4 --cast fp8 --direction push --quick --no-cpu-baseline
run n8_pairs70b_pull 8 --steps 20 --warmup 4 --topology pairs --model llama70b-tp4 --quick --no-cpu-baseline
run n8_mixtral64k_pull 8 --steps 5 --warmup 3 --model mixtral --ctx 65536 --quick --no-cpu-baseline
timeout 100 benchmarks/nccl_bcast_baseline --iters 5 --warmup 1 > gpurun_out/r02_nccl_bcast_n8.json 2> gpurun_out/r02_nccl_bcast_n8.err; echo "nccl bcast rc=$?"; cut -c1-500 gpurun_out/r02_nccl_bcast_n8.json
timeout 150 python benchmarks/mc_bench.py --hybrid --iters 8 --out gpurun_out/r02_mc_hybrid_n8.json > gpurun_out/r02_mc_hybrid_n8.log 2>&1; echo "mc hybrid rc=$?"; grep fraction_multicast gpurun_out/r02_mc_hybrid_n8.log | cut -c1-200
