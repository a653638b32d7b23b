#!/bin/bash
# two GPUs: every multi-GPU parity test (incl. the permuting pull) + the layout transform over NVLink, pushed and pulled
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m "gpu and multigpu" -x -q > gpurun_out/r02_pytest_gpu_multi_2gpu.log 2>&1; echo "pytest multi rc=$?"; tail -n 8 gpurun_out/r02_pytest_gpu_multi_2gpu.log
for side in dst src; do
  timeout 200 python benchmarks/permute_bench.py --peer 1 --peer-side $side > gpurun_out/r02_permute_bench_peer_$side.json 2> gpurun_out/r02_permute_bench_peer_$side.err; echo "permute $side rc=$?"; cat gpurun_out/r02_permute_bench_peer_$side.json; tail -n 3 gpurun_out/r02_permute_bench_peer_$side.err
done
