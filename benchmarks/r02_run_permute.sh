#!/bin/bash
# layout-transforming hand-off: parity suite + timing (1 GPU)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_permute.py -x -q > gpurun_out/r02_pytest_gpu_permute.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/r02_pytest_gpu_permute.log
timeout 200 python benchmarks/permute_bench.py > gpurun_out/r02_permute_bench.json 2> gpurun_out/r02_permute_bench.err; echo "bench rc=$?"; cat gpurun_out/r02_permute_bench.json; tail -n 5 gpurun_out/r02_permute_bench.err
