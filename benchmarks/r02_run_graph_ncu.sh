#!/bin/bash
# CUDA-graph capture test + one ncu --set full capture of the permuting kernel (1 GPU)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_paged.py -x -q -k "graph" > gpurun_out/r02_pytest_graph.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/r02_pytest_graph.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:paged_permute -s 4 -c 2 -f -o gpurun_out/r02_ncu_permute python benchmarks/permute_bench.py --iters 2 > gpurun_out/r02_ncu_permute.log 2>&1; echo "ncu rc=$?"; tail -n 3 gpurun_out/r02_ncu_permute.log
ls -la gpurun_out/*.ncu-rep
