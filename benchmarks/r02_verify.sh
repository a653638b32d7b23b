#!/bin/bash
# What the driver runs at round end, on one GPU: the GPU parity suite, smoke(), the default bench line and the reference arm.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_1gpu.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r02_pytest_gpu_1gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r02_smoke.log
timeout 400 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; cut -c1-2500 gpurun_out/r02_bench_n1.json; tail -n 5 gpurun_out/r02_bench_n1.err
timeout 400 python bench.py --impl reference > gpurun_out/r02_bench_n1_reference.json 2> gpurun_out/r02_bench_n1_reference.err; echo "ref rc=$?"; cut -c1-800 gpurun_out/r02_bench_n1_reference.json
