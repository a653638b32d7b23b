#!/bin/bash
# Round-2 single-GPU session: design study, parity suite, bench line, ncu evidence.  Everything lands in gpurun_out/.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
rm -f gpurun_out/r02_copylab_n1_c.jsonl
for w in lib ref perm; do timeout 120 benchmarks/copylab --iters 20 --only $w >> gpurun_out/r02_copylab_n1_c.jsonl 2>> gpurun_out/r02_copylab_n1_c.err; done
cat gpurun_out/r02_copylab_n1_c.jsonl; tail -n 3 gpurun_out/r02_copylab_n1_c.err
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_1gpu.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/r02_pytest_gpu_1gpu.log
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r02_bench_n1.json; tail -n 5 gpurun_out/r02_bench_n1.err
if [ "$1" != "quick" ]; then
  timeout 300 python benchmarks/kvbench.py --num-blocks 1,128 --tokens-per-block 16 --direction h2d,d2d --out gpurun_out/r02_kvbench_h2d_d2d.csv > gpurun_out/r02_kvbench.log 2>&1; echo "kvbench rc=$?"; tail -n 30 gpurun_out/r02_kvbench_h2d_d2d.csv
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_ncu_launches_bench_n1.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --quick > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
  for w in ours ref; do
    timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -c 3 -f -o gpurun_out/r02_ncu_n1_$w python benchmarks/profile_one.py --which $w > gpurun_out/r02_ncu_n1_$w.log 2>&1; echo "ncu $w rc=$?"
  done
fi
