cd /root/repo 2>/dev/null || cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/r02_bench_n1.json
timeout 300 python benchmarks/kvbench.py --num-blocks 1,128 --tokens-per-block 16 --direction h2d,d2d --out gpurun_out/r02_kvbench_h2d_d2d.csv > gpurun_out/r02_kvbench.log 2>&1; echo "kvbench rc=$?"; cat gpurun_out/r02_kvbench_h2d_d2d.csv | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_ncu_launches_bench_n1.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --quick > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
for w in ours ref; do
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -c 3 -f -o gpurun_out/r02_ncu_n1_$w python benchmarks/profile_one.py --which $w > gpurun_out/r02_ncu_n1_$w.log 2>&1; echo "ncu $w rc=$?"
done
ls -la gpurun_out/*.ncu-rep
