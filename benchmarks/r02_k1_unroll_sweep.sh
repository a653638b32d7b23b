cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; rm -f gpurun_out/r02_k1_unroll_sweep.txt
for T in 256 128; do for U in 8 4 2 1; do
  echo "== threads=$T unroll=$U" >> gpurun_out/r02_k1_unroll_sweep.txt
  KVBM_K1_THREADS=$T KVBM_K1_UNROLL=$U timeout 100 python benchmarks/kvbench.py --num-blocks 1,128 --tokens-per-block 16 --direction h2d,d2h,d2d --backend vectorized --pattern lw_to_fc --impl ours --iters 60 2>/dev/null | grep ours | cut -d, -f3,5,11,12 >> gpurun_out/r02_k1_unroll_sweep.txt
done; done
echo "== reference" >> gpurun_out/r02_k1_unroll_sweep.txt
timeout 100 python benchmarks/kvbench.py --num-blocks 1,128 --tokens-per-block 16 --direction h2d,d2h,d2d --backend vectorized --pattern lw_to_fc --impl reference --iters 60 2>/dev/null | grep reference | cut -d, -f3,5,11,12 >> gpurun_out/r02_k1_unroll_sweep.txt
cat gpurun_out/r02_k1_unroll_sweep.txt
