mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { name=$1; shift; timeout 240 "$@" 2>&1 | grep -E '^\{|Error|error|Traceback' | tail -3 | tee gpurun_out/$name.json; }
run bench_n8 $TR --nproc-per-node 8 --master-port 29601 bench.py --gpus 8 --steps 20 --warmup 5
run bench_n8_rep $TR --nproc-per-node 8 --master-port 29602 bench.py --gpus 8 --steps 20 --warmup 5 --replicate --no-cpu-baseline
run bench_n4 $TR --nproc-per-node 4 --master-port 29603 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline
run bench_n8_pairs70b $TR --nproc-per-node 8 --master-port 29604 bench.py --gpus 8 --steps 20 --warmup 5 --topology pairs --model llama70b-tp4 --no-cpu-baseline
run bench_n5_fp8 $TR --nproc-per-node 5 --master-port 29605 bench.py --gpus 5 --steps 20 --warmup 5 --cast fp8 --no-cpu-baseline
timeout 300 $TR --nproc-per-node 8 --master-port 29606 benchmarks/fanout_sweep.py --ctx 1024,4096,16384,65536 --hit 0,0.5,0.9 --out gpurun_out/fanout_n8.jsonl 2>&1 | grep -E '^\{|Error|Traceback' | tail -30
timeout 120 python benchmarks/overlap.py 2>&1 | tail -24
