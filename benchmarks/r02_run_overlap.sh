#!/bin/bash
# BASELINE configs[3] per pair: the layer-streamed hand-off behind a real tensor-core layer (cuBLAS bf16 GEMM per layer), round-2 engine.
# spin = warps of ONE launch wait on the ready flags (CUDA_MODULE_LOADING=EAGER); streamwait = what AUTO does under lazy loading.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
CUDA_MODULE_LOADING=EAGER timeout 250 python benchmarks/overlap.py --standin gemm --ctas 16 --out gpurun_out/r02_overlap_gemm_ctas16_spin.json > gpurun_out/r02_overlap_spin.log 2>&1; echo "spin rc=$?"; tail -n 3 gpurun_out/r02_overlap_spin.log
timeout 250 python benchmarks/overlap.py --standin gemm --ctas 16 --out gpurun_out/r02_overlap_gemm_ctas16_streamwait.json > gpurun_out/r02_overlap_streamwait.log 2>&1; echo "streamwait rc=$?"; tail -n 3 gpurun_out/r02_overlap_streamwait.log
python - <<'PY'
import json
for m in ("spin", "streamwait"):
    try:
        d = json.load(open(f"gpurun_out/r02_overlap_gemm_ctas16_{m}.json"))
        print(m, {k: round(v, 3) for k, v in d.items() if isinstance(v, float)}, d.get("bit_exact_probe"), d.get("peer"))
    except Exception as e:
        print(m, "failed", e)
PY
