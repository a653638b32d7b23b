#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
export CUDA_MODULE_LOADING=EAGER
i=0
IFS=";" read -ra LIST <<< "${CFGS:---ctas 16;--ctas 8;--ctas 16}"
for cfg in "${LIST[@]}"; do
  i=$((i+1))
  timeout 200 python benchmarks/overlap.py --standin gemm --iters ${ITERS:-15} $cfg --out gpurun_out/r02_overlap_peer_${TAG:-}$i.json > gpurun_out/r02_overlap_peer_$i.log 2>&1; echo "cfg[$cfg] rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_overlap_peer_*.json")):
    d = json.load(open(f))
    keys = ("t_compute_ms", "t_transfer_capped_ms", "t_overlapped_ours_ms", "t_overlapped_ours_events_ms", "t_overlapped_ours_prereleased_ms", "t_overlapped_ours_per_layer_ms", "t_overlapped_ref_style_ms")
    print(f[-7:], d["ctas"], d["ring"], d["peer"], {k[2:-3]: round(d[k], 3) for k in keys if k in d}, "ev/c=%.3f" % d["slowdown_vs_compute_ours_events"])
PY
