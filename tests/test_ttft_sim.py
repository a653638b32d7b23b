"""the TTFT harness (benchmarks/ttft_sim.py) keeps the invariants of the reference's hand-off model
(lib/mocker/src/replay/offline/disagg.rs:1336-1374 `test_handoff_delay_increases_decode_visible_ttft`): with everything
else equal, TTFT grows by the hand-off delay, and prefix hits found by the RadixTree shrink what is moved."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(tmp_path, *extra):
    out = tmp_path / "sim.json"
    subprocess.run([sys.executable, os.path.join(ROOT, "benchmarks", "ttft_sim.py"), "--requests", "120", "--out", str(out), *extra],
                   check=True, capture_output=True, timeout=120)
    return json.load(open(out))


def test_handoff_delay_is_what_separates_the_data_planes(tmp_path):
    r = _run(tmp_path, "--rate", "5", "--prompt-tokens", "4096", "--shared-fraction", "0")
    p = r["planes"]
    assert p["ours"]["prefix_hit_rate"] == 0.0
    # lightly loaded, one prompt size: TTFT = prefill + hand-off + first decode for every request
    kv = 4096 * r["config"]["kv_bytes_per_token"]
    prefill_ms = 4096 / r["config"]["prefill_tok_per_s"] * 1e3
    want_mocker = prefill_ms + kv / 64e9 * 1e3 + r["config"]["first_decode_ms"]
    assert abs(p["mocker64"]["ttft_ms"]["p50"] - want_mocker) < 0.5
    want_ours = prefill_ms + r["ours_model"]["latency_ms"] + kv / (r["ours_model"]["bandwidth_gbs"] * 1e9) * 1e3 + r["config"]["first_decode_ms"]
    assert abs(p["ours"]["ttft_ms"]["p50"] - want_ours) < 0.5
    assert p["ours"]["ttft_ms"]["p50"] < p["mocker64"]["ttft_ms"]["p50"] < p["cpu"]["ttft_ms"]["p50"]
    assert abs(r["ttft_drop_ms_vs_ours"]["mocker64"]["p50"] - (p["mocker64"]["handoff_ms"]["p50"] - p["ours"]["handoff_ms"]["p50"])) < 0.5


def test_shared_prefixes_are_found_by_the_router_and_not_moved(tmp_path):
    r = _run(tmp_path, "--rate", "5", "--prompt-tokens", "4096", "--shared-fraction", "0.5", "--families", "2", "--decode-workers", "2")
    hit = r["planes"]["ours"]["prefix_hit_rate"]
    assert 0.35 < hit <= 0.5      # half of every prompt is shared; the first request of a family on a worker misses
    assert r["planes"]["ours"]["blocks_moved"] == r["planes"]["mocker64"]["blocks_moved"]
