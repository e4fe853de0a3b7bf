"""The JSON lines bench.py has printed on B200 (committed under profiles/) carry every key the
driver contract names -- a cheap guard against a bench edit dropping one."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
             "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _lines(pattern):
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", pattern))):
        with open(path) as f:
            txt = f.read().strip()
        if txt:
            yield path, json.loads(txt.splitlines()[-1])


def test_latest_default_line_has_the_contract_keys():
    found = list(_lines("r1_pf_bench_default.json")) + list(_lines("r1_pf_scale_default_n2.json"))
    assert found
    for path, line in found:
        assert BASE_KEYS <= set(line), (path, BASE_KEYS - set(line))
        assert line["metric"].startswith("queries/sec exact cosine top-10 @10M") and line["unit"] == "queries/s"
        assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
        assert line["config"]["workload"].startswith("10000000x1024 bf16 corpus, batch-1024") and "model" not in line["config"]
        assert line["clocks"]["sm_mhz"] and "reasons" in line["clocks"]
        e2e = line["e2e"]
        assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] == 1024 * 1024 * 4 and e2e["d2h_bytes_per_step"] == 1024 * 10 * 12
        roof = line["roofline"]
        assert roof["bound"] == "tensor" and roof["unit"] == "TFLOP/s" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
        assert roof["traffic"] > 0 and line["gpu_launches"] > 0
        if line["n_gpus"] == 1:
            cpu = line["cpu_baseline"]
            assert cpu["kind"] == "port" and cpu["cores"] >= 1 and cpu["value"] > 0 and cpu["sample"]
        pf = line["secondary_prefiltered"]
        assert pf["identical_to_exact_scan"] is True and pf["value"] > line["secondary"]["value"]


def test_encode_line_has_the_contract_keys():
    """`bench.py --workload encode` (the embedding encoder) as printed on B200."""
    found = list(_lines("r2b_final_bench_encode.json"))
    assert found
    for path, line in found:
        assert BASE_KEYS <= set(line), (path, BASE_KEYS - set(line))
        assert line["unit"] == "tokens/s" and line["value"] > 5e5 and line["dtype"] == "f16" and "random-init" in line["data"]
        assert "model" not in line["config"] and line["config"]["workload"].startswith("64 chunks x 512 tokens")
        assert line["gpu_launches"] == line["steps"] * 170                      # 2 + 7 x 24 per forward pass
        roof = line["roofline"]
        assert roof["bound"] == "tensor" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
        e2e = line["e2e"]
        assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] == 64 * 1024 * 4
        assert e2e["from_text"]["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
        assert set(line["breakdown"]["per_gemm"]) == {"qkv", "attn_out", "ffn1_gelu", "ffn2"}


def test_default_line_carries_the_encoder_block():
    for _, line in _lines("r2b_final_bench_default.json"):
        enc = line["encoder"]
        assert enc["tokens_per_s"] > 5e5 and enc["gpu_launches"] == enc["steps"] * 170 and enc["query_latency_ms_16_tokens"] > 0


def test_reference_arm_line():
    found = list(_lines("r1_final2_bench_reference.json"))
    assert found
    for _, line in found:
        assert line["impl"] == "reference" and line["gpu_launches"] == 0
        assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        assert line["cpu_baseline"]["value"] == line["value"]


def _bench_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("sqe_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_labels_come_from_what_is_run():
    b = _bench_module()
    assert b.metric_name(10_000_000, 10) == "queries/sec exact cosine top-10 @10Mx1024"      # BASELINE.json's metric
    assert b.metric_name(12_500_000, 100) == "queries/sec exact cosine top-100 @12.5Mx1024"
    assert b.metric_name(32_717, 5) == "queries/sec exact cosine top-5 @32717x1024"
    assert "configs[2]" in b.baseline_tag(10_000_000, "bf16", 1024, 10)
    assert b.baseline_tag(1_000_000, "fp32", 1, 10) == "BASELINE configs[1]"
    assert b.baseline_tag(100_000_000, "fp16", 256, 100) == "BASELINE configs[3]"
    assert b.baseline_tag(2_000_000, "bf16", 1024, 10) == "not a BASELINE configuration"


def test_literal_reference_cpu_line_runs():
    lit = _bench_module().cpu_literal_reference(budget_s=0.5)
    assert lit["cores"] == 1 and lit["kind"] == "port"
    assert lit["lfu_cache_get_1000_entries"]["value"] > 0 and lit["lfu_cache_get_1000_entries"]["queries"] >= 3
    assert lit["row_by_row_cosine_top10_32717_chunks"]["value"] > 0
