"""The bench line's shape (driver contract): checked on the committed headline line and on the reference arm run live
on CPU with a tiny sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def test_committed_headline_line_has_every_contract_key():
    with open(os.path.join(ROOT, "profiles", "r02s3_bench_c2.json")) as fh:
        line = json.load(fh)
    with open(os.path.join(ROOT, "BASELINE.json")) as fh:
        base = json.load(fh)
    assert BASE_KEYS | {"roofline", "clocks"} <= set(line)
    assert line["metric"] == base["metric"] and line["unit"] == "images/s" and line["higher_is_better"] is True
    assert line["warmup"] >= 3 and line["n_gpus"] == 1 and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"] and "l2_policy" in line["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    # the batch crosses PCIe as bf16 (config.pixels says so): 2 bytes per value + the int64 labels
    assert "bf16" in line["config"]["pixels"]
    assert line["e2e"]["h2d_bytes_per_step"] == 256 * 3 * 224 * 224 * 2 + 256 * 8 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["e2e"]["value"] <= line["value"] * 1.02          # e2e includes the copies: never faster than resident
    roof = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(roof) and roof["bound"] in ("hbm", "tensor")
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9 and 0 < roof["frac"] < 1
    assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"]) and line["cpu_baseline"]["kind"] == "port"
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(line["clocks"]["reasons"])
    assert line["gpu_launches"] > 0
    # round 2: the line also carries north_star's denominators and BASELINE.json's second metric
    assert line["cpu_baseline"]["same_config"] is True
    eager = line["gpu_eager_baseline"]
    assert eager["fp32"] > 0 and eager["autocast_bf16"] > 0 and eager["speedup_vs_autocast_bf16"] >= 5.0
    assert 0 < line["logits_max_abs_err"] < 5e-2 and line["logits_parity"]["ratio_to_reference_bf16_floor"] < 1.0
    assert line["inference"]["images_per_s"] > line["value"] and line["inference"]["bytes_kept_after_forward"] <= 4 * line["inference"]["logits_bytes"]
    assert 0 < line["whole_step_tensor_frac"] < 1 and line["roofline"]["traffic"] is not None
    # session 3: the text tower is measured on the fused blocks beside its stock PyTorch path
    text = line["text_tower"]
    assert text["prompts_per_s"] > text["stock_pytorch_fp32_prompts_per_s"] and text["features_rel_err_vs_stock_fp32"] < 1e-2


def test_committed_multi_gpu_lines_use_the_fused_exchange():
    """2 / 4 / 8 GPUs: the exchange in the step is the one-shot peer-memory all-reduce fused with SGD, no hand-shake
    timed out, and the CUDA-IPC check found the ranks bit-identical and every replayed step equal to the update recomputed
    from the ranks' local gradients."""
    for name in ("r02s3_peer_bench_n2_fused.json", "r02s3_peer_bench_n4_fused.json", "r02s3_s3_weak_n8.json",
                 "r02s3_s3_strong_n8.json"):
        with open(os.path.join(ROOT, "profiles", name)) as fh:
            line = json.load(fh)
        assert BASE_KEYS - {"cpu_baseline"} <= set(line) and line["n_gpus"] in (2, 4, 8)
        assert "pevit_allreduce_sgd" in line["exchange"]["kind"] and line["exchange"]["timed_out"] is False
    for name in ("r02s3_peer_check_n4_run1.json", "r02s3_peer_check_n4_run2.json", "r02s3_peer_check_n4_run3.json"):
        with open(os.path.join(ROOT, "profiles", name)) as fh:
            chk = json.load(fh)
        assert chk["peer_path"] and chk["identical_across_ranks"] and chk["graph_identical_across_ranks"]
        assert chk["graph_max_rel_diff_vs_recomputed_update"] < 1e-6 and not chk["timed_out"]
    with open(os.path.join(ROOT, "profiles", "r02s3_peer_check_n8.json")) as fh:
        chk = json.load(fh)
    assert chk["world"] == 8 and chk["peer_path"] and chk["identical_across_ranks"] and not chk["timed_out"]


def test_reference_arm_prints_the_contract_line_on_cpu():
    res = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-batch", "2"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and BASE_KEYS <= set(line)
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["cores"] >= 1
    assert line["gpu_launches"] == 0 and "workload" in line["config"]
