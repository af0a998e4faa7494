"""The reference arm of bench.py runs without a GPU: check the JSON contract of its line here (keys the driver reads)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train samples/sec (fwd+bwd)" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("gpt_fusion_stage n_embd=512") and d["config"]["sample_batch"] == 2
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "oracle port" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_gpu_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback on the product path: without CUDA the default arm exits with an error instead of timing the oracle."""
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)


def test_hbm_family_bytes_follow_the_per_sample_figures_of_the_measurement_plan():
    """bench.py's algorithmic bytes of the HBM-bound families (roofline.families[*].gbps) against SURVEY.md §8(d) / BASELINE.md §3:
    E_t = T*C = 492 544 token elements and E_f = 3*5*C*H^2 = 491 520 feature elements per sample at stage 4 (C = 512, H = 8)."""
    sys.path.insert(0, ROOT)
    import bench
    b = 12
    hb = bench.hbm_family_bytes(b)
    e_t, e_f = 962 * 512 * b, 3 * 5 * 512 * 8 * 8 * b
    assert hb["tokens_fwd"] == e_f * 4 + e_t * 4 + 962 * 512 * 4          # read features, write tokens (+ pos_emb once)
    assert hb["upsample_add_fwd"] == e_t * 4 + 2 * e_f * 4                # read tokens + features, write features
    assert hb["upsample_add_bwd"] == e_f * 4 + e_t * 4
    assert hb["layernorm_fwd"] == 16 * e_t * 6 + e_t * 8                  # 16 x (fp32 in, bf16 out) + ln_f (fp32 out)
    assert hb["layernorm_bwd"] == 16 * e_t * 16 + e_t * 14
    assert hb["colsum"] == 8 * (962 * b) * (2048 + 1536) * 2
    assert hb["pack_block_weights"] == 8 * (4 * 512 * 512 + 2 * 2048 * 512) * 8
    assert all(v > 0 for v in hb.values())
