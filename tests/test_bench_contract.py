"""bench.py's reference arm (`--impl reference`) needs no GPU: it times the unmodified reference's camera::render
(oracle/_ref, or the oracle port where the reference was not compiled) on the host cores.  This pins the JSON line the
driver parses; the GPU arm's line carries the same keys plus roofline / clocks (checked on the GPU box by the driver)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_arm(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--width", "40"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    lines = run_arm()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "samples_px_per_s" and d["unit"] == "samples*px/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0  # nothing of ours runs in the reference arm


def test_reference_arm_other_ranks_exit_quietly():
    assert run_arm({"RANK": "1", "WORLD_SIZE": "2"}) == []


def test_ncu_figures_are_tied_to_the_kernel_code_not_to_its_comments(tmp_path, monkeypatch):
    """bench.py quotes DRAM traffic / issue utilisation / active lanes only from an ncu capture of THESE kernels
    (profiles/latest_ncu.json carries a hash of csrc/ with comments and whitespace removed): rewording a comment keeps the
    figures, changing a token drops them."""
    import importlib
    import sys

    sys.path.insert(0, ROOT)
    mod = importlib.import_module("raytracing-practice_b200.csrc_sha")
    a = "int f(int x) { // add one\n  return x + 1; /* here */ }\n"
    b = "int f(int x) {   // plus one, reworded\n\n  return x + 1; }\n"
    c = "int f(int x) { return x + 2; }\n"
    assert mod._code_only(a) == mod._code_only(b) != mod._code_only(c)
    assert mod._code_only('const char* s = "// not a comment";') == 'const char* s = "// not a comment";'
    import bench

    ncu, src = bench.latest_ncu()
    if ncu is not None:  # the committed capture matches the tree: its keys are the ones bench.py reads
        assert ncu["csrc_sha"] == bench.csrc_sha() and ncu["dram_bytes_per_sample"] > 0 and 0 < ncu["issue_slot_utilisation"] <= 1
    else:
        assert "omitted" in src or "missing" in src
