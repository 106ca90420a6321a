"""bench.py's output contract, checked on the CPU through the reference arm (the GPU arm needs a B200) and through its helpers."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--case", "mini_tripolar", "--ntr", "3",
                        "--steps", "2", "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["unit"] == "cell-updates/s" and d["higher_is_better"] is True and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert cb["multiblock_equals_singleblock"] is True          # verified before timing
    assert d["e2e"] == dict(value=d["value"], unit=d["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert d["config"]["same_grid_as_gpu_arm"] is True and d["config"]["global_grid"] == [40, 30, 12]
    # both arms name the workload with the same string (the driver compares the arms' configs)
    b = _bench()
    from mom5_b200.synthetic import CASES
    assert d["config"]["workload"] == b.workload_name("mini_tripolar", CASES["mini_tripolar"], 3)


def test_traffic_file_is_refused_when_measured_on_other_kernels(tmp_path, monkeypatch):
    b = _bench()
    t, why = b.traffic_record("k_sweby_xy_tma")
    cur = json.load(open(os.path.join(ROOT, "profiles", f"traffic_{b.ROUND}.json")))
    if cur["kernel_source_sha1"] == b.kernel_source_hash():
        assert t == cur["kernels"]["k_sweby_xy_tma"] and t > 1e11          # the fused pass moves > 100 GB per launch at 0.1 degree
    else:
        assert t is None and why.startswith("stale")
    # a file stamped with another hash is refused, whatever it contains
    monkeypatch.setattr(b, "kernel_source_hash", lambda: "0" * 16)
    t, why = b.traffic_record("k_sweby_xy_tma")
    assert t is None and why.startswith("stale")


def test_algorithmic_bytes_follow_survey_8d():
    b = _bench()
    assert b.b_alg(3) == 80.0 + 64.0 / 3 and b.b_alg(10) == 86.4
    s = b.sweep_bytes(3)
    assert (s["z"], s["x"], s["y"]) == (64.0, 88.0, 152.0) and abs(sum(s.values()) / 3 - b.b_alg(3)) < 1e-12
