"""bench.py contract on CPU: the reference arm (the oracle port of the reference on the host cores) prints one JSON line with
the keys the driver reads; non-zero ranks of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"]


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                          capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    out = _run({})
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.config_block("cfg3", 1) and d["scaling"] == "weak"      # same object as the GPU arm prints


def test_reference_arm_is_silent_on_other_ranks():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_peaks_reader_tolerates_the_driver_file(tmp_path, monkeypatch):
    """roofline denominators: flat keys, nested keys, a missing value and a broken file all give numbers."""
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.peaks() == (6650.0, 1400.0, "fallback (B200_PROFILING.md)")
    f = tmp_path / "MEASURED_PEAKS.json"
    f.write_text(json.dumps({"hbm_gbs": 7000.5, "bf16_tflops": 1600, "bf16_tflops_sustained": 1390}))
    assert bench.peaks()[:2] == (7000.5, 1390.0) and bench.peaks()[2].startswith("measured")
    f.write_text(json.dumps({"b200": {"hbm_gbs": 6900, "bf16_tflops": 1500}, "clocks": {"sm_mhz": 1965}}))
    assert bench.peaks()[:2] == (6900.0, 1500.0)
    f.write_text(json.dumps({"hbm_gbs": 6900}))
    hbm, tf, src = bench.peaks()
    assert (hbm, tf) == (6900.0, 1400.0) and "fallback for bf16_tflops" in src
    f.write_text("{not json")
    assert bench.peaks()[:2] == (6650.0, 1400.0)


def test_roofline_model_reproduces_the_committed_table():
    """profiles/r02_model_vs_measured.txt is `tools/roofline_model.py` over the two committed launch lists: the tool still runs and
    pairs every expected launch with a measured one (80 vocoder launches, 67 TTE launches), floors below the measured times."""
    for args, n_launches in ((["profiles/r02_launches_voc_cfg2.csv"], 80), (["profiles/r02_launches_tte_cfg3.csv", "--tte"], 67)):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "roofline_model.py")] + args, capture_output=True, text=True,
                             timeout=120, cwd=ROOT)
        assert out.returncode == 0, out.stderr
        lines = out.stdout.splitlines()
        split = lines.index("")
        assert split - 1 == n_launches
        total = lines[-1].split()
        assert total[0] == "total"
        meas, floor, frac = float(total[1]), float(total[4]), float(total[5])
        assert 0.3 < frac < 1.0 and floor < meas
        assert " ".join(lines[-1].split()) in " ".join(open(os.path.join(ROOT, "profiles", "r02_model_vs_measured.txt")).read().split())


def test_chain_schedule_is_hazard_free():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_chain_schedule.py")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr


def test_cpu_baseline_sample_is_bounded():
    sys.path.insert(0, ROOT)
    import bench
    d = bench.cpu_baseline_sample("cfg2", budget_s=0.5, max_steps=3, plan=[(0, 24)] * 3)
    assert d["kind"] == "port" and d["unit"] == "samples/s" and d["cores"] >= 1 and d["value"] > 0 and d["best_step"] >= d["value"] * 0.999
    assert "1 utterance x 24 units" in d["sample"] and "step(s)" in d["sample"]


def test_workload_selection_and_config_block():
    """N = 1 measures the single-GPU text -> wav configuration (configs[2]), N > 1 the sharded pool (configs[3]); the `config` object
    is the same function of (workload, world) in both arms, so the driver's same-config check compares like with like."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.resolve_workload("auto", 1) == "cfg3" and bench.resolve_workload("auto", 8) == "cfg4"
    assert bench.resolve_workload("cfg2", 8) == "cfg2"
    c3, c4 = bench.config_block("cfg3", 1), bench.config_block("cfg4", 8)
    assert c3["workload"].startswith("cfg3") and c3["units_per_rank_per_step"] == 32 * 2048
    assert c4["workload"].startswith("cfg4") and c4["pool_units"] == sum(bench.cfg4_lengths()) and "strong" in c4["sharding"]
    L = bench.cfg4_lengths()
    assert len(L) == 256 and min(L) >= 256 and max(L) <= 4096 and max(L) < bench.TTE_MAX_LEN
    plan = bench.cpu_sample_plan("cfg4", 5)
    assert len(plan) == 5 and all(S == (T + 7) // 8 for S, T in plan) and plan == sorted(plan)
    assert bench.cpu_sample_plan("cfg3", 2) == [(256, 2048)] * 2


def test_reference_arm_at_two_ranks_times_the_sharded_pool():
    out = _run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert out.returncode == 0, out.stderr
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.config_block("cfg4", 2) and d["scaling"] == "strong"
