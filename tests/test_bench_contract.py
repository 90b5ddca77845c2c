"""bench.py contract on the CPU box: the reference arm (the oracle port on the host cores) prints ONE JSON line with
the keys the driver reads, and under a multi-rank launch only rank 0 works and prints."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, *args):
    env = dict(os.environ, **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], env=env, capture_output=True,
                          text=True, timeout=600, cwd=ROOT)


import pytest


@pytest.mark.parametrize("port", [True, False])
def test_reference_arm_prints_one_json_line_with_the_contract_keys(port):
    """Both CPU arms (the numpy oracle port; the reference's own modules when baseline/_ref or
    $DIFFSPTK_REFERENCE_ROOT holds the package) print the contract's line with the native arm's config keys."""
    sys.path.insert(0, ROOT)
    import bench
    r = _run({}, "--impl", "reference", "--workload", "lpc2par", "--steps", "1", "--warmup", "1",
             *(["--port"] if port else []))
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "frames/s" and j["higher_is_better"] is True
    assert j["metric"].startswith("frames/sec") and j["value"] > 0 and j["n_gpus"] == 1
    assert j["vs_baseline"] is None and j["data"] == "synthetic"
    assert j["config"] == bench.workload_config("lpc2par")        # identical keys and values in both arms
    cb = j["cpu_baseline"]
    have_ref = bench.load_reference_package()[0] is not None
    assert cb["kind"] == ("port" if port or not have_ref else "reference")
    assert cb["cores"] >= 1 and cb["sample"] and cb["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--impl", "reference", "--gpus", "2", "--steps", "1")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workload_table_matches_the_scope_contract():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.n_frames(160000) == 2000 and bench.n_frames(80000) == 1000 and bench.n_frames(1) == 1
    cfg, B, T, rd, wr = bench.WORKLOADS["stft"]                       # BASELINE.json config 2, SURVEY.md section 8(d)
    assert (B, T, rd + wr) == (256, 160000, 1348)
    assert bench.WORKLOADS["lpc"][1:] == (1024, 80000, 320, 100)      # config 3: 420 B / frame
    assert bench.WORKLOADS["mcep"][3] + bench.WORKLOADS["mcep"][4] == 1128
    assert bench.WORKLOADS["mfcc"][3] + bench.WORKLOADS["mfcc"][4] == 372


def test_ncu_traffic_is_tied_to_the_kernel_sources():
    """roofline.traffic comes from profiles/traffic.json ONLY while the kernel sources are the ones the ncu capture was
    made from (it was a constant); a stale entry must read as None, not as a number."""
    sys.path.insert(0, ROOT)
    import bench
    db = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for wl, rec in db.items():
        if wl.startswith("_"):
            continue
        assert set(rec) >= {"dram_bytes_per_launch", "capture", "sources", "source_digest"}
        fresh = rec["source_digest"] == bench.kernel_source_digest(rec["sources"])
        assert (bench.ncu_record(wl) is not None) == fresh
    assert bench.ncu_record("no-such-workload") is None
