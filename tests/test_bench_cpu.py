"""CPU: the parts of bench.py that need no device -- the clock-sample reduction and the reference arm's JSON line
(`--impl reference`: the reference's own CPU path through oracle/_ref, or the C port where the reference headers were not
compiled), alone and as rank 0 / rank 1 of a world of two."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_clock_sampler_reduction():
    import bench
    cs = bench.ClockSampler(0)                                # no device here: neither NVML nor nvidia-smi delivers a sample
    cs.stop()
    assert cs.window(0.0, 1e18)["sm_mhz"] is None and cs.window(0.0, 1e18)["samples"] == 0      # says so instead of inventing a clock
    # (host time, SM MHz, max SM MHz, watts or None, NVML clocks-event-reasons bits): only samples inside the window count
    cs.rows = [(9.0, 600.0, 1965.0, 90.0, 0x8),               # before the timed region (an idle clock and a slowdown that is not ours)
               (10.0, 1965.0, 1965.0, 350.1, 0x4), (10.5, 1950.0, 1965.0, 360.2, 0x0), (11.0, 1965.0, 1965.0, None, 0x1),
               (12.5, 800.0, 1965.0, 100.0, 0x40)]            # after it
    r = cs.window(10.0, 11.0)
    assert r["sm_mhz"] == 1965.0 and r["sm_max_mhz"] == 1965.0 and r["samples"] == 3 and r["inside_timed_region"]
    assert r["reasons"] == ["sw_power_cap"] and r["power_w_max"] == 360.2
    assert cs.window(9.0, 12.5)["reasons"] == ["hw_slowdown", "hw_thermal_slowdown", "sw_power_cap"]


def _check_reference_line(line, n_gpus):
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "Mreads/s" and d["metric"].startswith("Mreads/s classified")
    assert d["n_gpus"] == n_gpus and d["steps"] == 1 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert "workload" in d["config"] and not any(k in d["config"] for k in ("model", "seq_len", "global_batch"))
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mreads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--reads", "20000"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    _check_reference_line(lines[0], 1)


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout, r.stderr)
