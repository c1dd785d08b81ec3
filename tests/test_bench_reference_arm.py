"""The reference arm of bench.py runs on CPU: JSON contract, and the thread count must not collapse to 1 under torchrun
(which exports OMP_NUM_THREADS=1 to its workers -- the round-1 SCALE ratios were inflated ~12x by exactly that)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                           "--ref-pages", "128", *args], capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))


def test_reference_arm_json_and_thread_count():
    r = run({"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"}, "--gpus", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "maxsim_query_page_pairs_per_s" and line["unit"] == "pairs/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 2 and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and "128 pages" in cb["sample"]
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    assert cb["cores"] == avail, "the reference arm must use every host core, whatever OMP_NUM_THREADS says"
    assert line["config"]["pages_per_gpu"] == 128 and "BASELINE configs[1]" in line["config"]["workload"]


def test_reference_arm_other_ranks_stay_silent():
    r = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""
