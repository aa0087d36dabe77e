"""bench.py's reference arm runs on CPU only (the reference's OpenMP program, machine-translated to C: oracle/_ref;
or the restated reference if that is not built), so its JSON line can be
checked here: the keys of the measurement contract, and that rank > 0 of a multi-rank launch prints nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None, *args):
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "s3",
                        "--iter-max", "5", "--steps", "1", "--warmup", "0", *args],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip()


def test_reference_arm_line():
    out = _run()
    line = json.loads(out.splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "cell_updates_per_s" and line["unit"] == "cell-updates/s"
    assert line["higher_is_better"] is True and line["dtype"] == "f64" and line["data"] == "synthetic"
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["gpu_launches"] == 0
    assert line["config"]["workload"] == "s3_64"
    cb = line["cpu_baseline"]
    from oracle import build_ref
    have_ref = os.path.exists(build_ref.lib_path("ibm_3d_uniform_omp_cpu", "omp", "b"))
    assert cb["kind"] == ("reference" if have_ref else "port")
    assert cb["cores"] >= 1 and "sample" in cb and cb["value"] == line["value"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_are_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2") == ""
