"""CPU: the `--impl reference` arm of bench.py (the restated reference CPU path) prints one JSON line with the contract's
keys, times the configuration it names in full (nothing extrapolated) and uses the host threads it is given -- also when
torchrun has exported OMP_NUM_THREADS=1 to its workers."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *flags):
    env = dict(os.environ); env.update(env_extra)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n-el", "12", "--steps", "2",
                          "--warmup", "1", "--no-trend", *flags], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])


def test_reference_arm_line():
    line = _run({"OMP_NUM_THREADS": "1"})           # what a torchrun worker sees
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "analysis+adjoint iters/s" and line["unit"] == "iters/s"
    assert line["config"]["workload"].startswith("cylinder_4x2_ne12") and line["config"]["dofs"] == 8340
    assert line["steps"] == 2 and abs(line["value"] * line["ms_per_step"] / 1e3 - 1.0) < 1e-9
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and "nothing extrapolated" in cb["sample"]
    assert cb["cores"] == len(os.sched_getaffinity(0))          # not the single thread torchrun would have left it
    assert line["e2e"] == {"value": line["value"], "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert set(line["cpu_phase_s"]) == {"assemble_RK", "lu_state", "linearize", "lu_adjoint", "gradients"}


def test_other_ranks_print_nothing():
    env = dict(os.environ); env.update({"RANK": "1", "WORLD_SIZE": "2"})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_watchdog_prints_the_line_and_exits_zero():
    """A post-measurement section that never returns (a collective some rank does not enter) must not cost the
    measured line: the watchdog prints it with a note and leaves with exit code 0."""
    code = ("import sys, time; sys.path.insert(0, %r); import bench\n"
            "wd = bench.Watchdog(0); wd.line = {'metric': 'm', 'value': 1.0}\n"
            "wd.start(1, 'hanging section'); time.sleep(30); print('not reached')\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "not reached" not in out.stdout
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["value"] == 1.0 and "hanging section" in line["notes"][0]
    code2 = ("import sys, time; sys.path.insert(0, %r); import bench\n"
             "wd = bench.Watchdog(1); wd.start(1, 'x'); wd.cancel(); time.sleep(2); print('finished')\n" % ROOT)
    out2 = subprocess.run([sys.executable, "-c", code2], capture_output=True, text=True, timeout=60)
    assert out2.returncode == 0 and out2.stdout.strip() == "finished"
