"""bench.py on a box without a GPU: the reference arm (the unmodified reference on the host cores) prints the JSON
contract, ranks other than 0 print nothing, and our own arm fails loudly instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*argv, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True,
                          cwd=ROOT, env=e, timeout=600)


@pytest.mark.parametrize("workload,metric,unit", [("cube127", "poisson_solve_gpts_per_s", "Gpts/s"),
                                                  ("nscube31", "ns_steps_per_s", "steps/s")])
def test_reference_arm_contract(ref, workload, metric, unit):
    r = run("--impl", "reference", "--workload", workload, "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == metric and d["unit"] == unit
    assert d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["dtype"] == "f64" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent(ref):
    r = run("--impl", "reference", "--workload", "cube127", "--steps", "1", "--warmup", "0", "--gpus", "2",
            env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_own_arm_needs_a_gpu():
    import fdm_b200
    if fdm_b200.lib().fdmb_device_count() > 0:
        pytest.skip("a GPU is present")
    r = run("--workload", "cube127", "--steps", "1", "--warmup", "1")
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]      # no number without the CUDA path


def test_numa_binding_never_raises(tmp_path):
    """bench.py's host-placement helper is an optimisation: whatever the box exposes, it returns a record."""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    before = os.sched_getaffinity(0)

    class Props:
        pci_domain_id, pci_bus_id, pci_device_id = 0, 0x1b, 0
    fake = types.SimpleNamespace(cuda=types.SimpleNamespace(get_device_properties=lambda i: Props()))
    r = b.bind_to_gpu_numa_node(fake, 0)
    assert isinstance(r, dict) and "bound" in r
    broken = types.SimpleNamespace(cuda=types.SimpleNamespace(get_device_properties=lambda i: object()))
    assert b.bind_to_gpu_numa_node(broken, 0)["bound"] is False
    os.sched_setaffinity(0, before)


def test_cpu_baseline_sample(ref):
    """cpu_baseline times the unmodified reference over a bounded sample and says what the sample was."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.CPU_SAMPLE_SECONDS = 0.3
    for wl, unit in (("cube127", "Gpts/s"), ("nscube31", "steps/s")):
        r = b.cpu_baseline(b.WORKLOADS[wl])
        assert r["kind"] == "reference" and r["unit"] == unit and r["value"] > 0 and r["cores"] >= 1
        assert "s of work" in r["sample"] and "mean" in r["sample"]
    calls = []
    reps, mean, best = b.timed_sample(lambda: calls.append(1), min_seconds=0.0)
    assert reps == len(calls) == 3 and best <= mean
