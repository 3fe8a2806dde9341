"""bench.py's contract on the CPU side: the reference arm prints ONE JSON line with the keys the driver reads, ranks
other than 0 exit quietly under torchrun, and the B200 arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e, cwd=ROOT,
                          timeout=600)


def test_reference_arm_prints_the_contract_line():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", "speech32"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "quantized vectors/sec (VQ fwd+bwd)" and d["unit"] == "vectors/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] >= 3 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("speech32")
    cb = d["cpu_baseline"]
    # the unmodified reference class where its tree is present (this container), the oracle port elsewhere (the GPU box)
    have_ref = os.path.isdir("/root/reference/src/acoustic_locating_vq_vae") or os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "src"))
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["config"]["rows_per_gpu"] == 16000 and d["config"]["K"] == 1024 and d["config"]["D"] == 128
    assert d["e2e"] == {"value": d["value"], "unit": "vectors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    p = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU")
def test_b200_arm_has_no_cpu_fallback():
    p = _run(["--steps", "1", "--warmup", "1"])
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)


def test_both_arms_describe_the_workload_identically():
    """The driver pairs the two arms by their `config`: the dictionaries must be equal key for key."""
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    for name in ("rir256", "speech32", "sweep_k512_d64"):
        for world in (1, 8):
            a, b = bench.make_config(name, world), bench.make_config(name, world)
            assert a == b and a["rows_per_gpu"] == bench.WORKLOADS[name][0] * bench.WORKLOADS[name][2]
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    d = json.loads(p.stdout.strip())
    assert d["config"] == bench.make_config("rir256", 1)
