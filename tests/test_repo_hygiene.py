"""CPU checks that keep the measurement and tooling scripts loadable: every Python file of the repo compiles, the shell
jobs under tools/ parse, bench.py's argument parser accepts the flags the docs quote, and the profile index only names
files that exist."""
import glob
import os
import py_compile
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "*.py")) + glob.glob(os.path.join(ROOT, "tools", "*.py")) +
                                        glob.glob(os.path.join(ROOT, "acoustic_locating_vq-vae_b200", "*.py")) +
                                        glob.glob(os.path.join(ROOT, "oracle", "*.py")) + glob.glob(os.path.join(ROOT, "tests", "golden", "*.py"))))
def test_python_file_compiles(path, tmp_path):
    py_compile.compile(path, cfile=str(tmp_path / "x.pyc"), doraise=True)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "tools", "*.sh"))))
def test_shell_job_parses(path):
    assert subprocess.run(["bash", "-n", path]).returncode == 0


def test_bench_help_lists_the_documented_flags():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--workload", "--no-onehot", "--sweep-all", "--no-sweep", "--no-module", "--skip-e2e",
                 "--skip-cpu", "--graph", "--no-graph", "--dp-mode", "--strong", "--nccl", "--emulate-dp", "--exact", "--no-screen"):
        assert flag in out.stdout, flag


def test_profile_index_names_existing_files():
    text = open(os.path.join(ROOT, "profiles", "README.md")).read()
    names = set(re.findall(r"`((?:r1|r2)_[A-Za-z0-9_.]+\.(?:json|csv|txt))`", text)) | {"traffic.json"}
    missing = sorted(n for n in names if not os.path.exists(os.path.join(ROOT, "profiles", n)))
    assert not missing, missing
