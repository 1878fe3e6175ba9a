"""The host thread pool behind fl_step_observe_host_compact (csrc/expand_pool.h), stressed on the CPU: many short jobs back to
back (the shape of eight environment ranges per step), with and without the thread sanitizer."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "expand_pool_stress.cpp")
INC = os.path.join(ROOT, "flatland-marl_b200", "csrc")


def _build(tmp_path, extra):
    exe = str(tmp_path / "expand_pool_stress")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-I", INC, SRC, "-o", exe] + extra)
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_every_index_runs_once_with_its_own_job(tmp_path):
    exe = _build(tmp_path, [])
    out = subprocess.run([exe, "30000", "8"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_thread_sanitizer_finds_no_race(tmp_path):
    try:
        exe = _build(tmp_path, ["-fsanitize=thread", "-g"])
    except subprocess.CalledProcessError:
        pytest.skip("g++ has no thread sanitizer runtime here")
    out = subprocess.run([exe, "3000", "6"], capture_output=True, text=True, timeout=600)
    if "FATAL: ThreadSanitizer" in out.stderr and "unexpected memory mapping" in out.stderr:
        pytest.skip("thread sanitizer cannot run in this container")
    assert out.returncode == 0 and "WARNING: ThreadSanitizer" not in out.stderr, out.stdout + out.stderr[-3000:]
