"""Host-side machinery of the .simlod streamer (long-lived loader threads, streaming copy) under the sanitizers —
CPU only; the GPU suite covers the streamer end to end (tests/test_stream_file.py)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "loader_pool_test.cpp")


@pytest.mark.parametrize("sanitizer", ["thread", "address,undefined"])
def test_loader_pool_and_streaming_copy_under_sanitizers(tmp_path, sanitizer):
    exe = str(tmp_path / "loader_pool_test")
    build = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-pthread", "-fsanitize=" + sanitizer, "-fno-sanitize-recover=all", SRC, "-o", exe],
                           capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr:
        pytest.skip("sanitizer runtime not installed: " + build.stderr.splitlines()[-1])
    assert build.returncode == 0, build.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "loader pool ok" in run.stdout
    assert "ThreadSanitizer" not in run.stderr and "AddressSanitizer" not in run.stderr and "runtime error" not in run.stderr, run.stderr
