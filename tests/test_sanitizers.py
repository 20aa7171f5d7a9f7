"""Sanitizer builds of the native CPU engine (the reference has none): a threaded
pull/push/update stress compiled with ThreadSanitizer and with Address+UB sanitizers."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = [os.path.join(ROOT, "tests", "cpp", "core_stress.cpp"), os.path.join(ROOT, "openembedding_b200", "csrc", "core", "exb_core.cpp")]


@pytest.mark.parametrize("san", ["thread", "address,undefined"])
def test_core_engine_under_sanitizer(san):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = os.path.join(tempfile.mkdtemp(), "core_stress")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=" + san, "-I" + os.path.dirname(SRC[1])] + SRC
                       + ["-lpthread", "-o", exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    if r.returncode != 0 and ("cannot find" in r.stdout or "unrecognized" in r.stdout):
        pytest.skip("sanitizer runtime not installed: " + r.stdout[-200:])
    assert r.returncode == 0, r.stdout[-2000:]
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                       env=dict(os.environ, TSAN_OPTIONS="halt_on_error=1", ASAN_OPTIONS="detect_leaks=1"))
    assert r.returncode == 0 and "CORE_STRESS_OK" in r.stdout, r.stdout[-3000:]
    assert "WARNING: ThreadSanitizer" not in r.stdout and "ERROR: AddressSanitizer" not in r.stdout and "runtime error" not in r.stdout
