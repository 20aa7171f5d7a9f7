"""Label-encode a Criteo TSV with the native tool (tools/criteo_preprocess.cpp) -- counterpart of
the reference's examples/criteo_preprocess.py + test/criteo_preprocess.cpp."""
import os
import subprocess
import sys

here = os.path.dirname(os.path.abspath(__file__))
src = os.path.join(here, "..", "tools", "criteo_preprocess.cpp")
exe = os.path.join(here, "..", "tools", "criteo_preprocess")
if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
    subprocess.check_call(["g++", "-O2", "-std=c++17", src, "-o", exe])
sys.exit(subprocess.call([exe] + sys.argv[1:]))
