#!/bin/bash
timeout 110 python - <<'PY'
import sys, runpy
import pytest
rc = pytest.main(["tests/test_gpu_ops.py", "tests/test_gpu_properties.py", "-m", "gpu", "-x", "-q", "-k", "pyr", "-p", "no:cacheprovider"])
print("pytest rc", rc, flush=True)
sys.argv = ["scripts/time_pyrdown.py"]
runpy.run_path("scripts/time_pyrdown.py", run_name="__main__")
PY
