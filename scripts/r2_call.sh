#!/bin/bash
timeout 100 python - <<'PY'
import runpy, sys
import __graft_entry__ as g
g.smoke(); print("smoke ok", flush=True)
sys.argv = ["scripts/e2e_pageable.py"]
runpy.run_path("scripts/e2e_pageable.py", run_name="__main__")
PY
