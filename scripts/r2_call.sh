#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/time_pyrdown.py 2>&1 | tail -3
MA_PYRDOWN_SIMPLE=1 python scripts/time_pyrdown.py 2>&1 | tail -3
