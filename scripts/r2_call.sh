#!/bin/bash
python scripts/time_nmi.py 2>&1 | grep -v Warn | tail -26
