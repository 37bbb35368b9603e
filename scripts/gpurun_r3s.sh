#!/bin/bash
timeout 300 python scripts/debug_e2e.py 4 2>&1 | grep -v Warn
timeout 300 python scripts/debug_e2e.py 1 2>&1 | grep -v Warn
