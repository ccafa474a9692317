#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q tests/test_e2pn_gpu.py 2>&1 | grep -E "^E|Error|assert|passed|failed" | head -20
