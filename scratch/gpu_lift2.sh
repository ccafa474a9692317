#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q tests/test_e2pn_gpu.py -k "lifted" 2>&1 | tail -40
