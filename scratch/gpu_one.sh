#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q tests/test_points_gpu.py -k "cap or batch" 2>&1 | tail -5
