#!/bin/bash
timeout 900 python -m pytest tests/test_e2pn_gpu.py -m gpu -x -q -k "rows or kpconv_matches_oracle" 2>&1 | tail -15
timeout 600 python scratch/bench_rows.py 16 2>&1 | tail -12
