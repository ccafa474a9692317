#!/bin/bash
timeout 1500 python -m pytest tests/test_training_gpu.py -m gpu -x -q -s 2>&1 | grep -E "PARITY|passed|failed|Error|error|assert" | head
python scratch/bench_train.py 2>&1 | grep -E "recompute|Error|error" | tail -4
