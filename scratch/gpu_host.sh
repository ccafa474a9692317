#!/bin/bash
python scratch/host_overhead.py 2>&1 | head -60
