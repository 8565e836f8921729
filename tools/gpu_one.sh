#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "enhance" 2>&1 | grep -E "enhance |passed|failed|^E " | cut -c1-250
