#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_boundary.py -m gpu -q 2>&1 | grep -E "passed|failed|^E  |FAILED" | head -30
