#!/bin/bash
# quick sanity of HEAD after host-side changes
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_engine.py tests/test_gpu_kernels.py -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -3
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -2
