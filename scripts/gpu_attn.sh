#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "spatial_attention" --timeout 120 -p no:cacheprovider 2>&1 | grep -v "^  \|^E   *where\|^E   *+" | tail -25
