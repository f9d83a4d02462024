#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -4
OMT_TEST_MATH=3xtf32,fp32 timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
