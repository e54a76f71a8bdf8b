#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_models.py -m gpu -q --tb=short -x -k "cnn1d" > gpurun_out/test_c1d.log 2>&1
echo "tests exit $? $(tail -n 1 gpurun_out/test_c1d.log)"
grep -h "FAILED\|Error" gpurun_out/test_c1d.log | head
timeout 300 python tools/c1d_rate.py
