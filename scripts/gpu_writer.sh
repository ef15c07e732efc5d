#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_model_gpu.py tests/test_pipeline_gpu.py -q -m gpu -p no:cacheprovider -x -s -k "writer or pipeline or banks" > gpurun_out/pytest_writer.log 2>&1
echo "== pytest writer/pipeline exit $?"; grep -E "passed|failed|Error|error|writer|pipeline \(" gpurun_out/pytest_writer.log | tail -14
