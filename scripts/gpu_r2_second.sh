#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_config5_gpu.py tests/test_pipeline_gpu.py tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_second.log 2>&1; echo "== pytest rc=$?"; tail -25 gpurun_out/pytest_second.log
for mw in 1 2; do for m in 0 4; do HV_ATTN_MW=$mw HV_ATTN_POLY=$m timeout -s KILL 200 python scripts/attn_bench.py 12 6912 8 40 2>&1 | tail -1 | sed "s/^/MW=$mw /"; done; done | tee gpurun_out/attn_ab2.log
for pp in 1 2; do for mw in 1 2; do HV_ATTN_PP=$pp HV_ATTN_MW=$mw timeout -s KILL 200 python scripts/attn_bench.py 12 6912 8 40 2>&1 | tail -1 | sed "s/^/PP=$pp MW=$mw /"; done; done | tee gpurun_out/attn_ab3.log
HV_ATTN_PP=1 HV_ATTN_MW=2 HV_ATTN_POLY=31 timeout -s KILL 200 python scripts/attn_bench.py 12 6912 8 40 > gpurun_out/attn_phases_tok_mw2.log 2>&1; tail -9 gpurun_out/attn_phases_tok_mw2.log
HV_ATTN_MW=2 timeout -s KILL 200 python scripts/attn_bench.py 4 9216 8 40 2>&1 | tail -1
timeout 900 python scripts/error_ladder.py --out gpurun_out/error_ladder_config2.txt > gpurun_out/ladder.log 2>&1; echo "== ladder rc=$?"; tail -4 gpurun_out/ladder.log
HV_TRACE=gpurun_out/trace_c2_b.csv timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/bench_c2_b.log 2>&1; echo "== bench rc=$?"; tail -n 1 gpurun_out/bench_c2_b.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['op_profile'])"
HV_ATTN_MW=2 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/bench_c2_mw2.log 2>&1; echo "== bench mw2 rc=$?"; tail -n 1 gpurun_out/bench_c2_mw2.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['op_profile'])"
