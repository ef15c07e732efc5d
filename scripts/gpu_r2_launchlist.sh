#!/bin/bash
# HBM ceilings (write / read / copy) and the ncu launch list of the bench command with DRAM bytes per launch (profiles/r02_dram_traffic.*,
# profiles/r02_ncu_launches_step.csv).  ncu replays every launch: numbers printed by this bench run are not bench values.
mkdir -p gpurun_out
timeout 120 python scripts/write_bw.py 2>&1 | tee gpurun_out/write_bw.log
timeout -s KILL 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --launch-skip 3000 \
  -k 'regex:gemm_kernel|attn_|gn_|layernorm|small_linear|nhwc|ncfhw|timestep|temporal|smallconv|pg_conv|unshuffle|conv3x3_direct|window_gather|cfg_ddim|advance_index' \
  --csv --log-file gpurun_out/dram.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager --no-extras > gpurun_out/ncu_dram_bench.log 2>&1
echo "== ncu exit $?"; tail -n 1 gpurun_out/ncu_dram_bench.log | cut -c1-160; wc -l gpurun_out/dram.csv; du -sh gpurun_out/dram.csv
python scripts/summarize_dram.py gpurun_out/dram.csv gpurun_out/r02_dram_traffic gpurun_out/r02_ncu_launches_step.csv | tail -20
