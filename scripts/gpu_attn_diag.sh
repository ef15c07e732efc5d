#!/bin/bash
mkdir -p gpurun_out
for c in "2 384 8 40" "2 128 8 80" "2 256 8 80" "2 300 8 80" "2 384 8 80" "2 1728 8 40" "2 108 8 16" "1 6912 8 40" "6 432 8 40 432 3 3" "4 200 8 80 200 2 2" "2 640 8 160"; do
  timeout -s KILL 120 python scripts/attn_diag.py $c 2>&1 | grep CASE | tee -a gpurun_out/attn_diag.log
done
echo "--- sanitizer on d=80 L=300" | tee -a gpurun_out/attn_diag.log
timeout -s KILL 300 compute-sanitizer --tool memcheck python scripts/attn_diag.py 2 300 8 80 2>&1 | grep -v "^$" | head -60 | tee -a gpurun_out/attn_sanitizer.log
