#!/bin/bash
for pe in 0 21 22; do HV_ATTN_POLY=$pe timeout -s KILL 200 python scripts/attn_bench.py 12 6912 8 40 2>&1 | tail -1; done
for pe in 0 21 22; do HV_ATTN_POLY=$pe timeout -s KILL 200 python scripts/attn_bench.py 48 1728 8 80 2>&1 | tail -1; done
