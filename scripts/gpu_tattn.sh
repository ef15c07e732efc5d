#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -k temporal > gpurun_out/pytest_tattn.log 2>&1
echo "== pytest temporal exit $?"; tail -8 gpurun_out/pytest_tattn.log
for v in 0 1; do HV_TATTN_TMA=$v timeout -s KILL 300 python - <<'PY'
import os, torch, sys
sys.path.insert(0, os.getcwd())
from humanvid_b200._native import check, i32, i64, lib, ptr, stream
for (B, Fr, HW, heads, d) in [(2, 24, 6912, 8, 40), (2, 24, 1728, 8, 80)]:
    Cc = heads * d
    qkv = torch.randn(B * Fr * HW, 3 * Cc, device="cuda").half()
    out = torch.zeros(B * Fr * HW, Cc, device="cuda", dtype=torch.half)
    run = lambda: check(lib().hv_op_temporal_attention(ptr(qkv), ptr(out), i64(B), i64(Fr), i64(HW), i32(heads), i32(d), stream()))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = (qkv.numel() + out.numel()) * 2 / 1e9
    print(f"HV_TATTN_TMA={os.environ.get('HV_TATTN_TMA')} temporal attention HW={HW} d={d}: {ms:.3f} ms  {gb / ms:.2f} TB/s", flush=True)
PY
done
