#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention" -p no:cacheprovider > gpurun_out/ops_attention.log 2>&1
echo "== ops attention exit $?"; tail -n 5 gpurun_out/ops_attention.log
for t in test_pose_guider_and_camera_encoder_golden test_zero_init_modules_are_exact_noops test_unet_narrow_parity test_unet_narrow_reference_banks_and_cfg test_unet_image_variant_no_motion_module test_unet_full_width_against_reference_golden test_pose_guider_config2_shape_vs_oracle_fp16 test_too_many_frames_is_an_error; do
  timeout -s KILL 600 python -m pytest tests/test_model_gpu.py -m gpu -q -s -k "$t" -p no:cacheprovider > gpurun_out/model_$t.log 2>&1
  echo "== $t exit $?"
  grep -E "passed|failed|Error|error|assert|narrow:|full-width" gpurun_out/model_$t.log | head -12
done
