#!/bin/bash
# memcheck over the small-scene GPU tests (goldens, fused pieces, losses, integrate fixtures)
mkdir -p gpurun_out
export CUDA_LAUNCH_BLOCKING=0
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 30 \
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_losses.py tests/test_gpu_integrate.py tests/test_gpu_fused.py -q -m gpu \
  -k "golden or fixture or autograd_module or c_abi or ssim or normal_consistency or densification or activate_forward or sh_sizes or ragged or long_tile or one_pixel or many_points" \
  > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|passed|failed" gpurun_out/sanitize_memcheck.log | sort | uniq -c | head -20
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_losses.py tests/test_gpu_parity.py -q -m gpu -k "ssim_and_l1 or depth_normal or golden" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/sanitize_racecheck.log | sort | uniq -c | head -20
