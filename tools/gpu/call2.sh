#!/usr/bin/env bash
# round-2 GPU call 2: which mixed operand formats does tcgen05.mma.kind::f16 accept? (each in its own process:
# an illegal-instruction trap kills the CUDA context) + the rest of the suite without the mixed-format tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P="python -m pytest -m gpu -q --tb=line -p no:cacheprovider"
timeout 300 $P tests/test_fp16_gpu.py -k "wgrad_mixed_formats and dhw0" > gpurun_out/r2c2_mixed_wgrad.log 2>&1   # A = f16, B = bf16
timeout 300 $P tests/test_fp16_gpu.py -k "swapped_roles" > gpurun_out/r2c2_mixed_swapped.log 2>&1                  # A = bf16, B = f16 (MN-major)
timeout 300 $P tests/test_fp16_gpu.py -k "dgrad_mixed_formats and 1-False" > gpurun_out/r2c2_dgrad_bf16w.log 2>&1  # bf16 x bf16 pack in fp16 mode
timeout 600 $P tests/test_fp16_gpu.py -k "not mixed and not swapped and not north_star and not forward_pass" > gpurun_out/r2c2_fp16_rest.log 2>&1
timeout 900 $P tests/test_kernels_gpu.py tests/test_golden.py tests/test_conv_gpu.py > gpurun_out/r2c2_kernels.log 2>&1
M1_DGRAD_W_BF16=1 timeout 900 $P tests/test_model_gpu.py -k "fp32 or adam or inference or fit" > gpurun_out/r2c2_model_fp32.log 2>&1
for f in gpurun_out/r2c2_*.log; do echo "== $f"; tail -4 $f; done
