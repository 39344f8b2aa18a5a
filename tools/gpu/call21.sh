#!/usr/bin/env bash
# round-2 GPU call 21: software-pipelined loads in the gate kernels: parity, isolated GB/s, step time
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P="python -m pytest -m gpu -q --tb=short -p no:cacheprovider -x"
timeout 200 $P tests/test_kernels_gpu.py tests/test_fp16_gpu.py -k "se_ or dropout or determin or north_star" > gpurun_out/r2c21_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/r2c21_kernels.log
timeout 100 python tools/bench_elementwise.py res0x32 res1x64 res2x128 > gpurun_out/r2c21_ew.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c21_bench.json 2> gpurun_out/r2c21_bench.err; echo "rc=$?" >> gpurun_out/r2c21_bench.err
grep -E "passed|failed|rc=" gpurun_out/r2c21_kernels.log | tail -3; grep se_gate gpurun_out/r2c21_ew.log; head -c 220 gpurun_out/r2c21_bench.json | cut -c60-220; echo
