#!/usr/bin/env bash
# round-2 GPU call 1: full GPU test suite (fp16 mode, deterministic reductions), smoke, epilogue experiment, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c1_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c1_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2c1_smoke.log
timeout 300 python tools/bench_conv.py conv3_r0 sersp0 sersp1 sersp2 sersp3 conv2_r0 --what fwd,dgrad --variant 1 > gpurun_out/r2c1_epi_base.log 2>&1
M1_EPI_TMA=1 timeout 300 python tools/bench_conv.py conv3_r0 sersp0 sersp1 sersp2 sersp3 conv2_r0 --what fwd,dgrad --variant 1 > gpurun_out/r2c1_epi_tma.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
echo "bench rc=$?" >> gpurun_out/r2c1_bench.err
M1_EPI_TMA=1 timeout 200 python -m pytest tests/test_conv_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "halo" > gpurun_out/r2c1_halo_epitma.log 2>&1
echo "halo epi_tma rc=$?" >> gpurun_out/r2c1_halo_epitma.log
tail -5 gpurun_out/r2c1_pytest.log; cat gpurun_out/r2c1_smoke.log | tail -3; cat gpurun_out/r2c1_bench.json | head -c 600
