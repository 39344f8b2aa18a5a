#!/usr/bin/env bash
# round-2 GPU call 16: the driver's own sequence on a fresh box (whole GPU suite in ONE pytest process with -x, smoke,
# default bench with cpu_baseline, reference arm), then the ncu launch list of one step (profiler range = eager pass)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r2c16_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2c16_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c16_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2c16_smoke.log
timeout 900 python bench.py > gpurun_out/r2c16_bench.json 2> gpurun_out/r2c16_bench.err; echo "rc=$?" >> gpurun_out/r2c16_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r2c16_bench_ref.json 2> gpurun_out/r2c16_bench_ref.err; echo "rc=$?" >> gpurun_out/r2c16_bench_ref.err
M1_CUDA_PROFILER_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --profile-from-start off --clock-control none --csv --log-file gpurun_out/r2c16_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2c16_ncu_bench.log 2>&1; echo "rc=$?" >> gpurun_out/r2c16_ncu_bench.log
tail -3 gpurun_out/r2c16_pytest_gpu.log; tail -2 gpurun_out/r2c16_smoke.log; head -c 300 gpurun_out/r2c16_bench.json; echo; head -c 300 gpurun_out/r2c16_bench_ref.json; echo; wc -l gpurun_out/r2c16_launches.csv; tail -2 gpurun_out/r2c16_ncu_bench.log | cut -c1-200
