#!/usr/bin/env bash
# round-2 GPU call 7: the driver's own sequence on a fresh box - whole GPU suite in ONE pytest process with -x,
# smoke, default bench (with cpu_baseline), reference arm, then the ncu launch list of a short bench run
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r2c7_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2c7_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c7_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2c7_smoke.log
timeout 900 python bench.py > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err; echo "rc=$?" >> gpurun_out/r2c7_bench.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2c7_bench_ref.json 2> gpurun_out/r2c7_bench_ref.err; echo "rc=$?" >> gpurun_out/r2c7_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2c7_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2c7_ncu_bench.log 2>&1; echo "rc=$?" >> gpurun_out/r2c7_ncu_bench.log
tail -3 gpurun_out/r2c7_pytest_gpu.log; tail -2 gpurun_out/r2c7_smoke.log; head -c 600 gpurun_out/r2c7_bench.json; echo; head -c 400 gpurun_out/r2c7_bench_ref.json; echo; tail -2 gpurun_out/r2c7_ncu_bench.log | head -c 300; wc -l gpurun_out/r2c7_launches.csv
