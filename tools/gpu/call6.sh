#!/usr/bin/env bash
# round-2 GPU call 6: cascade parity, cfg-5 bench, per-launch profile dump of the cfg-2 step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P="python -m pytest -m gpu -q --tb=short -p no:cacheprovider -s"
timeout 900 $P tests/test_cascade_gpu.py > gpurun_out/r2c6_test_cascade.log 2>&1; echo "rc=$?" >> gpurun_out/r2c6_test_cascade.log
timeout 900 $P tests/test_model_gpu.py tests/test_golden.py > gpurun_out/r2c6_test_model.log 2>&1; echo "rc=$?" >> gpurun_out/r2c6_test_model.log
timeout 900 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c6_bench_c5.json 2> gpurun_out/r2c6_bench_c5.err; echo "rc=$?" >> gpurun_out/r2c6_bench_c5.err
M1_DUMP_PROF=gpurun_out/r2c6_prof_dump.txt timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err; echo "rc=$?" >> gpurun_out/r2c6_bench.err
for f in gpurun_out/r2c6_*.log; do echo "== $f"; grep -E "passed|failed|rc=" $f | tail -3; done
head -c 300 gpurun_out/r2c6_bench_c5.json; echo; tail -n 5 gpurun_out/r2c6_bench_c5.err
