#!/usr/bin/env bash
# round-2 GPU call 5: bench cfg-2 with / without the multi-tile variant in the autotune, cfg-1, cfg-4
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err; echo "rc=$?" >> gpurun_out/r2c5_bench.err
M1_CONV_MULTI_TUNE=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c5_bench_nomulti.json 2> gpurun_out/r2c5_bench_nomulti.err
timeout 600 python bench.py --config c1 --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/r2c5_bench_c1.json 2> gpurun_out/r2c5_bench_c1.err; echo "rc=$?" >> gpurun_out/r2c5_bench_c1.err
timeout 600 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c5_bench_c4.json 2> gpurun_out/r2c5_bench_c4.err; echo "rc=$?" >> gpurun_out/r2c5_bench_c4.err
timeout 300 python -m pytest -m gpu -q --tb=short -p no:cacheprovider tests/test_model_gpu.py -k "inference" > gpurun_out/r2c5_infer.log 2>&1
for f in gpurun_out/r2c5_*.json; do echo "== $f"; head -c 250 $f; echo; done
tail -3 gpurun_out/r2c5_*.err; tail -3 gpurun_out/r2c5_infer.log
