#!/usr/bin/env bash
# round-2 GPU call 14: casts + ConvT bias gradients on the side stream: parity + bench; c1 / c4 / c5 bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P="python -m pytest -m gpu -q --tb=short -p no:cacheprovider"
timeout 900 $P tests/test_fp16_gpu.py tests/test_model_gpu.py tests/test_golden.py tests/test_zz_graph_gpu.py tests/test_cascade_gpu.py > gpurun_out/r2c14_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2c14_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c14_bench.json 2> gpurun_out/r2c14_bench.err; echo "rc=$?" >> gpurun_out/r2c14_bench.err
timeout 600 python bench.py --config c1 --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/r2c14_bench_c1.json 2> gpurun_out/r2c14_bench_c1.err; echo "rc=$?" >> gpurun_out/r2c14_bench_c1.err
timeout 600 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c14_bench_c4.json 2> gpurun_out/r2c14_bench_c4.err; echo "rc=$?" >> gpurun_out/r2c14_bench_c4.err
timeout 900 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c14_bench_c5.json 2> gpurun_out/r2c14_bench_c5.err; echo "rc=$?" >> gpurun_out/r2c14_bench_c5.err
for f in gpurun_out/r2c14_*.log; do echo "== $f"; grep -E "passed|failed|rc=|FAILED|Error" $f | tail -6; done
for f in gpurun_out/r2c14_bench.json gpurun_out/r2c14_bench_c1.json gpurun_out/r2c14_bench_c4.json gpurun_out/r2c14_bench_c5.json; do echo $f; head -c 230 $f | cut -c1-230; echo; done
tail -2 gpurun_out/r2c14_bench*.err
