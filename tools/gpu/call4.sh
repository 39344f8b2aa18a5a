#!/usr/bin/env bash
# round-2 GPU call 4: multi-tile conv variant with a deep TMA ring: parity, per-shape sweep of the three variants, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P="python -m pytest -m gpu -q --tb=short -p no:cacheprovider -s"
timeout 600 $P tests/test_conv_gpu.py > gpurun_out/r2c4_test_conv_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2c4_test_conv_gpu.log
timeout 600 $P tests/test_model_gpu.py -k "adam" tests/test_fp16_gpu.py -k "adam or deterministic" > gpurun_out/r2c4_misc.log 2>&1; echo "rc=$?" >> gpurun_out/r2c4_misc.log
timeout 900 python tools/sweep_conv.py --what fwd --iters 3 > gpurun_out/r2c4_sweep_fwd.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err; echo "bench rc=$?" >> gpurun_out/r2c4_bench.err
M1_CONV_MULTI_TUNE=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c4_bench_nomulti.json 2> gpurun_out/r2c4_bench_nomulti.err
for f in gpurun_out/r2c4_*.log; do echo "== $f"; grep -E "passed|failed|rc=|per step" $f | tail -4; done
head -c 300 gpurun_out/r2c4_bench.json; echo; head -c 300 gpurun_out/r2c4_bench_nomulti.json
