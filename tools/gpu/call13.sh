#!/usr/bin/env bash
# round-2 GPU call 13: weight gradients on a side stream (fork/join inside the step graph), thin weight gradients
# through partial tiles + parallel reduce, augmentation kernels; parity then A/B benches
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P="python -m pytest -m gpu -q --tb=short -p no:cacheprovider"
timeout 900 $P tests/test_augment.py tests/test_conv_gpu.py tests/test_fp16_gpu.py tests/test_model_gpu.py tests/test_golden.py tests/test_zz_graph_gpu.py tests/test_cascade_gpu.py > gpurun_out/r2c13_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2c13_tests.log
M1_WG_SCRATCH=2 timeout 600 $P tests/test_conv_gpu.py tests/test_fullsize_gpu.py > gpurun_out/r2c13_conv_scr2.log 2>&1; echo "rc=$?" >> gpurun_out/r2c13_conv_scr2.log
timeout 200 python tools/bench_conv.py att2_c1 conv3_r2 conv3_r1 conve0 conv2_r0 --what wgrad > gpurun_out/r2c13_wg.log 2>&1
M1_DUMP_PROF=gpurun_out/r2c13_prof_dump.txt timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c13_bench.json 2> gpurun_out/r2c13_bench.err; echo "rc=$?" >> gpurun_out/r2c13_bench.err
M1_WGRAD_STREAM=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c13_bench_noside.json 2> gpurun_out/r2c13_bench_noside.err
M1_WG_SCRATCH=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c13_bench_noscr.json 2> gpurun_out/r2c13_bench_noscr.err
M1_WGRAD_STREAM=0 M1_WG_SCRATCH=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c13_bench_neither.json 2> gpurun_out/r2c13_bench_neither.err
for f in gpurun_out/r2c13_*.log; do echo "== $f"; grep -E "passed|failed|rc=|FAILED|Error" $f | tail -6; done
cat gpurun_out/r2c13_wg.log
for f in gpurun_out/r2c13_bench.json gpurun_out/r2c13_bench_noside.json gpurun_out/r2c13_bench_noscr.json gpurun_out/r2c13_bench_neither.json; do echo $f; head -c 200 $f | cut -c60-200; echo; done
tail -3 gpurun_out/r2c13_bench.err
