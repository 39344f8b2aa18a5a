#!/usr/bin/env bash
# round-2 GPU call 11: split-K weight gradient through a scratch buffer + reduce kernel (no atomics on thin launches)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P="python -m pytest -m gpu -q --tb=short -p no:cacheprovider -x"
M1_WG_SCRATCH=2 timeout 600 $P tests/test_conv_gpu.py tests/test_fp16_gpu.py tests/test_fullsize_gpu.py > gpurun_out/r2c11_conv_scr2.log 2>&1; echo "rc=$?" >> gpurun_out/r2c11_conv_scr2.log
timeout 900 $P tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_golden.py tests/test_zz_graph_gpu.py > gpurun_out/r2c11_model.log 2>&1; echo "rc=$?" >> gpurun_out/r2c11_model.log
M1_DUMP_PROF=gpurun_out/r2c11_prof_dump.txt timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c11_bench.json 2> gpurun_out/r2c11_bench.err; echo "rc=$?" >> gpurun_out/r2c11_bench.err
M1_WG_SCRATCH=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c11_bench_noscr.json 2> gpurun_out/r2c11_bench_noscr.err
M1_WG_SCRATCH=2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c11_bench_scr2.json 2> gpurun_out/r2c11_bench_scr2.err
for f in gpurun_out/r2c11_*.log; do echo "== $f"; grep -E "passed|failed|rc=|FAILED|Error" $f | tail -4; done
for f in gpurun_out/r2c11_bench.json gpurun_out/r2c11_bench_noscr.json gpurun_out/r2c11_bench_scr2.json; do echo $f; head -c 200 $f | cut -c60-200; echo; done
tail -3 gpurun_out/r2c11_bench.err
