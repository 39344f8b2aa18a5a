#!/usr/bin/env bash
# round-2 GPU call 18: split-K partial tiles for EVERY weight gradient whose tiles fit the scratch (M1_WG_SCRATCH=2) vs
# thin launches only (default 1): parity + A/B bench + isolated timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P="python -m pytest -m gpu -q --tb=short -p no:cacheprovider"
M1_WG_SCRATCH=2 timeout 600 $P tests/test_conv_gpu.py tests/test_fullsize_gpu.py tests/test_fp16_gpu.py > gpurun_out/r2c18_conv_scr2.log 2>&1; echo "rc=$?" >> gpurun_out/r2c18_conv_scr2.log
S="convtd2 sersp3 sersp2 sersp0 serse2 att2_c1 conve0 conv2_r0"
echo "== scratch 1" > gpurun_out/r2c18_wg.log; timeout 200 python tools/bench_conv.py $S --what wgrad >> gpurun_out/r2c18_wg.log 2>&1
echo "== scratch 2" >> gpurun_out/r2c18_wg.log; M1_WG_SCRATCH=2 timeout 200 python tools/bench_conv.py $S --what wgrad >> gpurun_out/r2c18_wg.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c18_bench.json 2> gpurun_out/r2c18_bench.err; echo "rc=$?" >> gpurun_out/r2c18_bench.err
M1_WG_SCRATCH=2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c18_bench_scr2.json 2> gpurun_out/r2c18_bench_scr2.err
grep -E "passed|failed|rc=" gpurun_out/r2c18_conv_scr2.log | tail -3; cat gpurun_out/r2c18_wg.log
for f in gpurun_out/r2c18_bench.json gpurun_out/r2c18_bench_scr2.json; do echo $f; head -c 200 $f | cut -c60-200; echo; done
