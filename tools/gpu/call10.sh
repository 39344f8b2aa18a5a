#!/usr/bin/env bash
# round-2 GPU call 10: L2 prefetch in the conv / wgrad producers, fused attention kernels, bf16 twins from the conv
# epilogue, forward gate kernel back to run-time dropout dispatch: parity, then A/B benches
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P="python -m pytest -m gpu -q --tb=short -p no:cacheprovider -x"
timeout 600 $P tests/test_kernels_gpu.py tests/test_fp16_gpu.py tests/test_conv_gpu.py > gpurun_out/r2c10_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/r2c10_kernels.log
timeout 900 $P tests/test_model_gpu.py tests/test_golden.py tests/test_zz_graph_gpu.py tests/test_cascade_gpu.py tests/test_fullsize_gpu.py > gpurun_out/r2c10_model.log 2>&1; echo "rc=$?" >> gpurun_out/r2c10_model.log
M1_DUMP_PROF=gpurun_out/r2c10_prof_dump.txt timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c10_bench.json 2> gpurun_out/r2c10_bench.err; echo "rc=$?" >> gpurun_out/r2c10_bench.err
M1_CONV_PREFETCH=0 M1_WG_PREFETCH=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c10_bench_nopf.json 2> gpurun_out/r2c10_bench_nopf.err
M1_CONV_PREFETCH=2 M1_WG_PREFETCH=8 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c10_bench_pf2.json 2> gpurun_out/r2c10_bench_pf2.err
M1_ATTN_FUSED=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c10_bench_noattn.json 2> gpurun_out/r2c10_bench_noattn.err
timeout 300 python tools/bench_elementwise.py res0x32 res1x64 > gpurun_out/r2c10_ew.log 2>&1
for f in gpurun_out/r2c10_*.log; do echo "== $f"; grep -E "passed|failed|rc=|FAILED|Error" $f | tail -4; done
for f in gpurun_out/r2c10_bench.json gpurun_out/r2c10_bench_nopf.json gpurun_out/r2c10_bench_pf2.json gpurun_out/r2c10_bench_noattn.json; do echo $f; head -c 200 $f | cut -c60-200; echo; done
tail -3 gpurun_out/r2c10_bench.err; grep se_gate gpurun_out/r2c10_ew.log
