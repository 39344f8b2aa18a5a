#!/usr/bin/env bash
# round-2 GPU call 8: staged (bulk-copy ring) bandwidth kernels, shared stem+serse1 trunk, tiled weight re-pack,
# atomics-free SE excite backward: kernel parity first (short timeouts: a hang must not take the box), then
# isolated GB/s staged vs register skeletons, model parity, bench + per-launch dump
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P="python -m pytest -m gpu -q --tb=short -p no:cacheprovider -x"
timeout 300 $P tests/test_kernels_gpu.py > gpurun_out/r2c8_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/r2c8_kernels.log
timeout 400 $P tests/test_fp16_gpu.py > gpurun_out/r2c8_fp16.log 2>&1; echo "rc=$?" >> gpurun_out/r2c8_fp16.log
if grep -q "rc=0" gpurun_out/r2c8_kernels.log; then
  timeout 300 python tools/bench_elementwise.py > gpurun_out/r2c8_ew_staged.log 2>&1
  M1_STAGED=0 timeout 300 python tools/bench_elementwise.py > gpurun_out/r2c8_ew_regs.log 2>&1
fi
timeout 900 $P tests/test_model_gpu.py tests/test_golden.py tests/test_zz_graph_gpu.py > gpurun_out/r2c8_model.log 2>&1; echo "rc=$?" >> gpurun_out/r2c8_model.log
timeout 600 $P tests/test_cascade_gpu.py tests/test_fullsize_gpu.py > gpurun_out/r2c8_cascade_full.log 2>&1; echo "rc=$?" >> gpurun_out/r2c8_cascade_full.log
M1_DUMP_PROF=gpurun_out/r2c8_prof_dump.txt timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err; echo "rc=$?" >> gpurun_out/r2c8_bench.err
M1_STAGED=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c8_bench_nostaged.json 2> gpurun_out/r2c8_bench_nostaged.err
M1_SHARE_TRUNK=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c8_bench_noshare.json 2> gpurun_out/r2c8_bench_noshare.err
for f in gpurun_out/r2c8_*.log; do echo "== $f"; grep -E "passed|failed|rc=|FAILED" $f | tail -4; done
paste gpurun_out/r2c8_ew_staged.log gpurun_out/r2c8_ew_regs.log | awk '{print $1,$2,$3,$5,"|",$14,$16}' | head -40
for f in gpurun_out/r2c8_bench.json gpurun_out/r2c8_bench_nostaged.json gpurun_out/r2c8_bench_noshare.json; do head -c 230 $f | cut -c1-230; echo; done
tail -3 gpurun_out/r2c8_bench.err
