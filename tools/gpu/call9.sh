#!/usr/bin/env bash
# round-2 GPU call 9: gate kernels specialised on the dropout source (parity + isolated GB/s), cascade test, then the
# ncu --set full captures of the step's kernels (profiler range = the eager per-family pass of bench.py)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P="python -m pytest -m gpu -q --tb=short -p no:cacheprovider -x"
timeout 400 $P tests/test_kernels_gpu.py tests/test_fp16_gpu.py > gpurun_out/r2c9_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/r2c9_kernels.log
timeout 600 $P tests/test_cascade_gpu.py tests/test_model_gpu.py -s > gpurun_out/r2c9_model.log 2>&1; echo "rc=$?" >> gpurun_out/r2c9_model.log
timeout 300 python tools/bench_elementwise.py res0x32 res1x64 res2x128 > gpurun_out/r2c9_ew.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c9_bench.json 2> gpurun_out/r2c9_bench.err; echo "rc=$?" >> gpurun_out/r2c9_bench.err
export M1_CUDA_PROFILER_RANGE=1
for spec in "wgrad:wgrad_tc_kernel:170" "conv:conv_tc_kernel|conv_halo_kernel|conv_tc_multi_kernel:330" "bw:se_gate|inorm_|attn_|se_excite|colsum|cast_kernel|pack_tiled|adam_kernel|logits_focal:700"; do
  name="${spec%%:*}"; rest="${spec#*:}"; pat="${rest%:*}"; cnt="${rest##*:}"
  timeout 1200 ncu --set full --profile-from-start off --clock-control none -k "regex:$pat" -c "$cnt" -f -o /tmp/r2c9_$name \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2c9_ncu_$name.log 2>&1
  echo "rc=$?" >> gpurun_out/r2c9_ncu_$name.log
  ncu -i /tmp/r2c9_$name.ncu-rep --page raw --csv > gpurun_out/r2c9_full_$name.csv 2>> gpurun_out/r2c9_ncu_$name.log
  ls -la /tmp/r2c9_$name.ncu-rep >> gpurun_out/r2c9_ncu_$name.log
done
for f in gpurun_out/r2c9_*.log; do echo "== $f"; grep -E "passed|failed|rc=|FAILED|ncu-rep" $f | tail -4; done
cat gpurun_out/r2c9_ew.log | head -30
head -c 230 gpurun_out/r2c9_bench.json; echo
wc -c gpurun_out/r2c9_full_*.csv
