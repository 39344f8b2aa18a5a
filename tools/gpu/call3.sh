#!/usr/bin/env bash
# round-2 GPU call 3: whole GPU suite, one process per test file (a trap in one file cannot poison the others),
# smoke, a short bench in the fp16 mode
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P="python -m pytest -m gpu -q --tb=short -p no:cacheprovider -s"
for f in test_fp16_gpu test_model_gpu test_zz_graph_gpu test_fullsize_gpu test_kernels_gpu test_golden test_conv_gpu; do
  timeout 900 $P tests/$f.py > gpurun_out/r2c3_$f.log 2>&1
  echo "rc=$?" >> gpurun_out/r2c3_$f.log
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c3_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2c3_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err
echo "bench rc=$?" >> gpurun_out/r2c3_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision bf16 > gpurun_out/r2c3_bench_bf16.json 2> gpurun_out/r2c3_bench_bf16.err
for f in gpurun_out/r2c3_*.log; do echo "== $f"; grep -E "passed|failed|rc=" $f | tail -3; done
head -c 400 gpurun_out/r2c3_bench.json; echo; head -c 400 gpurun_out/r2c3_bench_bf16.json
