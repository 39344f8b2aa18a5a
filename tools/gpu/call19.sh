#!/usr/bin/env bash
# round-2 GPU call 19: final state - the driver's sequence (whole GPU suite in one pytest process with -x, smoke,
# default bench) + the other configs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r2c19_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2c19_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c19_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2c19_smoke.log
timeout 900 python bench.py > gpurun_out/r2c19_bench.json 2> gpurun_out/r2c19_bench.err; echo "rc=$?" >> gpurun_out/r2c19_bench.err
timeout 300 python bench.py --config c1 --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/r2c19_bench_c1.json 2> gpurun_out/r2c19_bench_c1.err
timeout 400 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c19_bench_c4.json 2> gpurun_out/r2c19_bench_c4.err
timeout 500 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c19_bench_c5.json 2> gpurun_out/r2c19_bench_c5.err
tail -3 gpurun_out/r2c19_pytest_gpu.log; tail -2 gpurun_out/r2c19_smoke.log
for f in gpurun_out/r2c19_bench.json gpurun_out/r2c19_bench_c1.json gpurun_out/r2c19_bench_c4.json gpurun_out/r2c19_bench_c5.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['value'],2), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), d['clocks'])
PY
done
