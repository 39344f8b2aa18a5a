#!/usr/bin/env bash
# round-2 GPU call 17 (2 GPUs): the driver's N=2 launch line with DEFAULT settings must finish (flat all-reduce
# between the graphs); short timeouts - a hang must not burn the budget
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 240 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c17_bench_n2.json 2> gpurun_out/r2c17_bench_n2.err; echo "rc=$?" >> gpurun_out/r2c17_bench_n2.err
timeout 200 $TR bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2c17_bench_ref_n2.json 2> gpurun_out/r2c17_bench_ref_n2.err; echo "rc=$?" >> gpurun_out/r2c17_bench_ref_n2.err
head -c 400 gpurun_out/r2c17_bench_n2.json; echo; tail -2 gpurun_out/r2c17_bench_n2.err | cut -c1-200; head -c 200 gpurun_out/r2c17_bench_ref_n2.json; echo; tail -1 gpurun_out/r2c17_bench_ref_n2.err
