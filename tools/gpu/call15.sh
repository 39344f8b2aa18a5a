#!/usr/bin/env bash
# round-2 GPU call 15 (2 GPUs): data-parallel correctness (R ranks x b == 1 process x R*b) with the side-stream weight
# gradients + bucketed all-reduce captured in the backward graph, then the 2-GPU bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR tools/check_dp.py > gpurun_out/r2c15_check_dp.log 2>&1; echo "rc=$?" >> gpurun_out/r2c15_check_dp.log
NCCL_DEBUG=WARN timeout 900 $TR bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2c15_bench_n2.json 2> gpurun_out/r2c15_bench_n2.err; echo "rc=$?" >> gpurun_out/r2c15_bench_n2.err
M1_CUDA_GRAPH_DP=flat timeout 900 $TR bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2c15_bench_n2_flat.json 2> gpurun_out/r2c15_bench_n2_flat.err
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2c15_bench_n1.json 2> gpurun_out/r2c15_bench_n1.err
tail -4 gpurun_out/r2c15_check_dp.log
for f in gpurun_out/r2c15_bench_n2.json gpurun_out/r2c15_bench_n2_flat.json gpurun_out/r2c15_bench_n1.json; do echo $f; head -c 230 $f | cut -c60-230; echo; done
grep -i "capturing\|flat all-reduce\|error" gpurun_out/r2c15_bench_n2.err | head -5
