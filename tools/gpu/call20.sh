#!/usr/bin/env bash
# round-2 GPU call 20: ncu --set full over the gate-backward kernels of one step (DRAM traffic of the dominant
# bandwidth family for roofline.bandwidth.traffic)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export M1_CUDA_PROFILER_RANGE=1
timeout 420 ncu --set full --profile-from-start off --clock-control none -k "regex:se_gate_bwd" -c 64 -f -o /tmp/r2c20_bw \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2c20_ncu_bw.log 2>&1
echo "rc=$?" >> gpurun_out/r2c20_ncu_bw.log
ncu -i /tmp/r2c20_bw.ncu-rep --page raw --csv > gpurun_out/r2c20_full_bw.csv 2>> gpurun_out/r2c20_ncu_bw.log
wc -c gpurun_out/r2c20_full_bw.csv; tail -2 gpurun_out/r2c20_ncu_bw.log | cut -c1-200
