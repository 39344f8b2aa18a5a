#!/usr/bin/env bash
# round-2 GPU call 12: augmentation kernels vs the oracle; thin weight-gradient experiments (what bounds them?)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P="python -m pytest -m gpu -q --tb=short -p no:cacheprovider"
timeout 600 $P tests/test_augment.py > gpurun_out/r2c12_augment.log 2>&1; echo "rc=$?" >> gpurun_out/r2c12_augment.log
timeout 600 $P tests/test_model_gpu.py -k "shared_trunk" > gpurun_out/r2c12_shared.log 2>&1; echo "rc=$?" >> gpurun_out/r2c12_shared.log
S="att2_c1 conv3_r2 conv3_r1 conve0 conv2_r0 convtd2"
echo "== default" > gpurun_out/r2c12_wg.log; timeout 200 python tools/bench_conv.py $S --what wgrad >> gpurun_out/r2c12_wg.log 2>&1
echo "== NOLOAD (MMA + barriers + epilogue only)" >> gpurun_out/r2c12_wg.log; M1_WG_NOLOAD=1 timeout 200 python tools/bench_conv.py $S --what wgrad >> gpurun_out/r2c12_wg.log 2>&1
for st in 3 4 6; do echo "== STAGES $st" >> gpurun_out/r2c12_wg.log; M1_WG_STAGES=$st timeout 200 python tools/bench_conv.py $S --what wgrad >> gpurun_out/r2c12_wg.log 2>&1; done
for kv in 32 64; do echo "== KV $kv STAGES 6" >> gpurun_out/r2c12_wg.log; M1_WG_KV=$kv M1_WG_STAGES=6 timeout 200 python tools/bench_conv.py $S --what wgrad >> gpurun_out/r2c12_wg.log 2>&1; done
echo "== fwd/dgrad of the same shapes" >> gpurun_out/r2c12_wg.log; timeout 200 python tools/bench_conv.py $S --what fwd,dgrad >> gpurun_out/r2c12_wg.log 2>&1
for f in gpurun_out/r2c12_*.log; do echo "== $f"; grep -E "passed|failed|rc=|FAILED|Error" $f | tail -4; done
cat gpurun_out/r2c12_wg.log
