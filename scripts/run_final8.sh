#!/bin/bash
# 8-GPU box: final scaling table (peer path), cut comparison, NCCL reference point, correctness at 8 ranks
mkdir -p gpurun_out
for cut in planes rows; do
DIST_CUT=$cut TB_P2P=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 tests/dist_check.py > gpurun_out/dist_check8_$cut.log 2>&1; echo "dist_check8 cut=$cut exit $?"
done
n=8
for cut in rows planes; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2958$n bench.py --gpus $n --steps 5 --warmup 3 --cut $cut > gpurun_out/scale_c5_${n}_$cut.log 2>&1
echo "N=$n cut=$cut exit $?"; grep '^{' gpurun_out/scale_c5_${n}_$cut.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  value %.4g ms/step %.2f iters %.1f e2e %.4g'%(d['value'],d['ms_per_step'],d['config']['cg']['iters_per_step_mean'],d['e2e']['value']), d['config']['parallelism'][-40:])"
done
TB_P2P=1 scripts/run_scale.sh "4 2 1" _final
