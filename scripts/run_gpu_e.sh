#!/bin/bash
# 2-GPU box: peer-memory path vs NCCL path
mkdir -p gpurun_out
run2() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
TB_P2P=1 run2 29611 tests/dist_check.py > gpurun_out/dist_check2_p2p.log 2>&1; echo "dist_check p2p exit $?"; grep -E "rank|peer|Error|error" gpurun_out/dist_check2_p2p.log | tail -6 | cut -c1-250
TB_P2P=0 run2 29612 tests/dist_check.py > gpurun_out/dist_check2_nccl.log 2>&1; echo "dist_check nccl exit $?"; grep -E "rank" gpurun_out/dist_check2_nccl.log | tail -3 | cut -c1-250
for mode in 1 0; do
TB_P2P=$mode run2 2962$mode bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_c5_2gpu_p2p$mode.log 2>&1; echo "bench 2gpu TB_P2P=$mode exit $?"
grep '^{' gpurun_out/bench_c5_2gpu_p2p$mode.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  value %.4g ms/step %.2f iters %.1f launches %d'%(d['value'],d['ms_per_step'],d['config']['cg']['iters_per_step_mean'],d['gpu_launches']), d['config']['parallelism'])"
grep -v '^{' gpurun_out/bench_c5_2gpu_p2p$mode.log | grep -iE "error|fail|peer" | tail -5 | cut -c1-300
done
# 1 GPU: assembly tests + bench after the kernel changes
timeout 300 python -m pytest tests/test_gpu_assembly_gather.py tests/test_gpu_assembly.py -m gpu -q -x --timeout=300 2>&1 | tail -3
CUDA_VISIBLE_DEVICES=0 timeout 600 python scripts/bench_assembly.py --cells hex,tet > gpurun_out/bench_assembly.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/bench_assembly.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['workload'], d['form'], 'mode', d['mode_requested'], d['mode_used'], 'chunks', d['chunks'], 'ms %.2f'%d['ms'], 'Mel/s %.1f'%(d['elements_per_s']/1e6), 'GB/s %.0f frac %.3f'%(d['achieved_gbs'], d['frac']))
    else: print(l.rstrip()[-300:])
PY
