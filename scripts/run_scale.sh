#!/bin/bash
# usage: scripts/run_scale.sh "8 4" [tag]   (inside gpurun --gpus 8); env TB_P2P selects the comm path
mkdir -p gpurun_out
tag=${2:-}
for n in ${1:-8}; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/scale_c5_$n$tag.log 2>&1
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_c5_$n$tag.log 2>&1
  fi
  echo "N=$n$tag exit $?"; grep '^{' gpurun_out/scale_c5_$n$tag.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  value %.4g ms/step %.2f iters %.1f launches %d setup %.1fs'%(d['value'],d['ms_per_step'],d['config']['cg']['iters_per_step_mean'],d['gpu_launches'],d['config']['setup_s']), d['config']['parallelism'][:60], '|', d['config'].get('assembly'))"
  grep -v '^{' gpurun_out/scale_c5_$n$tag.log | grep -iE "error|fail|peer|Traceback" | tail -5 | cut -c1-300
done
