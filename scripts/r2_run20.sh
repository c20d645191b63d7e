#!/bin/bash
# 2 GPUs: compute-sanitizer on the new kernels (1 GPU), then the driver's N = 2 command and the distributed checks after this
# session's changes (split element kernel, stream-ordered assembly)
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
{
echo "## memcheck: smoke (assembly split kernel, persistent CG variant 2, cell sweep)"; echo '```'
timeout 600 $S --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "smoke ok|ERROR SUMMARY|Invalid|Error" | head -8; echo '```'
echo "## memcheck: assembly (split element kernel + plane-major gather, all cell types), traced source programs"; echo '```'
timeout 900 $S --tool memcheck python -m pytest tests/test_gpu_assembly_gather.py tests/test_source_program.py -m gpu -q --timeout=800 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid" | head -8; echo '```'
echo "## racecheck: split element kernel (shared-memory phases A / B), hex + tet"; echo '```'
timeout 900 $S --tool racecheck python -m pytest tests/test_gpu_assembly_gather.py -m gpu -q --timeout=800 -k "matrices_bitwise" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | head -8; echo '```'
echo "## memcheck + racecheck: persistent CG variants (ticket barrier), block-Jacobi lower-triangle apply"; echo '```'
timeout 900 $S --tool memcheck python -m pytest tests/test_gpu_spmv_cg.py tests/test_gpu_precond.py -m gpu -q --timeout=800 -k "persistent_variants or block_jacobi or precond" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid" | head -8
timeout 900 $S --tool racecheck python -m pytest tests/test_gpu_precond.py -m gpu -q --timeout=800 -k "block_jacobi" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | head -8; echo '```'
} > gpurun_out/compute_sanitizer_r02.md 2>&1
cat gpurun_out/compute_sanitizer_r02.md
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $T --master-port 29631 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_c5_n2.log 2>&1; echo "c5 n2 exit $?"
timeout 600 $T --master-port 29632 bench.py --gpus 2 --workload c4 --precond block_jacobi --bj-rows 64 --steps 3 --warmup 3 --e2e-steps 0 > gpurun_out/bench_c4_n2.log 2>&1; echo "c4 n2 exit $?"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=700 -k "2]" > gpurun_out/pytest_multi2.log 2>&1; echo "multi exit $?"; tail -n 3 gpurun_out/pytest_multi2.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c*_n2.log')):
    ok=False
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); ok=True
            e=d.get('e2e') or {}
            print(f, 'N=%d value %.4g ms/step %.2f its %s e2e %s parity %s'%(d['n_gpus'], d['value'], d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], e.get('value'), (d.get('parity') or {}).get('ok')))
    if not ok: print(f, 'NO JSON', open(f).read()[-2500:])
PY
