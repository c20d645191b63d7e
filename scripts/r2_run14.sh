#!/bin/bash
# why is the split element kernel slow in a back-to-back loop when each launch is fast under ncu?
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active --format=csv,noheader -lms 200 > gpurun_out/clk_$name.csv 2>&1 &
  SMI=$!
  env "$@" timeout 600 python scripts/bench_assembly.py --modes 2 --cells hex --reps 20 > gpurun_out/asm_$name.log 2>&1
  kill $SMI
  python - "$name" <<'PY'
import json,sys
name=sys.argv[1]
for l in open(f'gpurun_out/asm_{name}.log'):
    if l.startswith('{'):
        d=json.loads(l)
        if d['form'] in ('mass','diffusion_tensor'): print('  %-28s %-18s %7.3f ms'%(name, d['form'], d['ms']))
clk=[l.split(',') for l in open(f'gpurun_out/clk_{name}.csv') if 'MHz' in l]
if clk:
    mhz=sorted(int(c[0].split()[0]) for c in clk); pw=max(float(c[1].split()[0]) for c in clk)
    print('     clocks min/median/max %d/%d/%d MHz, max power %.0f W, reasons %s'%(mhz[0], mhz[len(mhz)//2], mhz[-1], pw, sorted(set(c[2].strip() for c in clk))))
PY
}
run split0 TB_ELEMENT_SPLIT=0
run split1_planes1 TB_ELEMENT_SPLIT=1 TB_EA_PLANES=1
run split1_planes0 TB_ELEMENT_SPLIT=1 TB_EA_PLANES=0
run split1_planes0_qb2 TB_ELEMENT_SPLIT=1 TB_EA_PLANES=0 TB_ELEMENT_QB=2
run split1_planes1_again TB_ELEMENT_SPLIT=1 TB_EA_PLANES=1
for v in "1 1" "1 0" "0 0"; do set -- $v
TB_ELEMENT_SPLIT=$1 TB_EA_PLANES=$2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:'k_element_matrices|k_gather_rows' --csv --log-file gpurun_out/launches_asm_s$1p$2.csv python scripts/bench_assembly.py --modes 2 --cells hex --reps 5 > /dev/null 2>&1
python - $1 $2 <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(f'gpurun_out/launches_asm_s{sys.argv[1]}p{sys.argv[2]}.csv')) if len(r)>5 and r[0].isdigit()]
print('launch list split=%s planes=%s:'%(sys.argv[1],sys.argv[2]), ' '.join('%s=%.2f'%(r[4][:22].replace('void ',''), float(r[-1])/ (1e6 if 'ns' in r[-2] or float(r[-1])>1e4 else 1)) for r in rows))
PY
done
