tag=$1
out=gpurun_out/${tag}_rc.txt
: > $out
for mode in none 1 2; do
  for cfg in "16384 32" "65536 64"; do
    set -- $cfg
    echo "=== CHUB_RCW_BULK_TMA=$mode n=$1 K=$2 fuse=1" >> $out
    if [ $mode = none ]; then unset CHUB_RCW_BULK_TMA; else export CHUB_RCW_BULK_TMA=$mode; fi
    timeout 300 python bench.py --workload rowcyclic --rc-n $1 --rc-k $2 --rc-fuse 1 --steps 3 --warmup 1 --no-cpu 2>> gpurun_out/${tag}_rc.err | python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print('ms_per_update', d.get('ms_per_update'), 'frac', d['roofline'].get('frac'), 'parity', d['roofline'].get('parity'))
" >> $out
  done
done
