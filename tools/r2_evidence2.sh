#!/bin/bash
# Round-2 evidence, final build: per-layer igemm table, ncu launch list of one C3 step (time + DRAM bytes per launch), ncu --set full
# of the four-group small-head attention kernel and of the lean epilogue with a residual.  Same rules as tools/r2_evidence.sh.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
python tools/igemm_detail.py c3 > gpurun_out/r2_igemm_detail_c3.txt 2> gpurun_out/ev2_detail.err
echo "detail rc $?"
B="bench.py --workload c3 --resident-only --steps 1 --warmup 3 --no-cpu-baseline"
TOTAL=$(python -c "
import runpy, sys
sys.argv = '$B'.split()
runpy.run_path('bench.py', run_name='__main__')
from weatherconverter_b200 import ops
print('TOTAL_LAUNCHES', ops.launch_count())" 2>gpurun_out/ev2_plain_bench.err | grep TOTAL_LAUNCHES | awk '{print $2}')
echo "total launches $TOTAL"
SKIP=$((TOTAL - 470))
python $B > gpurun_out/ev2_plain_bench.log 2>&1 && $NCU --kernel-name-base demangled -k regex:wc:: -s $SKIP --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r2_launches_c3_final_raw.csv python $B > gpurun_out/ev2_ncu_bench.log 2>&1
echo "launch list rc $? ($(wc -l < gpurun_out/r2_launches_c3_final_raw.csv) lines)"
cap() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  "$@" > gpurun_out/ev_plain_$name.log 2>&1 && $NCU --set full --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/r2_$name "$@" > gpurun_out/ev_ncu_$name.log 2>&1
  local rc=$?
  if [ -f /tmp/r2_$name.ncu-rep ]; then
    ncu -i /tmp/r2_$name.ncu-rep --page raw --csv > gpurun_out/r2_ncu_$name.raw.csv 2>/dev/null
    ncu -i /tmp/r2_$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2_ncu_$name.source.csv.gz
  fi
  echo "$name rc $rc"
}
cap attn_small4_hd16 attention_small4 2 python tools/bench_attn.py 32,4,8192,16
cap igemm_1x1_64_256_res_lean igemm 2 python tools/bench_conv.py 32,64,256,64,128,1,1
cap igemm_3x3_64_64 igemm 2 python tools/bench_conv.py 32,64,64,64,128,3,0
du -sh gpurun_out
