#!/bin/bash
# Round-2 evidence: ncu launch list of one C3 step (time + DRAM bytes per launch) and ncu --set full captures of the kernels the
# roofline claims rest on.  Every ncu command runs only after the same command has exited 0 without ncu (&&).  The .ncu-rep
# files are exported to CSV (raw + source pages) on the box and removed: gpurun_out/ may carry at most 64 MiB back.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
B="bench.py --workload c3 --resident-only --steps 1 --warmup 3 --no-cpu-baseline"
# total number of library launches of that command (plan building + warm-up + the one timed step): skip all but the last ~470
TOTAL=$(python -c "
import runpy, sys
sys.argv = '$B'.split()
runpy.run_path('bench.py', run_name='__main__')
from weatherconverter_b200 import ops
print('TOTAL_LAUNCHES', ops.launch_count())" 2>gpurun_out/ev_plain_bench.err | grep TOTAL_LAUNCHES | awk '{print $2}')
echo "total launches $TOTAL"
SKIP=$((TOTAL - 470))
python $B > gpurun_out/ev_plain_bench.log 2>&1 && $NCU --kernel-name-base demangled -k regex:wc:: -s $SKIP --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r2_launches_c3_raw.csv python $B > gpurun_out/ev_ncu_bench.log 2>&1
echo "launch list rc $? ($(wc -l < gpurun_out/r2_launches_c3_raw.csv) lines)"
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
cap igemm_3x3_256_b16 igemm 2 python tools/bench_conv.py 16,256,256,64,128,3,0
cap igemm_3x3_256_b32 igemm 2 python tools/bench_conv.py 32,256,256,64,128,3,0
cap igemm_3x3_768 igemm 2 python tools/bench_conv.py 32,768,768,16,32,3,0
cap igemm_1x1_64_256_lean igemm 2 python tools/bench_conv.py 32,64,256,64,128,1,0
cap attn_hd64 attention_kernel 2 python tools/bench_attn.py 16,4,8192,64
cap attn_hd128 attention_kernel 2 python tools/bench_attn.py 16,4,2048,128
cap attn_small_hd16 attention_small 2 python tools/bench_attn.py 32,4,8192,16
cap gn_stats gn_stats 6 python tools/bench_elementwise.py
cap gn_apply gn_apply 6 python tools/bench_elementwise.py
cap ddpm_step ddpm_step 3 python tools/bench_elementwise.py
cap sgg_update sgg_update 3 python tools/bench_elementwise.py
du -sh gpurun_out
