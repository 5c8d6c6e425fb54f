#!/bin/bash
# Final-build record of round 2: full GPU suite, bench lines of every workload (+ reference arm), per-class tables, ncu launch list of
# one C3 step (time + DRAM bytes per launch) and one ncu --set full capture of the head_dim 32 four-group attention kernel.
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2_final_gputest.log 2>&1; echo "gpu tests rc $? : $(tail -1 gpurun_out/r2_final_gputest.log)"
python bench.py > gpurun_out/r2_bench_c3_final.json 2> gpurun_out/r2_bench_c3_final.err; echo "c3 rc $?"
python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r2_bench_c2_final.json 2> gpurun_out/r2_bench_c2_final.err; echo "c2 rc $?"
python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r2_bench_c5_final.json 2> gpurun_out/r2_bench_c5_final.err; echo "c5 rc $?"
python bench.py --workload ref --no-cpu-baseline > gpurun_out/r2_bench_ref_final.json 2> gpurun_out/r2_bench_ref_final.err; echo "ref rc $?"
python bench.py --workload c3 --batch 64 --no-cpu-baseline --steps 6 > gpurun_out/r2_bench_c3_batch64_final.json 2> /dev/null; echo "b64 rc $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_c3_reference_arm.json 2> gpurun_out/r2_bench_c3_reference_arm.err; echo "reference arm rc $?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc $?"
python tools/igemm_detail.py c3 > gpurun_out/r2_igemm_detail_c3_final.txt 2>/dev/null
python tools/class_detail.py c3 > gpurun_out/r2_class_detail_c3_final.txt 2>/dev/null
NCU="ncu --clock-control none"
B="bench.py --workload c3 --resident-only --steps 1 --warmup 3 --no-cpu-baseline"
TOTAL=$(python -c "
import runpy, sys
sys.argv = '$B'.split()
runpy.run_path('bench.py', run_name='__main__')
from weatherconverter_b200 import ops
print('TOTAL_LAUNCHES', ops.launch_count())" 2>gpurun_out/r2_final_plain_bench.err | grep TOTAL_LAUNCHES | awk '{print $2}')
echo "total launches $TOTAL"
SKIP=$((TOTAL - 460))
python $B > gpurun_out/r2_final_plain_bench.log 2>&1 && $NCU --kernel-name-base demangled -k regex:wc:: -s $SKIP --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r2_launches_c3_final2_raw.csv python $B > gpurun_out/r2_final_ncu_bench.log 2>&1
echo "launch list rc $? ($(wc -l < gpurun_out/r2_launches_c3_final2_raw.csv) lines)"
python tools/bench_attn_prescaled.py 32,4,8192,32 > gpurun_out/r2_final_attn_plain.log 2>&1 && $NCU --set full --import-source on -k regex:attention_small4 -s 2 -c 1 -f -o /tmp/r2_attn_small4_hd32 python tools/bench_attn_prescaled.py 32,4,8192,32 > gpurun_out/r2_final_attn_ncu.log 2>&1
echo "attn capture rc $?"
if [ -f /tmp/r2_attn_small4_hd32.ncu-rep ]; then
  ncu -i /tmp/r2_attn_small4_hd32.ncu-rep --page raw --csv > gpurun_out/r2_ncu_attn_small4_hd32.raw.csv 2>/dev/null
  ncu -i /tmp/r2_attn_small4_hd32.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2_ncu_attn_small4_hd32.source.csv.gz
fi
du -sh gpurun_out
