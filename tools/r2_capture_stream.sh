set -u
NCU="ncu --clock-control none"
python tools/bench_conv.py 32,64,256,64,128,1,1 > gpurun_out/ev_plain_stream.log 2>&1 && $NCU --set full --import-source on -k regex:igemm -s 2 -c 1 -f -o /tmp/r2_stream python tools/bench_conv.py 32,64,256,64,128,1,1 > gpurun_out/ev_ncu_stream.log 2>&1
echo rc $?
ncu -i /tmp/r2_stream.ncu-rep --page raw --csv > gpurun_out/r2_ncu_igemm_1x1_64_256_res_stream.raw.csv 2>/dev/null
ncu -i /tmp/r2_stream.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2_ncu_igemm_1x1_64_256_res_stream.source.csv.gz
cat gpurun_out/ev_plain_stream.log
