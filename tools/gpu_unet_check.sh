#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_unet.py -x -q -m gpu -s > gpurun_out/unet_test.log 2>&1; echo "unet tests rc=$?" | tee -a gpurun_out/unet_test.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -30 gpurun_out/unet_test.log; tail -5 gpurun_out/smoke.log
