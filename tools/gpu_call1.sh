#!/bin/bash
# round-1 GPU call 1: new tests (compress path, likelihood schedule 2), A/B timing, ncu capture of the new kernels
mkdir -p gpurun_out
rm -f gpurun_out/status.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 200 python -m pytest tests/test_gpu_zcompress.py tests/test_gpu_zlikelihood.py -q > gpurun_out/new_tests.log 2>&1
echo "new_tests rc=$?" >> gpurun_out/status.txt
timeout 90 python tools/likelihood_ab.py > gpurun_out/likelihood_ab.json 2> gpurun_out/likelihood_ab.err
echo "ab rc=$?" >> gpurun_out/status.txt
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"symbols|likelihood" -c 16 -f -o gpurun_out/r01_symbols_likelihood python tools/profile_symbols.py > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/status.txt
tail -5 gpurun_out/new_tests.log
cat gpurun_out/likelihood_ab.json
cat gpurun_out/status.txt
