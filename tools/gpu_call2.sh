#!/bin/bash
# round-1 GPU call 2: re-verify the compress-path kernels (binary-search index) and likelihood schedules 2/3, A/B timing,
# bench line, ncu capture of the final kernels, then the likelihood-touching part of the old suite
mkdir -p gpurun_out
rm -f gpurun_out/status.txt
timeout 150 python -m pytest tests/test_gpu_zcompress.py tests/test_gpu_zlikelihood.py -q > gpurun_out/new_tests.log 2>&1
echo "new_tests rc=$?" >> gpurun_out/status.txt
timeout 90 python tools/likelihood_ab.py > gpurun_out/likelihood_ab.json 2> gpurun_out/likelihood_ab.err
echo "ab rc=$?" >> gpurun_out/status.txt
timeout 240 python bench.py --steps 10 --warmup 3 --no-gan > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/status.txt
timeout 100 ncu --set full --clock-control none --import-source on -k regex:"symbols|likelihood" -c 18 -f -o gpurun_out/r01_symbols_likelihood_v2 python tools/profile_symbols.py > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/status.txt
timeout 150 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity.py -x -q > gpurun_out/old_tests.log 2>&1
echo "old_tests rc=$?" >> gpurun_out/status.txt
tail -3 gpurun_out/new_tests.log
cat gpurun_out/likelihood_ab.json
tail -c 3000 gpurun_out/bench.json
tail -3 gpurun_out/old_tests.log
cat gpurun_out/status.txt
