#!/bin/bash
# Round 2, GPU call 6: E1 bulk-copy window segments; thin level 2 through the parity tests; profiles for the record.
mkdir -p gpurun_out
S=gpurun_out/c6_status.txt
: > $S
T="tests/test_gpu_ops.py tests/test_gpu_parity.py tests/test_gpu_conv_modes.py tests/test_gpu_grad.py tests/test_gpu_gan.py"
timeout 900 python -m pytest $T -m gpu -q -rfEs > gpurun_out/c6_tests.log 2>&1; echo "tests rc=$?" >> $S
HFC_THIN_EPILOGUE=2 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity.py tests/test_gpu_conv_modes.py -m gpu -q -rfEs > gpurun_out/c6_tests_thin2.log 2>&1; echo "tests thin2 rc=$?" >> $S
timeout 90 python tools/profile_thin_layers.py > gpurun_out/c6_thin_layers_1.txt 2>&1
HFC_THIN_EPILOGUE=2 timeout 90 python tools/profile_thin_layers.py > gpurun_out/c6_thin_layers_2.txt 2>&1
B="--steps 20 --warmup 5 --no-train --no-cpu-baseline --no-compress --no-eager"
timeout 150 python bench.py $B > gpurun_out/c6_bench_fwd.json 2> gpurun_out/c6_bench_fwd.err; echo "bench fwd rc=$?" >> $S
HFC_THIN_EPILOGUE=2 timeout 150 python bench.py $B > gpurun_out/c6_bench_fwd_thin2.json 2> gpurun_out/c6_bench_fwd_thin2.err; echo "bench fwd thin2 rc=$?" >> $S
# launch list of the forward (cold-cache, serialised: shares, not absolutes)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_fwd.csv \
    python bench.py --profile --steps 2 --warmup 1 > gpurun_out/c6_launches.log 2>&1; echo "launch list rc=$?" >> $S
# ncu --set full: the fused residual conv (widenorm) inside the forward, and the thin layers
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel.*1.*2.*1 -s 4 -c 2 -f -o gpurun_out/r02_resconv_fused \
    python bench.py --profile --steps 1 --warmup 1 > gpurun_out/c6_ncu_res.log 2>&1; echo "ncu resconv rc=$?" >> $S
HFC_REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -c 12 -f -o gpurun_out/r02_bigmap_v4 \
    python tools/profile_thin_layers.py > gpurun_out/c6_ncu_thin.log 2>&1; echo "ncu thin rc=$?" >> $S
cat $S
tail -3 gpurun_out/c6_tests.log; tail -3 gpurun_out/c6_tests_thin2.log
cat gpurun_out/c6_thin_layers_1.txt gpurun_out/c6_thin_layers_2.txt
