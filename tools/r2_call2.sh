#!/bin/bash
# Round 2, GPU call 2: full suite on the fp16-gradient build, backward-precision tables (fp16 vs bf16), ncu --set full of the
# three epilogue-bound big-map layers (E1, G.up4, G3), full bench.
mkdir -p gpurun_out
S=gpurun_out/c2_status.txt
: > $S
python -c "from hific_b200 import _lib; print('abi', _lib.lib.hfc_abi_version())" >> $S 2>&1
timeout 700 python -m pytest tests -m gpu -q -rfEs > gpurun_out/c2_tests.log 2>&1; echo "tests rc=$?" >> $S
HFC_GRAD_FMT=fp16 timeout 240 python tools/grad_precision.py > gpurun_out/c2_grad_precision_fp16.txt 2>&1; echo "grad_precision fp16 rc=$?" >> $S
HFC_GRAD_FMT=bf16 timeout 240 python tools/grad_precision.py > gpurun_out/c2_grad_precision_bf16.txt 2>&1; echo "grad_precision bf16 rc=$?" >> $S
HFC_REPS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -c 24 -f -o gpurun_out/r2_bigmap \
    python tools/profile_thin_layers.py > gpurun_out/c2_ncu.log 2>&1; echo "ncu rc=$?" >> $S
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/c2_bench_full.json 2> gpurun_out/c2_bench_full.err; echo "bench full rc=$?" >> $S
HFC_GRAD_FMT=bf16 timeout 200 python bench.py --steps 20 --warmup 5 --no-gan --no-cpu-baseline --no-compress > gpurun_out/c2_bench_bf16.json 2> gpurun_out/c2_bench_bf16.err; echo "bench bf16 rc=$?" >> $S
cat $S
tail -4 gpurun_out/c2_tests.log
tail -12 gpurun_out/c2_grad_precision_fp16.txt
tail -12 gpurun_out/c2_grad_precision_bf16.txt
