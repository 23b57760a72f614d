#!/bin/bash
# Round 2, GPU call 3: thin epilogue v2 (16 epilogue warps / 4 accumulator stages) -- parity + timing + ncu; teacher-forced
# backward-precision tables; the bench with its new legs.
mkdir -p gpurun_out
S=gpurun_out/c3_status.txt
: > $S
timeout 700 python -m pytest tests -m gpu -q -rfEs > gpurun_out/c3_tests.log 2>&1; echo "tests rc=$?" >> $S
for t in 0 1 2; do
  HFC_THIN_EPILOGUE=$t timeout 90 python tools/profile_thin_layers.py > gpurun_out/c3_thin_layers_$t.txt 2>&1; echo "thin layers $t rc=$?" >> $S
done
B="--steps 20 --warmup 5 --no-train --no-cpu-baseline --no-compress --no-eager"
for t in 0 1 2; do
  HFC_THIN_EPILOGUE=$t timeout 150 python bench.py $B > gpurun_out/c3_bench_fwd_thin$t.json 2> gpurun_out/c3_bench_fwd_thin$t.err; echo "bench fwd thin=$t rc=$?" >> $S
done
HFC_GRAD_FMT=fp16 timeout 300 python tools/grad_precision.py > gpurun_out/c3_grad_precision_fp16.txt 2>&1; echo "grad_precision fp16 rc=$?" >> $S
HFC_GRAD_FMT=bf16 timeout 300 python tools/grad_precision.py > gpurun_out/c3_grad_precision_bf16.txt 2>&1; echo "grad_precision bf16 rc=$?" >> $S
HFC_REPS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -c 30 -f -o gpurun_out/r2_bigmap_thin \
    python tools/profile_thin_layers.py > gpurun_out/c3_ncu.log 2>&1; echo "ncu rc=$?" >> $S
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c3_bench_full.json 2> gpurun_out/c3_bench_full.err; echo "bench full rc=$?" >> $S
cat $S
tail -4 gpurun_out/c3_tests.log
cat gpurun_out/c3_thin_layers_*.txt
tail -14 gpurun_out/c3_grad_precision_fp16.txt
tail -14 gpurun_out/c3_grad_precision_bf16.txt
