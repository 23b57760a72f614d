#!/bin/bash
# Round 2, GPU call 4: lean TMA producer (warp-uniform, per-K-block table) + resident weights for the thin layers +
# bulk-staged likelihood schedule: parity (full suite), layer timings, bench.
mkdir -p gpurun_out
S=gpurun_out/c4_status.txt
: > $S
timeout 900 python -m pytest tests -m gpu -q -rfEs -x > gpurun_out/c4_tests.log 2>&1; echo "tests rc=$?" >> $S
timeout 90 python tools/profile_thin_layers.py > gpurun_out/c4_thin_layers_1.txt 2>&1; echo "thin layers rc=$?" >> $S
HFC_NO_BRES=1 timeout 90 python tools/profile_thin_layers.py > gpurun_out/c4_thin_layers_nobres.txt 2>&1
HFC_THIN_EPILOGUE=2 timeout 90 python tools/profile_thin_layers.py > gpurun_out/c4_thin_layers_2.txt 2>&1
HFC_THIN_EPILOGUE=0 timeout 90 python tools/profile_thin_layers.py > gpurun_out/c4_thin_layers_0.txt 2>&1
timeout 120 python tools/layer_times.py > gpurun_out/c4_layer_times.json 2> gpurun_out/c4_layer_times.err; echo "layer_times rc=$?" >> $S
timeout 700 python bench.py --steps 20 --warmup 5 > gpurun_out/c4_bench_full.json 2> gpurun_out/c4_bench_full.err; echo "bench full rc=$?" >> $S
HFC_THIN_EPILOGUE=2 timeout 150 python bench.py --steps 20 --warmup 5 --no-train --no-cpu-baseline --no-compress --no-eager > gpurun_out/c4_bench_fwd_thin2.json 2> gpurun_out/c4_bench_fwd_thin2.err; echo "bench thin2 rc=$?" >> $S
HFC_REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -c 12 -f -o gpurun_out/r2_bigmap_v3 \
    python tools/profile_thin_layers.py > gpurun_out/c4_ncu.log 2>&1; echo "ncu rc=$?" >> $S
cat $S
tail -5 gpurun_out/c4_tests.log
cat gpurun_out/c4_thin_layers_*.txt
