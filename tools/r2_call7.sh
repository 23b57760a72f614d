#!/bin/bash
# Round 2, GPU call 7: state of the tree after the register-border epilogues + thin level 2: full suite, layer times, bench,
# profiles for the record (fused residual conv ncu --set full, training-step kernel breakdown).
mkdir -p gpurun_out
S=gpurun_out/c7_status.txt
: > $S
timeout 60 python tools/profile_thin_layers.py > gpurun_out/c7_thin_layers.txt 2>&1; echo "thin layers rc=$?" >> $S
timeout 120 python tools/layer_times.py > gpurun_out/c7_layer_times.txt 2> gpurun_out/c7_layer_times.err; echo "layer_times rc=$?" >> $S
timeout 700 python bench.py --steps 20 --warmup 5 > gpurun_out/c7_bench_full.json 2> gpurun_out/c7_bench_full.err; echo "bench full rc=$?" >> $S
timeout 1100 python -m pytest tests -m gpu -q -rfEs > gpurun_out/c7_tests.log 2>&1; echo "tests rc=$?" >> $S
timeout 200 python tools/train_profile.py > gpurun_out/c7_train_profile.txt 2>&1; echo "train profile rc=$?" >> $S
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 12 -c 6 -f -o gpurun_out/r02_resconv_fused \
    python bench.py --profile --steps 1 --warmup 1 > gpurun_out/c7_ncu_res.log 2>&1; echo "ncu resconv rc=$?" >> $S
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c7_smoke.txt 2>&1; echo "smoke rc=$?" >> $S
cat $S
cat gpurun_out/c7_thin_layers.txt
tail -4 gpurun_out/c7_tests.log
tail -3 gpurun_out/c7_smoke.txt
