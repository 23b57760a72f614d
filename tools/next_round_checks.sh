#!/bin/bash
# First GPU call of the next round: decide the switches round 1 left opt-in.
#   gpurun --timeout 600 -- 'bash tools/next_round_checks.sh'            (1 GPU)
#   gpurun --gpus 2 --timeout 300 -- 'bash tools/next_round_checks.sh 2' (overlapped all-reduce, 2 GPUs)
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  # 1. full GPU suite (round 1 ended without a complete run after the last edits)
  HFC_RUN_UNVERIFIED=1 timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/nr_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/nr_status.txt   # incl. tests/test_gpu_zzdlmm.py (DLMM kernels: first run on hardware)
  # 2. bench: train_step.with_native_lpips_trunk vs train_step.ms_per_step decides HFC_LPIPS_TRUNK's default
  timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/nr_bench.json 2> gpurun_out/nr_bench.err; echo "bench rc=$?" >> gpurun_out/nr_status.txt
  # 2b. forward with the residual convs fused with their ChannelNorm (hfc_conv_forward_widenorm; parity: tests/test_gpu_zzwidenorm.py
  #     in step 1): compare "value" / "ms_per_step" with nr_bench.json -- expected ~0.5 ms less per forward if it works
  HFC_FUSE_RESNORM=1 timeout 120 python bench.py --steps 20 --warmup 5 --no-train --no-cpu-baseline --no-compress \
      > gpurun_out/nr_bench_fused_resnorm.json 2> gpurun_out/nr_bench_fused_resnorm.err; echo "bench fused rc=$?" >> gpurun_out/nr_status.txt
  # 2c. thin epilogue for the 60-channel big-map layers (own kernel instantiation; the switch is read once per process):
  #     parity through the existing op / model tests, then the forward time
  HFC_THIN_EPILOGUE=1 timeout 200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity.py tests/test_gpu_conv_modes.py -x -q \
      > gpurun_out/nr_tests_thin.log 2>&1; echo "tests thin rc=$?" >> gpurun_out/nr_status.txt
  HFC_THIN_EPILOGUE=1 timeout 120 python bench.py --steps 20 --warmup 5 --no-train --no-cpu-baseline --no-compress \
      > gpurun_out/nr_bench_thin.json 2> gpurun_out/nr_bench_thin.err; echo "bench thin rc=$?" >> gpurun_out/nr_status.txt
  HFC_THIN_EPILOGUE=1 HFC_FUSE_RESNORM=1 timeout 120 python bench.py --steps 20 --warmup 5 --no-train --no-cpu-baseline --no-compress \
      > gpurun_out/nr_bench_thin_fused.json 2> gpurun_out/nr_bench_thin_fused.err; echo "bench thin+fused rc=$?" >> gpurun_out/nr_status.txt
  timeout 60 python tools/profile_thin_layers.py > gpurun_out/nr_thin_layers_off.txt 2>&1
  HFC_THIN_EPILOGUE=1 timeout 60 python tools/profile_thin_layers.py > gpurun_out/nr_thin_layers_on.txt 2>&1
  # 3. re-profile the compress-path kernels (64-bit divisions removed after the last capture)
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:"symbols" -c 12 -f -o gpurun_out/nr_symbols \
      python tools/profile_symbols.py > gpurun_out/nr_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/nr_status.txt
else
  # 4. overlapped gradient all-reduce over NCCL: self-check + training-step time, against the plain path
  for mode in 0 1; do
    HFC_OVERLAP_ALLREDUCE=$mode timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 12 --warmup 3 --no-gan --no-cpu-baseline \
      --no-compress > gpurun_out/nr_bench_${N}gpu_overlap$mode.json 2> gpurun_out/nr_bench_${N}gpu_overlap$mode.err
    echo "bench N=$N overlap=$mode rc=$?" >> gpurun_out/nr_status.txt
  done
fi
cat gpurun_out/nr_status.txt
