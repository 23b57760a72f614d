#!/bin/bash
# How often does tests/test_gpu_parity.py::test_pipelined_forward_matches_direct_calls fail, and with which switches?
mkdir -p gpurun_out
out=gpurun_out/flaky_pipe.txt
: > $out
run() {  # label, env...
  label=$1; shift
  pass=0; fail=0
  for i in 1 2 3 4 5 6; do
    if env "$@" timeout 120 python -m pytest tests/test_gpu_parity.py -q -x -k "pipelined_forward" > /tmp/fp.log 2>&1; then pass=$((pass+1)); else fail=$((fail+1)); fi
  done
  echo "$label: pass $pass fail $fail" >> $out
}
run "default" HFC_X=0
run "thin off" HFC_THIN_EPILOGUE=0
run "launch blocking" CUDA_LAUNCH_BLOCKING=1
run "two-launch resnorm" HFC_FUSE_RESNORM=0
timeout 200 python tools/flaky_pipe_diag.py >> $out 2>&1
cat $out
