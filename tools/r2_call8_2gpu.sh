#!/bin/bash
# Round 2, second 2-GPU call: reducer diagnostic + the train_ddp launcher (with the synthetic-LPIPS opt-in it rightly demands).
mkdir -p gpurun_out
S=gpurun_out/c8_status.txt
: > $S
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  tools/ddp_selfcheck.py > gpurun_out/c8_selfcheck.txt 2> gpurun_out/c8_selfcheck.err; echo "selfcheck rc=$?" >> $S
HFC_LPIPS_SYNTHETIC=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  -m hific_b200.train_ddp --model_type compression_gan --regime low --batch_size 8 --n_steps 6 --log_interval 2 \
  --save gpurun_out/c8_ddp > gpurun_out/c8_train_ddp.log 2>&1; echo "train_ddp rc=$?" >> $S
ls -la gpurun_out/c8_ddp/checkpoints >> $S 2>&1
rm -rf gpurun_out/c8_ddp
cat $S
grep -v "^$" gpurun_out/c8_selfcheck.txt | tail -40
grep -v "Warning\|_lpips\|^$" gpurun_out/c8_train_ddp.log | tail -8
