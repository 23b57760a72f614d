"""2-GPU diagnostic of the in-backward gradient reducer over NCCL (run under torchrun): one training step's gradients
through (a) the plain after-backward all-reduce, (b) the in-backward reducer, (c) the reducer with a host synchronisation
after every bucket (HFC_REDUCER_SYNC=1 semantics) -- per-tensor relative differences, worst first."""
import logging
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HFC_LPIPS_SYNTHETIC", "1")
import torch
import torch.distributed as dist

from hific_b200 import synth
from hific_b200.config import ModelModes, ModelTypes, mse_lpips_args
from hific_b200.dist import InBackwardGradientReducer, allreduce_gradients
from hific_b200.model import Model


def main():
    rank, local, world = (int(os.environ[k]) for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B = int(os.environ.get("HFC_B", 8))
    cfg = mse_lpips_args()
    cfg.batch_size = B
    model = Model(cfg, logging.getLogger("chk"), model_mode=ModelModes.TRAINING, model_type=ModelTypes.COMPRESSION)
    model.load_state_dict(synth.synth_state_dict(0), strict=False)
    model.to(dev).train()
    x = synth.synth_image(B, 256, 256, seed=1 + rank).to(dev)
    named = [(k, p) for k, p in model.named_parameters()]
    params = [p for _, p in named]
    density = list(model.Hyperprior.hyperlatent_likelihood.parameters())
    reducer = InBackwardGradientReducer(dist, world)

    def run(mode):
        for p in params:
            p.grad = None
        torch.manual_seed(1234)
        loss = model(x, train_generator=True)["compression"]
        if mode == "plain":
            loss.backward()
            allreduce_gradients(params, dist, world)
        else:
            reducer.debug_sync = mode == "sync"
            with reducer:
                loss.backward()
            reducer.reduce_rest(density)
        torch.cuda.synchronize()
        return {k: p.grad.detach().clone() for k, p in named if p.grad is not None}

    run("overlap")                           # calibrates the loss scales
    res = {m: run(m) for m in ("plain", "overlap", "sync", "plain")}
    again = run("plain")
    for name, a, b in (("plain vs plain (run-to-run)", res["plain"], again), ("overlap vs plain", res["overlap"], res["plain"]),
                       ("sync vs plain", res["sync"], res["plain"])):
        diffs = sorted(((float((a[k] - b[k]).norm() / b[k].norm().clamp_min(1e-30)), k) for k in b), reverse=True)
        if rank == 0:
            print(f"== {name}: worst {diffs[0][0]:.3e} {diffs[0][1]}; tensors above 1e-3: {sum(d > 1e-3 for d, _ in diffs)} of {len(diffs)}")
            for d, k in diffs[:8]:
                print(f"   {d:.3e} {k}")
    if rank == 0:
        print("buckets launched in the last overlapped backward:", reducer.buckets_launched, "bytes", reducer.bytes_reduced)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
