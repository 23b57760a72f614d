"""Data-parallel training entry point: one process per GPU, gradients all-reduced over NCCL / NVLink.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
        -m hific_b200.train_ddp --model_type compression_gan --regime low --batch_size 32 --n_steps 1000 --save out/

The reference has no multi-GPU launcher (`train.py:303-308` raises NotImplementedError for more than one GPU), so this
is the `train_ddp.py`-style entry SURVEY.md 8e(i) asks for.  `train()` below is the reference's loop (`train.py:89-200`:
alternating generator / discriminator iterations, `optimize_compression_loss` / `optimize_loss`, the `update_lr` hook,
`utils.save_model`'s checkpoint layout) with what data parallelism adds:

  * identical initial weights on every rank (same seed before `Model(...)`), rank-offset seeds for data and noise;
  * `args.gpu = LOCAL_RANK` before the model is built, so `PerceptualLoss(gpu_ids=[args.gpu])` and every plan live on the
    rank's own device;
  * generator iterations: the Encoder / Hyperprior / Generator gradients are all-reduced (mean) in one coalesced in-place
    NCCL call after the backward; `--overlap` hands them over layer by layer FROM INSIDE the backward instead
    (`dist.InBackwardGradientReducer`, buckets of 32 MB on a side stream) -- equal results, measured slower on B200s
    because NCCL's kernels take SMs from the backward; discriminator iterations: one coalesced
    all-reduce of the discriminator's gradients;
  * the reference's stale-gradient quirk is kept: on generator iterations autograd also fills the discriminator's
    `.grad` (its parameters require grad), nobody zeroes it (`train.py:54-59` only zeroes the stepped optimizers), and the
    next discriminator iteration ACCUMULATES onto it.  The sum is all-reduced on that discriminator iteration, so N ranks
    on shards of a batch take the step one process takes on the whole batch (tests/test_dist_gloo.py);
  * `weighted_rate_loss` picks lambda from the rank's own `q_bpp.item()` (losses.py:21-25), i.e. per local batch, as the
    single-process code does per batch;
  * checkpoints (`utils.save_model`'s dict: `model_state_dict`, the three optimizer state_dicts, `epoch`, `steps`,
    `args`) are written by rank 0 only and load with the reference's `utils.load_model`.

Data: `--image_dir` (any folder of .png / .jpg files, random `crop_size` crops) or, by default, synthetic uniform-noise
images -- the reference's OpenImages loader (`src/helpers/datasets.py`) is a caller of the hot path, not part of it.
"""
import argparse
import datetime
import glob
import itertools
import logging
import os
import time

import torch

from . import dist as hdist
from .config import ModelModes, ModelTypes, hific_args, mse_lpips_args
from .model import Model


def optimize_compression_loss(model, compression_loss, amortization_opt, hyperlatent_likelihood_opt, reducer, density):
    """train.py:54-59 + gradient reduction."""
    if reducer is not None:
        with reducer:
            compression_loss.backward()
        reducer.reduce_rest(density)
    else:
        compression_loss.backward()
    if reducer is not None and hasattr(reducer, "then_step"):
        reducer.then_step(amortization_opt)         # all-reduce in buckets, Adam of bucket i right behind bucket i
    else:
        amortization_opt.step()
    hyperlatent_likelihood_opt.step()
    amortization_opt.zero_grad()
    hyperlatent_likelihood_opt.zero_grad()


def optimize_loss(loss, opt, params, dist, world):
    """train.py:49-52 + gradient reduction."""
    loss.backward()
    if world > 1:
        hdist.allreduce_gradients(params, dist, world)
    opt.step()
    opt.zero_grad()


def update_lr(args, optimizer, itr, logger):
    """src/helpers/utils.py:74-85."""
    vals, steps = args.lr_schedule['vals'], args.lr_schedule['steps']
    idx = sum(1 for s in steps if itr >= s)
    lr = args.learning_rate * vals[idx]
    for group in optimizer.param_groups:
        if group['lr'] != lr:
            logger.info('=============================')
            logger.info(f'Changing learning rate {group["lr"]} -> {lr}')
            group['lr'] = lr


def make_optimizers(model, args, adam=None):
    """train.py:287-300: Adam(lr) over the amortization models, the hyper-latent density and the discriminator."""
    if adam is None:
        from .optim import Adam as adam
    amort = itertools.chain.from_iterable([am.parameters() for am in model.amortization_models])
    opts = dict(amort=adam(amort, lr=args.learning_rate),
                hyper=adam(model.Hyperprior.hyperlatent_likelihood.parameters(), lr=args.learning_rate))
    if model.use_discriminator is True:
        opts['disc'] = adam(model.Discriminator.parameters(), lr=args.learning_rate)
    return opts


def save_model(model, optimizers, epoch, args, logger, rank=0):
    """utils.save_model's checkpoint dictionary (src/helpers/utils.py:125-167), written by rank 0 only."""
    if rank != 0:
        return None
    directory = args.checkpoints_save
    os.makedirs(directory, exist_ok=True)
    args_d = dict((n, getattr(args, n)) for n in dir(args) if not (n.startswith('_') or 'logger' in n))
    args_d['timestamp'] = '{:%Y_%m_%d_%H:%M}'.format(datetime.datetime.now())
    path = os.path.join(directory, '{}_epoch{}_idx{}_{:%Y_%m_%d_%H:%M:%S}.pt'.format(args.name, epoch, model.step_counter,
                                                                                     datetime.datetime.now()))
    save_dict = {'model_state_dict': {k: v.detach().cpu() for k, v in model.state_dict().items()},
                 'compression_optimizer_state_dict': optimizers['amort'].state_dict(),
                 'hyperprior_optimizer_state_dict': optimizers['hyper'].state_dict(),
                 'epoch': epoch, 'steps': model.step_counter, 'args': args_d}
    if model.use_discriminator is True:
        save_dict['discriminator_state_dict'] = {k: v.detach().cpu() for k, v in model.Discriminator.state_dict().items()}
        save_dict['discriminator_optimizer_state_dict'] = optimizers['disc'].state_dict()
    torch.save(save_dict, f=path)
    logger.info('Saved model at Epoch {}, step {} to {}'.format(epoch, model.step_counter, path))
    return path


def train(args, model, batches, device, logger, optimizers, dist=None, rank=0, world=1, overlap=False):
    """The reference's loop body (train.py:114-141, 175-193) over an iterable of (B, 3, H, W) batches in [0, 1]."""
    amortization_opt, hyperlatent_likelihood_opt = optimizers['amort'], optimizers['hyper']
    disc_opt = optimizers.get('disc')
    density = list(model.Hyperprior.hyperlatent_likelihood.parameters())
    disc_params = list(model.Discriminator.parameters()) if model.use_discriminator is True else []
    reducer = hdist.InBackwardGradientReducer(dist, world, group=hdist.reducer_group(dist)) if (world > 1 and overlap) else None
    if world > 1 and reducer is None:
        class _Plain:                                  # same interface, the all-reduce after backward
            def __enter__(self):
                return self

            def __exit__(self, *e):
                return False

            def then_step(self, opt):
                # the amortization group: coalesced in-place all-reduce in 64 MB buckets on a side stream with the Adam launch
                # of each bucket right behind it (plain all-reduce + opt.step() for optimizers that cannot step by bucket)
                hdist.allreduce_then_step(opt, dist, world)

            def reduce_rest(self, params):
                return hdist.allreduce_gradients(list(params), dist, world)
        reducer = _Plain()
    current_D_steps, train_generator = 0, True
    start, ckpt_path, last_loss = time.time(), None, None
    model.train()
    for idx, data in enumerate(batches):
        data = data.to(device, dtype=torch.float)
        if model.use_discriminator is True:
            losses = model(data, train_generator=train_generator)
            compression_loss, disc_loss = losses['compression'], losses['disc']
            if train_generator is True:
                optimize_compression_loss(model, compression_loss, amortization_opt, hyperlatent_likelihood_opt, reducer,
                                          density)
                train_generator = False
            else:
                optimize_loss(disc_loss, disc_opt, disc_params, dist, world)
                current_D_steps += 1
                if current_D_steps == args.discriminator_steps:
                    current_D_steps = 0
                    train_generator = True
                continue
        else:
            losses = model(data, train_generator=True)
            compression_loss = losses['compression']
            optimize_compression_loss(model, compression_loss, amortization_opt, hyperlatent_likelihood_opt, reducer, density)
        last_loss = compression_loss
        if model.step_counter % args.log_interval == 1:
            logger.info('[rank %d] step %d | compression loss %.4f | %.1f s', rank, model.step_counter,
                        compression_loss.item(), time.time() - start)
            update_lr(args, amortization_opt, model.step_counter, logger)
            update_lr(args, hyperlatent_likelihood_opt, model.step_counter, logger)
            if model.use_discriminator is True:
                update_lr(args, disc_opt, model.step_counter, logger)
        if (idx % args.save_interval == 1) and (idx > args.save_interval):
            ckpt_path = save_model(model, optimizers, 0, args, logger, rank)
        if model.step_counter > args.n_steps:
            logger.info('Reached step limit [args.n_steps = {}]'.format(args.n_steps))
            break
    ckpt_path = save_model(model, optimizers, 0, args, logger, rank) or ckpt_path
    if world > 1:
        dist.barrier()                                  # nobody leaves (or reads the checkpoint) before rank 0 has written it
    return model, ckpt_path, last_loss


def synthetic_batches(n, batch, size, seed):
    g = torch.Generator().manual_seed(seed)
    for _ in range(n):
        yield torch.rand((batch, 3, size, size), generator=g)


def folder_batches(image_dir, n, batch, size, seed):
    """Random `size` x `size` crops of the images in a folder (PIL), in [0, 1]."""
    import numpy as np
    from PIL import Image
    files = sorted(glob.glob(os.path.join(image_dir, '*.png')) + glob.glob(os.path.join(image_dir, '*.jpg')))
    if not files:
        raise FileNotFoundError(f'no .png / .jpg files in {image_dir}')
    rng = np.random.default_rng(seed)
    for _ in range(n):
        out = torch.empty((batch, 3, size, size))
        for b in range(batch):
            img = np.asarray(Image.open(files[rng.integers(len(files))]).convert('RGB'), dtype=np.float32) / 255.
            if img.shape[0] < size or img.shape[1] < size:
                img = np.pad(img, ((0, max(0, size - img.shape[0])), (0, max(0, size - img.shape[1])), (0, 0)), mode='reflect')
            y, x = rng.integers(img.shape[0] - size + 1), rng.integers(img.shape[1] - size + 1)
            out[b] = torch.from_numpy(img[y:y + size, x:x + size].copy()).permute(2, 0, 1)
        yield out


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    ap.add_argument('--model_type', default=ModelTypes.COMPRESSION, choices=[ModelTypes.COMPRESSION, ModelTypes.COMPRESSION_GAN])
    ap.add_argument('--regime', default='low', choices=['low', 'med', 'high'])
    ap.add_argument('--batch_size', type=int, default=8, help='per GPU')
    ap.add_argument('--crop_size', type=int, default=256)
    ap.add_argument('--n_steps', type=int, default=100)
    ap.add_argument('--learning_rate', type=float, default=1e-4)
    ap.add_argument('--log_interval', type=int, default=100)
    ap.add_argument('--save_interval', type=int, default=50000)
    ap.add_argument('--image_dir', default=None)
    ap.add_argument('--save', default='experiments/ddp')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--warmstart_ckpt', default=None, help='checkpoint (reference or hific_b200 format) to start from')
    ap.add_argument('--overlap', action='store_true',
                    help='all-reduce 32 MB gradient buckets from inside the backward (dist.InBackwardGradientReducer) instead of one '
                         'coalesced all-reduce after it; measured slower on B200s over NCCL, whose kernels take SMs from the '
                         'backward (DESIGN.md section 4)')
    a = ap.parse_args(argv)
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('LOCAL_RANK', 0), ('WORLD_SIZE', 1)))
    if not torch.cuda.is_available():
        raise SystemExit('hific_b200.train_ddp: no CUDA device -- the hific_b200 path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=device)
    logging.basicConfig(level=logging.INFO if rank == 0 else logging.WARNING, format='%(asctime)s %(levelname)s %(message)s')
    logger = logging.getLogger('train_ddp')
    args = hific_args() if a.model_type == ModelTypes.COMPRESSION_GAN else mse_lpips_args()
    args.regime, args.target_rate, args.lambda_A = a.regime, args.target_rate_map[a.regime], args.lambda_A_map[a.regime]
    args.batch_size, args.crop_size, args.image_dims = a.batch_size, a.crop_size, (3, a.crop_size, a.crop_size)
    args.latent_dims = (args.latent_channels, a.crop_size // 16, a.crop_size // 16)
    args.n_steps, args.learning_rate, args.log_interval, args.save_interval = a.n_steps, a.learning_rate, a.log_interval, a.save_interval
    args.gpu, args.multigpu = local_rank, False          # per-rank LPIPS / plan device; checkpoints hold the bare module's keys
    args.name = f'hific_b200_{a.model_type}_{a.regime}'
    args.checkpoints_save = os.path.join(a.save, 'checkpoints')
    torch.manual_seed(a.seed)                            # identical initial weights on every rank
    model = Model(args, logger, model_mode=ModelModes.TRAINING, model_type=a.model_type)
    if a.warmstart_ckpt:
        ck = torch.load(a.warmstart_ckpt, map_location='cpu')
        model.load_state_dict(ck['model_state_dict'], strict=False)
    model.to(device)
    optimizers = make_optimizers(model, args)
    torch.manual_seed(a.seed + 1000 * (rank + 1))        # quantisation noise differs per rank
    n_batches = a.n_steps * (1 + (args.discriminator_steps if model.use_discriminator else 0)) + 2
    seed = a.seed + 17 * (rank + 1)
    batches = folder_batches(a.image_dir, n_batches, a.batch_size, a.crop_size, seed) if a.image_dir else \
        synthetic_batches(n_batches, a.batch_size, a.crop_size, seed)
    t0 = time.time()
    model, ckpt, last = train(args, model, batches, device, logger, optimizers, dist, rank, world, overlap=a.overlap)
    torch.cuda.synchronize()
    if rank == 0:
        logger.info('Training complete. Time elapsed: %.3f s. Number of steps: %d. Global batch %d. Checkpoint: %s',
                    time.time() - t0, model.step_counter, world * a.batch_size, ckpt)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
