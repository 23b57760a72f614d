"""Generate tests/golden/*.npz by running the REAL reference modules (imported from /root/reference) on
synthetic weights/inputs, and check the oracle restatement against them -- TEST INFRASTRUCTURE.

    python oracle/make_golden.py            # regenerate fixtures (build container only)

Fixtures hold outputs only (strided subsets + full-tensor moments for the big ones); inputs and weights are
regenerated from seeds by hific_b200.synth, which the script also validates against the reference's own
state_dict keys and shapes.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> (batch, H, W, training, evaluation_mode)
CASES = {
    "train_128": (2, 128, 128, True, False),
    "eval_128": (2, 128, 128, False, False),
    "train_256": (1, 256, 256, True, False),
    "evalmode_100x144": (1, 100, 144, False, True),
}
SEED = 0


def summarize(name, t, out, full_limit=40000):
    """Store t in full when small, else a strided subset; always its first two moments."""
    a = t.detach().cpu().numpy().astype(np.float32)
    out[name + ".shape"] = np.array(a.shape, dtype=np.int64)
    out[name + ".sum"] = np.array(a.astype(np.float64).sum())
    out[name + ".sqsum"] = np.array((a.astype(np.float64) ** 2).sum())
    if a.size <= full_limit:
        out[name + ".full"] = a
    else:
        step = int(np.ceil(a.size / full_limit))
        out[name + ".stride"] = np.array(step, dtype=np.int64)
        out[name + ".sub"] = a.reshape(-1)[::step].copy()


def build_reference_model(gan=False):
    ref_shim.install()
    import logging
    from default_config import ModelModes, ModelTypes, hific_args, mse_lpips_args
    from src.model import Model

    class A(hific_args if gan else mse_lpips_args):
        pass

    args = A()
    args.image_dims = (3, 256, 256)
    args.latent_dims = (args.latent_channels, 16, 16)
    args.batch_size = 2
    logger = logging.getLogger("golden")
    model = Model(args, logger, model_mode=ModelModes.TRAINING,
                  model_type=ModelTypes.COMPRESSION_GAN if gan else ModelTypes.COMPRESSION)
    return model, ModelModes


def gan_case(synth, O):
    """COMPRESSION_GAN training forward of the real reference (discriminator, LPIPS, all loss terms) on a 2x128x128
    batch; also exports the vendored LPIPS lin weights (data, not code) for the product's PerceptualLoss."""
    import torchvision
    model, ModelModes = build_reference_model(gan=True)
    sd = synth.synth_state_dict(SEED, gan=True)
    model.load_state_dict(sd, strict=True)
    model.train(True)
    model.args.latent_dims = (220, 8, 8)
    b, h, w = 2, 128, 128
    x = synth.synth_image(b, h, w, SEED)
    noise_z = synth.synth_noise((b, 320, 2, 2), "zgan", SEED)
    noise_y = synth.synth_noise((b, 220, 8, 8), "ygan", SEED)
    u_before = {i: sd[f"Discriminator.conv{i}.weight_u"].clone() for i in range(1, 5)}
    with torch.no_grad(), ref_shim.NoiseFeeder([noise_z, noise_y]):
        losses, inter = model(x, train_generator=True, return_intermediates=True)
        disc = model.discriminator_forward(inter, train_generator=True)   # second power iteration (from updated u)
    pnet = model.perceptual_loss.model.net
    pnet = pnet.module if hasattr(pnet, "module") else pnet
    lins = [getattr(pnet, f"lin{k}").model[1].weight.detach().reshape(-1).clone() for k in range(5)]
    os.makedirs(os.path.join(ROOT, "high-fidelity-generative-compression_b200", "weights"), exist_ok=True)
    np.savez(os.path.join(ROOT, "high-fidelity-generative-compression_b200", "weights", "lpips_alex_lin_v0.1.npz"),
             **{f"lin{k}": l.numpy() for k, l in enumerate(lins)})
    out = {"compression_loss": np.array(float(losses["compression"])), "disc_loss": np.array(float(losses["disc"]))}
    with torch.no_grad():
        out["distortion"] = np.array(float(model.distortion_loss(inter.reconstruction, inter.input_image)))
        out["perceptual"] = np.array(float(model.perceptual_loss_wrapper(inter.reconstruction, inter.input_image)))
    summarize("recon", inter.reconstruction, out)
    out["n_bpp"], out["q_bpp"] = np.array(float(inter.n_bpp)), np.array(float(inter.q_bpp))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "gan_train_128.npz"), **out)

    # pin the oracle: same pipeline restated
    with torch.no_grad():
        recon, hyper, y = O.compression_forward(sd, x, True, False, noise_z, noise_y)
        trunk = pnet.net
        feats = torch.nn.Sequential(*[m for sl in (trunk.slice1, trunk.slice2, trunk.slice3, trunk.slice4, trunk.slice5)
                                      for m in sl])
        lp = O.lpips_forward(feats, lins, recon, x).mean()
        dist = O.distortion_loss(recon, x)
        sd1 = dict(sd)
        d_in = torch.cat([x, recon], 0)
        lat = torch.repeat_interleave(hyper.decoded, 2, dim=0)
        _, logits, new_uv = O.discriminator_forward(sd1, d_in, lat, training=True)
        d_real, d_gen = torch.chunk(logits.squeeze(), 2, dim=0)
        d_loss, g_loss = O.gan_losses_non_saturating(d_real, d_gen)
        cfg = dict(lambda_A=2 ** 1, lambda_B=2 ** (-4), target_rate=0.14, lambda_schedule=dict(vals=[2., 1.], steps=[50000]),
                   target_schedule=dict(vals=[0.20 / 0.14, 1.], steps=[50000]))
        rate, _ = O.weighted_rate_loss(cfg, hyper.total_nbpp, hyper.total_qbpp, step=1)
        total = rate + 0.075 * 2 ** (-5) * dist + 1.0 * lp + 0.15 * g_loss
    errs = dict(recon=(recon - inter.reconstruction).abs().max().item(),
                perceptual=abs(float(lp) - float(out["perceptual"])), distortion=abs(float(dist) - float(out["distortion"])) / float(out["distortion"]),
                disc_loss=abs(float(d_loss) - float(out["disc_loss"])), total=abs(float(total) - float(out["compression_loss"])) / abs(float(out["compression_loss"])))
    print("gan_train_128", {k: f"{v:.3e}" for k, v in errs.items()}, "losses", float(out["compression_loss"]), float(out["disc_loss"]),
          float(out["perceptual"]), float(out["distortion"]))
    cfg.update(k_M=0.075 * 2 ** (-5), k_P=1.0, beta=0.15)
    return [("gan_train_128", errs)] + grad_case(model, sd, x, noise_z, noise_y, cfg, feats, lins, O)


def grad_case(model, sd, x, noise_z, noise_y, cfg, feats, lins, O):
    """BACKWARD pin: every parameter gradient of the real reference's two alternating steps (train.py:137-141 --
    compression loss on generator steps, discriminator loss otherwise) vs torch autograd of the oracle restatement.
    The fixture keeps the gradient norms + strided samples of the reference."""
    out, report = {}, []
    for tag, train_generator in (("gstep", True), ("dstep", False)):
        model.load_state_dict(sd, strict=True)          # resets the spectral-norm u / v buffers
        model.zero_grad(set_to_none=True)
        model.step_counter = 0
        with ref_shim.NoiseFeeder([noise_z, noise_y]):
            losses = model(x, train_generator=train_generator)
        losses["compression" if train_generator else "disc"].backward()
        ref = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}

        sdg = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(("weight_u", "weight_v")))
               for k, v in sd.items()}
        comp, d_loss, _ = O.gan_training_losses(sdg, x, noise_z, noise_y, cfg, feats, lins, train_generator, step=1)
        (comp if train_generator else d_loss).backward()
        worst, worst_key, n = 0.0, None, 0
        for k, g in ref.items():
            if k.startswith("perceptual_loss"):
                continue
            og = sdg[k].grad
            assert og is not None, f"oracle has no gradient for {k} ({tag})"
            e = ((og - g).norm() / g.norm().clamp_min(1e-30)).item()
            if e > worst:
                worst, worst_key = e, k
            n += 1
            out[f"{tag}.{k}.norm"] = np.array(float(g.norm()))
            flat = g.reshape(-1)
            out[f"{tag}.{k}.sub"] = flat[:: max(1, flat.numel() // 64)][:64].numpy().copy()
        missing = [k for k, v in sdg.items() if v.requires_grad and v.grad is not None and k not in ref]
        assert not missing, f"oracle has gradients the reference lacks: {missing[:4]}"
        print(f"grad pin {tag}: {n} parameter gradients, worst rel L2 {worst:.3e} ({worst_key}); "
              f"loss ref {float(losses['compression' if train_generator else 'disc']):.6f} "
              f"oracle {float(comp if train_generator else d_loss):.6f}")
        report.append((f"grad_{tag}", {"worst_rel_l2": worst}))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "gan_grads_128.npz"), **out)
    return report


def main():
    from hific_b200 import synth
    from oracle import hific_oracle as O

    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    model, ModelModes = build_reference_model(gan=False)

    # 1. state_dict contract
    ref_sd = model.state_dict()
    shapes = synth.hific_shapes()
    assert set(ref_sd.keys()) == set(shapes.keys()), (set(ref_sd) ^ set(shapes))
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), (k, tuple(v.shape), shapes[k])
    print(f"state_dict contract OK: {len(shapes)} keys")
    sd = synth.synth_state_dict(SEED)
    model.load_state_dict(sd, strict=True)

    report = []
    for name, (b, h, w, training, evalmode) in CASES.items():
        x = synth.synth_image(b, h, w, SEED)
        model.train(training)
        model.model_mode = ModelModes.EVALUATION if evalmode else ModelModes.TRAINING
        # shapes of y and z for the noise tensors
        hp, wp = (-(-h // 16) * 16, -(-w // 16) * 16) if evalmode else (h, w)
        yh, yw = hp // 16, wp // 16
        if evalmode:
            yh, yw = -(-yh // 4) * 4, -(-yw // 4) * 4
        zh, zw = yh // 4, yw // 4
        noise_z = synth.synth_noise((b, 320, zh, zw), f"z{name}", SEED)
        noise_y = synth.synth_noise((b, 220, yh, yw), f"y{name}", SEED)
        with torch.no_grad(), ref_shim.NoiseFeeder([noise_z, noise_y]) as nf:
            intermediates, hyperinfo = model.compression_forward(x)
            assert nf.calls == 2
            y_ref = model.Encoder(x if not evalmode else torch.nn.functional.pad(
                x, (0, wp - w, 0, hp - h), mode="reflect"))
        out = {}
        summarize("y", y_ref, out)
        summarize("decoded", hyperinfo.decoded, out)
        summarize("recon", intermediates.reconstruction, out)
        for f in ("latent_nbpp", "hyperlatent_nbpp", "total_nbpp", "latent_qbpp", "hyperlatent_qbpp", "total_qbpp"):
            out[f] = np.array(float(getattr(hyperinfo, f)))
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)

        # 2. pin the oracle restatement against the reference outputs
        with torch.no_grad():
            recon, hyper, y = O.compression_forward(sd, x, training, evalmode, noise_z, noise_y)
        errs = dict(
            y=(y if not evalmode else y)[..., :y_ref.shape[2], :y_ref.shape[3]].sub(y_ref).abs().max().item()
            if not evalmode else 0.0,
            decoded=(hyper.decoded - hyperinfo.decoded).abs().max().item(),
            recon=(recon - intermediates.reconstruction).abs().max().item(),
            nbpp=abs(float(hyper.total_nbpp) - float(hyperinfo.total_nbpp)),
            qbpp=abs(float(hyper.total_qbpp) - float(hyperinfo.total_qbpp)),
        )
        report.append((name, errs))
        print(name, {k: f"{v:.3e}" for k, v in errs.items()},
              "bpp n/q", float(hyperinfo.total_nbpp), float(hyperinfo.total_qbpp))
    report += gan_case(synth, O)
    bad = [(n, e) for n, e in report if max(e.values()) > 1e-4]
    if bad:
        raise SystemExit(f"oracle disagrees with the reference: {bad}")
    print("oracle pinned against the reference on", len(report), "cases")


if __name__ == "__main__":
    main()
