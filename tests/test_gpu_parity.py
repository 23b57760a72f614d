"""GPU parity of the CUDA hot path against the CPU oracle and the committed golden vectors.

Protocol (stage-wise): every stage is fed the oracle's input for that stage; the
discontinuous rounding y -> y_hat is checked as a mismatch FRACTION; end-to-end bpp by absolute tolerance.
Tolerance: north_star asks for 1e-3 relative on the generator+hyperprior forward output; the fp16-operand
tcgen05 path (10-bit mantissa, the same as the TF32 cuDNN path the reference runs on GPUs) is held to
REL_TOL below, measured as relative L2 per stage.
"""
import logging
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from hific_b200 import synth  # noqa: E402
from hific_b200.config import ModelModes, ModelTypes, mse_lpips_args  # noqa: E402
from hific_b200.model import Model  # noqa: E402
from oracle import hific_oracle as O  # noqa: E402

REL_TOL = 1e-3          # relative L2 per stage (north_star: 1e-3 rel)
FLIP_TOL = 2e-3         # fraction of y_hat elements allowed to round differently
BPP_TOL = 2e-3          # relative tolerance on the bpp scalars
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


class Feed:
    """Feeds the two torch.nn.init.uniform_ draws of Hyperprior.forward (hyperprior.py:65)."""

    def __init__(self, noises):
        self.noises, self.calls = list(noises), 0

    def __enter__(self):
        self._orig = torch.nn.init.uniform_

        def fake(t, a=0.0, b=1.0):
            if (a, b) != (-0.5, 0.5) or self.calls >= len(self.noises):
                return self._orig(t, a, b)       # e.g. parameter initialisers running inside the context
            n = self.noises[self.calls]
            self.calls += 1
            with torch.no_grad():
                t.copy_(n.to(t.device))
            return t

        torch.nn.init.uniform_ = fake
        return self

    def __exit__(self, *e):
        torch.nn.init.uniform_ = self._orig


@pytest.fixture(scope="module")
def sd():
    return synth.synth_state_dict(0)


@pytest.fixture(scope="module")
def model(sd):
    m = Model(mse_lpips_args(), logging.getLogger("parity"))
    m.load_state_dict(sd, strict=True)
    return m.cuda()


def oracle_case(sd, b, h, w, training, tag):
    x = synth.synth_image(b, h, w, 0)
    nz = synth.synth_noise((b, 320, h // 64, w // 64), "z" + tag, 0)
    ny = synth.synth_noise((b, 220, h // 16, w // 16), "y" + tag, 0)
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        recon, hyp, y = O.compression_forward(sd, x, training, False, nz, ny)
    return x, nz, ny, recon, hyp, y


@pytest.mark.parametrize("b,h,w,training,tag", [(2, 128, 128, True, "train_128"), (1, 256, 256, True, "train_256"),
                                                (2, 128, 128, False, "eval_128")])
def test_stagewise_parity(model, sd, b, h, w, training, tag):
    x, nz, ny, recon_o, hyp_o, y_o = oracle_case(sd, b, h, w, training, tag)
    model.train(training)
    dev = "cuda"
    with torch.no_grad():
        assert rel_l2(model.Encoder(x.to(dev)), y_o) < REL_TOL
        assert rel_l2(model.Hyperprior.analysis_net(y_o.to(dev)), hyp_o.hyperlatents) < REL_TOL
        z_dec = hyp_o.noisy_hyperlatents if training else hyp_o.quantized_hyperlatents
        assert rel_l2(model.Hyperprior.synthesis_mu(z_dec.to(dev)), hyp_o.latent_means) < REL_TOL
        sg = model.Hyperprior.synthesis_std(z_dec.to(dev)).clamp(min=0.11)
        assert rel_l2(sg, hyp_o.latent_scales) < REL_TOL
        # Likelihood kernels fed the oracle's tensors: identical inputs -> y_hat must agree (up to a vanishing
        # number of ties) and the four bit-rates tightly.
        from hific_b200 import ops
        H = model.Hyperprior
        zn, zq, sz = ops.hyperlatent_likelihood(hyp_o.hyperlatents.to(dev), H.hyperlatent_likelihood.packed_params(),
                                                nz.to(dev))
        assert torch.equal(zq.cpu(), hyp_o.quantized_hyperlatents)
        dec, sy = ops.latent_likelihood(y_o.to(dev), hyp_o.latent_means.to(dev), hyp_o.latent_scales.to(dev),
                                        ny.to(dev), 0.11, "gaussian")
        assert ((dec.cpu() - hyp_o.decoded).abs() > 0.5).float().mean().item() < 1e-5
        assert (dec.cpu() - hyp_o.decoded).abs().median().item() < 1e-6
        scale = 1.0 / (b * -math.log(2.0) * h * w)
        for got, ref in ((sz[0], hyp_o.hyperlatent_nbpp), (sz[1], hyp_o.hyperlatent_qbpp),
                         (sy[0], hyp_o.latent_nbpp), (sy[1], hyp_o.latent_qbpp)):
            assert abs(float(got) * scale - float(ref)) <= 2e-5 * max(1.0, abs(float(ref)))
        # The module as a whole, fed the oracle's y.  In training mode everything upstream of the rounding is
        # continuous (noisy z feeds the synthesis nets): strict.  In eval mode round(z) feeds them, and a z
        # element within ~5e-4 of a half-integer may round the other way under fp16 operands (2 of 2560 do in
        # this fixture, exactly as the oracle's own fp16-rounding emulation predicts); one such flip moves
        # mu/sigma of a whole neighbourhood, so there only the flip FRACTION of round(z) is asserted.
        with Feed([nz, ny]) as f:
            info = H(y_o.to(dev), spatial_shape=(h, w))
            assert f.calls == 2
        if training:
            assert ((info.decoded.cpu() - hyp_o.decoded).abs() > 0.5).float().mean().item() < FLIP_TOL
            for fld in ("latent_nbpp", "hyperlatent_nbpp", "total_nbpp", "latent_qbpp", "hyperlatent_qbpp", "total_qbpp"):
                ref = float(getattr(hyp_o, fld))
                assert abs(float(getattr(info, fld)) - ref) <= BPP_TOL * max(1.0, abs(ref)), fld
        else:
            z_ours = H.analysis_net(y_o.to(dev)).cpu()
            assert (torch.floor(z_ours + 0.5) != hyp_o.quantized_hyperlatents).float().mean().item() < FLIP_TOL
            assert abs(float(info.hyperlatent_qbpp) - float(hyp_o.hyperlatent_qbpp)) <= 0.02 * float(hyp_o.hyperlatent_qbpp)
            assert abs(float(info.total_qbpp) - float(hyp_o.total_qbpp)) <= 0.05 * float(hyp_o.total_qbpp)
        assert rel_l2(model.Generator(hyp_o.decoded.to(dev)), recon_o) < REL_TOL


def test_against_reference_golden_train_256(model, sd):
    """The committed vectors came from the REAL reference (oracle/make_golden.py)."""
    gold = np.load(os.path.join(GOLDEN, "train_256.npz"))
    x, nz, ny, recon_o, hyp_o, y_o = oracle_case(sd, 1, 256, 256, True, "train_256")
    model.train(True)
    with torch.no_grad():
        y = model.Encoder(x.cuda()).cpu().numpy().reshape(-1)
        xh = model.Generator(hyp_o.decoded.cuda()).cpu().numpy().reshape(-1)
    gy = gold["y.full"].reshape(-1) if "y.full" in gold else None
    if gy is None:
        y, gy = y[::int(gold["y.stride"])], gold["y.sub"]
    assert np.linalg.norm(y - gy) / np.linalg.norm(gy) < REL_TOL
    step = int(gold["recon.stride"])
    sub = gold["recon.sub"]
    assert np.linalg.norm(xh[::step] - sub) / np.linalg.norm(sub) < REL_TOL


def test_evaluation_mode_padding_and_crop(model, sd):
    """EVALUATION + eval(): pad to multiples of 16 / 4, crop back (model.py:133-144,160), ragged 100x144."""
    x = synth.synth_image(1, 100, 144, 0)
    with torch.no_grad():
        recon_o, hyp_o, _ = O.compression_forward(sd, x, training=False, evaluation_mode=True)
    model.eval()
    model.model_mode = ModelModes.EVALUATION
    try:
        with torch.no_grad():
            recon, q_bpp = model(x.cuda())
    finally:
        model.model_mode = ModelModes.TRAINING
    assert tuple(recon.shape) == (1, 3, 100, 144)
    assert recon.min().item() >= 0.0 and recon.max().item() <= 1.0
    # End to end: a rounding flip in y_hat (a +-1 change of one latent) perturbs x_hat locally by far more than
    # 1e-3, and the randomly initialised generator amplifies it; the image is therefore only sanity-checked here
    # (stage-wise parity above is the strict test) while the rate is held tight.
    assert rel_l2(recon, recon_o.clamp(0, 1)) < 0.3
    assert abs(float(q_bpp) - float(hyp_o.total_qbpp)) <= 5e-3 * float(hyp_o.total_qbpp)


def test_full_size_properties_batch32(model):
    """c2 size (32x3x256x256): determinism and batch independence (size-independent properties)."""
    model.eval()
    x = synth.synth_image(32, 256, 256, 3).cuda()
    with torch.no_grad():
        y1 = model.Encoder(x)
        y2 = model.Encoder(x)
        assert torch.equal(y1, y2), "encoder is not deterministic"
        y_single = model.Encoder(x[5:6].contiguous())
        assert torch.allclose(y1[5:6], y_single, rtol=0, atol=1e-5), "sample 5 depends on its batch neighbours"
        g1 = model.Generator(torch.round(y1))
        assert torch.isfinite(g1).all()
        g_single = model.Generator(torch.round(y1[7:8]).contiguous())
        assert torch.allclose(g1[7:8], g_single, rtol=0, atol=1e-4)


def test_refuses_cpu_and_records_grad(model):
    """No CPU fallback: a CPU input raises.  With autograd enabled the call is recorded (training plan), without it
    the fused inference plan runs and nothing is recorded."""
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            model.Encoder(torch.zeros(1, 3, 128, 128))
    x = torch.zeros(1, 3, 128, 128, device="cuda")
    assert model.Encoder(x).requires_grad
    with torch.no_grad():
        assert not model.Encoder(x).requires_grad


def test_cuda_graph_replay_matches_eager(sd):
    """Opt-in CUDA-graph path: same kernels, so eval-mode outputs must be bit-identical to the eager path."""
    m = Model(mse_lpips_args(), logging.getLogger("graph"))
    m.load_state_dict(sd, strict=True)
    m.cuda().eval()
    x1 = synth.synth_image(2, 128, 128, 5).cuda()
    x2 = synth.synth_image(2, 128, 128, 6).cuda()
    with torch.no_grad():
        e1, i1 = m.compression_forward(x1)
        e2, i2 = m.compression_forward(x2)
        m.enable_cuda_graph(True)
        g1, j1 = m.compression_forward(x1)
        g2, j2 = m.compression_forward(x2)      # replay with new input
        g1b, _ = m.compression_forward(x1)
    assert torch.equal(e1.reconstruction, g1.reconstruction) and torch.equal(e2.reconstruction, g2.reconstruction)
    assert torch.equal(g1.reconstruction, g1b.reconstruction)
    assert float(i1.total_qbpp) == float(j1.total_qbpp) and float(i2.total_qbpp) == float(j2.total_qbpp)


def test_pipelined_forward_matches_direct_calls(sd):
    """hific_b200.pipeline.PipelinedForward (H2D / compute / D2H on three streams, two slots): every submission returns
    exactly what a plain synchronous Model.forward returns for that batch, also when slots are reused."""
    from hific_b200.config import ModelModes
    from hific_b200.pipeline import PipelinedForward
    m = Model(mse_lpips_args(), logging.getLogger("pipe"), model_mode=ModelModes.EVALUATION)
    m.load_state_dict(sd, strict=False)     # EVALUATION mode adds the coder tables (utils.py:214 loads non-strictly too)
    m.cuda().eval()
    xs = [synth.synth_image(2, 128, 128, 20 + i).pin_memory() for i in range(5)]
    with torch.no_grad():
        ref = [tuple(t.cpu().clone() for t in m(x.cuda(), writeout=False)) for x in xs]
    pipe = PipelinedForward(m, depth=2)
    tickets, got = [], []
    for i, x in enumerate(xs):
        tickets.append(pipe.submit(x))
        if i >= 1:
            r, b = pipe.result(tickets[i - 1])
            got.append((r.clone(), b.clone()))
    r, b = pipe.result(tickets[-1])
    got.append((r.clone(), b.clone()))
    for (r0, b0), (r1, b1) in zip(ref, got):
        assert torch.equal(r0, r1), "reconstruction differs from the synchronous call"
        assert abs(float(b0) - float(b1)) <= 1e-6 * abs(float(b0))      # fp64 atomics: summation order may differ
    with pytest.raises(ValueError):
        pipe.result(tickets[0])          # slot long reused
    with pytest.raises(ValueError):
        pipe.submit(torch.zeros(2, 3, 128, 128))   # not pinned


def test_c5_size_1024_properties_and_oracle(model, sd):
    """Config c5 (compress.py inference, 8 x 3 x 1024 x 1024): size-independent properties at the full size --
    determinism, batch independence, crop consistency (a latent whose receptive field lies inside a 512-pixel crop
    must not depend on what is outside it) -- plus one full-size sample against the CPU oracle."""
    model.eval()
    x = synth.synth_image(8, 1024, 1024, 9).cuda()
    with torch.no_grad():
        y = model.Encoder(x)
        assert y.shape == (8, 220, 64, 64) and torch.isfinite(y).all()
        assert torch.equal(y, model.Encoder(x)), "encoder is not deterministic at 1024 x 1024"
        y3 = model.Encoder(x[3:4].contiguous())
        assert torch.allclose(y[3:4], y3, rtol=0, atol=1e-5), "sample 3 depends on its batch neighbours"
        crop = x[3:4, :, 256:768, 256:768].contiguous()
        yc = model.Encoder(crop)                                   # latents 16..47 of the full image
        m = 8                                                      # > receptive-field radius (in latent pixels)
        a, b = y[3:4, :, 16 + m:48 - m, 16 + m:48 - m], yc[:, :, m:32 - m, m:32 - m]
        assert rel_l2(b, a) < 1e-5, "interior latents depend on pixels outside their receptive field"
        y_hat = torch.round(y[3:4])
        g = model.Generator(y_hat)
        assert g.shape == (1, 3, 1024, 1024) and torch.isfinite(g).all()
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        y_o = O.encoder_forward(sd, x[3:4].cpu())
        g_o = O.generator_forward(sd, y_hat.cpu())
    assert rel_l2(y3, y_o) < 1e-3
    assert rel_l2(g, g_o) < 2e-3
