"""GPU: the InstanceNorm2d variant (use_channel_norm = False; src/normalisation/instance.py:7-15, encoder.py:41-44,
generator.py:21-24, 81-84) on the real kernels.
  * hfc_instancenorm / hfc_instancenorm_bwd against torch.nn.functional.instance_norm and its autograd on the same device
    (fp32, TF32 off): every output form the plans use -- fp32 rows, bordered fp16 act buffer with reflected border and
    zeroed channel padding, residual adds, ReLU; dz as fp32 rows and as the 16-bit operand, dgamma / dbeta / dbias;
  * Encoder / Generator built with channel_norm=False against the oracle (which tests/test_instancenorm_cpu.py pins to
    golden vectors of the real reference modules): forward without autograd, forward in training mode and every
    parameter gradient.
Tolerances: fp32 outputs 1e-4 relative (summation order); fp16 buffers at fp16 resolution; network gradients 6e-2
relative L2 per tensor against the operand-matched oracle and 1e-1 against the fp32 oracle (fp16 operands + ReLU-mask flips on
small maps; measured 4.0e-2 / 7.2e-2)."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hific_b200 import grad as _grad, ops, synth, train_plan  # noqa: E402
from hific_b200.network import encoder, generator  # noqa: E402
from hific_b200.ops import ACT_NONE, ACT_RELU, Geom, round_up  # noqa: E402
from oracle import hific_oracle as O  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def rows(t):
    return t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]).contiguous()


CASES = [  # n, c, h, w, cpad, border (pt, pl, pb, pr), act, residuals, mean offset
    (2, 60, 64, 48, 64, (1, 0, 0, 1), ACT_RELU, 0, 0.0),      # Encoder block 1 (asymmetric reflect border)
    (3, 960, 16, 16, 960, (1, 1, 1, 1), ACT_NONE, 2, 0.5),    # residual trunk: + identity + head
    (2, 220, 9, 7, 256, (1, 1, 1, 1), ACT_NONE, 0, 40.0),     # Generator input norm, ragged map, |mean| >> std
    (1, 480, 32, 32, 512, (0, 0, 0, 0), ACT_RELU, 0, 0.0),    # channel padding, no border
    (2, 120, 130, 70, 128, (3, 3, 3, 3), ACT_RELU, 1, -3.0),  # wide border, strip tail (130 * 70 is no multiple of 8)
]


@pytest.mark.parametrize("n,c,h,w,cpad,border,act,nres,offset", CASES)
def test_instancenorm_forward_kernel(n, c, h, w, cpad, border, act, nres, offset):
    g = torch.Generator().manual_seed(c + h)
    x = (torch.randn(n, c, h, w, generator=g) * (1 + torch.rand(1, c, 1, 1, generator=g) * 3) + offset).to(DEV)
    gamma = (1 + 0.2 * torch.randn(c, generator=g)).to(DEV)
    beta = (0.2 * torch.randn(c, generator=g)).to(DEV)
    res = [torch.randn(n, c, h, w, generator=g).to(DEV) for _ in range(nres)]
    geom = Geom(n, h, w, c, cpad, *border)
    reflect = any(border)
    out_act, out_f32 = ops.instancenorm(rows(x), geom, gamma, beta, act=act, reflect=reflect,
                                        res1=rows(res[0]) if nres > 0 else None, res2=rows(res[1]) if nres > 1 else None,
                                        want_f32=True)
    ref = F.instance_norm(x, weight=gamma, bias=beta, eps=1e-5)
    ref = F.relu(ref) if act == ACT_RELU else ref
    for r in res:
        ref = ref + r
    assert rel(out_f32, rows(ref)) < 1e-4 * (1 + abs(offset))
    want = F.pad(ref, (border[1], border[3], border[0], border[2]), mode="reflect") if reflect else ref
    got = out_act[..., :c].permute(0, 3, 1, 2).float()
    assert torch.allclose(got, want, rtol=2e-3, atol=2e-3 * (1 + abs(offset) * 1e-2))
    if cpad > c:
        assert out_act[..., c:].float().abs().max().item() == 0.0


@pytest.mark.parametrize("n,c,h,w,cpad,border,act,nres,offset", CASES)
def test_instancenorm_backward_kernel(n, c, h, w, cpad, border, act, nres, offset):
    g = torch.Generator().manual_seed(7 * c + h)
    x = (torch.randn(n, c, h, w, generator=g) * (1 + torch.rand(1, c, 1, 1, generator=g) * 3) + offset).to(DEV)
    gamma = (1 + 0.2 * torch.randn(c, generator=g)).to(DEV)
    beta = (0.2 * torch.randn(c, generator=g)).to(DEV)
    up = torch.randn(n, c, h, w, generator=g).to(DEV)
    xo, go, bo = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.instance_norm(xo, weight=go, bias=bo, eps=1e-5)
    y = F.relu(y) if act == ACT_RELU else y
    (y * up).sum().backward()
    tol = 2e-4 * (1 + abs(offset))
    # fp32 rows
    dz, dg, db, dbias = train_plan.norm_bwd(rows(x), rows(up), gamma, beta, act, as_operand=False, kind="instance", n=n)
    assert rel(dz[:, :c], rows(xo.grad)) < tol
    assert rel(dg, go.grad) < tol and rel(db, bo.grad) < tol
    # the bias of the conv in front of the norm: sum of dz over the pixels = 0 up to rounding
    assert dbias.abs().max().item() <= 1e-3 * rows(xo.grad).abs().sum(0).max().item() + 1e-6
    # 16-bit operand form (pitch round_up(c, 64), channel padding zeroed)
    dz16, dg2, db2, _ = train_plan.norm_bwd(rows(x), rows(up), gamma, beta, act, as_operand=True, kind="instance", n=n)
    op = dz16.view(torch.bfloat16 if _grad.GRAD_BF16 else torch.float16).float()
    assert op.shape[1] == round_up(c, 64)
    assert rel(op[:, :c], rows(xo.grad)) < (8e-3 if _grad.GRAD_BF16 else 1e-3)
    if op.shape[1] > c:
        assert op[:, c:].abs().max().item() == 0.0
    assert rel(dg2, go.grad) < tol and rel(db2, bo.grad) < tol


def _modules(n_res=2):
    sd = synth.instance_norm_variant(synth.synth_state_dict(3, n_residual_blocks=n_res))
    enc = encoder.Encoder((3, 128, 128), 2, C=220, channel_norm=False)
    gen = generator.Generator((220, 8, 8), 2, C=220, n_residual_blocks=n_res, channel_norm=False)
    enc.load_state_dict({k[8:]: v for k, v in sd.items() if k.startswith("Encoder.")}, strict=True)
    gen.load_state_dict({k[10:]: v for k, v in sd.items() if k.startswith("Generator.")}, strict=True)
    return sd, enc.to(DEV), gen.to(DEV)


def test_networks_with_instance_norm_against_oracle():
    n_res = 2
    sd, enc, gen = _modules(n_res)
    g = torch.Generator().manual_seed(21)
    x = synth.synth_image(2, 192, 160, 5)
    y_hat = torch.round(torch.randn((2, 220, 12, 10), generator=g) * 2)
    w_enc = torch.randn((2, 220, 12, 10), generator=g)
    w_gen = torch.randn((2, 3, 192, 160), generator=g)

    def oracle(rnd):
        sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith(("Encoder.", "Generator."))}
        y_o = O.encoder_forward(sdg, x, rnd=rnd)
        (y_o * w_enc).sum().backward()
        yo = y_hat.clone().requires_grad_(True)
        x_o = O.generator_forward(sdg, yo, n_residual_blocks=n_res, rnd=rnd)
        (x_o * w_gen).sum().backward()
        return sdg, y_o.detach(), x_o.detach(), yo.grad

    # fp32 oracle (the reference's arithmetic) and the operand-matched oracle (conv operands rounded to fp16 in the forward,
    # as the kernels do: the same ReLU masks up to rounding, so what remains is the backward's own arithmetic -- DESIGN 3.13)
    sdg, y_o, x_o, dy_o = oracle(O._ident)
    sdm, _, _, dy_m = oracle(O.round_fp16)
    # inference (no autograd): the same layer-by-layer plan
    enc.eval(), gen.eval()
    with torch.no_grad():
        assert rel(enc(x.to(DEV)).cpu(), y_o) < 5e-3
        assert rel(gen(y_hat.to(DEV)).cpu(), x_o) < 5e-3
    # training: forward + every gradient
    enc.train(), gen.train()
    yp = y_hat.to(DEV).requires_grad_(True)
    y_p = enc(x.to(DEV))
    (y_p * w_enc.to(DEV)).sum().backward()
    x_p = gen(yp)
    (x_p * w_gen.to(DEV)).sum().backward()
    assert rel(y_p.detach().cpu(), y_o) < 5e-3 and rel(x_p.detach().cpu(), x_o) < 5e-3
    assert rel(yp.grad.cpu(), dy_m) < 5e-2 and rel(yp.grad.cpu(), dy_o) < 1e-1
    worst_m, worst_o = ("", 0.0), ("", 0.0)
    for prefix, mod in (("Encoder.", enc), ("Generator.", gen)):
        for name, p in mod.named_parameters():
            want = sdg[prefix + name].grad
            if name.endswith(".bias") and want.abs().max() < 1e-3 * sdg[prefix + name[:-5] + ".weight"].grad.abs().max():
                # conv bias in front of an InstanceNorm: mathematically zero gradient, rounding noise on both sides
                assert p.grad.abs().max().item() < 2e-2 * sdg[prefix + name[:-5] + ".weight"].grad.abs().max().item(), name
                continue
            worst_m = max(worst_m, (prefix + name, rel(p.grad.cpu(), sdm[prefix + name].grad)), key=lambda t: t[1])
            worst_o = max(worst_o, (prefix + name, rel(p.grad.cpu(), want)), key=lambda t: t[1])
    print(f"instance-norm networks, worst parameter gradient: vs operand-matched oracle {worst_m}, vs fp32 oracle {worst_o}")
    # measured on a B200 (profiles/r02_instancenorm_tests.log): 4.0e-2 matched (Generator.resblock_1.norm1.bias, a 120-pixel map),
    # 7.2e-2 against the fp32 oracle (Encoder.conv_block2.2.bias) -- ReLU-mask flips of a 2-image batch, cf. DESIGN.md 3.13
    assert worst_m[1] < 6e-2, worst_m
    assert worst_o[1] < 1e-1, worst_o


def test_model_level_use_channel_norm_false():
    """Model(args.use_channel_norm = False) (src/model.py:69-72 passes the flag to both networks): the evaluation forward
    runs end to end on the instance-norm plans; rate held tight against the oracle, image sanity-checked (a rounding flip
    of one latent moves x_hat by far more than the kernel tolerances -- see test_gpu_parity)."""
    import logging
    from hific_b200.config import ModelModes, ModelTypes, mse_lpips_args
    from hific_b200.model import Model
    cfg = mse_lpips_args()
    cfg.use_channel_norm = False
    cfg.n_residual_blocks = 2
    model = Model(cfg, logging.getLogger("in-test"), model_mode=ModelModes.EVALUATION, model_type=ModelTypes.COMPRESSION)
    sd = synth.instance_norm_variant(synth.synth_state_dict(0, n_residual_blocks=2))
    model.load_state_dict(sd, strict=False)          # EVALUATION mode adds the coder tables
    model.to(DEV).eval()
    x = synth.synth_image(2, 128, 128, 9)
    with torch.no_grad():
        recon, q_bpp = model(x.to(DEV), writeout=False)
        recon_o, hyp_o, _ = O.compression_forward(sd, x, training=False, evaluation_mode=True, n_residual_blocks=2)
    assert tuple(recon.shape) == (2, 3, 128, 128)
    assert rel(recon.cpu(), recon_o.clamp(0, 1)) < 0.3
    assert abs(float(q_bpp) - float(hyp_o.total_qbpp)) <= 5e-3 * float(hyp_o.total_qbpp)
