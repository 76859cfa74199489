"""GPU parity of the UNet generator and PatchGAN discriminator engines against the oracle and reference goldens."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import nets

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return z, json.loads(str(z["meta"]))


def _x(m, c=3):
    g = torch.Generator().manual_seed(m["x_seed"])
    return torch.rand((m["n"], c, m["hw"], m["hw"]), generator=g) * 2 - 1


@pytest.fixture(scope="module")
def engine_mod():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from deepliif_b200 import engine
    return engine


@pytest.mark.parametrize("name", ["unet256_batch_256", "unet512_batch_512", "unet128_inst_128"])
@pytest.mark.parametrize("precision,tol", [("bf16x3", 1e-3), ("fp16x3", 1e-3)])
def test_unet_vs_golden_and_oracle(engine_mod, name, precision, tol):
    z, m = _load(name)
    sd = nets.make_state_dict(nets.unet_param_shapes(m["num_downs"], 64, 3, 3, m["norm"]), m["seed"], m["init"])
    x = _x(m)
    eng = engine_mod.UnetEngine(sd, num_downs=m["num_downs"], norm=m["norm"], precision=precision)
    y = eng.forward(x.cuda()).cpu()
    with torch.no_grad():
        y_orc = nets.unet_forward(x, sd, num_downs=m["num_downs"], norm=m["norm"], norm_mode="sample")
    s = m["subsample"]
    err_gold = np.abs(y.numpy()[:, :, ::s, ::s] - z["y"]).max()
    err_orc = (y - y_orc).abs().max().item()
    print(f"{name} {precision}: max|d| vs golden {err_gold:.3e}, vs oracle {err_orc:.3e}")
    assert err_gold <= tol and err_orc <= tol


@pytest.mark.parametrize("name", ["dbasic_batch_128", "dn4_inst_128"])
def test_discriminator_vs_golden_and_oracle(engine_mod, name):
    z, m = _load(name)
    sd = nets.make_state_dict(nets.nlayer_d_param_shapes(m["n_layers"], 64, 6, m["norm"]), m["seed"], m["init"])
    x = _x(m, 6)
    eng = engine_mod.NLayerDEngine(sd, n_layers=m["n_layers"], norm=m["norm"], norm_mode="batch")
    y = eng.forward(x.cuda()).cpu()
    with torch.no_grad():
        y_orc = nets.nlayer_d_forward(x, sd, n_layers=m["n_layers"], norm=m["norm"], norm_mode="batch")
    err_gold = np.abs(y.numpy() - z["y"]).max()
    err_orc = (y - y_orc).abs().max().item()
    print(f"{name}: max|d| vs golden {err_gold:.3e}, vs oracle {err_orc:.3e} (|y| max {y_orc.abs().max():.2f})")
    assert err_gold <= 1e-3 * max(1.0, float(np.abs(z["y"]).max())) and err_orc <= 1e-3 * max(1.0, y_orc.abs().max().item())


def test_discriminator_512_basic(engine_mod):
    """BASELINE config 4 shape: 70x70 PatchGAN on [N,6,512,512] -> [N,1,62,62]."""
    sd = nets.make_state_dict(nets.nlayer_d_param_shapes(3, 64, 6, "instance"), 41, "stress")
    g = torch.Generator().manual_seed(5)
    x = torch.rand((2, 6, 512, 512), generator=g) * 2 - 1
    eng = engine_mod.NLayerDEngine(sd, n_layers=3, norm="instance", norm_mode="batch")
    y = eng.forward(x.cuda()).cpu()
    assert tuple(y.shape) == (2, 1, 62, 62)
    torch.set_num_threads(min(32, os.cpu_count()))
    with torch.no_grad():
        y_orc = nets.nlayer_d_forward(x, sd, n_layers=3, norm="instance", norm_mode="batch")
    err = (y - y_orc).abs().max().item()
    print(f"D basic 512: max|d| vs oracle {err:.3e} (|y| max {y_orc.abs().max():.2f})")
    assert err <= 1e-3 * max(1.0, y_orc.abs().max().item())


def test_unet256_single_pass_bf16_tolerance(engine_mod):
    """BASELINE configs[4] precision: single-pass bf16 operands.  Not an fp32-parity mode: operand rounding alone gives
    ~1e-2 (SURVEY.md 8d measured 9e-3 by emulation on UNet-512); the stated tolerance for this mode is 5e-2 max-abs."""
    z, m = _load("unet256_batch_256")
    sd = nets.make_state_dict(nets.unet_param_shapes(8, 64, 3, 3, "batch"), m["seed"], m["init"])
    x = _x(m)
    y = engine_mod.UnetEngine(sd, num_downs=8, norm="batch", precision="bf16").forward(x.cuda()).cpu()
    err = np.abs(y.numpy()[:, :, ::4, ::4] - z["y"]).max()
    print(f"unet256 single-pass bf16: max|d| vs reference golden {err:.3e}")
    assert err <= 5e-2
