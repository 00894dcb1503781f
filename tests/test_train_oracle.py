"""CPU tests of the train-step oracle (oracle/train_oracle.py): the restated loss against the fixture generated from
the reference's own utils/loss_utils.py (tests/golden/make_golden_loss.py), Adam group layout, relocation algebra."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_ssim.npz")


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_loss_oracle_matches_reference_fixture(tag):
    from oracle import train_oracle as T

    z = np.load(GOLD)
    img = torch.from_numpy(z[f"{tag}_img"]).requires_grad_(True)
    gt = torch.from_numpy(z[f"{tag}_gt"])
    assert abs(T.l1_loss(img, gt).item() - float(z[f"{tag}_l1"])) < 1e-7
    assert abs(T.ssim(img[None], gt[None]).item() - float(z[f"{tag}_ssim"])) < 1e-6
    loss = T.photometric_loss(img, gt, 0.2)
    assert abs(loss.item() - float(z[f"{tag}_loss"])) < 1e-6
    loss.backward()
    ref = torch.from_numpy(z[f"{tag}_grad"])
    assert (img.grad - ref).abs().max().item() <= 1e-6 * ref.abs().max().item() + 1e-10


def test_relocate_oracle_properties():
    from oracle import train_oracle as T

    torch.manual_seed(0)
    N, D = 50, 6
    params = [torch.randn(N, w) for w in (3, 3, 3, 1, 4, 6, 15)]
    before = [p.clone() for p in params]
    dead = torch.tensor([3, 7, 11, 20])
    src = torch.tensor([5, 5, 9, 5])
    moments = [(torch.ones_like(p), torch.ones_like(p)) for p in params]
    T.relocate(params, moments, dead, src)
    # dead rows are copies of their sources except for the opacity
    for k in (0, 1, 2, 4, 5, 6):
        assert torch.equal(params[k][dead], before[k][src])
    o = torch.sigmoid(before[3][5, 0])
    want = (1 - (1 - o) ** (1.0 / 4.0)).clamp(0.005, 1 - torch.finfo(torch.float32).eps)
    got = torch.sigmoid(params[3][torch.tensor([3, 7, 20, 5]), 0])
    assert torch.allclose(got, want.expand(4), atol=1e-6)
    assert moments[0][0][5].abs().sum() == 0 and moments[0][0][9].abs().sum() == 0
    assert moments[0][0][3].abs().sum() > 0  # moments of the relocated (dead) rows are NOT reset by the reference
    grown, gm = T.add_new(params, moments, torch.tensor([1, 1, 2]))
    assert grown[0].shape[0] == N + 3 and gm[0][0].shape[0] == N + 3
    assert torch.equal(grown[0][N:], params[0][torch.tensor([1, 1, 2])])
    assert gm[3][0][N:].abs().sum() == 0 and gm[3][0][1].abs().sum() == 0


def test_expon_lr_matches_reference_formula():
    from ubs_b200.training import expon_lr

    # utils/general_utils.py:52-65 evaluated by hand
    for step in (0, 1, 100, 15000, 30000, 40000):
        t = min(max(step / 30000, 0), 1)
        want = np.exp(np.log(1.6e-4) * (1 - t) + np.log(1.6e-6) * t)
        assert abs(expon_lr(step, 1.6e-4, 1.6e-6, lr_delay_mult=0.01, max_steps=30000) - want) < 1e-18
    assert expon_lr(-1, 1.0, 1.0) == 0.0


def test_packed_adam_lr_columns_follow_record_layout():
    from ubs_b200 import fused
    from ubs_b200.training import DEFAULT_LR

    for D in (6, 7):
        sl = fused.record_slices(D)
        assert sl["opacity"].start == D + 3 and sl["scale"].start == 2 * D + 2  # the columns csrc/optim.cu assumes
        assert sl["l_triangle"].stop <= fused.record_stride(D)
    assert set(DEFAULT_LR) == set(sl)


def test_sgld_noise_oracle_properties():
    """train.py:156-163: dead primitives (opacity -> 0) take the full covariance-shaped noise, opaque ones none; the
    displacement is Sigma_xyz @ (noise * w) with the K1/K2 oracle that the _torch_impl fixtures pin."""
    from oracle import train_oracle as T
    from oracle import ubs_oracle as O

    torch.manual_seed(3)
    N, D = 64, 6
    params = [torch.randn(N, w) * 0.3 for w in (3, 3, 3, 1, 4, 6, 15)]
    params[3][:32] = -12.0   # (1 - sigmoid)^100 ~ 1
    params[3][32:] = 3.0     # (1 - 0.95)^100 underflows to 0 in FP32
    noise = torch.randn(N, 3)
    new = T.sgld_noise(params, noise, 5e5, 1.6e-4)
    assert torch.equal(new[32:], params[0][32:])
    cov = O.rot_scale_l_triangle_to_covar(O.l_triangle_to_rotmat(params[6][:, :3]), torch.nn.functional.softplus(params[5]),
                                          params[6], spatial_block=True)
    w = (1 - torch.sigmoid(params[3][:32])) ** 100
    want = torch.einsum("nij,nj->ni", cov[:32], noise[:32] * w * 5e5 * 1.6e-4)
    assert torch.allclose(new[:32] - params[0][:32], want, rtol=1e-4, atol=1e-6)
    assert torch.allclose(cov, cov.transpose(1, 2))


GROUPS = ("xyz", "mean", "rgb", "opacity", "beta", "scale", "l_triangle")


def _load_relocate_fixture(D):
    z = np.load(os.path.join(os.path.dirname(GOLD), f"relocate_D{D}.npz"))

    def state(tag):
        params = [torch.from_numpy(z[f"{tag}_{n}"]).clone() for n in GROUPS]
        moments = [(torch.from_numpy(z[f"{tag}_{n}_m"]).clone(), torch.from_numpy(z[f"{tag}_{n}_v"]).clone()) for n in GROUPS]
        return params, moments

    return z, state


@pytest.mark.parametrize("D", [6, 7])
def test_relocation_oracle_matches_reference_fixture(D):
    """oracle/train_oracle.py relocate / add_new against tests/golden/relocate_D*.npz = outputs of the reference's OWN
    relocate_gs / add_new_gs (scene/beta_model.py:575-657), run on CPU by tests/golden/make_golden_relocate.py with the
    indices torch.multinomial drew recorded.  Pins the restatement (parameters AND Adam moments)."""
    from oracle import train_oracle as T

    z, state = _load_relocate_fixture(D)
    params, moments = state("in")
    dead = torch.from_numpy(z["dead_mask"]).nonzero(as_tuple=True)[0]
    reinit = torch.from_numpy(z["reinit_idx"])
    assert dead.numel() == reinit.numel() > 10 and torch.bincount(reinit).max() >= 2
    T.relocate(params, moments, dead, reinit)
    want_p, want_m = state("relocated")
    for n, p, w in zip(GROUPS, params, want_p):
        assert torch.allclose(p, w, rtol=1e-6, atol=1e-7), n
    for n, (m, v), (wm, wv) in zip(GROUPS, moments, want_m):
        assert torch.equal(m, wm) and torch.equal(v, wv), n
    add_idx = torch.from_numpy(z["add_idx"])
    assert add_idx.numel() == int(z["n_added"]) > 0
    grown, grown_m = T.add_new(params, moments, add_idx)
    want_p, want_m = state("grown")
    for n, p, w in zip(GROUPS, grown, want_p):
        assert p.shape == w.shape and torch.allclose(p, w, rtol=1e-6, atol=1e-7), n
    for n, (m, v), (wm, wv) in zip(GROUPS, grown_m, want_m):
        assert torch.equal(m, wm) and torch.equal(v, wv), n


@pytest.mark.parametrize("D", [6, 7])
def test_sgld_noise_oracle_matches_reference_fixture(D):
    """oracle/train_oracle.py sgld_noise against tests/golden/sgld_D*.npz = the reference's own statements of
    train.py:156-163 executed by tests/golden/make_golden_sgld.py (with its pure-torch K1 / K2 for the covariance)."""
    from oracle import train_oracle as T

    z = np.load(os.path.join(os.path.dirname(GOLD), f"sgld_D{D}.npz"))
    N = z["xyz"].shape[0]
    params = [torch.from_numpy(z["xyz"]), torch.zeros(N, D - 3), torch.zeros(N, 3), torch.from_numpy(z["opacity"]),
              torch.zeros(N, D - 2), torch.from_numpy(z["scale"]), torch.from_numpy(z["l_triangle"])]
    got = T.sgld_noise(params, torch.from_numpy(z["noise"]), float(z["noise_lr"]), float(z["xyz_lr"]))
    want = torch.from_numpy(z["xyz_out"])
    moved = (want - params[0]).abs()
    assert moved.max().item() > 1e-7
    assert bool(((got - want).abs() <= 1e-5 * moved + 1e-7 * params[0].abs() + 1e-12).all())
