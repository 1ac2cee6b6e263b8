"""GPU parity of the batch assembly: dataset.DataAugment / dataset.assemble_batch (tru_augment_fwd,
tru_mix_crop) against the reference's DataAugment outputs (tests/golden/augment_ref.npz, made by executing
dataset.py:79-126 with seeded draws) and against the oracle (torchaudio on the CPU) on seeded inputs.

Tolerance: 1e-5 absolute on signals bounded by the biquads' clamp to [-1, 1] (north_star: <= 1e-4 on features).  The
kernel filters 63-sample chunks in parallel and stitches them by superposition, so it rounds differently from the
sequential fp32 loop of torchaudio (measured difference ~1.5e-6); the crop / copy part is bit-exact."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import tru_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "augment_ref.npz"))


def test_augment_matches_reference_golden_with_the_reference_draws(golden_dir):
    """Seeding ``random`` like the generator did must pick the same (low-pass, high-pass, gain) and give the same audio."""
    from tinyrecurrentunet_b200 import dataset
    g = golden(golden_dir)
    noise = torch.from_numpy(g["noise"]).cuda()
    aug = dataset.DataAugment()
    for k in range(4):
        random.seed(100 + k)
        out = aug(noise)
        want = torch.from_numpy(g["out%d" % k])
        assert out.shape == want.shape
        assert (out.cpu() - want).abs().max().item() <= TOL, k
    # the batched call with explicit parameters gives the same rows
    params = []
    for k in range(4):
        random.seed(100 + k)
        params.append(aug.sample_params())
        gain, lp, hp = g["par%d" % k]
        assert (float(params[-1][0]), float(params[-1][1]), float(params[-1][2])) == (gain, lp, hp)
    out = aug(noise.repeat(4, 1), params)
    for k in range(4):
        assert (out[k].cpu() - torch.from_numpy(g["out%d" % k][0])).abs().max().item() <= TOL


@pytest.mark.parametrize("n", [5, 62, 63, 64, 1000, 16127, 16128, 16129, 64000])
def test_augment_matches_oracle_ragged_lengths_and_clamping(n):
    """Lengths around the chunk (63) and tile (16,128) sizes; rows 1 and 2 are loud enough for both clamps to act."""
    from tinyrecurrentunet_b200 import dataset
    g = torch.Generator().manual_seed(n)
    noise = torch.randn(3, n, generator=g) * torch.tensor([[0.2], [2.0], [6.0]])
    aug = dataset.DataAugment()
    params = [(aug.gains[7], aug.lp_freqs[0], aug.hp_freqs[0]), (aug.gains[200], aug.lp_freqs[29], aug.hp_freqs[7]),
              (-5.5, 8150.0, 999.0)]                              # python numbers are accepted too
    out = aug(noise.cuda(), params).cpu()
    for b in range(3):
        want = O.augment(noise[b:b + 1], *params[b])[0]
        assert (out[b] - want).abs().max().item() <= TOL, b
    if n >= 1000:
        assert (out[2].abs() >= 1.0).sum().item() > 0             # the clamp really was exercised
    # 1-D and (1,N) call shapes of the reference
    one = aug(noise[0].cuda(), params[:1])
    assert one.shape == (n,) and torch.equal(one.cpu(), out[0])


def test_assemble_batch_matches_oracle():
    from tinyrecurrentunet_b200 import dataset
    g = torch.Generator().manual_seed(5)
    B, n_clean, crop = 4, 70000, 64000
    clean = torch.randn(B, n_clean, generator=g) * 0.1
    aug = dataset.DataAugment()
    random.seed(3)
    params = [aug.sample_params() for _ in range(B)]
    # (a) the reference's case: noise exactly as long as the crop, random clean start drawn like dataset.py:371
    noise = torch.randn(B, crop, generator=g) * 0.5
    np.random.seed(11)
    starts = [int(np.random.randint(low=0, high=n_clean - crop + 1)) for _ in range(B)]
    np.random.seed(11)
    c, y = dataset.assemble_batch(clean.cuda(), noise.cuda(), crop, aug, params)
    c_ref, y_ref = O.assemble_batch(clean, noise, params, starts, [0] * B, crop)
    assert torch.equal(c.cpu(), c_ref)
    assert (y.cpu() - y_ref).abs().max().item() <= TOL
    # (b) short noise rows that wrap, explicit offsets
    noise = torch.randn(B, 30000, generator=g) * 0.5
    ns = [0, 29999, 12345, 7]
    c, y = dataset.assemble_batch(clean.cuda(), noise.cuda(), crop, aug, params, clean_start=starts, noise_start=ns)
    c_ref, y_ref = O.assemble_batch(clean, noise, params, starts, ns, crop)
    assert torch.equal(c.cpu(), c_ref)
    assert (y.cpu() - y_ref).abs().max().item() <= TOL


def test_full_size_batch_is_deterministic_and_feeds_the_front_end():
    """BASELINE.json's batch (32 x 4 s): identical rows with identical parameters come out bit-identical from 32
    different CTAs, equal the oracle, and the assembled pair goes straight into loss_fn's front end."""
    from tinyrecurrentunet_b200 import dataset, ops
    clean, noise = O.synthetic_batch(1, n=64000)
    aug = dataset.DataAugment()
    params = [(aug.gains[100], aug.lp_freqs[10], aug.hp_freqs[3])] * 32
    noise32 = (noise * 10).repeat(32, 1).cuda()
    c, y = dataset.assemble_batch(clean.repeat(32, 1).cuda(), noise32, 64000, aug, params, clean_start=[0] * 32)
    assert all(torch.equal(y[0], y[b]) for b in range(1, 32))
    c_ref, y_ref = O.assemble_batch(clean, noise * 10, params[:1], [0], [0], 64000)
    assert (y[0].cpu() - y_ref[0]).abs().max().item() <= TOL
    feats = ops.frontend(y)
    assert feats.shape == (32, 501, 4, 257) and torch.isfinite(feats).all()


def test_augment_c_abi_argument_checks():
    from tinyrecurrentunet_b200 import _lib as L, dataset
    x = torch.zeros(2, 100, device="cuda")
    coef = dataset.DataAugment().coefficients([(-6.0, 8000.0, 1000.0)] * 2).cuda()
    out = torch.empty_like(x)
    st = L.stream_ptr()
    assert L.lib.tru_augment_fwd(2, 100, x.data_ptr(), coef.data_ptr(), out.data_ptr(), st) == 0
    assert L.lib.tru_augment_fwd(0, 100, x.data_ptr(), coef.data_ptr(), out.data_ptr(), st) == -1
    assert L.lib.tru_augment_fwd(2, 100, None, coef.data_ptr(), out.data_ptr(), st) == -1
    assert L.lib.tru_mix_crop(2, 100, 100, 101, x.data_ptr(), x.data_ptr(), None, None, out.data_ptr(), out.data_ptr(), st) == -1
    with pytest.raises(L.TruError):
        dataset.assemble_batch(x, x, 101)
    with pytest.raises(L.TruError):
        dataset.assemble_batch(x, x, 50, clean_start=[0, 51])
    with pytest.raises(L.TruError):
        dataset.DataAugment()(torch.zeros(2, 100))                # CPU tensor: no fallback
    torch.cuda.synchronize()
