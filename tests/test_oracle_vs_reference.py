"""CPU: the oracle restatement reproduces outputs of the reference's own code
(tests/golden/*.npz, written by oracle/make_golden.py which executes
/root/reference).  This is what pins the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import tru_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_frontend_channels_match_dataset_py(golden_dir):
    g = _load(golden_dir, "dataset_ref.npz")
    audio = torch.from_numpy(g["audio"]).view(1, -1)
    feats = O.frontend(audio)[0]                      # (T', 4, F)
    ref3 = torch.from_numpy(g["feats3"])              # (T', 3, F): logmag, sin, cos
    assert feats.shape == (1 + audio.shape[1] // 128, 4, 257)
    assert torch.equal(feats[:, 0], ref3[:, 0])
    assert torch.equal(feats[:, 2], ref3[:, 1])
    assert torch.equal(feats[:, 3], ref3[:, 2])
    # ch1 = pcenfunc(|X|) (dataset.py:56-76)
    torch.testing.assert_close(feats[:, 1], torch.from_numpy(g["pcen"])[0], rtol=1e-6, atol=1e-6)


def test_framing_counts():
    # SURVEY section 4 known answers: T' = 1 + N//128
    for n, t in ((96000, 751), (64000, 501), (160000, 1251)):
        assert 1 + n // O.HOP == t


def test_mod_phase_and_istft_match_dataset_py(golden_dir):
    g = _load(golden_dir, "dataset_ref.npz")
    sp = O.mod_phase(torch.from_numpy(g["mp_m"]), torch.from_numpy(g["mp_s"]),
                     torch.from_numpy(g["mp_c"]))
    assert torch.equal(sp.real, torch.from_numpy(g["mp_re"])[0])
    assert torch.equal(sp.imag, torch.from_numpy(g["mp_im"])[0])
    z = torch.complex(torch.from_numpy(g["ist_re"]), torch.from_numpy(g["ist_im"]))
    assert torch.equal(O.istft_rect(z), torch.from_numpy(g["ist_out"]))
    # ProcessAudio.backward(forward(x)) (dataset.py:275-298)
    audio = torch.from_numpy(g["audio"]).view(1, -1)
    f4 = O.frontend(audio)
    rec = O.features_to_audio(f4[:, :, [0, 2, 3]])
    torch.testing.assert_close(rec, torch.from_numpy(g["backward"]), rtol=0, atol=1e-6)


def test_phase_aware_mask_matches_phm_py(golden_dir):
    g = _load(golden_dir, "phm_ref.npz")
    mix = torch.complex(torch.from_numpy(g["mix_re"]), torch.from_numpy(g["mix_im"]))
    est = torch.complex(torch.from_numpy(g["est_re"]), torch.from_numpy(g["est_im"]))
    assert torch.equal(O.phase_aware_mask(mix, est, 0.5), torch.from_numpy(g["out"]))


def test_mrstft_loss_matches_stft_loss_py(golden_dir):
    g = _load(golden_dir, "stft_loss_ref.npz")
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    y = torch.from_numpy(g["y"])
    sc, mg = O.mrstft_loss(x, y)
    (sc + mg).backward()
    torch.testing.assert_close(sc.detach(), torch.from_numpy(g["sc"]), rtol=1e-6, atol=0)
    torch.testing.assert_close(mg.detach(), torch.from_numpy(g["mag"]), rtol=1e-6, atol=0)
    torch.testing.assert_close(x.grad, torch.from_numpy(g["grad_x"]), rtol=1e-5, atol=1e-9)


def test_layer_blocks_match_network_py(golden_dir):
    g = _load(golden_dir, "network_blocks_ref.npz")
    i = 0
    while f"b{i}_name" in g:
        name = str(g[f"b{i}_name"])
        args = [int(a) for a in g[f"b{i}_args"]]
        if name == "GRUBlock":
            args[3] = bool(args[3])
        mod = getattr(O, name)(*args)
        pre = f"b{i}_sd_"
        sd = {k[len(pre):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(pre)}
        assert set(sd) == set(mod.state_dict())          # same state-dict keys
        mod.load_state_dict(sd)
        ins = [torch.from_numpy(g[f"b{i}_in{j}"]) for j in range(2) if f"b{i}_in{j}" in g]
        mod.eval()
        assert torch.equal(mod(*[t.clone() for t in ins]), torch.from_numpy(g[f"b{i}_eval"]))
        mod.train()
        assert torch.equal(mod(*[t.clone() for t in ins]), torch.from_numpy(g[f"b{i}_train"]))
        i += 1
    assert i == 12


def test_parameter_counts_match_figure():
    # docs/net.jpg: 1D-CNN 81,344; FGRU 82,880; TGRU 82,880
    net = O.TRUNet(3, 64, 3, 128, [5, 3], [2, 1], 192)
    cnt = lambda m: sum(p.numel() for p in m.parameters())
    assert cnt(net.encoder) == 81344
    assert cnt(net.FGRU) == 82880
    assert cnt(net.TGRU) == 82880
    assert cnt(net.decoder) == 134368
    assert cnt(net) == 381472
    sd = net.state_dict()
    assert len(sd) == 177 and len(list(net.parameters())) == 108
    for k in ("encoder.0.StandardConv1d.0.weight", "encoder.3.DepthwiseSeparableConv1d.4.running_var",
              "FGRU.GRU.weight_hh_l0_reverse", "TGRU.conv.1.num_batches_tracked",
              "decoder.0.FirstTrCNN.3.weight", "decoder.2.TrCNN.0.bias", "decoder.5.LastTrCNN.3.bias"):
        assert k in sd


def test_model_shapes_batched_equals_flat_and_streaming():
    torch.manual_seed(0)
    net = O.randomize_bn(O.TRUNet()).eval()
    x = torch.randn(2, 6, 4, 257)
    with torch.no_grad():
        y = net(x)
        assert y.shape == (2, 6, 8, 257)
        # D10: batch element b in eval mode == the 3-D call on x[b]
        torch.testing.assert_close(y[1], net(x[1]), rtol=1e-5, atol=1e-5)
        # D11: streaming (one frame at a time, carried h) == offline
        h = None
        outs = []
        for t in range(6):
            o, h = net(x[:, t:t + 1], h0=h, return_state=True)
            outs.append(o)
        torch.testing.assert_close(torch.cat(outs, 1), y, rtol=1e-4, atol=1e-5)


def test_loss_fn_runs_and_is_differentiable():
    torch.manual_seed(0)
    net = O.randomize_bn(O.TRUNet()).train()
    clean, noisy = O.synthetic_batch(1, n=128 * 24)
    loss, d, den = O.loss_fn(net, clean, noisy)
    assert den.shape == clean.shape
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
    assert set(d) == {"l1", "stft_sc", "stft_mag"}


def test_augment_oracle_and_parameter_draws_match_reference(golden_dir):
    """oracle.augment (torchaudio calls in the order of dataset.py:121-125) against the reference's DataAugment output,
    and the drop-in DataAugment's seeded draws + coefficient rows (host logic only, no GPU) against the same run."""
    import random
    from tinyrecurrentunet_b200 import dataset
    g = np.load(os.path.join(golden_dir, "augment_ref.npz"))
    noise = torch.from_numpy(g["noise"])
    aug = dataset.DataAugment()
    for k in range(4):
        random.seed(100 + k)
        gain, lp, hp = aug.sample_params()
        assert (float(gain), float(lp), float(hp)) == tuple(g["par%d" % k])
        out = O.augment(noise, gain, lp, hp)
        assert (out - torch.from_numpy(g["out%d" % k])).abs().max().item() <= 1e-7
        row = aug.coefficients([(gain, lp, hp)])
        assert row.shape == (1, 59) and torch.isfinite(row).all()
        # the chunk matrix is the 63rd power of the recurrence's companion matrix
        a1, a2 = row[0, 4].double().item(), row[0, 5].double().item()
        A = np.array([[-a1, -a2], [1.0, 0.0]])
        assert np.allclose(np.linalg.matrix_power(A, 63).reshape(-1), row[0, 6:10].double().numpy(), rtol=1e-5, atol=1e-12)
        assert np.allclose(np.linalg.matrix_power(A, 63 * 8).reshape(-1), row[0, 10:14].double().numpy(), rtol=1e-5, atol=1e-12)


def test_cos_sim_loss_oracle_matches_reference(golden_dir):
    """oracle.cos_sim_loss against the reference's CosSimLoss (cos_loss.py:41-56) on the one-row input it can run."""
    g = np.load(os.path.join(golden_dir, "cos_loss_ref.npz"))
    out = O.cos_sim_loss(torch.from_numpy(g["x"]), torch.from_numpy(g["y"]))
    assert abs(out.item() - float(g["out"])) <= 1e-6 * abs(float(g["out"]))
