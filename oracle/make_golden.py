"""Generate tests/golden/*.npz by EXECUTING the reference's own code.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):   python oracle/make_golden.py

What is pinned:
  dataset.py   ProcessAudio.forward / backward / mod_phase, pcenfunc   (imported
               unmodified; the unused ``import librosa`` is stubbed, X11)
  stft_loss.py MultiResolutionSTFTLoss forward + autograd backward      (imported
               unmodified)
  network.py   layer classes, lines 9-120 (the file has a SyntaxError at :134,
               so its first 120 lines are exec'd as they lie on disk; nothing is
               copied into this repo)
  phm.py       formula :34-44, executed with the two misspelt names aliased (X7)
  dataset.py   DataAugment (gain + low/high-pass biquads of torchaudio) with seeded draws
  cos_loss.py  CosSimLoss on one row (imported unmodified)
  util.py      LinearWarmupCosineDecay, lines 81-156 (the file has a SyntaxError
               further down, X8; this slice is exec'd as it lies on disk)
Everything is small (a few hundred kB in total) and committed.
"""
import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_reference():
    sys.modules.setdefault("librosa", types.ModuleType("librosa"))
    sys.path.insert(0, REF)
    import dataset as ref_dataset           # noqa
    import stft_loss as ref_stft_loss       # noqa
    sys.path.pop(0)
    src = open(os.path.join(REF, "network.py")).read().split("class TRUNet")[0]
    src = src.replace("from phm import PhaseAwareMask", "")
    ns = {}
    exec(compile(src, "reference/network.py[:120]", "exec"), ns)
    return ref_dataset, ref_stft_loss, ns


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_dataset, ref_stft_loss, net_ns = _import_reference()
    torch.manual_seed(1234)

    # ---- front end / back end (dataset.py) --------------------------------
    n = 128 * 40
    t = torch.arange(n) / 16000.0
    audio = (0.1 * torch.randn(n) * (0.5 + 0.5 * torch.sin(2 * np.pi * 5 * t))
             + 0.05 * torch.sin(2 * np.pi * 440 * t)).float()
    dp = ref_dataset.ProcessAudio()
    feats3 = dp(audio.view(1, 1, n).clone())                       # (T', 3, F)
    spec = torch.stft(audio.view(1, n), n_fft=512, hop_length=128, normalized=False,
                      return_complex=True)
    mag = spec.abs()                                                # (1, F, T')
    pcen_ref = ref_dataset.pcenfunc(mag.transpose(1, 2).clone(), training=True)  # (1,T',F)
    back = dp.backward(feats3.clone())                              # (1, n)
    m = torch.rand(257, 9) * 2.4 - 1.2
    s = torch.randn(257, 9)
    c = torch.randn(257, 9)
    modp = dp.mod_phase(m, s, c)                                    # (1, F, T) c64
    ist_in = torch.randn(1, 257, 12, dtype=torch.complex64)
    ist = torch.istft(ist_in, n_fft=512, hop_length=128, normalized=False)
    np.savez_compressed(
        os.path.join(OUT, "dataset_ref.npz"),
        audio=audio.numpy(), feats3=feats3.numpy(), pcen=pcen_ref.numpy(),
        backward=back.numpy(), mp_m=m.numpy(), mp_s=s.numpy(), mp_c=c.numpy(),
        mp_re=modp.real.numpy(), mp_im=modp.imag.numpy(),
        ist_re=ist_in.real.numpy(), ist_im=ist_in.imag.numpy(), ist_out=ist.numpy())

    # ---- multi-resolution STFT loss (stft_loss.py) -------------------------
    x = (0.1 * torch.randn(2, 6000)).requires_grad_(True)
    y = 0.1 * torch.randn(2, 6000)
    mr = ref_stft_loss.MultiResolutionSTFTLoss(
        fft_sizes=[512, 1024, 2048], hop_sizes=[50, 120, 240],
        win_lengths=[240, 600, 1200], sc_lambda=0.5, mag_lambda=0.5, band="full")
    sc, mg = mr(x, y)
    (sc + mg).backward()
    np.savez_compressed(
        os.path.join(OUT, "stft_loss_ref.npz"),
        x=x.detach().numpy(), y=y.numpy(), sc=sc.detach().numpy(), mag=mg.detach().numpy(),
        grad_x=x.grad.numpy())

    # ---- layer classes (network.py:9-120) ----------------------------------
    blocks = {}
    specs = [("StandardConv1d", (4, 64, 5, 2), (2, 4, 257), None),
             ("DepthwiseSeparableConv1d", (64, 128, 3, 1), (2, 64, 128), None),
             ("DepthwiseSeparableConv1d", (128, 128, 5, 2), (2, 128, 128), None),
             ("DepthwiseSeparableConv1d", (128, 128, 3, 2), (2, 128, 32), None),
             ("GRUBlock", (128, 64, 64, True), (2, 16, 128), None),
             ("GRUBlock", (64, 128, 64, False), (16, 7, 64), None),
             ("FirstTrCNN", (64, 64, 3, 2), (2, 64, 16), None),
             ("TrCNN", (192, 64, 5, 2), (2, 64, 31), (2, 128, 32)),
             ("TrCNN", (192, 64, 3, 1), (2, 64, 65), (2, 128, 64)),
             ("TrCNN", (192, 64, 5, 2), (2, 64, 66), (2, 128, 64)),
             ("TrCNN", (192, 64, 3, 1), (2, 64, 129), (2, 128, 128)),
             ("LastTrCNN", (128, 8, 5, 2), (2, 64, 130), (2, 64, 128))]
    for i, (name, args, shp1, shp2) in enumerate(specs):
        torch.manual_seed(100 + i)
        mod = net_ns[name](*args)
        for mm in mod.modules():
            if isinstance(mm, torch.nn.BatchNorm1d):
                with torch.no_grad():
                    mm.weight.uniform_(0.5, 1.5)
                    mm.bias.normal_(0, 0.1)
                    mm.running_mean.normal_(0, 0.1)
                    mm.running_var.uniform_(0.5, 1.5)
        x1 = torch.randn(*shp1)
        inputs = (x1,) if shp2 is None else (x1, torch.randn(*shp2))
        sd = {k: v.clone() for k, v in mod.state_dict().items()}
        mod.eval()
        y_eval = mod(*[t.clone() for t in inputs])
        mod.train()
        y_train = mod(*[t.clone() for t in inputs])
        pre = f"b{i}_"
        blocks[pre + "name"] = np.array(name)
        blocks[pre + "args"] = np.array(args, dtype=np.int64)
        for j, t_in in enumerate(inputs):
            blocks[pre + f"in{j}"] = t_in.numpy()
        blocks[pre + "eval"] = y_eval.detach().numpy()
        blocks[pre + "train"] = y_train.detach().numpy()
        for k, v in sd.items():
            blocks[pre + "sd_" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "network_blocks_ref.npz"), **blocks)

    # ---- phase-aware mask (phm.py:34-44) ----------------------------------
    mix = torch.randn(1, 257, 7, dtype=torch.complex64)
    est = torch.randn(1, 257, 7, dtype=torch.complex64)
    src = open(os.path.join(REF, "phm.py")).read()
    body = src.split("def forward(self, mixture, estimated):")[1]
    lines = [ln.strip() for ln in body.splitlines() if "=" in ln and not ln.strip().startswith("#")]
    env = {"torch": torch, "mixture": mix, "estimated": est,
           "self": types.SimpleNamespace(beta=0.5)}
    for ln in lines:
        if ln.startswith("soft_mask"):
            env["phase_mix"] = env["phase_mixture"]      # X7: the two misspelt names
            env["phase_est"] = env["phase_estimated"]
        exec(ln, env)
    np.savez_compressed(os.path.join(OUT, "phm_ref.npz"),
                        mix_re=mix.real.numpy(), mix_im=mix.imag.numpy(),
                        est_re=est.real.numpy(), est_im=est.imag.numpy(),
                        out=env["estimated"].numpy())
    # ---- learning-rate schedule (util.py:81-156) ---------------------------
    # util.py does not parse as a whole (X8); the schedule's own lines do, so that slice is executed as it lies on disk.
    usrc = open(os.path.join(REF, "util.py")).read()
    sl = usrc[usrc.index("def anneal_linear"):usrc.index("def std_normal")]
    uns = {}
    exec(compile("from math import cos, pi\n" + sl, "reference/util.py[81:156]", "exec"), uns)
    cases = []
    for lr_max, n_iter, it0, warm, steps in ((4e-4, 1000, 0, 0.05, 2100), (4e-4, 1000, 30, 0.05, 1200),
                                             (4e-4, 1000, 500, 0.05, 700), (1e-3, 77, 0, 0.3, 200), (2e-4, 40, 39, 0.3, 90)):
        opt = types.SimpleNamespace(param_groups=[{"lr": None}])
        sch = uns["LinearWarmupCosineDecay"](opt, lr_max=lr_max, n_iter=n_iter, iteration=it0, divider=25,
                                             warmup_proportion=warm)
        lrs = [sch.step() for _ in range(steps)]
        assert opt.param_groups[0]["lr"] == lrs[-1]
        cases.append(dict(lr_max=lr_max, n_iter=n_iter, iteration=it0, warmup_proportion=warm,
                          lr_hex=[float(v).hex() for v in lrs]))
    with open(os.path.join(OUT, "lr_schedule_ref.json"), "w") as fh:
        json.dump(cases, fh)

    # ---- augmentation (dataset.py:79-126) ---------------------------------
    import random
    aug = ref_dataset.DataAugment()
    g = torch.Generator().manual_seed(99)
    na = 3 * 16128 + 777                                  # a few kernel tiles plus a ragged tail
    tt = torch.arange(na) / 16000.0
    noise = (0.6 * torch.randn(1, na, generator=g) * (0.3 + 0.7 * torch.sin(2 * np.pi * 1.5 * tt) ** 2)).float()
    noise[:, 5000:5400] *= 8.0                            # loud burst: the per-biquad clamp to [-1, 1] is active
    rec = {"noise": noise.numpy()}
    for k in range(4):
        random.seed(100 + k)
        out = aug(noise.clone())
        random.seed(100 + k)                              # replay the three draws (order: low-pass, high-pass, gain)
        lp = random.choice(aug.lp_freqs); hp = random.choice(aug.hp_freqs); gain = random.choice(aug.gains)
        rec["out%d" % k] = out.numpy()
        rec["par%d" % k] = np.array([float(gain), float(lp), float(hp)], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "augment_ref.npz"), **rec)

    # ---- cosine-similarity loss (cos_loss.py, imported unmodified; it only runs for one row) ----
    sys.path.insert(0, REF)
    import cos_loss as ref_cos_loss         # noqa
    sys.path.pop(0)
    gc = torch.Generator().manual_seed(77)
    cy = torch.randn(1, 4500, generator=gc) * 0.2
    cx = cy + 0.1 * torch.randn(1, 4500, generator=gc)
    cout = ref_cos_loss.CosSimLoss()(cx, cy)
    np.savez_compressed(os.path.join(OUT, "cos_loss_ref.npz"), x=cx.numpy(), y=cy.numpy(), out=np.float64(float(cout)))

    print("golden fixtures written to", os.path.normpath(OUT))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
