"""CPU-only: the C-ABI library loads and exports every symbol include/tru_b200.h declares;
host-side logic (module tree, state-dict keys, parameter order, error paths)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    import importlib.util
    spec = importlib.util.spec_from_file_location("tru_build", os.path.join(ROOT, "tinyrecurrentunet_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build_library()


def test_library_exports_every_declared_symbol():
    lib_path = _ensure_built()
    lib = ctypes.CDLL(lib_path)
    header = open(os.path.join(ROOT, "include", "tru_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(tru_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    for name in sorted(declared):
        assert hasattr(lib, name), "missing export: %s" % name
    lib.tru_abi_version.restype = ctypes.c_int
    assert lib.tru_abi_version() == 1
    # the test / tuning entry points are declared too (include/tru_b200_debug.h), and nothing is exported undeclared
    dbg = open(os.path.join(ROOT, "include", "tru_b200_debug.h")).read()
    dbg = re.sub(r"/\*.*?\*/", "", dbg, flags=re.S)
    declared_dbg = set(re.findall(r"\b(tru_[a-z0-9_]+)\s*\(", dbg))
    for name in sorted(declared_dbg):
        assert hasattr(lib, name), "missing export: %s" % name
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in syms.splitlines() if ln.split() and ln.split()[-1].startswith("tru_")}
    assert exported == declared | declared_dbg, (exported ^ (declared | declared_dbg))


def test_binding_table_matches_header():
    from tinyrecurrentunet_b200 import _lib
    header = open(os.path.join(ROOT, "include", "tru_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(tru_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)


def test_no_gpu_means_loud_failure_not_fallback():
    """Without a CUDA device every hot-path op must raise (no CPU path)."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tinyrecurrentunet_b200 import _lib, dataset, network, ops, stft_loss
    with pytest.raises(Exception):
        ops.frontend(torch.zeros(1, 4096))
    with pytest.raises(Exception):
        network.TRUNet()(torch.zeros(3, 4, 257))
    with pytest.raises(Exception):
        dataset.ProcessAudio()(torch.zeros(1, 1, 4096))
    with pytest.raises(Exception):
        stft_loss.MultiResolutionSTFTLoss()(torch.zeros(1, 4096), torch.zeros(1, 4096))
    assert _lib.lib.tru_init() != 0            # no device: error code, message set, no crash
    assert _lib.lib.tru_last_error()


def test_module_tree_matches_reference_schema():
    """State-dict keys / shapes of SURVEY Appendix C; the 7 reference kwargs are accepted."""
    from oracle import tru_oracle as O
    from tinyrecurrentunet_b200 import network
    net = network.TRUNet(input_size=3, channels_input=64, channels_output=3, channels_hidden=128,
                         kernel_sizes=[5, 3], strides=[2, 1], tr_channels_input=192)
    ref = O.TRUNet()
    sd, sr = net.state_dict(), ref.state_dict()
    assert list(sd) == list(sr) and len(sd) == 177
    assert all(sd[k].shape == sr[k].shape for k in sd)
    assert sum(p.numel() for p in net.parameters()) == 381472
    assert [k for k, _ in net.named_parameters()] == network.PARAM_ORDER
    bns = dict(net.named_modules())
    assert all(isinstance(bns[n], torch.nn.BatchNorm1d) for n in network.BN_ORDER)
    net.load_state_dict(sr)                      # checkpoints are interchangeable
    with pytest.raises(NotImplementedError):
        network.TRUNet(in_channels=3)


def test_reference_api_names_exist():
    from tinyrecurrentunet_b200 import cos_loss, dataset, distributed, network, optim, phm, stft_loss, util
    for mod, names in ((network, ["StandardConv1d", "DepthwiseSeparableConv1d", "GRUBlock", "FirstTrCNN", "TrCNN",
                                  "LastTrCNN", "TRUNet"]),
                       (phm, ["PhaseAwareMask"]),
                       (stft_loss, ["stft", "SpectralConvergenceLoss", "LogSTFTMagnitudeLoss", "STFTLoss",
                                    "MultiResolutionSTFTLoss"]),
                       (dataset, ["ProcessAudio", "pcenfunc", "unwrap", "diff", "DataAugment"]),
                       (util, ["loss_fn", "sampling", "find_max_epoch", "LinearWarmupCosineDecay", "denoise", "denoise_host_batches",
                               "StreamingDenoiser", "GraphedDenoise", "stream_host_frames", "CudaPrefetcher"]),
                       (cos_loss, ["CosSimLoss"]),
                       (optim, ["FlatAdamW"]),
                       (distributed, ["init_distributed", "apply_gradient_allreduce", "reduce_tensor"])):
        for n in names:
            assert hasattr(mod, n), (mod.__name__, n)


def test_cpu_helpers_match_oracle():
    """The small elementwise helpers kept for API compatibility (not the hot path)."""
    from oracle import tru_oracle as O
    from tinyrecurrentunet_b200 import dataset, phm
    dp = dataset.ProcessAudio()
    m, s, c = torch.rand(257, 5) * 2.4 - 1.2, torch.randn(257, 5), torch.randn(257, 5)
    assert torch.equal(dp.mod_phase(m, s, c)[0], O.mod_phase(m, s, c))
    x = torch.rand(1, 9, 257)
    torch.testing.assert_close(dataset.pcenfunc(x.clone()), O.pcen(x)[0])
    a = torch.randn(1, 257, 4, dtype=torch.complex64)
    b = torch.randn(1, 257, 4, dtype=torch.complex64)
    torch.testing.assert_close(phm.PhaseAwareMask(0.5)(a, b), O.phase_aware_mask(a, b, 0.5))
    with pytest.raises(NotImplementedError):
        dataset.ProcessAudio(n_fft=1024)


def test_checkpoint_round_trip_in_the_reference_format(tmp_path):
    """train.py:155-162 / :70-95: a checkpoint {iter, model_state_dict, optimizer_state_dict, training_time_seconds} written by
    the reference layer list (the oracle's TRUNet is network.py:9-150 verbatim) loads into the drop-in module and back, key for
    key, and util.find_max_epoch-style resume (largest <iter>.pkl) picks the right file."""
    import torch
    from oracle import tru_oracle as O
    from tinyrecurrentunet_b200 import network
    torch.manual_seed(3)
    ref = O.randomize_bn(O.TRUNet())
    opt = torch.optim.AdamW(ref.parameters(), lr=4e-4)
    ckdir = tmp_path / "checkpoint"
    ckdir.mkdir()
    for it in (5000, 10000):
        torch.save({"iter": it, "model_state_dict": ref.state_dict(), "optimizer_state_dict": opt.state_dict(),
                    "training_time_seconds": 12.5}, str(ckdir / ("%d.pkl" % it)))
    newest = max(int(f.name[:-4]) for f in ckdir.iterdir() if f.name.endswith(".pkl"))
    assert newest == 10000
    ck = torch.load(str(ckdir / ("%d.pkl" % newest)), map_location="cpu")
    net = network.TRUNet(3, 64, 3, 128, [5, 3], [2, 1], 192)
    missing = net.load_state_dict(ck["model_state_dict"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    sd = net.state_dict()
    assert list(sd.keys()) == list(ref.state_dict().keys())
    for k, v in ref.state_dict().items():
        assert torch.equal(sd[k], v), k
    # and back: a checkpoint written from the drop-in module restores the reference model and its optimizer state
    opt2 = torch.optim.AdamW(net.parameters(), lr=4e-4)
    opt2.load_state_dict(ck["optimizer_state_dict"])
    torch.save({"iter": 15000, "model_state_dict": net.state_dict(), "optimizer_state_dict": opt2.state_dict(),
                "training_time_seconds": 20.0}, str(ckdir / "15000.pkl"))
    back = O.TRUNet()
    back.load_state_dict(torch.load(str(ckdir / "15000.pkl"), map_location="cpu")["model_state_dict"], strict=True)
    for (k, a), (_, b) in zip(back.state_dict().items(), ref.state_dict().items()):
        assert torch.equal(a, b), k


def test_lr_schedule_matches_reference_golden(golden_dir):
    """util.LinearWarmupCosineDecay vs the reference's own class (util.py:109-156, executed by oracle/make_golden.py):
    every learning rate bit-identical, including resumed starts and the wrap-around after n_iter steps."""
    import json
    import types
    from tinyrecurrentunet_b200 import util
    cases = json.load(open(os.path.join(golden_dir, "lr_schedule_ref.json")))
    assert len(cases) >= 5
    for c in cases:
        opt = types.SimpleNamespace(param_groups=[{"lr": None}, {"lr": None}])
        sch = util.LinearWarmupCosineDecay(opt, lr_max=c["lr_max"], n_iter=c["n_iter"], iteration=c["iteration"],
                                           divider=25, warmup_proportion=c["warmup_proportion"])
        for i, want in enumerate(c["lr_hex"]):
            got = sch.step()
            assert got == float.fromhex(want), (c["n_iter"], c["iteration"], i, got, float.fromhex(want))
            assert opt.param_groups[0]["lr"] == got and opt.param_groups[1]["lr"] == got


def test_flat_adamw_keeps_the_adamw_interface_and_has_no_cpu_path():
    """optim.FlatAdamW is a torch.optim.AdamW (same defaults / param_groups keys, so train.py:68 and the checkpoint
    code work unchanged); stepping CPU parameters must fail loudly, not fall back."""
    from tinyrecurrentunet_b200 import _lib, optim
    p = torch.nn.Parameter(torch.randn(5, 3))
    opt = optim.FlatAdamW([p], lr=4e-4)
    ref = torch.optim.AdamW([torch.nn.Parameter(torch.randn(5, 3))], lr=4e-4)
    assert isinstance(opt, torch.optim.AdamW)
    assert set(opt.param_groups[0]) == set(ref.param_groups[0])
    for k in ("lr", "betas", "eps", "weight_decay"):
        assert opt.param_groups[0][k] == ref.param_groups[0][k]
    p.grad = torch.randn(5, 3)
    with pytest.raises(_lib.TruError):
        opt.step()
    with pytest.raises(_lib.TruError):
        optim.FlatAdamW([{"params": [p]}, {"params": [torch.nn.Parameter(torch.zeros(2))]}])
    with pytest.raises(_lib.TruError):
        optim.FlatAdamW([p], amsgrad=True)


def test_checkpoint_helpers_follow_train_py(tmp_path):
    """util.find_max_epoch (util.py:30-49) and save / load in the layout of train.py:155-161 / :70-95."""
    from oracle import tru_oracle as O
    from tinyrecurrentunet_b200 import network, util
    d = tmp_path / "ck"
    d.mkdir()
    assert util.find_max_epoch(str(d)) == -1
    for name in ("7.txt", "abc.pkl", ".pkl", "3000.pkl.bak", "-5.pkl"):
        (d / name).write_text("")
    assert util.find_max_epoch(str(d)) == -1
    torch.manual_seed(4)
    ref = O.randomize_bn(O.TRUNet())
    opt = torch.optim.AdamW(ref.parameters(), lr=4e-4)
    for it in (10, 200):
        path = util.save_checkpoint(str(d), it, ref, opt, 33.9)
        assert path.endswith("%d.pkl" % it)
    assert util.find_max_epoch(str(d)) == 200
    ck = torch.load(str(d / "200.pkl"), map_location="cpu")
    assert set(ck) == {"iter", "model_state_dict", "optimizer_state_dict", "training_time_seconds"}
    assert ck["iter"] == 200 and ck["training_time_seconds"] == 33
    net = network.TRUNet()
    it, secs = util.load_checkpoint(str(d), "max", net, torch.optim.AdamW(net.parameters(), lr=4e-4))
    assert (it, secs) == (200, 33)
    for k, v in ref.state_dict().items():
        assert torch.equal(net.state_dict()[k], v), k
    assert util.load_checkpoint(str(d), 11, net) == (-1, 0)
    assert util.load_checkpoint(str(tmp_path), "max", net) == (-1, 0)


def test_fold_batchnorm_export_keeps_the_eval_outputs():
    """util.fold_batchnorm: same keys, BatchNorm entries become the identity, and the reference layer list (oracle TRUNet)
    loaded with the folded weights reproduces the eval-mode output of the original (<= 1e-5 of the output scale)."""
    from oracle import tru_oracle as O
    from tinyrecurrentunet_b200 import util
    torch.manual_seed(8)
    ref = O.randomize_bn(O.TRUNet(), 8).eval()
    x = O.frontend(O.synthetic_batch(1, n=128 * 12)[1])
    with torch.no_grad():
        y = ref(x)
    sd = ref.state_dict()
    folded = util.fold_batchnorm(sd)
    assert list(folded.keys()) == list(sd.keys())
    n_bn = 0
    for k, v in folded.items():
        if k.endswith("running_mean"):
            n_bn += 1
            assert torch.count_nonzero(v) == 0 and torch.count_nonzero(folded[k[:-12] + "bias"]) == 0
            assert torch.all(folded[k[:-12] + "weight"] == 1)
    assert n_bn == 23
    ref2 = O.TRUNet()
    ref2.load_state_dict(folded)
    ref2.eval()
    with torch.no_grad():
        y2 = ref2(x)
    assert ((y2 - y).abs().max() / y.abs().max()).item() <= 1e-5
    assert not torch.equal(folded["encoder.1.DepthwiseSeparableConv1d.0.weight"], sd["encoder.1.DepthwiseSeparableConv1d.0.weight"])


def test_cuda_graph_threshold_is_the_first_frame_with_a_full_overlap_add_window():
    """StreamingDenoiser replays its step as a CUDA graph from frame GRAPH_FROM on.  The only step argument that depends
    on the frame index is the back end's 1 / (number of frames overlapping the emitted block) (backend.cu,
    backend_step_kernel; block frame_index - 2 of the centre-framed iSTFT of dataset.py:293-296): it must be constant
    from GRAPH_FROM on and not before.  Checked against the overlap-add envelope torch.istft itself divides by."""
    from tinyrecurrentunet_b200 import util
    T = 12
    spec = torch.stft(torch.ones(128 * (T - 1)), 512, hop_length=128, window=torch.ones(512), center=True,
                      pad_mode="reflect", return_complex=True)
    assert spec.shape[1] == T
    frames = torch.ones(T, 512)                                         # every frame contributes 1 to each sample it covers
    env = torch.nn.functional.fold(frames.t().unsqueeze(0), (1, 128 * (T - 1) + 512), (1, 512), stride=(1, 128))[0, 0, 0]
    env = env[256:256 + 128 * (T - 1)]                                  # centre trimming
    per_block = env.view(T - 1, 128)
    assert all(len(set(b.tolist())) == 1 for b in per_block[1:T - 3]), "interior blocks have one count each"
    counts = [int(b[0]) for b in per_block]                             # frames overlapping block q (first sample)
    first_full = next(q for q, c in enumerate(counts) if c == 4)
    assert util.StreamingDenoiser.GRAPH_FROM == first_full + 2          # block q is emitted by the step of frame q + 2
    assert all(c == 4 for c in counts[first_full:T - 3])
    # the kernel's own formula for an added frame: newest - max(frame_index - 3, 0) + 1
    kernel_cnt = [fi - max(fi - 3, 0) + 1 for fi in range(T)]
    assert kernel_cnt[util.StreamingDenoiser.GRAPH_FROM - 1] != 4 and all(c == 4 for c in kernel_cnt[util.StreamingDenoiser.GRAPH_FROM:])
