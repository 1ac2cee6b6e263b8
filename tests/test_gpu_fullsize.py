"""GPU parity at the sizes the benchmarks run (not toy shapes), through the C ABI vs the CPU oracle.

The small-shape tests in test_gpu_network.py take the single-wave code paths of the persistent kernels; the cases here have
enough rows for the multi-wave / k-split / multi-CTA-per-clip paths (tc_igemm_kernel, tc_wgrad_stream_kernel, the PCEN
look-back chain, the TGRU scan over a full clip) and cover BASELINE.json's configurations:

  configs[1]/[2]  4-s clips (T' = 501): per-layer forward, per-parameter gradients
  configs[3]      4096 concurrent streams: streaming == offline
  configs[4]      10-s clips (T' = 1251): front end -> eval network -> mask + iSTFT

Tolerances are north_star's: <= 1e-4 relative on features / activations / audio, <= 1e-3 on gradients."""
import pytest
import torch

from oracle import tru_oracle as O
from test_gpu_network import (FORWARD_MODES, GRAD_TOL, OUT_TOL, compare_intermediate_grads, compare_intermediates, feats_like,
                              force_relu_masks, forward_mode, gpu_relu_masks, make_pair, oracle_intermediates, rel,
                              relu_mask_mismatches)
from tinyrecurrentunet_b200 import _lib as L
from test_gpu_dsp import check_feats

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", FORWARD_MODES)
def test_forward_per_layer_at_4s_clip_length(mode):
    """every pre-BN conv output and GRU output of a B = 4, T' = 501 forward (2004 frames: 16 to 2004 row
    tiles per layer, several waves of the persistent GEMM grid) against the oracle's, layer by layer."""
    ref, net = make_pair(11)
    B, T = 4, 501
    x = feats_like(B, T, 20)
    training, skip = forward_mode(net, ref, mode)
    net._debug_keep_ws = True
    try:
        with torch.no_grad():
            y_ref, inter = oracle_intermediates(ref, x)
            y = net(x.cuda())
    finally:
        L.lib.tru_debug_set_eval_fusion(1)
    rows = compare_intermediates(net, inter, B, T, skip)
    print("\n".join("%-6s %.3e" % r for r in rows))
    bad = [r for r in rows if not r[1] <= OUT_TOL]
    assert not bad, bad
    assert rel(y, y_ref) <= OUT_TOL
    if training:
        sd_ref, sd = ref.state_dict(), net.state_dict()
        for k in sd_ref:
            if "running" in k:
                assert rel(sd[k], sd_ref[k]) <= OUT_TOL, k


def test_gradients_per_parameter_at_multi_wave_size():
    """all 108 parameter gradients and every intermediate dZ (<= 1e-3) at B = 2, T' = 126: 252 frames =
    enough rows for several tiles per CTA in the GEMM / weight-gradient kernels and a 126-step TGRU BPTT.

    Among the ~4e7 ReLU inputs of this shape ~30 sit within rounding distance of zero and take the other branch in one of the
    two forwards (counted below); with only 252 frames those flips alone move the parameter gradients by ~1e-2, which says
    nothing about the kernels (test_backward_matches_oracle sidesteps it by looking for a seed without flips - impossible at
    this size).  Here the oracle is made to differentiate the same piecewise-linear function instead: its ReLUs are replaced
    by the branch masks the CUDA forward took (force_relu_masks), which changes its forward by ~1e-6 at those ~30 elements
    and nothing else."""
    gradient_parity(2, 126, 7, 31)


@pytest.mark.parametrize("B", [10, 19])
def test_gradients_with_two_and_four_tgru_sequences_per_cta(B):
    """The TGRU recurrence kernels carry 1, 2 or 4 sequences per CTA depending on the batch (gru.cu tgru_seqs_per_cta: up to
    148 / 296 / more sequences; 16 sequences per clip).  Every other parity test runs B <= 5 (1 per CTA) or B = 32 (4 per CTA,
    compared with a replicated small batch); here B = 10 takes the 2-per-CTA and B = 19 the 4-per-CTA kernels, forward and
    BPTT, against the oracle - all intermediates and all 108 parameter gradients, same method as the test above."""
    gradient_parity(B, 5, 9, 61)


def gradient_parity(B, T, seed, xseed):
    ref, net = make_pair(seed)
    x = feats_like(B, T, xseed)
    ref.train()
    net.train()
    net._debug_keep_ws = True
    torch.manual_seed(5)
    w = torch.randn(B, T, 8, 257)
    y = net(x.cuda())
    flips, y_plain = relu_mask_mismatches(ref, net, lambda: ref(x).detach(), B, T)
    print("ReLU mask mismatches between the two plain forwards:", flips)
    assert flips <= 200, flips
    assert rel(y, y_plain) <= OUT_TOL
    ref2, _ = make_pair(seed)                               # fresh running statistics
    ref2.train()
    force_relu_masks(ref2, gpu_relu_masks(net, B, T))
    y_ref, inter = oracle_intermediates(ref2, x, keep_graph=True)
    assert rel(y, y_ref) <= OUT_TOL
    (y_ref * w).sum().backward()
    (y * w.cuda()).sum().backward()
    rows = compare_intermediate_grads(net, inter, B, T)
    print("\n".join("d%-6s %.3e" % r for r in rows))
    assert not [r for r in rows if not r[1] <= GRAD_TOL], rows
    gref = dict(ref2.named_parameters())
    gmax = max(p.grad.abs().max().item() for p in gref.values())
    zero_bias = {"encoder.%d.DepthwiseSeparableConv1d.%d.bias" % (i, j) for i in range(1, 6) for j in (0, 3)}
    zero_bias |= {"decoder.%d.%s.0.bias" % (d, c) for d, c in enumerate(["FirstTrCNN"] + ["TrCNN"] * 4 + ["LastTrCNN"])}
    zero_bias |= {"decoder.%d.%s.3.bias" % (d, c) for d, c in enumerate(["FirstTrCNN"] + ["TrCNN"] * 4)}
    zero_bias |= {"FGRU.conv.0.bias", "TGRU.conv.0.bias"}
    rows = []
    for k, p in net.named_parameters():
        assert p.grad is not None, k
        g, r = p.grad.cpu(), gref[k].grad
        if k in zero_bias:       # exactly-zero true gradient (conv bias in front of a training-mode BN): both sides hold the rounding
            # noise of a fully cancelling sum over ~10^6 elements; require it to stay below 1e-4 of the largest gradient
            rows.append((k, max(g.abs().max().item(), r.abs().max().item()) / (1e-4 * gmax) * GRAD_TOL))
            continue
        scale = max(r.abs().max().item(), 1e-3 * gmax)
        rows.append((k, (g - r).abs().max().item() / scale))
    print("\n".join("%-55s %.3e" % r for r in rows))
    bad = [r for r in rows if not r[1] <= GRAD_TOL]
    assert not bad, bad


def test_ten_second_clips_front_end_network_back_end():
    """VERDICT r01 (iii), BASELINE.json configs[4]: N = 160,000 samples -> T' = 1251 frames.  The PCEN smoother's look-back
    runs across 79 chunks per clip, the TGRU scans 1251 steps, the overlap-add covers 44 chunks.  Stage by stage on the
    oracle's own inputs (<= 1e-4 each), then the whole chain (the two front ends differ by ~2e-3 in the phase features of
    near-silent bins - check_feats - so the chained audio is compared in the L2 sense)."""
    from tinyrecurrentunet_b200 import ops, util
    B, N = 2, 160000
    ref, net = make_pair(13)
    ref.eval()
    net.eval()
    _, noisy = O.synthetic_batch(B, n=N, first=50)
    with torch.no_grad():
        feats_ref, m_ref = O.frontend(noisy, return_state=True)
        assert feats_ref.shape == (B, 1251, 4, 257)
        feats, m = ops.frontend(noisy.cuda(), return_state=True)
        check_feats(feats, feats_ref, noisy)
        assert rel(m, m_ref) <= OUT_TOL
        out_ref, h_ref = ref(feats_ref, return_state=True)
        out, h = net(feats_ref.cuda(), return_state=True)
        assert rel(out, out_ref) <= OUT_TOL
        assert rel(h, h_ref[0]) <= OUT_TOL
        audio_ref = O.backend(out_ref)
        audio = ops.mask_istft(out_ref.cuda())
        assert audio.shape == audio_ref.shape == (B, N)
        assert rel(audio, audio_ref) <= OUT_TOL
        chained, _ = util.denoise(net, noisy.cuda())
        err = ((chained.cpu().double() - audio_ref.double()).norm() / audio_ref.double().norm()).item()
        print("chained 10-s denoise, relative L2 error: %.3e" % err)
        assert err <= 5e-3, err


def test_streaming_4096_streams_equals_offline():
    """VERDICT r01 (iv), BASELINE.json configs[3]: S = 4096 concurrent streams take the batched code paths (the TGRU hidden
    projection of all 65,536 sequences as one tensor-core GEMM, 1024 CTAs in the front / back end steps).  Stage by stage on
    identical inputs, every stream, <= 1e-4:
      (1) front-end step vs offline front end (log-mag / PCEN wherever the bin is well conditioned, |X| >= 1e-3 max|X| of its
          frame: an fp32 FFT leaves ~1e-7 max|X| of noise in every bin, whatever the implementation - see check_feats);
      (2) TRUNet.step (carried TGRU state) vs the offline network on the SAME features;
      (3) back-end step (carried overlap-add) vs the offline back end on the SAME network outputs.
    Then the whole chain: per stream relative to its own level.  Two properties of the reference formula itself let a 1e-6
    perturbation through at isolated bins - near-silent bins (above) and the UN-wrapped phase difference of the mask
    (phm.py:41, SURVEY D7: atan2 jumps by 2 pi at the negative real axis) - so among ~10^7 bins a few streams differ more;
    they are counted (<= 2 %), the median stream must agree to 1e-5, and a sample is checked against the CPU oracle."""
    from tinyrecurrentunet_b200 import ops
    S, T = 4096, 10
    ref, net = make_pair(17)
    ref.eval()
    net.eval()
    g = torch.Generator().manual_seed(99)
    audio = 0.1 * torch.randn(S, 128 * (T - 1), generator=g)
    audio[::7] *= 0.01                                     # some quiet streams next to loud ones
    x = audio.cuda()
    with torch.no_grad():
        feats_off = ops.frontend(x)
        out_off = net(feats_off)
        offline = ops.mask_istft(out_off)
        xp = torch.nn.functional.pad(x.unsqueeze(1), (256, 256), mode="reflect").squeeze(1)
        mag = torch.stft(x, 512, 128, window=torch.ones(512, device="cuda"), return_complex=True).abs().transpose(1, 2)   # (S,T,F)
        well = mag >= 1e-3 * mag.amax(dim=2, keepdim=True)
        pcen = torch.zeros(S, 257, device="cuda")
        h, h2 = torch.zeros(S * 16, 128, device="cuda"), torch.zeros(S * 16, 128, device="cuda")
        ola, ola2 = torch.zeros(S, 384, device="cuda"), torch.zeros(S, 384, device="cuda")
        blocks, blocks2, e_feat, e_net = [], [], 0.0, 0.0
        for t in range(T):
            f = ops.frontend_step(xp[:, 128 * t:128 * t + 512].contiguous(), pcen)
            for ch in (0, 1):
                d = (f[:, ch] - feats_off[:, t, ch]).abs() * well[:, t]
                e_feat = max(e_feat, d.max().item() / feats_off[:, t, ch].abs().max().item())
            o, h = net.step(f, h)
            o2, h2 = net.step(feats_off[:, t].contiguous(), h2)                        # (2): same features as the offline network
            e_net = max(e_net, rel(o2, out_off[:, t]))
            blocks.append(ops.mask_istft_step(o, ola, t))
            blocks2.append(ops.mask_istft_step(out_off[:, t].contiguous(), ola2, t))   # (3): same inputs as the offline back end
        blocks.append(ops.mask_istft_step(None, ola, T, flush=True))
        blocks2.append(ops.mask_istft_step(None, ola2, T, flush=True))
        streamed, streamed2 = torch.cat(blocks[2:], dim=1), torch.cat(blocks2[2:], dim=1)
        print("front-end step %.2e  network step %.2e" % (e_feat, e_net))
        assert e_feat <= OUT_TOL and e_net <= OUT_TOL
        assert streamed.shape == offline.shape
        scale = offline.abs().amax(dim=1, keepdim=True).clamp_min(1e-12)     # per stream: quiet streams must not hide behind loud ones
        e_back = ((streamed2 - offline).abs() / scale).max().item()
        print("back-end step %.2e" % e_back)
        assert e_back <= OUT_TOL
        err = ((streamed - offline).abs() / scale).amax(dim=1)
        bad = int((err > OUT_TOL).sum())
        print("chained: streams beyond 1e-4: %d of %d (median %.2e, worst %.2e)" % (bad, S, err.median().item(), err.max().item()))
        assert err.median().item() <= 1e-5 and bad <= S // 50, (bad, err.median().item())
        pick = [s_ for s_ in (1, 7, 1023, 2048, 3000, 4095) if err[s_] <= OUT_TOL]
        den_ref = O.backend(ref(O.frontend(audio[pick])))
    for j, s_ in enumerate(pick):
        e = ((streamed[s_].cpu().double() - den_ref[j].double()).norm() / den_ref[j].double().norm()).item()
        assert e <= 5e-3, (s_, e)                           # chained front ends: see test_ten_second_clips_...


def _train_grads(net, mr, clean, noisy):
    from tinyrecurrentunet_b200 import util
    for p in net.parameters():
        p.grad = None
    loss, _ = util.loss_fn(net, (clean, noisy), ell_p=1, ell_p_lambda=1, stft_lambda=1, mrstftloss=mr)
    loss.backward()
    torch.cuda.synchronize()
    return loss.detach().clone(), [p.grad.detach().clone() for p in net.parameters()]


def test_two_identical_training_steps_give_identical_gradients():
    """run-to-run reproducibility of loss and gradients of a B = 8 x 4-s training step.  The overlap-adds and
    the weight-gradient reduction are ordered; BatchNorm statistics are fp64 atomic sums of fp32 partials (the order can move
    the fp64 sum by ~1e-16 relative, i.e. an fp32 coefficient by at most one ulp once in a while) and a few fp32 atomics remain
    (overlap-add of the loss gradient: 4e-7 run to run on d loss / d network output; bias / depthwise weight-gradient column sums), so
    the bound is 1e-5 of each tensor's scale (measured 1.8e-6; 100 times inside the 1e-3 gradient tolerance).  The conv biases in front
    of a training-mode BatchNorm have an exactly-zero true gradient: what is stored there is cancellation noise (~1e-8 of the largest
    gradient), different on every run, and reported separately."""
    from tinyrecurrentunet_b200 import stft_loss
    _, net = make_pair(3)
    net.train()
    mr = stft_loss.MultiResolutionSTFTLoss(fft_sizes=[512, 1024, 2048], hop_sizes=[50, 120, 240],
                                           win_lengths=[240, 600, 1200], sc_lambda=0.5, mag_lambda=0.5).cuda()
    clean, noisy = O.synthetic_batch(8, n=64000, first=70)
    clean, noisy = clean.cuda(), noisy.cuda()
    state = {k: v.clone() for k, v in net.state_dict().items()}
    l1, g1 = _train_grads(net, mr, clean, noisy)
    net.load_state_dict(state)                              # running statistics back to where they were
    l2, g2 = _train_grads(net, mr, clean, noisy)
    identical = sum(int(torch.equal(a, b)) for a, b in zip(g1, g2))
    gmax = max(a.abs().max().item() for a in g1)
    # relative to the tensor's own scale, floored at 1e-3 of the largest gradient (a conv bias in front of a training-mode
    # BatchNorm has an exactly-zero true gradient: what is stored there is cancellation noise, different on every run)
    per = [((a - b).abs().max() / max(a.abs().max().item(), 1e-3 * gmax)).item() for a, b in zip(g1, g2)]
    names = [k for k, _ in net.named_parameters()]
    zero_bias = {"encoder.%d.DepthwiseSeparableConv1d.%d.bias" % (i, j) for i in range(1, 6) for j in (0, 3)}
    zero_bias |= {"decoder.%d.%s.0.bias" % (d, c) for d, c in enumerate(["FirstTrCNN"] + ["TrCNN"] * 4 + ["LastTrCNN"])}
    zero_bias |= {"decoder.%d.%s.3.bias" % (d, c) for d, c in enumerate(["FirstTrCNN"] + ["TrCNN"] * 4)}
    zero_bias |= {"FGRU.conv.0.bias", "TGRU.conv.0.bias"}
    for v, k in sorted(zip(per, names), reverse=True)[:8]:
        print("  %-55s %.3e%s" % (k, v, "  (exactly-zero true gradient: cancellation noise)" if k in zero_bias else ""))
    worst = max(v for v, k in zip(per, names) if k not in zero_bias)
    worst_zero = max(v for v, k in zip(per, names) if k in zero_bias)
    print("bit-identical gradient tensors: %d / %d, worst relative difference %.3e (zero-gradient biases: %.3e of 1e-3 of the largest "
          "gradient), loss %r vs %r" % (identical, len(g1), worst, worst_zero, l1.item(), l2.item()))
    assert abs(l1.item() - l2.item()) <= 2e-7 * abs(l1.item())
    assert worst <= 1e-5 and worst_zero <= 1.0, (worst, worst_zero)
