"""Objective glue: drop-in for the hot-path part of the reference's util.py
(loss_fn :186-251, sampling :178-183, LinearWarmupCosineDecay :109-156, find_max_epoch :30-49;
checkpoint save / resume of train.py:70-95,155-161), repaired per SURVEY D5-D9 (the reference
file does not parse: X8).  Four fused CUDA stages: front end -> TRU-Net ->
mask + iSTFT -> L1 + multi-resolution STFT loss."""
import math
import os
import re

import torch

from . import ops


@torch.no_grad()
def sampling(net, noisy_features):
    """util.py:178-183."""
    return net(noisy_features)


def denoise(net, noisy_audio, beta=0.5):
    """noisy audio (B,N) -> (denoised audio (B,128*(N//128)), network output)."""
    feats = ops.frontend(noisy_audio)
    out = net(feats)
    return ops.mask_istft(out, beta), out


class GraphedDenoise:
    """``denoise`` for ONE input shape, captured once as a CUDA graph and replayed: the ~45 launches of a small batch
    (rt.py:80-89: one clip at a time) cost more on the host than on the device.  ``g = GraphedDenoise(net, B, N)``;
    ``audio, net_out = g(noisy)`` - both results live in the graph's own buffers and are overwritten by the next call.
    The graph holds the addresses of the weights (see StreamingDenoiser.reset_graph): build a new object after they moved."""

    def __init__(self, net, batch, n_samples, beta=0.5, device="cuda"):
        if net.training:
            raise ValueError("GraphedDenoise needs net.eval()")
        device = torch.device(device)
        self.noisy = torch.zeros(batch, n_samples, device=device)
        with torch.no_grad():
            denoise(net, self.noisy, beta)                     # lazy initialisation and shared-memory opt-ins happen here
            torch.cuda.current_stream(device).synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.audio, self.net_out = denoise(net, self.noisy, beta)

    def __call__(self, noisy_audio):
        if noisy_audio.shape != self.noisy.shape:
            raise ValueError("GraphedDenoise was captured for input shape %s" % (tuple(self.noisy.shape),))
        self.noisy.copy_(noisy_audio, non_blocking=True)
        self.graph.replay()
        return self.audio, self.net_out


class CudaPrefetcher:
    """Double-buffered host -> device staging on a side stream, so that the copy of batch i+1 overlaps the training step of
    batch i (train.py:124-125 copies synchronously).  Wraps any iterable of host batches (tuples of tensors; pin them for
    the copies to be asynchronous) and yields every batch exactly once, in order, as a tuple of device tensors:

        for clean, noisy in CudaPrefetcher(loader, device):
            step(clean, noisy)

    A yielded batch stays valid until the loop asks for the next one (the work the loop body enqueued on the current
    stream is fenced with an event before its buffer is overwritten)."""

    def __init__(self, batches, device="cuda"):
        self.batches = batches
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.bufs = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [None, None]                  # event after the last consumer of buffer k (None: never used)

    def _issue(self, k, host):
        cur = torch.cuda.current_stream(self.device)
        old = self.bufs[k]
        if old is None or len(old) != len(host) or any(o.shape != h.shape or o.dtype != h.dtype for o, h in zip(old, host)):
            if old is not None:                   # a ragged batch: the old block may still be the target of an in-flight copy
                for t in old:
                    t.record_stream(self.stream)
            self.bufs[k] = tuple(torch.empty(h.shape, device=self.device, dtype=h.dtype) for h in host)
            fence = torch.cuda.Event()            # the block may be recycled memory that earlier work on `cur` still reads
            fence.record(cur)
            self.stream.wait_event(fence)
        if self.free[k] is not None:
            self.stream.wait_event(self.free[k])  # the step that last read this buffer has finished
        with torch.cuda.stream(self.stream):
            for d, h in zip(self.bufs[k], host):
                d.copy_(h, non_blocking=True)
            self.ready[k].record(self.stream)

    def __iter__(self):
        it = iter(self.batches)
        try:
            self._issue(0, tuple(next(it)))
        except StopIteration:
            return
        k, more = 0, True
        while more:
            try:
                self._issue(k ^ 1, tuple(next(it)))      # the next batch travels while this one is used
            except StopIteration:
                more = False
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self.ready[k])
            yield self.bufs[k]
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self.free[k] = ev
            k ^= 1


def _host_pipeline(fn, batches, device):
    """For every pinned host tensor of ``batches``: device copy (one item ahead, side stream: CudaPrefetcher) -> ``fn`` on the
    current stream -> copy of its result into one of two pinned staging buffers on a second side stream.  Yields the staged
    results in order, each complete when yielded (it is overwritten two items later); item i is yielded after item i+1 has
    been enqueued, so both copies overlap the kernels of the neighbouring items."""
    device = torch.device(device)
    out_stream = torch.cuda.Stream(device=device)
    host = [None, None]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    pending = None
    k = 0
    with torch.no_grad():
        for (x,) in CudaPrefetcher(((b,) for b in batches), device):
            y = fn(x)
            cur = torch.cuda.current_stream(device)
            if host[k] is None or host[k].shape != y.shape:
                host[k] = torch.empty(y.shape, dtype=y.dtype, pin_memory=True)
            out_stream.wait_stream(cur)
            with torch.cuda.stream(out_stream):
                host[k].copy_(y, non_blocking=True)
                done[k].record(out_stream)
            y.record_stream(out_stream)
            if pending is not None:
                done[pending].synchronize()
                yield host[pending]
            pending = k
            k ^= 1
    if pending is not None:
        done[pending].synchronize()
        yield host[pending]


def denoise_host_batches(net, batches, device="cuda", beta=0.5):
    """Offline batch denoising with host buffers on both sides (denoise.py:82-95's loop: load clip -> net -> write wav).
    ``batches`` is an iterable of pinned host tensors (B,N); yields, for every batch in order, a pinned host tensor
    (B,128*(N//128)) holding the denoised audio - complete when it is yielded.  Host->device copies run one batch ahead
    on a side stream (CudaPrefetcher), device->host copies on a second side stream, so both overlap the kernels of the
    neighbouring batches; the yielded tensor is one of two staging buffers and is overwritten two batches later."""
    return _host_pipeline(lambda noisy: denoise(net, noisy, beta)[0], batches, device)


def stream_host_frames(denoiser, frames, device="cuda"):
    """The serving loop of a StreamingDenoiser with host buffers on both sides (stream.py:83-109: hop in from the sound
    card -> net -> hop out): ``frames`` is an iterable of pinned host tensors (S,512), one analysis frame per stream and
    hop; yields the pinned host block (S,128) of every step in order.  The upload of hop t+1 and the download of hop t-1
    overlap the kernels of hop t (one extra hop of pipeline latency for the copies to hide behind)."""
    return _host_pipeline(denoiser.step, frames, device)


class StreamingDenoiser:
    """Stateful frame-by-frame inference for S concurrent streams (rt.py:20-27 / stream.py:83-109 intent):
    carries the PCEN smoother (S,257), the TGRU hidden state (S*16,128) and the overlap-add tail (S,384).
    ``step(frames)`` takes the next 512-sample analysis frame of every stream (hop 128, the reference's centre /
    reflect framing) and returns the 128 output samples that became final (block t-2; zeros for the first two
    steps); ``flush()`` returns the last block after the final frame.  Output equals the offline ``denoise``."""

    def __init__(self, net, n_streams, beta=0.5, device="cuda", cuda_graph=False):
        """``cuda_graph=True``: from the fourth frame on (the back end's normalisation depends on the frame index until then)
        the whole step - ~35 kernels of three C calls - is replayed as ONE captured CUDA graph instead of being launched
        one by one: the step of a few streams is launch-bound (rt.py:20-27: one stream, one frame per call).  The graphs
        hold the addresses of the weights: call ``reset_graph()`` after anything that re-allocates them
        (``optim.FlatAdamW`` re-points ``p.data``; ``load_state_dict`` copies in place and is fine).  The graphs also hold the
        addresses of the state tensors: reset a stream's state by writing INTO ``pcen`` / ``h`` / ``ola`` (``sd.h[rows].zero_()``), not
        by rebinding the attributes."""
        if net.training:
            raise ValueError("StreamingDenoiser needs net.eval()")
        self.net, self.beta, self.t = net, beta, 0
        self.pcen = torch.zeros(n_streams, 257, device=device)
        self.h = torch.zeros(n_streams * 16, 128, device=device)
        self.ola = torch.zeros(n_streams, 384, device=device)
        self.cuda_graph, self._graphs = bool(cuda_graph), None

    GRAPH_FROM = 3          # backend_step_kernel: 1 / (number of frames overlapping the emitted block) is constant from here on

    @torch.no_grad()
    def step(self, frames):
        if self.cuda_graph and self.t >= self.GRAPH_FROM:
            return self._graph_step(frames)
        feats = ops.frontend_step(frames, self.pcen)
        out, self.h = self.net.step(feats, self.h)
        audio = ops.mask_istft_step(out, self.ola, self.t, self.beta)
        self.t += 1
        return audio

    # -- CUDA-graph replay of the steady-state step --------------------------------------------------------------
    def _capture_one(self, graph, pool, h_in, h_out):
        with torch.cuda.graph(graph, pool=pool, capture_error_mode="thread_local"):
            feats = ops.frontend_step(self._g_frames, self.pcen)              # PCEN state updated in place
            out, _ = self.net.step(feats, h_in, h_out=h_out)
            audio = ops.mask_istft_step(out, self.ola, self.GRAPH_FROM, self.beta)   # overlap-add tail updated in place
        return audio                    # lives in the graph's pool; everything else allocated in the capture is released to it

    def _graph_step(self, frames):
        if frames.shape != (self.pcen.shape[0], 512) or frames.dtype != torch.float32 or not frames.is_cuda:
            raise ValueError("step() takes the (S,512) float32 CUDA frames of the S streams")
        if self._graphs is None:
            # two graphs, because the TGRU state ping-pongs between two buffers: graph k reads hs[k] and writes hs[1 - k].
            # The eager steps before GRAPH_FROM have already run every kernel once (lazy per-device initialisation, dynamic
            # shared-memory opt-ins), so nothing but launches and memsets is recorded here.
            self._g_frames = torch.empty_like(frames)
            self._g_h = (self.h.contiguous(), torch.empty_like(self.h))
            torch.cuda.current_stream(frames.device).synchronize()
            graphs, pool = [], None
            for k in range(2):
                g = torch.cuda.CUDAGraph()
                audio = self._capture_one(g, pool, self._g_h[k], self._g_h[1 - k])
                pool = g.pool()
                graphs.append((g, audio))
            self._graphs, self._g_k = graphs, 0
            self.h = self._g_h[0]
        g, audio = self._graphs[self._g_k]
        self._g_frames.copy_(frames, non_blocking=True)
        g.replay()
        self._g_k ^= 1
        self.h = self._g_h[self._g_k]
        self.t += 1
        return audio.clone()            # the graph's own output buffer is overwritten by its next replay

    def reset_graph(self):
        """Drop the captured graphs (the next step captures again): needed after the weights moved to other addresses."""
        self._graphs = None

    @torch.no_grad()
    def flush(self):
        audio = ops.mask_istft_step(None, self.ola, self.t, self.beta, flush=True)
        self.t += 1
        return audio

    # -- the real-time loop of stream.py:83-109: raw audio in, hop by hop, through a 512-sample window per stream --
    @torch.no_grad()
    def feed(self, block):
        """Next 128 input samples of every stream, (S,128) float32 CUDA.  Keeps the sliding analysis window (the
        reference's ring buffer), builds the centre / reflect framing of the offline front end on the fly and returns
        the list of 128-sample output blocks that became final (none for the first three hops: one hop to fill the
        window's look-ahead plus the two hops of overlap-add look-ahead; one per hop afterwards)."""
        if block.shape[-1] != 128:
            raise ValueError("feed() takes one hop (128 samples) per stream")
        if not hasattr(self, "win"):
            self.win = torch.zeros(block.shape[0], 512, device=block.device)      # the last 512 samples received
            self.nblk = 0
        self.win = torch.cat((self.win[:, 128:], block), dim=1)
        self.nblk += 1
        k = self.nblk - 1                                    # index of the block just received
        frames = []
        if k == 2:                                           # samples 0..383 are in win[:, 128:]: frames 0 and 1 exist now
            a = self.win[:, 128:]
            frames.append(torch.cat((a[:, 1:257].flip(1), a[:, :256]), dim=1))   # reflect-padded start
            frames.append(torch.cat((a[:, 1:129].flip(1), a), dim=1))
        elif k > 2:
            frames.append(self.win)                          # frame k-1 = samples 128(k-1)-256 .. 128(k-1)+255
        out = []
        for fr in frames:
            y = self.step(fr.contiguous())
            if self.t > 2:                                   # steps 0 and 1 emit the look-ahead zeros
                out.append(y)
        return out

    @torch.no_grad()
    def finish(self):
        """End of the streams: the last two frames need the reflected tail, then the overlap-add is flushed.  Returns the
        remaining output blocks; in total as many blocks come out as went in."""
        if getattr(self, "nblk", 0) < 3:
            raise ValueError("finish() needs at least three hops of audio (the reflect padding reads 257 samples)")
        w = self.win                                         # samples N-512 .. N-1
        tail = w[:, 255:511].flip(1)                         # samples N-2, N-3, ..., N-257: the reflection of the end
        frames = [torch.cat((w[:, 128:], tail[:, :128]), dim=1), torch.cat((w[:, 256:], tail), dim=1)]
        out = []
        for fr in frames:
            y = self.step(fr.contiguous())
            if self.t > 2:
                out.append(y)
        out.append(self.flush())
        return out


def find_max_epoch(path):
    """util.py:30-49: the largest <iter> among the files ``<iter>.pkl`` in ``path``; -1 if there is none."""
    iters = [int(m.group(1)) for m in (re.fullmatch(r"([+-]?\d+)\.pkl", f) for f in os.listdir(path)) if m]
    return max([-1] + iters)


def save_checkpoint(ckpt_directory, n_iter, net, optimizer, training_time_seconds):
    """train.py:155-161: ``<n_iter>.pkl`` = {iter, model_state_dict, optimizer_state_dict, training_time_seconds}, the
    reference's on-disk format (a FlatAdamW state dict has torch.optim.AdamW's layout, so either side can load it)."""
    name = os.path.join(ckpt_directory, "{}.pkl".format(n_iter))
    torch.save({"iter": n_iter, "model_state_dict": net.state_dict(), "optimizer_state_dict": optimizer.state_dict(),
                "training_time_seconds": int(training_time_seconds)}, name)
    return name


def load_checkpoint(ckpt_directory, ckpt_iter, net, optimizer=None):
    """train.py:70-95: resume from ``<ckpt_iter>.pkl`` (``"max"`` = the newest).  Returns (ckpt_iter, seconds trained), or
    (-1, 0) when there is no usable checkpoint - the reference then trains from initialisation."""
    if ckpt_iter == "max":
        ckpt_iter = find_max_epoch(ckpt_directory)
    if ckpt_iter < 0:
        return -1, 0
    path = os.path.join(ckpt_directory, "{}.pkl".format(ckpt_iter))
    if not os.path.exists(path):
        return -1, 0
    checkpoint = torch.load(path, map_location="cpu")
    net.load_state_dict(checkpoint["model_state_dict"])
    if optimizer is not None:
        optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
    return ckpt_iter, checkpoint["training_time_seconds"]


def fold_batchnorm(state_dict, eps=1e-5):
    """Inference export (the artefact onnx.py:15-30 / rt.py:13-18 want to load): a state dict with the SAME 177 keys in
    which every eval-mode BatchNorm is folded into the convolution in front of it - ``w' = w * g / sqrt(var + eps)`` per
    output channel, ``b' = (b - mean) * g / sqrt(var + eps) + beta`` - and the BatchNorm entries are set to the identity
    (weight 1, bias 0, running_mean 0, running_var 1 - eps).  Loads into TRUNet (this package's or the reference's layer
    list) and gives the eval-mode outputs of the original; the CUDA kernels gain nothing from it (they apply the
    eval-mode affine while loading), it exists for exchange with other runtimes."""
    out = type(state_dict)((k, v.clone()) for k, v in state_dict.items())
    for key in state_dict:
        if not key.endswith(".running_mean"):
            continue
        bn = key[:-len(".running_mean")]
        parent, idx = bn.rsplit(".", 1)
        conv = "%s.%d" % (parent, int(idx) - 1)
        g, beta = state_dict[bn + ".weight"].double(), state_dict[bn + ".bias"].double()
        mean, var = state_dict[bn + ".running_mean"].double(), state_dict[bn + ".running_var"].double()
        scale = g / torch.sqrt(var + eps)
        w = state_dict[conv + ".weight"].double()
        transposed = parent.startswith("decoder.") and int(idx) - 1 == 3          # ConvTranspose1d: (C_in, C_out, k)
        shape = (1, -1, 1) if transposed else (-1, 1, 1)
        out[conv + ".weight"] = (w * scale.view(shape)).to(state_dict[conv + ".weight"].dtype)
        out[conv + ".bias"] = ((state_dict[conv + ".bias"].double() - mean) * scale + beta).to(state_dict[conv + ".bias"].dtype)
        out[bn + ".weight"] = torch.ones_like(state_dict[bn + ".weight"])
        out[bn + ".bias"] = torch.zeros_like(state_dict[bn + ".bias"])
        out[bn + ".running_mean"] = torch.zeros_like(state_dict[bn + ".running_mean"])
        out[bn + ".running_var"] = torch.full_like(state_dict[bn + ".running_var"], 1.0 - eps)
    return out


def _ramp_linear(a, b, x):
    return a + x * (b - a)


def _ramp_cosine(a, b, x):
    c = math.cos(math.pi * x) + 1
    return b + (a - b) / 2 * c


class LinearWarmupCosineDecay:
    """util.py:109-156 (train.py:98-104 builds it, :139 steps it before every optimizer step): a warm-up leg from
    ``lr_max / divider`` to ``lr_max`` over the first ``int(n_iter * warmup_proportion)`` steps, then a decay leg down
    to ``lr_max / divider / 1e4``; after both legs the schedule starts over.  ``iteration`` resumes mid-schedule.
    ``step()`` writes the new rate into every ``optimizer.param_groups[i]["lr"]`` (a host float: FlatAdamW passes it
    to the kernel by value, nothing synchronises) and returns it."""

    def __init__(self, optimizer, lr_max, n_iter, iteration=0, divider=25, warmup_proportion=0.3,
                 phase=("linear", "cosine")):
        self.optimizer = optimizer
        ramps = {"linear": _ramp_linear, "cosine": _ramp_cosine}
        n_warm = int(n_iter * warmup_proportion)
        lr_min = lr_max / divider
        # one (start, end, length, ramp) tuple per leg and the number of steps already taken on it
        self._legs = [(lr_min, lr_max, n_warm, ramps[phase[0]]),
                      (lr_max, lr_min / 1e4, n_iter - n_warm, ramps[phase[1]])]
        self._taken = [iteration, max(0, iteration - n_warm)]
        self.phase = 0 if iteration < n_warm else 1

    def step(self):
        k = self.phase
        start, end, length, ramp = self._legs[k]
        self._taken[k] += 1
        lr = ramp(start, end, self._taken[k] / length)
        for group in self.optimizer.param_groups:
            group["lr"] = lr
        if self._taken[k] >= length:
            self.phase += 1
            if self.phase == len(self._legs):
                self._taken = [0] * len(self._legs)
                self.phase = 0
        return lr


def loss_fn(net, X, ell_p=1, ell_p_lambda=1, stft_lambda=1, mrstftloss=None, **kwargs):
    """util.py:186-251.  X = (clean_audio, noisy_audio); shapes (1,1,N)/(1,N) as the
    reference's loader yields them, or batched (B,N).  Returns
    (loss, {"l1", "stft_sc", "stft_mag"}).  ``ell_p`` / ``ell_p_lambda`` are accepted
    and unused exactly like the reference (D9)."""
    clean_audio, noisy_audio = X
    clean = clean_audio.reshape(-1, clean_audio.shape[-1])
    noisy = noisy_audio.reshape(-1, noisy_audio.shape[-1])
    denoised, _ = denoise(net, noisy)
    n = denoised.shape[-1]
    if clean.shape[-1] != n:
        clean = clean[..., :n]
    clean = clean.contiguous()
    output_dic = {}
    if stft_lambda > 0:
        if mrstftloss is None:
            raise ValueError("stft_lambda > 0 needs mrstftloss (train.py:114)")
        l1, sc_loss, mag_loss = mrstftloss.forward_with_l1(denoised, clean)
        l1 = torch.abs(l1)
        loss = l1 + (sc_loss + mag_loss) * stft_lambda
        output_dic["l1"] = l1.data
        output_dic["stft_sc"] = sc_loss.data * stft_lambda
        output_dic["stft_mag"] = mag_loss.data * stft_lambda
    else:
        l1 = torch.abs(torch.nn.functional.l1_loss(denoised, clean))
        loss = l1
        output_dic["l1"] = l1.data
    return loss, output_dic
