"""Host-side mirror of the reference's model plug-in interface.

``Model(d_args)`` has the constructor, ``forward(x, Freq_aug=False, speaker_embedding=None)``
signature, return tuple and ``state_dict`` key layout of the reference's
``models/AASIST.py::Model`` (reference models/AASIST.py:728-921, with the checkpoint's (2,3)
``Residual_block`` encoder, models/RawNetGatSpoofST.py:225-278) -- so
``model.load_state_dict(torch.load("models/weights/AASIST.pth"))`` works unchanged -- but
owns no compute: parameters are plain ``nn.Parameter`` containers and ``forward`` hands raw
device pointers to libaasist_b200.so.  PyTorch is only the tensor container / stream owner.

Encoder selection.  The fork's ``Model`` always builds the Res2Net+SE encoder
(models/AASIST.py:766-772), which cannot load the shipped checkpoints.  Here the
encoder follows the ``d_args`` keys the fork reads: a config that carries ``res2net_width`` /
``res2net_scale`` (config/AASIST2.conf:29-30) -- or ``"encoder": "res2net"`` -- gets the
Res2Net+SE encoder with the fork's exact state_dict; a config without them (config/AASIST.conf,
AASIST-L.conf: the shipped checkpoints) gets the (2,3) ``Residual_block`` encoder.

Input range (precision ``"f16x3"``): waveforms are expected in [-1, 1] like the reference's
``soundfile`` floats; fp16 operand pairs saturate at +-65504, so an un-normalised (e.g. int16-scale)
waveform must be scaled first or scored with ``precision="fp32"``.  Utterances longer than ~18 s
(> 132 temporal nodes, ~291 000 samples) exceed the graph kernel's shared memory and raise.
"""
from __future__ import annotations

import ctypes as C
import random
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib

Tensor = torch.Tensor


def _bn(n: int) -> nn.Module:
    # parameter/buffer container with the key names of nn.BatchNorm{1,2}d
    return nn.BatchNorm1d(n)


class _Conv(nn.Module):
    """Container for a Conv2d's weight/bias (keys ``weight`` / ``bias``); never called."""

    def __init__(self, ci: int, co: int, kh: int, kw: int):
        super().__init__()
        conv = nn.Conv2d(ci, co, (kh, kw))          # same init as the reference's layers
        self.weight, self.bias = conv.weight, conv.bias


class _Linear(nn.Module):
    def __init__(self, i: int, o: int):
        super().__init__()
        lin = nn.Linear(i, o)
        self.weight, self.bias = lin.weight, lin.bias


def _xavier(*size: int) -> nn.Parameter:
    p = nn.Parameter(torch.empty(*size))
    nn.init.xavier_normal_(p)                        # AASIST.py:107-110
    return p


class _ResidualBlockParams(nn.Module):
    """state_dict layout of Residual_block (RawNetGatSpoofST.py:226-256)."""

    def __init__(self, nb_filts, first: bool = False):
        super().__init__()
        ci, co = nb_filts
        if not first:
            self.bn1 = nn.BatchNorm2d(ci)            # present in the checkpoint, dead in forward
        self.conv1 = _Conv(ci, co, 2, 3)
        self.bn2 = nn.BatchNorm2d(co)
        self.conv2 = _Conv(co, co, 2, 3)
        if ci != co:
            self.conv_downsample = _Conv(ci, co, 1, 3)


def res2net_splits(nb_filts, width: int, scale: int) -> Tuple[List[int], int]:
    """Split sizes / effective scale of Res2NetBlock.__init__ (AASIST.py:528-565)."""
    w = min(width, nb_filts[0])
    s = min(scale, w)
    base = max(1, nb_filts[0] // w)
    rem = nb_filts[0] - base * (w - 1)
    return [max(1, base if i < w - 1 else rem) for i in range(w)], s


class _SEParams(nn.Module):
    """state_dict layout of SELayer (AASIST.py:508-516): fc.0.weight, fc.2.weight (no biases)."""

    def __init__(self, channel: int, reduction: int = 16):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(channel, channel // reduction, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(channel // reduction, channel, bias=False), nn.Sigmoid())


class _Res2NetBlockParams(nn.Module):
    """state_dict layout of Res2NetBlock (AASIST.py:527-601)."""

    def __init__(self, nb_filts, first: bool = False, width: int = 14, scale: int = 8):
        super().__init__()
        ci, co = nb_filts
        if not first:
            self.bn1 = nn.BatchNorm2d(ci)            # LIVE in this block (AASIST.py:611-613)
        sizes, _ = res2net_splits(nb_filts, width, scale)
        self.convs = nn.ModuleList([_Conv(n, n, 3, 3) for n in sizes])
        self.bn2 = nn.BatchNorm2d(ci)
        self.conv_cat = _Conv(ci, co, 3, 3)
        self.se = _SEParams(co)
        if ci != co:
            self.conv_downsample = _Conv(ci, co, 1, 3)


class _ResidualBlock33Params(nn.Module):
    """state_dict layout of the fork's 3x3 Residual_block (AASIST.py:672-701)."""

    def __init__(self, nb_filts, first: bool = False):
        super().__init__()
        ci, co = nb_filts
        if not first:
            self.bn1 = nn.BatchNorm2d(ci)            # dead in forward (AASIST.py:706-712)
        self.conv1 = _Conv(ci, co, 3, 3)
        self.bn2 = nn.BatchNorm2d(co)
        self.conv2 = _Conv(co, co, 3, 3)
        if ci != co:
            self.conv_downsample = _Conv(ci, co, 1, 3)


class _SpeakerConditioningParams(nn.Module):
    """state_dict layout of SpeakerConditioningModule (AASIST.py:330-367)."""

    def __init__(self, spk_emb_dim: int, target_dim: int, use_attention: bool):
        super().__init__()
        self.proj = nn.Linear(spk_emb_dim, target_dim)
        if use_attention:
            self.attention = nn.Sequential(nn.Linear(target_dim * 2, target_dim), nn.Tanh(),
                                           nn.Linear(target_dim, 1), nn.Softmax(dim=1))
        self.fusion = nn.Sequential(nn.Linear(target_dim * 2, target_dim), nn.ReLU())


class _DenoisingParams(nn.Module):
    """state_dict layout of FeatureDenoising (AASIST_Robust.py:45-62); training-only, never applied here."""

    def __init__(self, c: int):
        super().__init__()
        self.g, self.theta, self.phi, self.W = (nn.Conv1d(c, c, 1) for _ in range(4))
        self.bn = nn.BatchNorm1d(c)


class _GaussianNoiseParams(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("noise", torch.tensor(0.))   # AASIST_Robust.py:35


class _GatParams(nn.Module):
    """state_dict layout of GraphAttentionLayer (AASIST.py:18-41)."""

    def __init__(self, in_dim: int, out_dim: int):
        super().__init__()
        self.att_proj = _Linear(in_dim, out_dim)
        self.att_weight = _xavier(out_dim, 1)
        self.proj_with_att = _Linear(in_dim, out_dim)
        self.proj_without_att = _Linear(in_dim, out_dim)
        self.bn = _bn(out_dim)


class _HtrgGatParams(nn.Module):
    """state_dict layout of HtrgGraphAttentionLayer (AASIST.py:114-148)."""

    def __init__(self, in_dim: int, out_dim: int):
        super().__init__()
        self.proj_type1 = _Linear(in_dim, in_dim)
        self.proj_type2 = _Linear(in_dim, in_dim)
        self.att_proj = _Linear(in_dim, out_dim)
        self.att_projM = _Linear(in_dim, out_dim)
        self.att_weight11 = _xavier(out_dim, 1)
        self.att_weight22 = _xavier(out_dim, 1)
        self.att_weight12 = _xavier(out_dim, 1)
        self.att_weightM = _xavier(out_dim, 1)
        self.proj_with_att = _Linear(in_dim, out_dim)
        self.proj_without_att = _Linear(in_dim, out_dim)
        self.proj_with_attM = _Linear(in_dim, out_dim)
        self.proj_without_attM = _Linear(in_dim, out_dim)
        self.bn = _bn(out_dim)


class _PoolParams(nn.Module):
    def __init__(self, in_dim: int):
        super().__init__()
        self.proj = _Linear(in_dim, 1)


def _encoder(filts, block=_ResidualBlockParams, **kw) -> nn.Sequential:
    # same nn.Sequential(nn.Sequential(block)) nesting as AASIST.py:766-772 -> keys encoder.N.0.*
    return nn.Sequential(
        nn.Sequential(block(filts[1], first=True, **kw)),
        nn.Sequential(block(filts[2], **kw)),
        nn.Sequential(block(filts[3], **kw)),
        nn.Sequential(block(filts[4], **kw)),
        nn.Sequential(block(filts[4], **kw)),
        nn.Sequential(block(filts[4], **kw)))


def draw_freq_mask(n_filters: int) -> Tuple[int, int]:
    """The reference's Freq_aug draws, in its order and from the same global generators
    (AASIST.py:487-489): ``A = int(np.random.uniform(0, 20))``, ``A0 = random.randint(0, F - A)``."""
    a = int(np.random.uniform(0, 20))
    a0 = random.randint(0, n_filters - a)
    return a0, a


class _NativeModel(nn.Module):
    """Shared machinery: handle life cycle, parameter hand-over, forward through the C ABI."""

    _kind = _lib.KIND_AASIST

    def __init__(self, d_args: dict, precision: Optional[str] = None):
        super().__init__()
        self.d_args = d_args
        self.precision = precision or d_args.get("precision", DEFAULT_PRECISION)
        if self.precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        self._handle: Optional[C.c_void_p] = None
        self._packed_key = None
        self._workspace: Optional[Tensor] = None
        self.last_topk: Optional[Tensor] = None
        self.last_pool_weights: Optional[Tensor] = None
        self.record_topk = False

    # -- configuration -> struct aasist_config -------------------------------------------------
    def _config(self) -> _lib.AasistConfig:
        d = self.d_args
        filts = d["filts"]
        cfg = _lib.AasistConfig()
        cfg.kind = self._kind
        cfg.precision = _lib.PRECISIONS[self.precision]
        cfg.first_conv = int(d["first_conv"])
        cfg.n_filters = int(filts[0])
        if self._kind == _lib.KIND_ROBUST:
            # CONV(out_channels=d_args['first_conv'], kernel_size=1024, stride=256) (AASIST_Robust.py:96-102):
            # `first_conv` is the NUMBER of sinc filters there and filts[0] is never read
            cfg.n_filters, cfg.first_conv = int(d["first_conv"]), 1024
        blocks = [filts[1], filts[2], filts[3], filts[4], filts[4], filts[4]]
        for i, (ci, co) in enumerate(blocks):
            cfg.enc_channels[i][0], cfg.enc_channels[i][1] = int(ci), int(co)
        if self._kind != _lib.KIND_RAWGAT_ST:
            cfg.gat_dims[0], cfg.gat_dims[1] = int(d["gat_dims"][0]), int(d["gat_dims"][1])
            for i in range(4):
                cfg.pool_ratios[i] = float(d["pool_ratios"][i])
                cfg.temperatures[i] = float(d["temperatures"][i])
        cfg.sample_rate = 16000
        cfg.encoder = getattr(self, "_encoder_kind", _lib.ENC_RESIDUAL23)
        cfg.res2net_width = int(d.get("res2net_width", 14))             # AASIST.py:739-740
        cfg.res2net_scale = int(d.get("res2net_scale", 8))
        if getattr(self, "use_speaker_conditioning", False):
            cfg.spk_emb_dim = int(self.spk_emb_dim)
            cfg.spk_level = 0 if self.conditioning_level == "frame" else 1
            cfg.spk_use_attention = 1 if self.use_attention else 0
        return cfg

    # -- parameter hand-over ---------------------------------------------------------------------
    # parameters are handed to the library once and re-packed only when they change: `_apply`
    # (.to/.cuda/.float) and `load_state_dict` mark the handle stale, in-place edits are caught by the
    # tensors' version counters (a few microseconds per forward).
    def _apply(self, fn, *args, **kwargs):
        self._stale = True
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._stale = True
        return super().load_state_dict(*args, **kwargs)

    def _state_key(self):
        tensors = self.__dict__.get("_flat")
        if tensors is None or self.__dict__.get("_stale", True):
            tensors = [v for _, v in self.state_dict(keep_vars=True).items()]
            self.__dict__["_flat"] = tensors
            self.__dict__["_stale"] = False
            self.__dict__["_flat_id"] = self.__dict__.get("_flat_id", 0) + 1
        return (self.__dict__["_flat_id"], sum(t._version for t in tensors))

    def _ensure_handle(self, device: torch.device):
        lib = _lib.load()
        key = (str(device), self.precision, self._state_key())
        if self._handle is not None and key == self._packed_key:
            return
        if self._handle is not None:
            lib.aasist_destroy(self._handle)
            self._handle = None
        with torch.cuda.device(device):
            cfg = self._config()
            handle = C.c_void_p()
            _lib.check(lib.aasist_create(C.byref(cfg), C.byref(handle)))
            try:
                for name, t in self.state_dict().items():
                    if name.endswith("num_batches_tracked"):
                        continue
                    t = t.detach().to(torch.float32).contiguous()
                    _lib.check(lib.aasist_set_param(handle, name.encode(), t.data_ptr(), t.numel()))
                _lib.check(lib.aasist_finalize(handle))
            except Exception:
                lib.aasist_destroy(handle)
                raise
        self._handle, self._packed_key = handle, key

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None and _lib._lib is not None:
                _lib._lib.aasist_destroy(self._handle)
        except Exception:
            pass

    def _workspace_for(self, nbytes: int, device: torch.device) -> Tensor:
        ws = self._workspace
        if ws is None or ws.device != device or ws.numel() < nbytes:
            self._workspace = ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return ws

    @property
    def hidden_dim(self) -> int:
        if self._kind == _lib.KIND_ROBUST:
            return 2
        return 5 * int(self.d_args["gat_dims"][1]) if self._kind == _lib.KIND_AASIST else 7

    def topk_layout(self, length: int):
        """[(n_nodes_in, k)] per GraphPool, in the order the indices are written."""
        lib = _lib.load()
        dev = next(self.parameters()).device
        self._ensure_handle(dev)
        n = C.c_int32()
        nk = (C.c_int32 * 12)()
        _lib.check(lib.aasist_topk_layout(self._handle, length, C.byref(n), nk))
        return [(nk[2 * i], nk[2 * i + 1]) for i in range(n.value)]

    def check_input_range(self) -> bool:
        """True (and a RuntimeWarning) when a forward that has completed since the last check saw samples outside
        the f16x3 operand range (|x| > 63.96: an un-normalised waveform).  Called after every synchronising entry
        point and, for the asynchronous ``forward``, at the start of the next call."""
        if self._handle is None or self.precision == "fp32":
            return False
        if _lib.load().aasist_input_range_exceeded(self._handle, 1):
            import warnings
            warnings.warn("aasist_b200: input samples exceed the fp16 operand range of precision "
                          f"'{self.precision}' (|x| > 63.96); the logits of that batch are wrong. Normalise the "
                          "waveform to [-1, 1] or use precision='fp32'.", RuntimeWarning, stacklevel=3)
            return True
        return False

    def _forward_native(self, x: Tensor, freq_mask: Optional[Tuple[int, int]] = None,
                        speaker_embedding: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        if x.dim() == 3 and x.size(1) == 1:                  # AASIST.py:816-817 accepts (B,1,L)
            x = x[:, 0]
        if x.dim() != 2:
            raise RuntimeError(f"expected input of shape (batch, samples), got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("aasist_b200 runs on CUDA (sm_100a) only and has no CPU fallback: "
                               "move the model and the input to a CUDA device")
        if self.training:
            raise NotImplementedError("aasist_b200 implements the eval-mode scoring forward only; "
                                      "call model.eval() (reference main.py:354)")
        lib = _lib.load()
        dev = x.device
        p0 = next(self.parameters())
        if p0.device != dev:
            raise RuntimeError(f"model parameters are on {p0.device}, input on {dev}")
        self._ensure_handle(dev)
        self.check_input_range()                  # of the forwards that have completed so far
        x = x.detach().to(torch.float32).contiguous()
        B, L = x.shape
        opts = None
        if freq_mask is not None or speaker_embedding is not None:
            opts = _lib.ForwardOpts()
            if freq_mask is not None:
                opts.freq_mask_start, opts.freq_mask_count = int(freq_mask[0]), int(freq_mask[1])
            if speaker_embedding is not None:
                emb = speaker_embedding.detach().to(device=dev, dtype=torch.float32).contiguous()
                want = (B, int(getattr(self, "spk_emb_dim", emb.shape[-1])))
                if getattr(self, "use_speaker_conditioning", False) and tuple(emb.shape) != want:
                    raise RuntimeError(f"speaker_embedding must have shape {want}, got {tuple(emb.shape)}")
                opts.speaker_embedding = emb.data_ptr()
        with torch.cuda.device(dev):
            _lib.check(lib.aasist_workspace_bytes(self._handle, B, L))       # shape errors surface here
            last_hidden = torch.empty(B, self.hidden_dim, dtype=torch.float32, device=dev)
            output = torch.empty(B, 2, dtype=torch.float32, device=dev)
            topk_ptr = scores_ptr = None
            if self.record_topk:
                layout = self.topk_layout(L)
                self.last_topk = torch.empty(B, sum(k for _, k in layout), dtype=torch.int32, device=dev)
                self.last_pool_weights = torch.empty(B, sum(n for n, _ in layout), dtype=torch.float32,
                                                     device=dev)
                topk_ptr, scores_ptr = self.last_topk.data_ptr(), self.last_pool_weights.data_ptr()
            stream = torch.cuda.current_stream(dev).cuda_stream
            # workspace = NULL: the handle's own scratch, the same buffer score_host / the scoring stream use
            _lib.check(lib.aasist_forward_ex(self._handle, x.data_ptr(), B, L,
                                             C.byref(opts) if opts is not None else None, last_hidden.data_ptr(),
                                             output.data_ptr(), topk_ptr, scores_ptr, None, 0, stream))
        return last_hidden, output

    # -- host-buffer entry (reference main.py:372-377: .to(device) ... .cpu()) ----------------------
    def score_host(self, x_host: Tensor, device: Optional[torch.device] = None) -> Tuple[Tensor, Tensor]:
        """x_host: CPU (ideally pinned) (B,L) fp32.  Returns CPU (last_hidden, output); the H2D
        copy, the forward and the D2H copy all run inside the C-ABI call."""
        lib = _lib.load()
        if not isinstance(x_host, torch.Tensor) or x_host.is_cuda or x_host.dim() != 2:
            raise RuntimeError("score_host expects a CPU tensor of shape (batch, samples); "
                               "use forward() for tensors that already live on the device")
        dev = device or next(self.parameters()).device
        self._ensure_handle(dev)
        x_host = x_host.to(torch.float32).contiguous()
        B, L = x_host.shape
        last_hidden = torch.empty(B, self.hidden_dim, dtype=torch.float32)
        output = torch.empty(B, 2, dtype=torch.float32)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.aasist_forward_host(self._handle, x_host.data_ptr(), B, L,
                                               last_hidden.data_ptr(), output.data_ptr(), stream))
        self.check_input_range()
        return last_hidden, output

    # -- input staging (reference data_utils.py:45-52 `pad`, applied per utterance at :208) -----------
    def pad_batch(self, utterances, max_len: int = 64600) -> Tensor:
        """Ragged 1-D waveforms -> (B, max_len) on the model's device by repeat-tiling / cropping,
        in one kernel (the reference does this on the host, one utterance at a time)."""
        lib = _lib.load()
        dev = next(self.parameters()).device
        self._ensure_handle(dev)
        lengths = [int(u.numel()) for u in utterances]
        if any(n < 1 for n in lengths):
            raise ZeroDivisionError("empty utterance (reference `pad` divides by the length)")
        flat = torch.cat([u.reshape(-1).to(device=dev, dtype=torch.float32, non_blocking=True) for u in utterances])
        B = len(lengths)
        offs = (C.c_int64 * B)()
        lens = (C.c_int32 * B)()
        acc = 0
        for i, n in enumerate(lengths):
            offs[i], lens[i] = acc, n
            acc += n
        out = torch.empty(B, max_len, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.aasist_pad_batch(self._handle, flat.data_ptr(), offs, lens, B, max_len,
                                            out.data_ptr(), stream))
        return out

    def stage_batch(self, utterances, row_len: int, starts: Optional[Sequence[int]] = None,
                    targets: Optional[Sequence[int]] = None) -> Tensor:
        """``out[b][i] = i < target_b ? x_b[(start_b + i) mod len_b] : 0`` on the model's device, one kernel
        (the deterministic core of data_utils.py ``pad`` / ``pad_random`` / ``dynamic_chunk_size`` /
        ``pad_sequence``; the random draws stay with the caller)."""
        lib = _lib.load()
        dev = next(self.parameters()).device
        self._ensure_handle(dev)
        lengths = [int(u.numel()) for u in utterances]
        if any(n < 1 for n in lengths):
            raise ZeroDivisionError("empty utterance (the reference divides by the length)")
        flat = torch.cat([u.reshape(-1).to(device=dev, dtype=torch.float32, non_blocking=True) for u in utterances])
        B = len(lengths)
        offs, lens = (C.c_int64 * B)(), (C.c_int32 * B)()
        acc = 0
        for i, n in enumerate(lengths):
            offs[i], lens[i] = acc, n
            acc += n
        st = (C.c_int32 * B)(*[int(v) for v in starts]) if starts is not None else None
        tg = (C.c_int32 * B)(*[int(v) for v in targets]) if targets is not None else None
        out = torch.empty(B, int(row_len), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.aasist_stage_batch(self._handle, flat.data_ptr(), offs, lens, st, tg, B, int(row_len),
                                              out.data_ptr(), stream))
        return out

    def pad_sequence(self, batch):
        """Device version of the reference collate function (data_utils.py:100-119): ``batch`` is a list of
        ``(x, label, duration)`` items; returns ``(X_padded, y, durations)`` with X zero-padded to the batch
        maximum rounded up to a multiple of 4, on the model's device."""
        lib = _lib.load()
        X = [item[0] for item in batch]
        y = torch.LongTensor([item[1] for item in batch])
        durations = torch.FloatTensor([item[2] for item in batch])
        B = len(X)
        lens = (C.c_int32 * B)(*[int(x.numel()) for x in X])
        max_len = _lib.check(lib.aasist_pad_sequence_length(lens, B))
        return self.stage_batch(X, max_len, None, [min(int(x.numel()), max_len) for x in X]), y, durations

    def dynamic_chunks(self, utterances, min_samples: int = 16000, max_samples: int = 96000):
        """``dynamic_chunk_size`` (data_utils.py:68-97) for a batch: the target length and crop start of every
        utterance are drawn from numpy's global generator in the reference's order, the crop / repeat-tile and
        the collate padding (``pad_sequence``) run on the device.  Returns ``(X_padded, durations)``."""
        targets, starts = [], []
        for u in utterances:
            n = int(u.numel())
            t = int(np.random.randint(min_samples, max_samples + 1))
            s = int(np.random.randint(0, n - t + 1)) if n >= t else 0
            targets.append(t)
            starts.append(s)
        B = len(targets)
        lens = (C.c_int32 * B)(*targets)
        row = _lib.check(_lib.load().aasist_pad_sequence_length(lens, B))
        return self.stage_batch(utterances, row, starts, targets), torch.FloatTensor([t / 16000 for t in targets])

    # -- pipelined scoring (reference main.py:364-378 without the per-batch round trip) -----------------
    def score_begin(self, capacity: int, max_batch: int, length: int, device: Optional[torch.device] = None):
        lib = _lib.load()
        dev = device or next(self.parameters()).device
        self._ensure_handle(dev)
        self._score_dev = dev
        self.__dict__["_score_keep"], self.__dict__["_score_n"] = [], 0
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.aasist_score_begin(self._handle, int(capacity), int(max_batch), int(length), stream))

    def score_submit(self, x_host: Tensor) -> None:
        """Queue one (B, L) CPU batch: staging copy, H2D on the copy stream, forward behind it; returns without
        waiting for the device.  A pinned ``x_host`` is read in place and must stay alive until score_finish."""
        if not isinstance(x_host, torch.Tensor) or x_host.is_cuda or x_host.dim() != 2:
            raise RuntimeError("score_submit expects a CPU tensor of shape (batch, samples)")
        x_host = x_host.to(torch.float32).contiguous()
        self.__dict__.setdefault("_score_keep", []).append(x_host)
        _lib.check(_lib.load().aasist_score_submit(self._handle, x_host.data_ptr(), x_host.shape[0]))
        self.__dict__["_score_n"] = self.__dict__.get("_score_n", 0) + x_host.shape[0]

    def score_finish(self, want_hidden: bool = False, on_device: bool = False):
        """Wait once for everything submitted; returns ``output`` (n, 2) -- a CUDA tensor when ``on_device``,
        else a CPU tensor -- or ``(last_hidden, output)`` when ``want_hidden``."""
        lib = _lib.load()
        dev = self._score_dev
        n = int(self.__dict__.get("_score_n", 0))
        where = dev if on_device else torch.device("cpu")
        out = torch.empty(n, 2, dtype=torch.float32, device=where)
        hid = torch.empty(n, self.hidden_dim, dtype=torch.float32, device=where) if want_hidden else None
        with torch.cuda.device(dev):
            got = _lib.check(lib.aasist_score_finish(self._handle, out.data_ptr(),
                                                     hid.data_ptr() if hid is not None else None, None))
        assert got == n, (got, n)
        self.check_input_range()
        self.__dict__["_score_keep"], self.__dict__["_score_n"] = [], 0
        return (hid, out) if want_hidden else out

    def _require_handle(self) -> None:
        if self._handle is None:
            raise RuntimeError("the native handle does not exist yet: run a forward (or move the model to its "
                               "CUDA device and call _ensure_handle) first")

    def profile(self, enable: bool = True) -> None:
        """Bracket every kernel launch with CUDA events on the launching stream."""
        self._require_handle()
        _lib.check(_lib.load().aasist_profile_enable(self._handle, 1 if enable else 0))

    def profile_report(self, reset: bool = True):
        """[{kernel, launches, ms}] accumulated since the last reset (synchronises the device)."""
        import json
        self._require_handle()
        buf = C.create_string_buffer(1 << 16)
        _lib.check(_lib.load().aasist_profile_report(self._handle, buf, len(buf), 1 if reset else 0))
        return json.loads(buf.value.decode())

    def launch_count(self) -> int:
        return int(_lib.load().aasist_launch_count(self._handle)) if self._handle is not None else 0


DEFAULT_PRECISION = "f16x3"


class Model(_NativeModel):
    """Drop-in for reference ``models/AASIST.py::Model`` (scoring forward).

    ``d_args`` keys read: ``filts, gat_dims, pool_ratios, temperatures, first_conv`` (AASIST.py:733-736,758) and
    the fork's optional ``res2net_width, res2net_scale, speaker_conditioning, spk_emb_dim, conditioning_level,
    use_attention`` (:739-747).  See the module docstring for how the encoder type follows from them."""

    _kind = _lib.KIND_AASIST

    def __init__(self, d_args: dict, precision: Optional[str] = None):
        super().__init__(d_args, precision)
        filts, gat_dims = d_args["filts"], d_args["gat_dims"]
        enc = d_args.get("encoder")
        if enc is None:
            enc = "res2net" if ("res2net_width" in d_args or "res2net_scale" in d_args) else "residual"
        if enc not in ("res2net", "residual"):
            raise ValueError("d_args['encoder'] must be 'res2net' or 'residual'")
        self._encoder_kind = _lib.ENC_RES2NET if enc == "res2net" else _lib.ENC_RESIDUAL23
        # speaker conditioning (AASIST.py:743-755): one module, applied to the fused T and S nodes
        self.use_speaker_conditioning = bool(d_args.get("speaker_conditioning", False))
        if self.use_speaker_conditioning:
            self.spk_emb_dim = int(d_args.get("spk_emb_dim", 256))
            self.conditioning_level = d_args.get("conditioning_level", "frame")
            self.use_attention = bool(d_args.get("use_attention", True))
            self.spk_cond_gat = _SpeakerConditioningParams(self.spk_emb_dim, gat_dims[1], self.use_attention)
        self.first_bn = nn.BatchNorm2d(1)                                    # AASIST.py:760
        if self._encoder_kind == _lib.ENC_RES2NET:                           # :766-772
            self.encoder = _encoder(filts, _Res2NetBlockParams, width=int(d_args.get("res2net_width", 14)),
                                    scale=int(d_args.get("res2net_scale", 8)))
        else:                                                                # the checkpoints' (2,3) blocks
            self.encoder = _encoder(filts)
        self.pos_S = nn.Parameter(torch.randn(1, 23, filts[-1][-1]))         # :774
        self.master1 = nn.Parameter(torch.randn(1, 1, gat_dims[0]))          # :775
        self.master2 = nn.Parameter(torch.randn(1, 1, gat_dims[0]))          # :776
        self.GAT_layer_S = _GatParams(filts[-1][-1], gat_dims[0])            # :778
        self.GAT_layer_T = _GatParams(filts[-1][-1], gat_dims[0])            # :781
        self.HtrgGAT_layer_ST11 = _HtrgGatParams(gat_dims[0], gat_dims[1])   # :785
        self.HtrgGAT_layer_ST12 = _HtrgGatParams(gat_dims[1], gat_dims[1])   # :787
        self.HtrgGAT_layer_ST21 = _HtrgGatParams(gat_dims[0], gat_dims[1])   # :790
        self.HtrgGAT_layer_ST22 = _HtrgGatParams(gat_dims[1], gat_dims[1])   # :793
        self.pool_S = _PoolParams(gat_dims[0])                               # :796
        self.pool_T = _PoolParams(gat_dims[0])
        self.pool_hS1 = _PoolParams(gat_dims[1])
        self.pool_hT1 = _PoolParams(gat_dims[1])
        self.pool_hS2 = _PoolParams(gat_dims[1])
        self.pool_hT2 = _PoolParams(gat_dims[1])
        self.out_layer = _Linear(5 * gat_dims[1], 2)                         # :804
        self.last_freq_mask: Optional[Tuple[int, int]] = None

    def forward(self, x: Tensor, Freq_aug: bool = False,
                speaker_embedding: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        """Returns ``(last_hidden (B, 5*gat_dims[1]), output (B, 2))`` (AASIST.py:921).

        ``Freq_aug=True`` zeroes ``A`` consecutive rows of the sinc bank starting at ``A0`` for this call, with
        ``A`` and ``A0`` drawn like the reference does (AASIST.py:486-490; ``last_freq_mask`` keeps the draw).
        ``speaker_embedding`` (B, spk_emb_dim) conditions the fused node features when the model was built with
        ``speaker_conditioning`` (frame level); without the module it is ignored, like in the reference."""
        mask = None
        if Freq_aug:
            mask = self.last_freq_mask = draw_freq_mask(int(self.d_args["filts"][0]))
        if speaker_embedding is not None and not self.use_speaker_conditioning:
            speaker_embedding = None                                         # AASIST.py:895 `use_... and ...`
        return self._forward_native(x, mask, speaker_embedding)


class RobustModel(_NativeModel):
    """Drop-in for reference ``models/AASIST_Robust.py::Model`` (eval forward, AASIST_Robust.py:198-303).

    Returns ``(ensemble_logits (B,2), logits (B,2))`` like the reference (:303).  The Gaussian-noise layer and
    the feature-denoising branch are training-only (:203-204, :230-235): their tensors are part of the
    state_dict and are loaded, never applied.  NOTE the reference model only runs when ``first_conv`` (its
    number of sinc filters) pools to 23 bands (69..71) -- with config/AASIST-Robust.conf's 128 it fails at
    ``e_S + pos_S`` for every input -- and for inputs of at least 560 641 samples (1025 taps at stride 256,
    seven 3x poolings); the same inputs raise ``RuntimeError`` here."""

    _kind = _lib.KIND_ROBUST
    _encoder_kind = _lib.ENC_RESIDUAL33

    def __init__(self, d_args: dict, precision: Optional[str] = None):
        super().__init__(d_args, precision)
        filts, gat_dims = d_args["filts"], d_args["gat_dims"]
        c = filts[-1][-1]
        self.encoder = _encoder(filts, _ResidualBlock33Params)               # AASIST_Robust.py:108-115
        self.first_bn = nn.BatchNorm2d(1)                                    # :118
        self.gaussian_noise = _GaussianNoiseParams()                         # :122
        self.denoising = _DenoisingParams(c)                                 # :125
        self.pos_S = nn.Parameter(torch.randn(1, 23, c))                     # :128
        self.GAT_layer_S = _GatParams(c, gat_dims[0])                        # :131
        self.GAT_layer_T = _GatParams(c, gat_dims[0])                        # :137
        self.master1 = nn.Parameter(torch.randn(1, 1, gat_dims[0]))          # :144
        self.master2 = nn.Parameter(torch.randn(1, 1, gat_dims[0]))          # :145 (unused by forward)
        self.HtrgGAT_layer_ST1 = _HtrgGatParams(gat_dims[0], gat_dims[1])    # :148
        self.HtrgGAT_layer_ST2 = _HtrgGatParams(gat_dims[1], gat_dims[1])    # :154
        self.pool_S = _PoolParams(gat_dims[0])                               # :161
        self.pool_T = _PoolParams(gat_dims[0])
        self.pool_hS = _PoolParams(gat_dims[1])
        self.pool_hT = _PoolParams(gat_dims[1])
        self.out_layer = _Linear(4 * gat_dims[1], 2)                         # :190
        self.aux_out_layer = _Linear(c, 2)                                   # :193
        self.ensemble_weight = nn.Parameter(torch.tensor([0.8, 0.2]))        # :196
        self.last_freq_mask: Optional[Tuple[int, int]] = None

    def forward(self, x: Tensor, Freq_aug: bool = False) -> Tuple[Tensor, Tensor]:
        if x.dim() == 4:                                                     # :207-214 input shapes
            x = x.squeeze(1).squeeze(1)
        if x.dim() == 1:
            x = x.unsqueeze(0)
        mask = None
        if Freq_aug:
            mask = self.last_freq_mask = draw_freq_mask(int(self.d_args["first_conv"]))
        return self._forward_native(x, mask, None)


class RawGATSTModel(_NativeModel):
    """Drop-in for reference ``models/RawNetGatSpoofST.py::Model`` (scoring forward)."""

    _kind = _lib.KIND_RAWGAT_ST

    def __init__(self, d_args: dict, precision: Optional[str] = None):
        super().__init__(d_args, precision)
        filts = d_args["filts"]
        self.first_bn = nn.BatchNorm2d(1)                       # RawNetGatSpoofST.py:291
        self.encoder_T = _encoder(filts)                        # :295-301
        self.encoder_S = _encoder(filts)                        # :303-309
        self.GAT_layer_T = _GatParams(64, 32)                   # :311-313
        self.GAT_layer_S = _GatParams(64, 32)
        self.GAT_layer_ST = _GatParams(32, 16)
        self.pool_T = _PoolParams(32)                           # :315-317
        self.pool_S = _PoolParams(32)
        self.pool_ST = _PoolParams(16)
        self.proj_T = _Linear(14, 12)                           # :319-322
        self.proj_S = _Linear(23, 12)
        self.proj_ST = _Linear(16, 1)
        self.out_layer = _Linear(7, 2)

    def forward(self, x: Tensor, Freq_aug: bool = False) -> Tuple[Tensor, Tensor]:
        """Returns ``(proj_ST (B,7), output (B,2))`` (RawNetGatSpoofST.py:356)."""
        mask = None
        if Freq_aug:                                            # RawNetGatSpoofST.py:326 conv_time(x, mask=Freq_aug)
            mask = self.last_freq_mask = draw_freq_mask(int(self.d_args["filts"][0]))
        return self._forward_native(x, mask, None)
