"""Host-side mirror of the reference's model plug-in interface.

``Model(d_args)`` has the constructor, ``forward(x, Freq_aug=False, speaker_embedding=None)``
signature, return tuple and ``state_dict`` key layout of the reference's
``models/AASIST.py::Model`` (reference models/AASIST.py:728-921, with the checkpoint's (2,3)
``Residual_block`` encoder, models/RawNetGatSpoofST.py:225-278) -- so
``model.load_state_dict(torch.load("models/weights/AASIST.pth"))`` works unchanged -- but
owns no compute: parameters are plain ``nn.Parameter`` containers and ``forward`` hands raw
device pointers to libaasist_b200.so.  PyTorch is only the tensor container / stream owner.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

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


def _encoder(filts) -> nn.Sequential:
    # same nn.Sequential(nn.Sequential(block)) nesting as AASIST.py:766-772 -> keys encoder.N.0.*
    return nn.Sequential(
        nn.Sequential(_ResidualBlockParams(filts[1], first=True)),
        nn.Sequential(_ResidualBlockParams(filts[2])),
        nn.Sequential(_ResidualBlockParams(filts[3])),
        nn.Sequential(_ResidualBlockParams(filts[4])),
        nn.Sequential(_ResidualBlockParams(filts[4])),
        nn.Sequential(_ResidualBlockParams(filts[4])))


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
        blocks = [filts[1], filts[2], filts[3], filts[4], filts[4], filts[4]]
        for i, (ci, co) in enumerate(blocks):
            cfg.enc_channels[i][0], cfg.enc_channels[i][1] = int(ci), int(co)
        if self._kind == _lib.KIND_AASIST:
            cfg.gat_dims[0], cfg.gat_dims[1] = int(d["gat_dims"][0]), int(d["gat_dims"][1])
            for i in range(4):
                cfg.pool_ratios[i] = float(d["pool_ratios"][i])
                cfg.temperatures[i] = float(d["temperatures"][i])
        cfg.sample_rate = 16000
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

    def _forward_native(self, x: Tensor) -> Tuple[Tensor, Tensor]:
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
        x = x.detach().to(torch.float32).contiguous()
        B, L = x.shape
        with torch.cuda.device(dev):
            nbytes = _lib.check(lib.aasist_workspace_bytes(self._handle, B, L))
            ws = self._workspace_for(int(nbytes), dev)
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
            _lib.check(lib.aasist_forward(self._handle, x.data_ptr(), B, L, last_hidden.data_ptr(),
                                          output.data_ptr(), topk_ptr, scores_ptr, ws.data_ptr(),
                                          ws.numel(), stream))
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
    """Drop-in for reference ``models/AASIST.py::Model`` (scoring forward)."""

    _kind = _lib.KIND_AASIST

    def __init__(self, d_args: dict, precision: Optional[str] = None):
        super().__init__(d_args, precision)
        filts, gat_dims = d_args["filts"], d_args["gat_dims"]
        if d_args.get("speaker_conditioning", False):
            raise NotImplementedError("speaker conditioning (AASIST.py:743-755) is outside the scoring "
                                      "path: no shipped weights use it")
        self.first_bn = nn.BatchNorm2d(1)                                    # AASIST.py:760
        self.encoder = _encoder(filts)                                       # :766-772 (2,3 blocks)
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

    def forward(self, x: Tensor, Freq_aug: bool = False,
                speaker_embedding: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        """Returns ``(last_hidden (B, 5*gat_dims[1]), output (B, 2))`` (AASIST.py:921)."""
        if Freq_aug:
            raise NotImplementedError("Freq_aug filter masking (AASIST.py:486-490) is training-only")
        if speaker_embedding is not None:
            raise NotImplementedError("speaker conditioning is outside the scoring path")
        return self._forward_native(x)


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
        if Freq_aug:
            raise NotImplementedError("Freq_aug filter masking is training-only")
        return self._forward_native(x)
