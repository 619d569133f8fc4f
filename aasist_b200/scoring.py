"""Callers of the hot path: the model factory and the eval-set scoring loop of the reference
driver, restated for one process per GPU.

* ``get_model``        -- reference main.py:251-259 (``import_module("models.<architecture>")``).
* ``score_utterances`` -- reference main.py:347-381 (``produce_evaluation_file``'s loop), sharded:
  utterances are independent in eval mode, so rank r scores the contiguous block
  ``[r*ceil(N/W), min(N,(r+1)*ceil(N/W)))`` and one all-gather over NCCL collects the scores
  (SURVEY 8(e)).  No other communication exists on this path.
* ``write_score_file`` -- reference main.py:383-387 (``"utt_id src key score"`` lines).
"""
from __future__ import annotations

from importlib import import_module
from typing import Callable, Iterable, Optional, Sequence, Tuple, Union

import torch
import torch.distributed as dist

Tensor = torch.Tensor


def get_model(model_config: dict, device: Union[str, torch.device], precision: Optional[str] = None):
    """Mirror of reference ``get_model`` (main.py:251-259)."""
    module = import_module("aasist_b200.models.{}".format(model_config["architecture"]))
    _model = getattr(module, "Model")
    model = _model(model_config, precision=precision).to(device)
    nb_params = sum([param.view(-1).size()[0] for param in model.parameters()])
    print("no. model params:{}".format(nb_params))
    return model


def shard_bounds(n_total: int, world_size: int, rank: int) -> Tuple[int, int, int]:
    """Contiguous block shard: returns (start, stop, per_rank) with per_rank = ceil(N/W)."""
    per = (n_total + world_size - 1) // world_size
    start = min(n_total, rank * per)
    return start, min(n_total, start + per), per


def score_utterances(model, source: Union[Tensor, Callable[[int, int], Tensor]], n_total: int,
                     batch_size: int = 512, group=None, device: Optional[torch.device] = None) -> Tensor:
    """Score ``n_total`` utterances; returns the (n_total,) fp32 bona-fide logits ``output[:,1]``
    (reference main.py:377) on ``device`` -- identical on every rank and for every world size.

    ``source`` is either an ``(n_total, L)`` tensor or a callable ``source(start, stop) -> (stop-start, L)``
    tensor, on the device or on the HOST.  Host batches go through the library's scoring stream
    (``aasist_score_begin/submit/finish``): pinned double-buffered staging, H2D of batch n+1 under the forward
    of batch n, one device->host hop at the very end instead of main.py:372-377's per-batch round trip.  With ``torch.distributed`` initialised, each rank scores its block and the scores
    are exchanged with ONE all-gather (NCCL on GPUs; gloo in the CPU tests).
    """
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    if device is None:
        device = next(model.parameters()).device
    start, stop, per = shard_bounds(n_total, world, rank)
    # the all-gather send buffer: every rank contributes `per` scores (the last rank's tail stays zero)
    local = torch.zeros(per, dtype=torch.float32, device=device)
    model.eval()
    first = None
    if stop > start:
        first = source(start, min(stop, start + batch_size)) if callable(source) else source[start:min(stop, start + batch_size)]
    host_pipeline = first is not None and not first.is_cuda and hasattr(model, "score_begin") and device.type == "cuda"
    with torch.no_grad():
        if host_pipeline:
            # HOST utterances: two staging buffers, the H2D of batch n+1 runs under the forward of batch n, scores
            # accumulate on the device and nothing waits until the end (replaces main.py:372-377's round trip)
            model.score_begin(stop - start, batch_size, first.shape[-1], device)
            for b0 in range(start, stop, batch_size):
                b1 = min(stop, b0 + batch_size)
                x = first if b0 == start else (source(b0, b1) if callable(source) else source[b0:b1])
                model.score_submit(x)
            local[:stop - start] = model.score_finish(on_device=True)[:, 1]
        else:
            for b0 in range(start, stop, batch_size):
                b1 = min(stop, b0 + batch_size)
                x = first if b0 == start else (source(b0, b1) if callable(source) else source[b0:b1])
                if x.device != device:
                    x = x.to(device, non_blocking=True)
                _, out = model(x)
                local[b0 - start:b1 - start] = out[:, 1]
    if not distributed:
        return local[:n_total]
    gathered = torch.empty(world * per, dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(gathered, local, group=group)
    return gathered[:n_total]


def write_score_file(save_path: str, utt_ids: Sequence[str], scores: Iterable[float],
                     trial_lines: Sequence[str]) -> None:
    """Reference main.py:382-387: one ``"utt_id src key score"`` line per trial."""
    # the reference formats Python floats (`batch_score.tolist()`, main.py:377-380): a Tensor / ndarray
    # (what score_utterances returns) is converted the same way, so "{}" never prints "tensor(...)"
    if isinstance(scores, torch.Tensor):
        scores = scores.detach().to(torch.float32).cpu().numpy().ravel().tolist()
    elif hasattr(scores, "ravel") and hasattr(scores, "tolist"):
        scores = scores.ravel().tolist()
    else:
        scores = [float(s) for s in scores]
    assert len(trial_lines) == len(utt_ids) == len(scores)
    with open(save_path, "w") as fh:
        for fn, sco, trl in zip(utt_ids, scores, trial_lines):
            _, utt_id, _, src, key = trl.strip().split(" ")
            assert fn == utt_id
            fh.write("{} {} {} {}\n".format(utt_id, src, key, sco))
