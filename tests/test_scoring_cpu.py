"""Host-side scoring loop: sharding arithmetic, score-file format, and the world_size-2
all-gather path over gloo (CPU) with a stand-in model."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aasist_b200.scoring import score_utterances, shard_bounds, write_score_file


class _StandIn(torch.nn.Module):
    """Deterministic per-utterance 'model' with the plug-in forward signature."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.tensor([0.5]))

    def forward(self, x, Freq_aug=False, speaker_embedding=None):
        s = (x * x).sum(dim=1) * self.w
        return x[:, :5], torch.stack([-s, s], dim=1)


def test_shard_bounds_cover_everything_once():
    for n in (1, 7, 512, 71237):
        for w in (1, 2, 4, 8):
            seen = []
            for r in range(w):
                a, b, per = shard_bounds(n, w, r)
                assert per == -(-n // w) and b - a <= per
                seen += list(range(a, b))
            assert seen == list(range(n))
    assert shard_bounds(71237, 8, 7) == (62335, 71237, 8905)


def test_single_process_scores_in_order():
    x = torch.randn(11, 32)
    s = score_utterances(_StandIn(), x, 11, batch_size=4, device=torch.device("cpu"))
    assert torch.allclose(s, (x * x).sum(1) * 0.5)
    s2 = score_utterances(_StandIn(), lambda a, b: x[a:b], 11, batch_size=3, device=torch.device("cpu"))
    assert torch.equal(s, s2)


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, 16, generator=g)
    s = score_utterances(_StandIn(), x, n, batch_size=3, device=torch.device("cpu"))
    q.put((rank, s.clone()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [9, 10])
def test_world_size_2_all_gather_is_shard_invariant(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, 16, generator=g)
    ref = score_utterances(_StandIn(), x, n, batch_size=4, device=torch.device("cpu"))
    assert torch.equal(got[0], ref) and torch.equal(got[1], ref)    # byte-identical for W=1 and W=2


def test_score_file_format(tmp_path):
    trials = ["LA_0001 LA_E_1 - A07 spoof\n", "LA_0002 LA_E_2 - - bonafide\n"]
    p = tmp_path / "scores.txt"
    write_score_file(str(p), ["LA_E_1", "LA_E_2"], [-1.5, 2.25], trials)
    assert p.read_text() == "LA_E_1 A07 spoof -1.5\nLA_E_2 - bonafide 2.25\n"


def test_score_file_accepts_the_tensor_score_utterances_returns(tmp_path):
    """a Tensor / ndarray of scores must print as plain floats, like the reference's
    `batch_score.tolist()` (main.py:377-380), so that calculate_tDCF_EER can parse the file."""
    import numpy as np
    trials = ["LA_0001 LA_E_1 - A07 spoof\n", "LA_0002 LA_E_2 - - bonafide\n"]
    x = torch.tensor([[1.0, 1.0], [1.5, 0.0]])
    scores = score_utterances(_StandIn(), x, 2, batch_size=1, device=torch.device("cpu"))
    assert isinstance(scores, torch.Tensor)
    for s in (scores, scores.numpy(), scores.double().numpy()):
        p = tmp_path / "scores.txt"
        write_score_file(str(p), ["LA_E_1", "LA_E_2"], s, trials)
        assert p.read_text() == "LA_E_1 A07 spoof 1.0\nLA_E_2 - bonafide 1.125\n"
        parsed = np.genfromtxt(str(p), dtype=str)           # what evaluation.py:26 does with the file
        assert parsed[:, 3].astype(np.float64).tolist() == [1.0, 1.125]
