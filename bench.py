"""bench.py -- AASIST batched utterance scoring on B200: utterances/sec, roofline, CPU baseline.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference --steps K --warmup W    (CPU oracle port on the host cores)

A "step" is one pass of the hot path (Model.forward) over one batch of synthetic 4 s / 16 kHz
waveforms per GPU.  Workload at N=1 = BASELINE.json configs[1]: AASIST (config/AASIST.conf,
models/weights/AASIST.pth), batch 512, L=64600.  N>1: every rank scores its own 512-utterance
shard; the bona-fide scores accumulate on the device and ONE NCCL all-gather at the end of the timed
region collects them (weak scaling).

Measurement layout of the native arm:
  1. `value`   K forwards over inputs resident in HBM, timed with CUDA events, nothing else inside the region;
  2. `e2e`     the same K batches from pinned HOST memory through the scoring stream
               (aasist_score_begin/submit/finish: H2D of batch n+1 under the forward of batch n), with the final
               device->host read of the scores inside the region;
  3. an UNTIMED pass with per-launch CUDA events for the per-kernel table and the roofline of the dominant kernel;
  4. baselines: the CPU oracle port (AASIST sample + BASELINE C1 = AASIST-L batch 24) and the oracle run eagerly
     on the GPU in fp32 (the reference's real deployment; TF32 off).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L_SAMPLES = 64600
# algorithmic FLOPs per utterance (2 x MAC of the reference fp32 ops, SURVEY 8(d) / BASELINE.md 4)
FLOPS_PER_UTT = {"AASIST": 19.1241e9, "AASIST-L": 13.2063e9, "RawGAT-ST": 37.044e9}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--model", default="AASIST", choices=["AASIST", "AASIST-L", "RawGAT-ST", "AASIST2"],
                    help="AASIST2 = the fork's Res2Net+SE model (config/AASIST2.conf; seeded weights, no checkpoint exists)")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--samples", type=int, default=64600,
                    help="utterance length (SURVEY 8(d) C5 length sweep; the headline metric is quoted at 64600)")
    ap.add_argument("--precision", default=None,
                    help="fp32 | f16x3 (default) | f16x2 (opt-in reduced-product mode; its error against the reference "
                         "goldens is measured and printed in the line)")
    ap.add_argument("--workload", default="batch", choices=["batch", "evalset"],
                    help="batch: BASELINE configs[1] (default, the contract line); evalset: configs[2], one pass over "
                         "71,237 synthetic utterances sharded across the ranks with one score all-gather")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    args = ap.parse_args()
    global L_SAMPLES
    L_SAMPLES = args.samples
    return args


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (the reference is pure Python/torch, nothing compiles into oracle/_ref)
# ------------------------------------------------------------------------------------------------
def gpu_eager_throughput(model_name: str, dev, batch: int = 64, repeats: int = 3):
    """The oracle's functional forward executed by torch on the GPU in fp32 with TF32 disabled -- what the
    reference's own `main.py --eval` does on a CUDA device (cuDNN / cuBLAS kernels, ~250 launches per forward)."""
    import torch
    from oracle import aasist_oracle as O
    import aasist_b200
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = {k: v.to(dev) for k, v in torch.load(aasist_b200.weights_path(model_name), map_location="cpu").items()}
    cfg, fwd = _oracle_forward(model_name)
    x = O.white_noise(batch, L_SAMPLES, 1234).to(dev)
    bank = O.sinc_filterbank(cfg["filts"][0], cfg["first_conv"]).to(dev)
    fwd(sd, x, bank)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(repeats):
        fwd(sd, x, bank)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / repeats
    del sd, x
    torch.cuda.empty_cache()
    return batch / (ms * 1e-3), ms


def _oracle_forward(model_name: str):
    """(config, forward(sd, x, bank)) of the CPU/GPU-eager oracle for a bench model."""
    from oracle import aasist_oracle as O
    if model_name == "AASIST2":
        from oracle import aasist2_oracle as O2
        cfg = O2.CONFIGS[model_name]
        return cfg, lambda sd, x, bank: O2.aasist2_forward(sd, cfg, x, None, None, bank)
    cfg = O.CONFIGS[model_name]
    return cfg, lambda sd, x, bank: O.forward(model_name, sd, cfg, x, None, bank)


def cpu_oracle_throughput(model_name: str, n_utt: int, repeats: int, warmup: int = 1):
    import torch
    from oracle import aasist_oracle as O
    import aasist_b200
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    sd = torch.load(aasist_b200.weights_path(model_name), map_location="cpu")
    cfg, fwd = _oracle_forward(model_name)
    x = O.white_noise(n_utt, L_SAMPLES, 1234)
    bank = O.sinc_filterbank(cfg["filts"][0], cfg["first_conv"])
    times = []
    for i in range(warmup + repeats):
        t0 = time.perf_counter()
        fwd(sd, x, bank)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return n_utt / min(times), n_utt * repeats / sum(times), cores, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: size each step so that the whole run stays within a few minutes
    probe_rate, _, cores, _ = cpu_oracle_throughput(args.model, 4, 1, warmup=1)
    budget_s = 150.0
    n_steps = args.steps + args.warmup
    per_step = int(max(2, min(24, budget_s * probe_rate / max(1, n_steps))))
    best, mean, cores, times = cpu_oracle_throughput(args.model, per_step, args.steps, warmup=args.warmup)
    ms = 1e3 * sum(times) / len(times)
    line = {
        "impl": "reference", "metric": "AASIST utterances/sec (4 s, 64600 samples)", "value": mean,
        "unit": "utt/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model} (config/{args.model}.conf + shipped weights) eval forward, "
                               f"L={L_SAMPLES}; CPU sample of {per_step} utterances per step "
                               f"(the GPU arm's step is {args.batch} utterances per GPU)",
                   "model_name": args.model, "batch_per_step": per_step, "samples": L_SAMPLES},
        "cpu_baseline": {"value": mean, "unit": "utt/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} utterances x {args.steps} steps, torch {cores} threads, "
                                   "oracle/aasist_oracle.py (functional restatement of the reference forward)"},
        "e2e": {"value": mean, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                    power.append(float(f[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        return out


def run_evalset(args, model, dev, world, rank):
    """BASELINE configs[2]: eval-set-sized scoring (71,237 utterances = ASVspoof2019-LA eval count), contiguous
    block shards, local batches of 512, ONE all-gather of the scores.  Utterances are generated on device in
    fixed chunks of 512 seeded 1234 + chunk_id, so the content does not depend on the world size."""
    import torch
    import torch.distributed as dist
    from aasist_b200.scoring import score_utterances
    n_total, chunk = 71237, 512

    from aasist_b200.scoring import shard_bounds
    lo_r, hi_r, _ = shard_bounds(n_total, world, rank)
    # the rank's whole shard is generated BEFORE the timed region and stays resident (18.4 GB at N=1)
    shard = torch.empty(hi_r - lo_r, L_SAMPLES, device=dev)
    for cid in range(lo_r // chunk, (hi_r - 1) // chunk + 1):
        g = torch.Generator(device=dev).manual_seed(1234 + cid)
        xc = 0.05 * torch.randn(chunk, L_SAMPLES, device=dev, generator=g)
        lo, hi = max(lo_r, cid * chunk), min(hi_r, (cid + 1) * chunk)
        shard[lo - lo_r:hi - lo_r] = xc[lo - cid * chunk:hi - cid * chunk]
    del xc

    def source(start, stop):
        return shard[start - lo_r:stop - lo_r]

    with torch.no_grad():
        for _ in range(2):
            model(shard[:chunk])                                                      # warm-up
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        launches0 = model.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        scores = score_utterances(model, source, n_total, batch_size=chunk)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item())
        _emit({
            "metric": "AASIST utterances/sec (4 s, 64600 samples)", "value": n_total / (ms * 1e-3), "unit": "utt/s",
            "n_gpus": world, "steps": 1, "warmup": 1, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f16x3(split)+f32acc" if model.precision != "fp32" else "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.model} eval-set-sized scoring: {n_total} utterances, L={L_SAMPLES}, contiguous "
                                   f"block shards over {world} GPU(s), local batch {chunk}, one all-gather of scores; "
                                   "the synthetic shard is generated on the device before the timed region",
                       "n_utterances": n_total, "precision": model.precision},
            "gpu_launches": int(model.launch_count() - launches0),
            "scores_checksum": float(scores.double().sum().item()), "scores_head": scores[:4].tolist()})
    if world > 1:
        dist.destroy_process_group()


def golden_parity(model, name: str, dev):
    """Error of `model` against the committed reference goldens (tests/golden/<name>_speech*.npz: logits and ordered
    GraphPool indices produced by the unmodified reference): max |logit error| and the fraction of utterances whose
    ordered top-k indices all equal the reference's.  Reported for any non-default precision mode."""
    import glob
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    from tests.util import golden_input, load_golden, pools_of
    worst, n_utt, n_ok, n_pos, n_pos_ok = 0.0, 0, 0, 0, 0
    model.record_topk = True
    try:
        for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", f"{name}_speech*.npz"))):
            tag = os.path.basename(path)[len(name) + 1:-4]
            gold, meta = load_golden(name, tag)
            x = golden_input(meta).to(dev)
            _, out = model(x)
            torch.cuda.synchronize()
            worst = max(worst, float(np.abs(out.cpu().numpy() - gold["output"]).max()))
            ref = np.concatenate([gold[p + ".idx"] for p in pools_of(name)], axis=1)
            got = model.last_topk.cpu().numpy()
            same = got == ref
            n_utt += same.shape[0]
            n_ok += int(same.all(axis=1).sum())
            n_pos += same.size
            n_pos_ok += int(same.sum())
    finally:
        model.record_topk = False
    return {"logits_max_abs_err": worst, "utterances": n_utt, "ordered_topk_match_rate": n_ok / max(1, n_utt),
            "topk_positions": n_pos, "topk_positions_equal": n_pos_ok,
            "fixtures": f"tests/golden/{name}_speech*.npz (reference-generated)",
            "ships": bool(worst <= 1e-3 and n_ok == n_utt)}


def _lib_sha16() -> str:
    """Identity of the build the numbers come from: the digest of the CUDA sources + compiler flags that
    aasist_b200/build.py stamps the library with (a hash of the binary would change with the checkout path that
    -lineinfo embeds).  profiles/ncu_traffic.json carries the same digest."""
    try:
        from aasist_b200 import build as _build
        return _build._digest()[:16]
    except Exception:
        return ""


def run_native(args):
    import torch
    import torch.distributed as dist

    import aasist_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        raise SystemExit("for --gpus N>1 launch with: python -m torch.distributed.run --nnodes=1 "
                         "--nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    name, B = args.model, args.batch
    cls = aasist_b200.RawGATSTModel if name == "RawGAT-ST" else aasist_b200.Model
    model = cls(aasist_b200.CONFIGS[name], precision=args.precision)
    model.load_state_dict(torch.load(aasist_b200.weights_path(name), map_location="cpu"), strict=True)
    model = model.to(dev).eval()
    precision = model.precision

    if args.workload == "evalset":
        return run_evalset(args, model, dev, world, rank)

    # synthetic 4 s waveforms, already resident in HBM for the device-timed region.
    # 512 x 64600 fp32 = 132 MB per GPU > the 126 MB L2, and every intermediate is far larger,
    # so no timed iteration can be served from L2.
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = 0.05 * torch.randn(B, L_SAMPLES, device=dev, generator=g)
    K = args.steps
    W = max(3, args.warmup)
    local = torch.empty(max(K, W) * B, device=dev)               # this rank's scores of the timed region (the warm-up reuses it)
    gathered = torch.empty(world * max(K, W) * B, device=dev) if world > 1 else None

    def run_steps(k):
        for i in range(k):
            _, out = model(x)
            local[i * B:(i + 1) * B] = out[:, 1]
        if world > 1:                                            # ONE all-gather at the end of the shard
            dist.all_gather_into_tensor(gathered, local)

    with torch.no_grad():
        run_steps(W)
        torch.cuda.synchronize()
        launches0 = model.launch_count()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        run_steps(K)
        ev1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        elapsed_ms = ev0.elapsed_time(ev1)
        clocks = sampler.stop() if rank == 0 else None
        launches = model.launch_count() - launches0
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())

        # end to end through the public scoring loop with HOST buffers: pinned host batch -> staging -> H2D on the
        # copy stream -> forward, K times without waiting (the H2D of batch n+1 runs under the forward of batch n),
        # then ONE wait, the score all-gather (N>1) and the device->host read of the scores
        # (replaces reference main.py:364-378, which round-trips every batch)
        e2e = None
        if not args.no_e2e:
            xh = x.cpu().pin_memory()
            KE = max(K, 2)                                        # the e2e warm-up submits two batches
            host_scores = torch.empty(world * KE * B).pin_memory()

            def run_e2e(k):
                model.score_begin(KE * B, B, L_SAMPLES, dev)     # same capacity in the warm-up: no allocation when timed
                for _ in range(k):
                    model.score_submit(xh)
                out = model.score_finish(on_device=True)
                sc = out[:, 1].contiguous()
                if world > 1:
                    dist.all_gather_into_tensor(gathered[:world * k * B], sc)
                    sc = gathered[:world * k * B]
                host_scores[:sc.numel()].copy_(sc, non_blocking=True)

            run_e2e(2)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run_e2e(K)
            e1.record()
            torch.cuda.synchronize()
            te = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            e2e = {"value": world * B * K / (float(te.item()) * 1e-3), "unit": "utt/s",
                   "h2d_bytes_per_step": B * L_SAMPLES * 4, "d2h_bytes_per_step": world * B * 4,
                   "steps": K, "ms_per_step": float(te.item()) / K,
                   "api": "aasist_b200.Model.score_begin/score_submit/score_finish (C ABI aasist_score_*), pinned host "
                          "input, scores read back to the host once at the end of the timed region"}

        # UNTIMED pass with per-launch CUDA events (on the launching stream) for the per-kernel table
        n_prof = max(2, min(K, 5))
        model.profile(True)
        model.profile_report(reset=True)
        for _ in range(n_prof):
            model(x)
        prof = model.profile_report(reset=True)
        model.profile(False)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * B * K / (elapsed_ms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # roofline of the dominant kernel (largest share of the profiled pass, CUDA events per launch)
    prof.sort(key=lambda r: -r["ms"])
    total_kernel_ms = sum(r["ms"] for r in prof) or 1.0
    top = prof[0] if prof else None
    roofline = None
    from aasist_b200 import workmodel
    if top is not None:
        flops_step = workmodel.kernel_flops(name, top["kernel"], B, L_SAMPLES)   # algorithmic FLOPs / step
        n_launch = max(1, top["launches"] // n_prof)                              # launches of it per step
        avg_launch_ms = top["ms"] / max(1, top["launches"])
        peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
        peak_src = ("MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks
                    else "fallback 1.4 PF sustained (B200_PROFILING.md)")
        ach = (flops_step / n_launch) / (avg_launch_ms * 1e-3) / 1e12
        # ncu DRAM bytes of this kernel: only from a capture of THIS build (the table records the library hash)
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            per_utt = tr.get("bytes_per_utterance", {}).get(name, {})
            if tr.get("source_digest16") == _lib_sha16() and top["kernel"] in per_utt:
                traffic = int(per_utt[top["kernel"]] * (B / n_launch))
                traffic_src = tr.get("source")
            elif top["kernel"] in per_utt:
                traffic_src = "stale: profiles/ncu_traffic.json was captured on another build (%s)" % tr.get("source_digest16")
        except Exception:
            pass
        # tensor-pipe products per reference MAC: 3 everywhere in f16x3; f16x2 keeps 3 in the sinc stage and block 0
        executed = 1.0 if precision == "fp32" else (
            2.0 if precision == "f16x2" and not top["kernel"].startswith(("enc0", "sinc")) else 3.0)
        roofline = {"bound": "tensor", "kernel": top["kernel"], "achieved": ach, "peak": peak_tf,
                    "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic, "traffic_source": traffic_src,
                    "executed_tflops": executed * ach, "executed_frac": executed * ach / peak_tf,
                    "peak_source": peak_src, "share_of_step": top["ms"] / total_kernel_ms,
                    "flops_per_launch": flops_step / n_launch, "launches_per_step": n_launch,
                    "avg_launch_ms": avg_launch_ms,
                    "note": "achieved = ALGORITHMIC FLOPs per launch (2xMAC of the reference fp32 ops, "
                            "aasist_b200/workmodel.py) / mean CUDA-event launch time, measured in a separate untimed "
                            "pass; the f16x3 path executes 3 tcgen05 MMAs per reference MAC (f16x2: 2 in blocks "
                            "1-5), so tensor-pipe work is that multiple of this figure (executed_tflops / executed_frac)"}
    line = {
        "metric": "AASIST utterances/sec (4 s, 64600 samples)", "value": value, "unit": "utt/s",
        "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if precision == "fp32" else f"{precision}(split)+f32acc",
        "data": "synthetic",
        "config": {"workload": f"{name} (config/{name}.conf, {aasist_b200.WEIGHTS[name]}) eval scoring forward, "
                               f"batch {B} per GPU, L={L_SAMPLES}",
                   "model_name": name, "batch_per_gpu": B, "global_batch": B * world, "samples": L_SAMPLES,
                   "precision": precision,
                   "parallelism": f"utterance-sharded dp{world}, one NCCL all-gather of the scores at the end of the "
                                  "timed region",
                   "l2_policy": "inputs (132 MB/GPU) and every intermediate exceed the 126 MB L2"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": roofline,
        "kernels": [{"kernel": r["kernel"], "launches": r["launches"], "ms_per_step": r["ms"] / n_prof,
                     "share": r["ms"] / total_kernel_ms} for r in prof],
        "kernels_note": f"per-launch CUDA events from a separate untimed pass of {n_prof} steps",
        "algorithmic_tflops": value * (FLOPS_PER_UTT[name] if L_SAMPLES == 64600 and name in FLOPS_PER_UTT else
                                       2.0 * sum(workmodel.stage_macs(name, L_SAMPLES).values())) / 1e12,
        "source_digest16": _lib_sha16(),
    }
    if precision != aasist_b200.model.DEFAULT_PRECISION and name in ("AASIST", "AASIST-L"):
        line["parity_vs_reference_goldens"] = golden_parity(model, name, dev)
    if not args.no_eager_baseline:
        try:
            rate, ms = gpu_eager_throughput(name, dev)
            line["gpu_eager_baseline"] = {
                "value": rate, "unit": "utt/s", "batch": 64, "ms_per_batch": ms,
                "what": "oracle/aasist_oracle.py forward executed eagerly by torch on this GPU, fp32, TF32 disabled "
                        "(cuDNN/cuBLAS): the reference's own deployment (main.py requires CUDA), batch 64"}
        except Exception as e:                                    # never lose the line over a baseline
            line["gpu_eager_baseline"] = {"error": str(e)[:200]}
    if not args.no_cpu_baseline and world >= 1:
        n_cpu = 8
        best, mean, cores, times = cpu_oracle_throughput(name, n_cpu, 2, warmup=1)
        line["cpu_baseline"] = {"value": best, "unit": "utt/s", "cores": cores, "kind": "port",
                                "sample": f"{n_cpu} utterances, best of 2 after 1 warm-up, torch CPU fp32, "
                                          f"{cores} threads (oracle/aasist_oracle.py)"}
        if L_SAMPLES == 64600:
            # BASELINE.json configs[0] / BASELINE.md section 3 (C1): AASIST-L forward on CPU, batch 24
            b1, m1, cores, _ = cpu_oracle_throughput("AASIST-L", 24, 2, warmup=1)
            line["cpu_baseline"]["c1_aasist_l_batch24"] = {
                "value": b1, "unit": "utt/s", "cores": cores,
                "sample": "AASIST-L, batch 24, L=64600, best of 2 after 1 warm-up (BASELINE configs[0])"}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def _emit(obj) -> None:
    """The contract is ONE json line on stdout; native libraries (NCCL prints its version) write to fd 1
    directly, so fd 1 is pointed at stderr for the whole run and the result goes to the saved descriptor."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)

if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
