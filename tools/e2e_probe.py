"""GPU-side probe: where does the end-to-end (host -> scores) time go?  python tools/e2e_probe.py [K]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aasist_b200

K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")
m = aasist_b200.Model(aasist_b200.CONFIGS["AASIST"])
m.load_state_dict(torch.load(aasist_b200.weights_path("AASIST"), map_location="cpu"))
m = m.to(dev).eval()
B, L = 512, 64600
x = 0.05 * torch.randn(B, L, device=dev)
xh = x.cpu().pin_memory()
xd = torch.empty_like(x)


def timed(fn, reps=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    host_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, host_ms / reps


with torch.no_grad():
    for _ in range(3):
        m(x)
    ms, _ = timed(lambda: xd.copy_(xh, non_blocking=True), 5)
    print(f"H2D 132 MB pinned: {ms:.2f} ms = {B * L * 4 / ms / 1e6:.1f} GB/s")
    ms, hms = timed(lambda: m(x), K)
    print(f"forward resident: {ms:.2f} ms/step (host enqueue {hms:.2f} ms/step)")

    def serial():
        xd.copy_(xh, non_blocking=True)
        m(xd)
    ms, hms = timed(serial, K)
    print(f"copy + forward, one stream: {ms:.2f} ms/step (host {hms:.2f})")

    def pipe(src):
        m.score_begin(K * B, B, L, dev)
        for _ in range(K):
            m.score_submit(src)
        return m.score_finish(on_device=True)
    pipe(xh)
    ms, hms = timed(lambda: pipe(xh))
    print(f"scoring stream, pinned source: {ms / K:.2f} ms/step (host total {hms:.1f} ms)")
    xp = xh.clone()          # pageable
    pipe(xp)
    ms, hms = timed(lambda: pipe(xp))
    print(f"scoring stream, pageable source: {ms / K:.2f} ms/step (host total {hms:.1f} ms)")
    ms, hms = timed(lambda: m.score_host(xh), K)
    print(f"forward_host (r1 path): {ms:.2f} ms/step")
    print("---- second round (order check)")
    ms, hms = timed(lambda: m.score_host(xh), K)
    print(f"forward_host again: {ms:.2f} ms/step")
    for nb in (128, 384, 512):
        xs = x[:nb].contiguous()
        m(xs)
        ms, hms = timed(lambda: m(xs), K)
        print(f"forward resident B={nb}: {ms:.2f} ms/step")
    ms, hms = timed(lambda: pipe(xh))
    print(f"scoring stream, pinned source again: {ms / K:.2f} ms/step")
    import subprocess
    print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout)
