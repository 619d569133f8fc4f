"""Summarise an `ncu --page raw --csv` export: per launch duration, DRAM bytes, tensor-pipe / HBM utilisation.
    python tools/ncu_summarize.py raw.csv [--traffic-json out.json --model AASIST --batch 512 --names k1,k2,...]"""
import csv, json, sys, hashlib, os

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[0]
units = rows[1]
col = {n: i for i, n in enumerate(hdr)}


def get(r, name, default=0.0):
    i = col.get(name)
    if i is None:
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return default


want = [("ms", "gpu__time_duration.sum"), ("dram_rd", "dram__bytes_read.sum"), ("dram_wr", "dram__bytes_write.sum"),
        ("tensor%", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("tensor_op%", "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("sm%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l1%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("fp32%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("regs", "launch__registers_per_thread"), ("smem", "launch__shared_mem_per_block_dynamic")]
scale = {}
for key, name in want:
    i = col.get(name)
    u = units[i] if i is not None else ""
    scale[key] = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
out = []
print(f"{'kernel':44s} {'ms':>8s} {'dramMB':>9s} {'tensor%':>8s} {'dram%':>7s} {'sm%':>6s} {'l1%':>6s} {'fp32%':>6s} {'regs':>5s}")
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[col["Kernel Name"]]
    d = {k: get(r, n) * scale[k] for k, n in want}
    d["kernel"] = name
    d["grid"] = r[col.get("Grid Size", 0)]
    out.append(d)
    print(f"{name[:44]:44s} {d['ms']:8.3f} {(d['dram_rd'] + d['dram_wr']) / 1e6:9.1f} {d['tensor%']:8.1f} {d['dram%']:7.1f} "
          f"{d['sm%']:6.1f} {d['l1%']:6.1f} {d['fp32%']:6.1f} {int(d['regs']):5d}")
print(f"total ms {sum(d['ms'] for d in out):.3f}   total DRAM MB {sum(d['dram_rd'] + d['dram_wr'] for d in out) / 1e6:.1f}")
if "--traffic-json" in sys.argv:
    a = sys.argv
    dst, model, batch = a[a.index("--traffic-json") + 1], a[a.index("--model") + 1], int(a[a.index("--batch") + 1])
    names = a[a.index("--names") + 1].split(",")
    assert len(names) == len(out), (len(names), len(out))
    # "name:utterances" overrides --batch for one launch; launches with the same name (the two half-pass launches of
    # a kernel) are summed first
    merged = {}
    for n, d in zip(names, out):
        nm, _, nb = n.partition(":")
        m = merged.setdefault(nm, {"bytes": 0.0, "utts": 0})
        m["bytes"] += d["dram_rd"] + d["dram_wr"]
        m["utts"] += int(nb) if nb else batch
    try:
        tr = json.load(open(dst))
    except Exception:
        tr = {}
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from aasist_b200 import build as _build
    sha = _build._digest()[:16]                  # sources + flags of the build the capture was taken on
    if tr.get("source_digest16") != sha:
        tr = {"bytes_per_utterance": {}}
    tr["source_digest16"] = sha
    tr["source"] = (f"ncu --set full, {len(out)} launches of one forward over {batch} utterances ({os.path.basename(path)}), launches of the "
                    "same kernel summed; bytes = dram__bytes_read.sum + dram__bytes_write.sum")
    tr["bytes_per_utterance"][model] = {n: int(m["bytes"] / m["utts"]) for n, m in merged.items()}
    tr["bytes_per_utterance"][model]["_total"] = sum(tr["bytes_per_utterance"][model].values())
    json.dump(tr, open(dst, "w"), indent=1)
    print("wrote", dst)
