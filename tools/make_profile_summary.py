"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/.

    python tools/make_profile_summary.py launches gpurun_out/launches.csv profiles/r01_launches.txt
    python tools/make_profile_summary.py kernel   gpurun_out/prof.ncu-rep profiles/r01_gemm_qkv   [kernel-index]
"""
import csv, json, os, re, subprocess, sys, collections


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    agg, tot = collections.OrderedDict(), 0.0
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us; tot += us
    with open(dst, "w") as out:
        out.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, one bench step (tools/profile_step.py)\n")
        out.write(f"# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        out.write(f"total {tot/1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches\n")
        for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write(f"{us/1e3:9.3f} ms  {100*us/tot:5.1f}%  n={c:6d}  avg={us/c:9.2f} us  {name[:120]}\n")
    print(open(dst).read())


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "sm__cycles_elapsed.max"]


def kernel(src, dst_prefix, idx=0):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    r = data[idx]
    out = {"kernel": r[hdr.index("Kernel Name")], "source": os.path.basename(src), "launch_index": idx}
    for h, u, v in zip(hdr, units, r):
        if any(h.endswith(k) or h == k for k in KEYS):
            out[h.split("TriageCompute.")[-1] + (f" [{u}]" if u else "")] = v
    rd, wr = None, None
    for h, u, v in zip(hdr, units, r):
        if h == "dram__bytes_read.sum":
            rd = float(v) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        if h == "dram__bytes_write.sum":
            wr = float(v) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(u, 1)
    out["dram_bytes_per_launch"] = (rd or 0) + (wr or 0)
    src_csv = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src_csv.splitlines()))
    blocks, cur = [], None
    for rr in rows:
        if rr and rr[0] == "Kernel Name":
            cur = []; blocks.append(cur)
        elif cur is not None:
            cur.append(rr)
    top = []
    if blocks:
        b = blocks[min(idx, len(blocks) - 1)]
        h2 = b[0]; d2 = [x for x in b[1:] if len(x) == len(h2)]
        ix = {k: i for i, k in enumerate(h2)}
        tot = sum(int(x[ix["# Samples"]] or 0) for x in d2) or 1
        stalls = [k for k in h2 if k.startswith("stall_") and "Not Issued" not in k]
        for x in sorted(d2, key=lambda x: -int(x[ix["# Samples"]] or 0))[:12]:
            s = int(x[ix["# Samples"]] or 0)
            st = max(((int(x[ix[k]] or 0), k) for k in stalls))
            top.append(f"{100*s/tot:5.1f}%  {x[ix['Source']][:80]:80s} {st[1]}")
        sass = " ".join(x[ix["Source"]] for x in d2)
        out["sass_has"] = {m: (m in sass) for m in ("UTCHMMA", "UTMALDG", "LDTM", "HMMA.16816", "SYNCS", "UTCBAR")}
    out["top_stall_instructions"] = top
    with open(dst_prefix + ".json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
