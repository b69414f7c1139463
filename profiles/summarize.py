#!/usr/bin/env python
"""Turn raw Nsight Compute outputs (gpurun_out/, scratch) into the small text summaries that
are committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv  profiles/rNN_launches.txt
  python profiles/summarize.py full     gpurun_out/prof.ncu-rep  profiles/rNN_<kernel>_full.txt

`launches`: the CSV written by `ncu --metrics gpu__time_duration.sum --csv`: per-kernel launch
count, total device time and share of the profiled region (cold-cache, serialised launches:
compare SHARES, not absolutes).
`full`: key metrics of every launch in an `ncu --set full` report (tensor pipe, DRAM bytes,
L2 -> SM bytes, occupancy limits) plus the ten most-sampled SASS instructions of the first launch.
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg ",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second",
    "l1tex__m_xbar2l1tex_read_bytes.sum ", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread ",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "smsp__inst_executed.sum ",
]


def launches(src, dst):
    rows = list(csv.reader(open(src, errors="replace")))
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(d["Metric Value"].replace(",", ""))
            unit = d.get("Metric Unit", "ns")
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)  # -> us
            name = d["Kernel Name"].split("(")[0]
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += v
    tot = sum(v[1] for v in agg.values()) or 1.0
    with open(dst, "w") as f:
        f.write(f"# source: {src}  (ncu --metrics gpu__time_duration.sum --clock-control none)\n")
        f.write("# per-launch times under ncu are cold-cache and serialised: compare shares\n")
        f.write(f"{'kernel':48s} {'launches':>8s} {'total us':>12s} {'avg us':>9s} {'share':>7s}\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:48s} {n:8d} {us:12.1f} {us / n:9.2f} {100 * us / tot:6.1f}%\n")
        f.write(f"{'TOTAL':48s} {sum(v[0] for v in agg.values()):8d} {tot:12.1f}\n")


def traffic(src, dst, batch="16", note=""):
    """Per-family DRAM bytes and serialised time of ONE step from the CSV of
    `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` over the
    NVTX range of tools/one_step.py -> the JSON bench.py reads for roofline.traffic."""
    import json
    rows = list(csv.reader(open(src, errors="replace")))
    hdr = None
    fam = collections.OrderedDict()

    def family(name):
        if "gemm_tc" in name or "attn_flash" in name or "thin_in_conv" in name:
            return "gemm_tc"
        if "gn_apply" in name:
            return "gn_apply"
        return "other"

    launches_seen = collections.defaultdict(set)
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            name = d["Kernel Name"].split("(")[0]
            f_ = fam.setdefault(family(name), collections.defaultdict(float))
            v = float(d["Metric Value"].replace(",", ""))
            unit = d.get("Metric Unit", "")
            m = d["Metric Name"]
            if m.startswith("dram__bytes"):
                v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
                f_["dram_bytes_read" if "read" in m else "dram_bytes_write"] += v
            elif m == "gpu__time_duration.sum":
                v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
                f_["ncu_time_ms_serialised"] += v
                launches_seen[family(name)].add(d["ID"])
            f_.setdefault("kernels", 0)
    out = {"source": note or f"ncu over the NVTX range of tools/one_step.py; raw CSV: {src}",
           "batch_per_gpu": int(batch)}
    for k, f_ in fam.items():
        out[k] = {"launches_per_step": len(launches_seen[k]),
                  "dram_bytes_read": f_["dram_bytes_read"], "dram_bytes_write": f_["dram_bytes_write"],
                  "dram_bytes": f_["dram_bytes_read"] + f_["dram_bytes_write"],
                  "ncu_time_ms_serialised": f_["ncu_time_ms_serialised"]}
    json.dump(out, open(dst, "w"), indent=1)


def _ncu(args):
    return subprocess.run(["ncu"] + args, check=True, capture_output=True, text=True).stdout


def full(src, dst):
    raw = list(csv.reader(io.StringIO(_ncu(["-i", src, "--page", "raw", "--csv"]))))
    hdr, units, data = raw[0], raw[1], raw[2:]
    with open(dst, "w") as f:
        f.write(f"# source: {src}  (ncu --set full --clock-control none --import-source on)\n")
        name_i = hdr.index("Kernel Name")
        f.write(f"# {len(data)} launches: " + ", ".join(r[name_i].split('(')[0] for r in data) + "\n")
        for i, h in enumerate(hdr):
            if any(h == k.strip() or (k.endswith(" ") and h == k.strip()) for k in KEYS):
                f.write(f"{h:72s} {units[i]:16s} " + "  ".join(r[i] for r in data) + "\n")
        src_csv = _ncu(["-i", src, "--page", "source", "--csv", "--kernel-id", ":::1"])
        rows = list(csv.reader(io.StringIO(src_csv)))
        h2 = next((r for r in rows if "Source" in r and "# Samples" in r), None)
        if h2:
            isrc, isamp = h2.index("Source"), h2.index("# Samples")
            stall = [i for i, h in enumerate(h2) if h.startswith("stall_") and "Not Issued" not in h]
            body = [r for r in rows[rows.index(h2) + 1:] if len(r) == len(h2) and r[isamp].isdigit()]
            tot = sum(int(r[isamp] or 0) for r in body) or 1
            f.write(f"\n# top sampled SASS instructions of launch 1 ({tot} samples)\n")
            for r in sorted(body, key=lambda r: -int(r[isamp] or 0))[:12]:
                top = sorted(((int(r[i] or 0), h2[i]) for i in stall), reverse=True)[:2]
                f.write(f"{100 * int(r[isamp] or 0) / tot:5.1f}%  {r[isrc][:64]:64s} "
                        f"{top[0][1]}={top[0][0]} {top[1][1]}={top[1][0]}\n")


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
