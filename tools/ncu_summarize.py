"""Turn ncu CSV dumps into the committed summaries under profiles/.

  python tools/ncu_summarize.py launches <launches.csv> <out.md> [steps]
      per-kernel launch count / total / share from `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --csv`
      (+ profiles/<stem>_traffic.json with the conv kernel's average DRAM bytes per launch when present)
  python tools/ncu_summarize.py full <raw.csv> <out.md>
      key metrics of one `ncu --set full` capture (`ncu -i rep --page raw --csv`)
"""
import collections
import csv
import json
import sys
from pathlib import Path


def launches(src, out, steps=None):
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= iv or not r[iid].isdigit():
            continue
        d = per.setdefault(int(r[iid]), {"name": r[ik].split("(")[0].replace("void ", "")})
        d[r[im]] = float(r[iv].replace(",", ""))
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d["name"], {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1
        a["ns"] += d.get("gpu__time_duration.sum", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a["ns"] for a in agg.values())
    lines = [f"# ncu launch list summary ({Path(src).name})", "",
             "`gpu__time_duration.sum` per launch, serialised and cold-cache (ncu replays every kernel), so only the",
             "SHARES are comparable with the live CUDA-event numbers of bench.py.", "",
             "| kernel | launches | total us | avg us | share | avg DRAM MB/launch (rd+wr) |", "|---|---:|---:|---:|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        mb = (a["rd"] + a["wr"]) / a["n"] / 1e6
        lines.append(f"| `{k}` | {a['n']} | {a['ns']/1e3:.1f} | {a['ns']/a['n']/1e3:.1f} | {100*a['ns']/tot:.1f}% | "
                     f"{mb:.1f} |" if mb else
                     f"| `{k}` | {a['n']} | {a['ns']/1e3:.1f} | {a['ns']/a['n']/1e3:.1f} | {100*a['ns']/tot:.1f}% | - |")
    if steps:
        lines += ["", f"{steps} forward passes captured."]
    Path(out).write_text("\n".join(lines) + "\n")
    conv = {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0}          # every instantiation of the conv kernel template
    for k, a in agg.items():
        if k.split("<")[0].split("::")[-1] == "conv_tc_kernel":
            for f in conv:
                conv[f] += a[f]
    if conv["n"] and conv["rd"] + conv["wr"] > 0:
        t = {"kernel": "conv_tc_kernel", "launches": conv["n"], "dram_bytes_per_launch": (conv["rd"] + conv["wr"]) / conv["n"],
             "dram_read_bytes_per_launch": conv["rd"] / conv["n"], "dram_write_bytes_per_launch": conv["wr"] / conv["n"],
             "source": Path(src).name}
        Path(out).with_name(Path(out).stem + "_traffic.json").write_text(json.dumps(t, indent=1) + "\n")
    print("\n".join(lines))


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__inst_executed.sum"]


def full(src, out):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full summary ({Path(src).name})", ""]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines += [f"## `{name}`", "", "| metric | value | unit |", "|---|---:|---|"]
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"| {w} | {r[i]} | {units[i]} |")
        # warp-state breakdown: sampled stall reasons (smsp__pcsamp_warps_issue_stalled_*), share of all samples
        samp = {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    samp[h[len("smsp__pcsamp_warps_issue_stalled_"):]] = float(r[i].replace(",", ""))
                except ValueError:
                    pass
        tot = sum(samp.values())
        if tot > 0:
            lines += ["", "warp stall samples (all warps of the kernel; producer / MMA / epilogue roles pooled):", "",
                      "| stall reason | samples | share |", "|---|---:|---:|"]
            for k, v in sorted(samp.items(), key=lambda kv: -kv[1]):
                if v > 0:
                    lines.append(f"| {k} | {int(v)} | {100 * v / tot:.1f}% |")
        lines.append("")
    Path(out).write_text("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        full(sys.argv[2], sys.argv[3])
