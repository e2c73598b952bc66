"""Summaries from the CSV pages of an ncu report (exported on the GPU box: the .ncu-rep itself is too large to travel back).

    python scripts/ncu_csv_report.py raw   RAW.csv OUT.md [TRAFFIC.json]     per-launch table (duration, DRAM bytes, throughputs, occupancy) and per-kernel DRAM traffic
    python scripts/ncu_csv_report.py source SRC.csv OUT.md [title]           instruction mix, stall reasons and hot source lines of ONE kernel (ncu --page source --csv --print-source cuda,sass)
"""
import csv
import json
import sys
from collections import Counter


def short(name):
    return name.split("(")[0].replace("void ", "").replace("is::", "").strip()


def raw(path, out_md, out_json=None):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    col = {h: i for i, h in enumerate(hdr)}
    want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
            ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"), ("launch__registers_per_thread", "regs"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
    lines = ["| " + " | ".join(n for _, n in want) + " |", "|" + "---|" * len(want)]
    traffic = {}
    for r in data:
        if len(r) <= col["Kernel Name"]:
            continue
        vals = []
        for key, _ in want:
            v = r[col[key]] if key in col else ""
            u = units[col[key]] if key in col else ""
            if key == "Kernel Name":
                v = short(v)[:56]
            elif v:
                try:
                    f = float(v.replace(",", ""))
                    v = f"{f:.3f} {u}" if u not in ("", "%") else f"{f:.1f}"
                except ValueError:
                    pass
            vals.append(v)
        lines.append("| " + " | ".join(vals) + " |")
        b = sum(float(r[col[k]].replace(",", "")) * scale.get(units[col[k]], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        t = float(r[col["gpu__time_duration.sum"]].replace(",", "")) * tscale.get(units[col["gpu__time_duration.sum"]], 1e-3)
        a = traffic.setdefault(short(r[col["Kernel Name"]]), {"launches": 0, "dram_bytes": 0.0, "ms": 0.0})
        a["launches"] += 1; a["dram_bytes"] += b; a["ms"] += t
    open(out_md, "w").write("\n".join(lines) + "\n")
    if out_json:
        json.dump({"steps": 1, "note": "one whole C2 step under ncu --set full --clock-control none (cold caches, serialised launches)", "kernels": traffic},
                  open(out_json, "w"), indent=1)
    tot = sum(a["dram_bytes"] for a in traffic.values())
    print(f"{len(data)} launches, DRAM traffic {tot / 1e9:.3f} GB")


def source(path, out_md, title=""):
    rows = list(csv.reader(open(path, errors="replace")))
    his = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    his = his[:1] + [len(rows) if len(his) < 2 else his[1]]   # the first kernel of the export only
    out = [f"# {title}", ""] if title else []
    sass, src = [], []
    seen = set()
    ix = None
    for n, hi in enumerate(his[:1]):
        h = rows[hi]
        ix = {k: i for i, k in enumerate(h)}             # "Source" appears twice: the later index (SASS text) wins
        end = his[1]
        for r in rows[hi + 1:end]:
            if len(r) != len(h):                          # a source line with commas / quotes in it: not needed for the totals
                continue
            if r[0] == "" and r[2].startswith("0x"):
                if r[2] not in seen:                      # an instruction is listed under every source line it is attributed to (inlining)
                    seen.add(r[2])
                    sass.append(r)
            elif r[0] != "" and r[2] == "-":
                src.append(r)
    if ix is None or "Instructions Executed" not in ix:
        open(out_md, "w").write("\n".join(out + ["(no source page in the export)"]) + "\n")
        return
    ie, isamp = ix["Instructions Executed"], ix["# Samples"]
    num = lambda v: float(v.replace(",", "")) if v not in ("", "-") else 0.0   # noqa: E731
    ti = sum(num(r[ie]) for r in sass) or 1.0
    ts = sum(num(r[isamp]) for r in sass) or 1.0
    c, sm = Counter(), Counter()
    for r in sass:
        t = r[3].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        c[op] += num(r[ie]); sm[op] += num(r[isamp])
    out += [f"SASS instructions {len(sass)}, executed warp instructions {ti:.0f}, stall samples {ts:.0f}", "",
            "instruction mix (inst % / sample %): " + "  ".join(f"{op} {v / ti * 100:.1f}/{sm[op] / ts * 100:.1f}" for op, v in c.most_common(16)), ""]
    st = [k for k in ix if k.startswith("stall_") and "Not Issued" not in k]
    tot = {k: sum(num(r[ix[k]]) for r in sass) for k in st}
    out += ["stall reasons (% of samples): " + "  ".join(f"{k[6:]} {v / ts * 100:.1f}" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:9]), "",
            "| line | inst % | sample % | source |", "|---|---|---|---|"]
    for r in src:
        a_, b_ = num(r[ie]) / ti * 100, num(r[isamp]) / ts * 100
        if a_ >= 2.0 or b_ >= 2.5:
            out.append(f"| {r[0]} | {a_:.1f} | {b_:.1f} | `{r[1].strip()[:110].replace('|', '/')}` |")
    open(out_md, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:8]))


if __name__ == "__main__":
    if sys.argv[1] == "raw":
        raw(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        source(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
