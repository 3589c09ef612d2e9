"""Turns the ncu artefacts that scripts/profile.sh left in gpurun_out/ into the tracked summaries under profiles/:
  profiles/<tag>_launches.csv      per-launch device time of one bench step (ncu --metrics gpu__time_duration.sum)
  profiles/<tag>_kernels.md        per-kernel share of the step + key counters of the --set full captures
  profiles/r1_traffic.json         dram bytes per launch of the dominant kernel (read by bench.py roofline.traffic)
Runs here (no GPU): `python scripts/summarize_profiles.py <tag> <batch>`."""
import collections, csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "sm__cycles_elapsed.max"]


def short(name):
    name = name.replace("void ", "").replace("ttk::", "")
    return name.split("(CUtensorMap")[0].split("(")[0] if "gemm_kernel" not in name else name.split("(CUtensorMap")[0]


def raw_page(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


def main(tag, batch):
    os.makedirs(PROF, exist_ok=True)
    md = [f"# ncu summary `{tag}` -- `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (batch {batch} clips 3x16x168x168 per step, 1 x B200)", ""]
    lc = os.path.join(OUT, f"launches_{tag}.csv")
    if os.path.exists(lc):
        lines = [l for l in open(lc) if not l.startswith("==")]
        rows = list(csv.reader(lines))
        hdr = rows[0]
        ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
        agg = collections.OrderedDict()
        for r in rows[1:]:
            if len(r) <= iv:
                continue
            try:
                v = float(r[iv].replace(",", ""))
            except ValueError:
                continue
            a = agg.setdefault(short(r[ik]), [0.0, 0])
            a[0] += v
            a[1] += 1
        unit = rows[1][hdr.index("Metric Unit")] if len(rows) > 1 else "ns"
        scale = 1e-3 if unit in ("ns", "nsecond") else 1.0
        tot = sum(a[0] for a in agg.values())
        md += ["## Launch list (cold-cache, serialised: shares, not absolutes)", "",
               f"{sum(a[1] for a in agg.values())} launches captured (about 1.6-2.5 steps incl. torch copy kernels).", "",
               "| kernel | launches | total us | share |", "|---|---|---|---|"]
        for k, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            md.append(f"| `{k}` | {n} | {v * scale:.1f} | {100 * v / tot:.1f} % |")
        md.append("")
        with open(os.path.join(PROF, f"{tag}_launches.csv"), "w") as f:
            f.writelines(lines)
    traffic = {}
    for kind in ("attn", "gemm", "vq"):
        rep = os.path.join(OUT, f"{kind}_{tag}.ncu-rep")
        if not os.path.exists(rep):
            continue
        hdr, units, data = raw_page(rep)
        ik = hdr.index("Kernel Name")
        md += [f"## `ncu --set full` capture: {kind} ({len(data)} launches)", ""]
        cols = [hdr.index(k) for k in KEYS if k in hdr]
        md.append("| kernel | " + " | ".join(f"{hdr[c]} [{units[c]}]" for c in cols) + " |")
        md.append("|---|" + "---|" * len(cols))
        for r in data:
            md.append(f"| `{short(r[ik])}` | " + " | ".join(r[c] for c in cols) + " |")
            if kind == "attn" and "attn" in r[ik]:
                def val(key):
                    c = hdr.index(key)
                    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[c]]
                    return float(r[c]) * mult
                traffic["ttk_attn_varlen_fwd"] = {"batch": int(batch), "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                                                  "source": f"profiles/{tag}_kernels.md (ncu --set full, one launch)"}
        md.append("")
    open(os.path.join(PROF, f"{tag}_kernels.md"), "w").write("\n".join(md) + "\n")
    if traffic:
        json.dump(traffic, open(os.path.join(PROF, f"{tag[:2]}_traffic.json"), "w"), indent=1)
    print("\n".join(md[:40]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else 64)
