"""profiles/<tag>_summary.md from a bench JSON line and an ncu launch list (CSV of `--metrics gpu__time_duration.sum,
dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed`), plus
profiles/gemm_dram_traffic.json (DRAM bytes per tcgen05 launch).  One forward = the launches between two set_io_kernel launches.

usage: make_profile_summary.py <tag> <bench.json> <launches.csv> [notes.md]"""
import collections, csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TC = ("gemm_bf16_tcgen05_kernel", "mlp_cluster_tcgen05_kernel", "mlp_fused_tcgen05_kernel", "convffn_tcgen05_kernel", "attention_umma_kernel",
      "repmixer_umma_kernel")


def short(name):
    n = name.replace("fvhd::", "")
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"<.*", "", n) if n.startswith("convffn") else n
    n = re.sub(r"\(.*", "", n)
    return n[:40]


def main():
    tag, bench_path, csv_path = sys.argv[1:4]
    notes = open(sys.argv[4]).read() if len(sys.argv) > 4 else ""
    bench = json.loads([l for l in open(bench_path) if l.startswith("{")][-1])
    rows = [r for r in csv.reader(l for l in open(csv_path) if l.startswith('"')) if len(r) > 14 and r[0] != "ID"]
    launches = collections.OrderedDict()
    for r in rows:
        d = launches.setdefault(int(r[0]), {"name": short(r[4]), "grid": r[8]})
        d[r[12]] = float(r[14].replace(",", ""))
    seq = list(launches.values())
    marks = [i for i, d in enumerate(seq) if d["name"].startswith("set_io_kernel")]
    if len(marks) < 2:
        raise SystemExit("need two set_io_kernel launches in the capture window")
    fwd = seq[marks[0]:marks[1]]
    agg = collections.OrderedDict()
    for d in fwd:
        a = agg.setdefault(d["name"], {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0, "tw": 0.0})
        a["n"] += 1
        a["ns"] += d.get("gpu__time_duration.sum", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        a["tw"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * d.get("gpu__time_duration.sum", 0.0)
    tot = sum(a["ns"] for a in agg.values())
    out = [f"# {tag}: bench line + ncu launch list", ""]
    out.append(f"* value **{bench['value']:.1f} {bench['unit']}** ({bench['ms_per_step']:.3f} ms/step), e2e {bench['e2e']['value']:.1f} "
               f"({bench['e2e']['ms_per_step']:.3f} ms incl. {bench['e2e']['h2d_bytes_per_step'] / 1e6:.1f} MB H2D + {bench['e2e']['d2h_bytes_per_step'] / 1e6:.2f} MB D2H), "
               f"clocks {bench.get('clocks')}")
    rf = bench["roofline"]
    out.append(f"* roofline ({rf['kernel']}; {rf['launches_per_step']} launches, live CUDA events): {rf['achieved']} TFLOP/s = "
               f"{100 * rf['frac']:.1f} % of {rf['peak']} TFLOP/s ({rf['peak_source']})")
    if bench.get("cpu_baseline"):
        cb = bench["cpu_baseline"]
        out.append(f"* cpu_baseline ({cb['kind']}, {cb['cores']} threads): {cb['value']:.2f} {cb['unit']}")
    tcl = [d for d in fwd if d["name"] in TC]
    if tcl:
        per = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in tcl) / len(tcl)
        alg = None
        ks = {k["kernel"]: k for k in bench.get("kernels", [])}
        if all(n in ks for n in set(d["name"] for d in tcl)):
            alg = sum(ks[n]["gbs"] * ks[n]["ms"] * 1e6 for n in set(d["name"] for d in tcl)) / len(tcl)
        json.dump({"bytes_per_launch": per, "algorithmic_bytes_per_launch": alg, "launches": len(tcl),
                   "source": f"profiles/{os.path.basename(csv_path)}: mean dram__bytes_read.sum + dram__bytes_write.sum over the tcgen05 launches "
                             "(GEMM + fused ConvFFN kernels) of one forward (B=1, 1024 px; ncu flushes caches between kernels, so operands come from "
                             "HBM once; outputs stay in L2)"}, open(os.path.join(ROOT, "profiles", "r02_tc_dram_traffic.json"), "w"), indent=1)
        out.append(f"* DRAM traffic of the tcgen05 launches (ncu, cold): {per / 1e6:.2f} MB per launch" + (f" vs algorithmic {alg / 1e6:.2f} MB" if alg else ""))
    out += ["", f"## ncu launch list, one forward ({len(fwd)} kernels between two set_io_kernel launches; cold-cache, serialised -> compare shares)", "",
            "| kernel | launches | total us | share | avg us | DRAM rd MB | DRAM wr MB | tensor-pipe active % (time-weighted) |", "|---|---|---|---|---|---|---|---|"]
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        out.append(f"| {n} | {a['n']} | {a['ns'] / 1e3:.1f} | {100 * a['ns'] / tot:.1f} % | {a['ns'] / a['n'] / 1e3:.1f} | {a['rd'] / 1e6:.1f} | {a['wr'] / 1e6:.1f} | {a['tw'] / max(a['ns'], 1):.1f} |")
    out += ["", "## live CUDA-event kernel table from bench.py (one step, every launch bracketed by events: includes ~4 us of event overhead per launch)", "",
            "| kernel | launches | ms | share | TFLOP/s | GB/s (algorithmic) |", "|---|---|---|---|---|---|"]
    for k in bench.get("kernels", []):
        out.append(f"| {k['kernel']} | {k['launches']} | {k['ms']} | {100 * k['share']:.1f} % | {k['tflops']} | {k['gbs']} |")
    if notes:
        out += ["", notes]
    open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:12]))


if __name__ == "__main__":
    main()
