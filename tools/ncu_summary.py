"""Summarise gpurun_out ncu artefacts into profiles/ (launch shares + key metrics of a full capture)."""
import collections, csv, io, subprocess, sys

def launch_shares(src, dst, header, only=None):
    """only: substring a kernel name must contain (bench.py also runs torch.matmul peak measurements: not ours)."""
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0]
        if only and only not in row["Kernel Name"]:
            continue
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(row["Metric Value"].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    out = header + ["kernel,launches,total_us,share_pct"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"\"{k}\",{a[0]},{a[1] / 1e3:.1f},{a[1] / tot * 100:.1f}")
    open(dst, "w").write("\n".join(out) + "\n")
    return out

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]

def full_summary(rep, dst, header):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    with open(dst, "w") as f:
        for h in header:
            f.write(h + "\n")
        w = csv.writer(f)
        for r in rows:
            w.writerow([r[i] for i in idx])
    return [[r[i] for i in idx] for r in rows]

def round_summary(tag, cmd):
    """profiles/<tag>_launch_shares.csv, profiles/<tag>_<kernel>_ncu_full_summary.csv for every gpurun_out/<tag>_prof_*.ncu-rep
    and profiles/<tag>_ncu_traffic.json (dram bytes per launch per kernel, read by bench.py's roofline.traffic)."""
    import glob, json, os

    launch_shares(f"gpurun_out/{tag}_launches.csv", f"profiles/{tag}_launch_shares.csv",
                  [f"# {tag} ncu launch list of: {cmd}",
                   "# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)"])
    traffic = {}
    for rep in sorted(glob.glob(f"gpurun_out/{tag}_prof_*.ncu-rep")):
        name = os.path.basename(rep)[len(tag) + 6:-8]
        rows = full_summary(rep, f"profiles/{tag}_{name}_ncu_full_summary.csv",
                            [f"# {tag} ncu --set full --clock-control none --import-source on ({name}) of: {cmd}"])
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))

            def to_bytes(k):
                v = float(d[k].replace(",", ""))
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[k]]

            traffic.setdefault(name + ":" + d["Kernel Name"].split("(")[0], []).append(
                to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))
            print(name, d["Kernel Name"][:60], d["gpu__time_duration.sum"], u["gpu__time_duration.sum"],
                  "tensor %", d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                  "dram %", d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"))
    json.dump({"how": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full; key = capture:kernel; layer "
                      "launches cover one wave of 37 888 points", "points_per_layer_launch": 37888,
               "bytes_per_launch": {k: sum(v) / len(v) for k, v in traffic.items()}},
              open(f"profiles/{tag}_ncu_traffic.json", "w"), indent=1)


if __name__ == "__main__":
    if len(sys.argv) == 3:
        round_summary(sys.argv[1], sys.argv[2])
        sys.exit(0)
    tag = sys.argv[1]
    cmd = sys.argv[2]
    for l in launch_shares(f"gpurun_out/launches_{tag}.csv", f"profiles/{tag}_launch_shares.csv",
                           [f"# {tag} ncu launch list of: {cmd}",
                            "# ncu --metrics gpu__time_duration.sum --clock-control none (one precompute+solve step; cold-cache, serialised: compare shares)"])[:14]:
        print(l)
    rep = sys.argv[3]
    for r in full_summary(rep, f"profiles/{tag}_layer_tc_ncu_full_summary.csv",
                          [f"# {tag} ncu --set full --clock-control none --import-source on -k regex:layer_tc_pair_kernel of: {cmd}"]):
        print(r)
