"""Summarise the ncu exports of scripts/profile_round.sh (gpurun_out/<tag>_*) into small tracked files under profiles/."""
import csv, collections, gzip, json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1b"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def raw_summary(name):
    f = os.path.join(G, f"{tag}_{name}_raw.csv")
    if not os.path.isfile(f):
        return None
    rows = list(csv.reader(open(f)))
    hdr, units, val = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, val)}
    out = [("Kernel Name", d.get("Kernel Name", ("?", ""))[0], "")]
    for k in KEEP:
        if k in d:
            out.append((k, d[k][0], d[k][1]))
    stalls = []
    for h, (v, u) in d.items():
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                stalls.append((int(v.replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(s for s, _ in stalls) or 1
    for s, h in sorted(stalls, reverse=True)[:8]:
        out.append((f"stall_samples.{h}", f"{100.0 * s / tot:.1f}", "% of pc samples"))
    with open(os.path.join(P, f"{tag}_{name}_ncu_summary.csv"), "w") as o:
        w = csv.writer(o)
        w.writerow(["metric", "value", "unit"])
        w.writerows(out)
    return d


def opcode_mix(name, per=None):
    f = os.path.join(G, f"{tag}_{name}_src.csv.gz")
    if not os.path.isfile(f):
        return
    rows = list(csv.reader(gzip.open(f, "rt")))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr, data = rows[hi], rows[hi + 1:]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    by = collections.defaultdict(lambda: [0, 0])
    tot_ex = tot_s = 0
    for r in data:
        try:
            ex, s = int(r[iex]), int(r[isamp])
        except (ValueError, IndexError):
            continue
        toks = [t for t in r[isrc].split() if not t.startswith("@")]
        op = toks[0].split(".")[0] if toks else "?"
        by[op][0] += ex; by[op][1] += s
        tot_ex += ex; tot_s += s
    with open(os.path.join(P, f"{tag}_{name}_opcode_mix.csv"), "w") as o:
        w = csv.writer(o)
        w.writerow(["opcode", "warp_instructions", "share_of_instructions_pct", "share_of_stall_samples_pct"] + (["per_unit"] if per else []))
        for op, (ex, s) in sorted(by.items(), key=lambda x: -x[1][0])[:25]:
            w.writerow([op, ex, f"{100.0 * ex / max(tot_ex, 1):.2f}", f"{100.0 * s / max(tot_s, 1):.2f}"] + ([f"{ex / per:.1f}"] if per else []))
        w.writerow(["TOTAL", tot_ex, "100", "100"] + ([f"{tot_ex / per:.1f}"] if per else []))


def launches(fn, out):
    f = os.path.join(G, fn)
    if not os.path.isfile(f):
        return None
    lines = [l for l in open(f) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        k = re.sub(r"\(.*", "", row["Kernel Name"])[:90]
        a = agg.setdefault(k, [0, 0.0, row["Grid Size"], row["Block Size"]])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1.0
    with open(os.path.join(P, out), "w") as o:
        w = csv.writer(o)
        w.writerow(["kernel", "launches", "total_us", "avg_us", "share_pct", "grid(last)", "block"])
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            w.writerow([k, a[0], f"{a[1]:.1f}", f"{a[1] / a[0]:.1f}", f"{100 * a[1] / tot:.2f}", a[2], a[3]])
    return agg


if __name__ == "__main__":
    os.makedirs(P, exist_ok=True)
    d = raw_summary("nuts")
    evals = None
    log = os.path.join(G, f"{tag}_nuts.log")
    if os.path.isfile(log):
        m = re.search(r"grad_evals_incl_start (\d+)", open(log).read())
        evals = int(m.group(1)) if m else None
    opcode_mix("nuts", per=evals)
    if d and evals:
        rd = float(d["dram__bytes_read.sum"][0].replace(",", "")) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}[d["dram__bytes_read.sum"][1]]
        wr = float(d["dram__bytes_write.sum"][0].replace(",", "")) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}[d["dram__bytes_write.sum"][1]]
        json.dump({"kernel": "nuts_warp_kernel<RadonModel,6,1>", "launch": "2048 chains x (150 tune + 50 draws), scripts/ncu_target_radon.py",
                   "grad_evals_incl_start_state": evals, "dram_bytes_read": rd, "dram_bytes_write": wr,
                   "dram_bytes_per_grad_eval": (rd + wr) / evals, "algorithmic_bytes_per_grad_eval": 28180},
                  open(os.path.join(P, f"{tag}_traffic.json"), "w"), indent=1)
    raw_summary("stochvol"); opcode_mix("stochvol")
    raw_summary("logistic"); opcode_mix("logistic")
    raw_summary("gemm"); opcode_mix("gemm")
    raw_summary("logistic_tc"); opcode_mix("logistic_tc")
    raw_summary("gemm_tc"); opcode_mix("gemm_tc")
    launches(f"{tag}_launches_logistic_tc.csv", f"{tag}_launches_logistic_tc_summary.csv")
    launches(f"{tag}_launches_mvgauss_tc.csv", f"{tag}_launches_mvgauss_tc_summary.csv")
    launches(f"{tag}_launches.csv", f"{tag}_launches_bench_summary.csv")
    launches(f"{tag}_launches_logistic.csv", f"{tag}_launches_logistic_summary.csv")
    launches(f"{tag}_launches_mvgauss.csv", f"{tag}_launches_mvgauss_summary.csv")
    print("summaries written to", P)
