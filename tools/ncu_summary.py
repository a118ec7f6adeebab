"""Summarise .ncu-rep captures + a launch list (gpu__time_duration) into small tracked files under profiles/."""
import collections, csv, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__shared_mem_per_block_dynamic"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return {k: (v, u) for k, u, v in zip(rows[0], rows[1], rows[2])}


def main(tag, launches_csv, *reps):
    summ = {}
    for rep in reps:
        d = raw(rep)
        name = d.get("Kernel Name", ("?", ""))[0].split("(")[0]
        summ[name] = {k: list(d[k]) for k in KEYS if k in d}
    json.dump(summ, open("profiles/%s_ncu_summary.json" % tag, "w"), indent=1)
    rows = list(csv.reader(open(launches_csv)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) > mv:
            try:
                v = float(r[mv].replace(",", ""))
            except ValueError:
                continue
            n = r[kn].split("(")[0].split("<")[0]
            agg[n][0] += 1
            agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    with open("profiles/%s_ncu_launch_shares.csv" % tag, "w") as fh:
        fh.write("kernel,launches,total_ns,share\n")
        for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            fh.write("%s,%d,%.0f,%.4f\n" % (n, c, t, t / tot))
    print(json.dumps(summ, indent=1))
    print(open("profiles/%s_ncu_launch_shares.csv" % tag).read()[:900])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], *sys.argv[3:])
