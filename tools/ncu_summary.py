"""Summarise .ncu-rep captures (+ optionally a gpu__time_duration launch list) into small tracked files under profiles/.

    python tools/ncu_summary.py <tag> <launches.csv|-> <rep> [<rep> ...]

profiles/<tag>_ncu_summary.json: one entry per captured launch ("<kernel>#<n>"): the metrics of KEYS, each [value, unit].
profiles/<tag>_ncu_launch_shares.csv: per-kernel launch count, total device time and share of the launch list."""
import collections, csv, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "launch__shared_mem_per_block_dynamic", "sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum",
        "gcc__cache_requests_type_instruction.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        # the L1TEX data pipe (one 128-B wavefront per cycle and SM): LSU side (global/local loads, TMA fills = "lgds",
        # LDS/STS = "shared") and tensor-core operand reads from shared memory
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed"]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        yield {k: (v, u) for k, u, v in zip(hdr, units, r)}


def main(tag, launches_csv, *reps):
    summ, seen = {}, collections.Counter()
    for rep in reps:
        for d in rows_of(rep):
            name = d.get("Kernel Name", ("?", ""))[0].split("(")[0]
            seen[name] += 1
            e = {k: list(d[k]) for k in KEYS if k in d}
            try:  # achieved DRAM GB/s of this launch (HBM-bound kernels: compare with MEASURED_PEAKS.json hbm_gbs)
                sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                tu = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
                b = sum(float(d[m][0].replace(",", "")) * sc[d[m][1]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                t = float(d["gpu__time_duration.sum"][0].replace(",", "")) * tu[d["gpu__time_duration.sum"][1]]
                e["dram_bytes_total"] = b
                e["dram_GBps"] = b / t / 1e9
            except Exception:
                pass
            summ["%s#%d" % (name, seen[name])] = e
    json.dump(summ, open("profiles/%s_ncu_summary.json" % tag, "w"), indent=1)
    print(json.dumps(summ, indent=1)[:6000])
    if launches_csv == "-":
        return
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
    print(open("profiles/%s_ncu_launch_shares.csv" % tag).read()[:1500])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], *sys.argv[3:])
