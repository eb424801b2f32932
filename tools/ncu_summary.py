"""Key metrics of every kernel launch in an ncu report (--set full): duration, DRAM bytes and throughput, FP64 / DMMA pipe
utilisation, occupancy, registers.  usage: ncu_summary.py report.ncu-rep [...]"""
import csv, subprocess, sys
WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "DMMA pipe %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 inst %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU wavefronts %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("launch__registers_per_thread", "registers"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block")]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    name_i = h.index("Kernel Name")
    for r in rows[2:]:
        print("== %s :: %s" % (rep.split("/")[-1], r[name_i][:90]))
        vals = {}
        for key, label in WANT:
            if key in h:
                i = h.index(key)
                vals[key] = (r[i], units[i])
                print("   %-24s %s %s" % (label, r[i], units[i]))
        try:
            t = float(vals["gpu__time_duration.sum"][0].replace(",", ""))
            tu = vals["gpu__time_duration.sum"][1]
            t_s = t * {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "msecond": 1e-3, "ms": 1e-3, "second": 1.0, "nsecond": 1e-9}.get(tu, 1e-9)
            def b(k):
                v, u = vals[k]
                return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            tot = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
            print("   %-24s %.1f GB/s  (%.3f GB in %.3f ms)" % ("DRAM read+write", tot / t_s * 1e-9, tot * 1e-9, t_s * 1e3))
        except Exception as e:
            print("   (no bandwidth figure: %s)" % e)
