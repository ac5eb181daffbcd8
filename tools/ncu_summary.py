"""Turn the two ncu captures of the profiling recipe into the tables kept under profiles/.

  python tools/ncu_summary.py <launches.csv> <raw.csv> <crops per launch> > profiles/<round>_summary.md

launches.csv : ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv  (launch list of bench.py)
raw.csv      : ncu -i <rep> --page raw --csv of an `ncu --set full --clock-control none --import-source on -k regex:tc_` capture
"""
import csv
import sys
from collections import OrderedDict


def launch_table(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].strip()
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[-1]) / 1e3
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | avg us | share |", "|---|---|---|---|"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.1f | %.3f |" % (k, n, us / n, us / tot))
    return "\n".join(out)


def full_table(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = rows[0]
    body = rows[2:]

    def col(name):
        i = hdr.index(name)
        return [r[i] for r in body]

    names = [n.split("(")[0].replace("void hp::", "") for n in col("Kernel Name")]
    f = lambda xs: [float(x.replace(",", "")) for x in xs]
    t = f(col("gpu__time_duration.sum"))
    tens = f(col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")) if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in hdr else [float("nan")] * len(t)
    rd, wr = f(col("dram__bytes_read.sum")), f(col("dram__bytes_write.sum"))
    issue = f(col("smsp__issue_active.avg.pct_of_peak_sustained_active")) if "smsp__issue_active.avg.pct_of_peak_sustained_active" in hdr else [float("nan")] * len(t)
    cyc = f(col("sm__cycles_elapsed.max"))
    regs = col("launch__registers_per_thread")
    shm = f(col("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")) if "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum" in hdr else [float("nan")] * len(t)
    out = ["| kernel | time us | tensor pipe active % | DRAM read MB | DRAM write MB | issue active % | SM cycles elapsed | regs/thread | LSU shared wavefronts |",
           "|---|---|---|---|---|---|---|---|---|"]
    traffic = {}
    for i, n in enumerate(names):
        out.append("| %s | %.1f | %.1f | %.1f | %.1f | %.1f | %d | %s | %.3g |" % (n, t[i], tens[i], rd[i], wr[i], issue[i], cyc[i], regs[i], shm[i]))
        traffic[n] = (rd[i] + wr[i]) * 1e6
    return "\n".join(out), traffic


if __name__ == "__main__":
    print(launch_table(sys.argv[1]))
    print()
    tab, traffic = full_table(sys.argv[2])
    print(tab)
    print()
    print("traffic (dram read + write, bytes per launch):", traffic)
