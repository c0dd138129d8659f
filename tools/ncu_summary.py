"""Summarise an ncu report (or its `--page raw --csv` export) per kernel instantiation -> markdown for profiles/.

    ncu -i gpurun_out/all.ncu-rep --page raw --csv > /tmp/all.csv     (done here when a .ncu-rep is given)
    python tools/ncu_summary.py gpurun_out/all.ncu-rep [--title "..."] > profiles/r02_ncu_per_kernel.md

Per kernel (demangled name incl. template arguments): launches captured, mean duration, DRAM bytes read / written per
launch, achieved DRAM GB/s, tensor-pipe activity, SM busy, registers, dynamic shared memory.
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

WANT = {
    "gpu__time_duration.sum": "dur",
    "dram__bytes_read.sum": "rd",
    "dram__bytes_write.sum": "wr",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "launch__registers_per_thread": "regs",
    "launch__shared_mem_per_block_dynamic": "smem",
    "launch__grid_size": "grid",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
}
UNIT_SCALE = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6, "second": 1e6,
              "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "KB": 1e3, "MB": 1e6, "GB": 1e9, "B": 1.0}


def load(path):
    if path.endswith(".ncu-rep"):
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv", "--print-kernel-base", "demangled"], stdout=subprocess.PIPE, text=True).stdout
    else:
        txt = open(path).read()
    lines = txt.splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))
    rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
    return rows[0], rows[1], rows[2:]


def main():
    path = sys.argv[1]
    title = sys.argv[sys.argv.index("--title") + 1] if "--title" in sys.argv else path
    head, units, rows = load(path)
    col = {h: i for i, h in enumerate(head)}
    for m in WANT:      # some ncu versions prefix section metrics ("SM_A.TriageCompute.<metric>"): match by suffix, plain name wins
        if m not in col:
            hit = [i for i, h in enumerate(head) if h.endswith("." + m)]
            if hit:
                col[m] = hit[0]
    name_i = col["Kernel Name"]
    agg = OrderedDict()
    for r in rows:
        if len(r) <= name_i:
            continue
        k = r[name_i]
        a = agg.setdefault(k, {"n": 0})
        a["n"] += 1
        for m, key in WANT.items():
            if m in col and r[col[m]] not in ("", "n/a"):
                try:
                    v = float(r[col[m]].replace(",", ""))
                except ValueError:
                    continue
                v *= UNIT_SCALE.get(units[col[m]], 1.0)
                a[key] = a.get(key, 0.0) + v
    print("# %s\n" % title)
    print("| kernel | launches | mean us | DRAM read MB | DRAM write MB | DRAM GB/s | DRAM % | tensor pipe % | SM % | regs | dyn smem KB |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1].get("dur", 0.0)):
        n = a["n"]
        g = lambda key: a.get(key, 0.0) / n
        dur = g("dur")
        gbs = (g("rd") + g("wr")) / dur / 1e3 if dur else 0.0
        print("| `%s` | %d | %.1f | %.1f | %.1f | %.0f | %.1f | %.1f | %.1f | %d | %.0f |" % (
            k.replace("se::", "").replace("|", "\\|"), n, dur, g("rd") / 1e6, g("wr") / 1e6, gbs, g("dram_pct"), g("tensor_pct"), g("sm_pct"), g("regs"), g("smem") / 1024))


if __name__ == "__main__":
    main()
