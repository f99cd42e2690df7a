"""Condense an `ncu -i X.ncu-rep --page raw --csv` dump into the few columns DESIGN.md / bench.py quote.
Usage: python tools/ncu_summary.py raw.csv [hbm_peak_GBs] > profiles/rNN_..._summary.txt"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("lts__t_sector_hit_rate.pct", "L2hit%"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "mem%"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("sm__cycles_elapsed.avg.per_second", "sm_clk"),
]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * m.get(unit, 1)


def to_sec(v, unit):
    m = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "second": 1, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}
    return float(v) * m[unit]


def main():
    path = sys.argv[1]
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
    if peak is None:
        p = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak = json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def find(name):
        if name in ix:
            return ix[name]
        for h, i in ix.items():
            if h.endswith(name):
                return i
        return None
    kn = find("Kernel Name")
    print("# %s  (HBM peak used for the fraction: %.0f GB/s)" % (os.path.basename(path), peak))
    print("%-46s %9s %9s %9s %8s %6s %7s %6s %6s %5s %6s %6s" % ("kernel", "time_us", "rd_MB", "wr_MB", "GB/s", "ofHBM", "L2hit%", "tens%", "sm%", "regs", "grid", "GHz"))
    for d in data:
        def val(name):
            i = find(name)
            return (d[i], units[i]) if i is not None and d[i] != "" else (None, None)
        t, tu = val("gpu__time_duration.sum")
        rd, ru = val("dram__bytes_read.sum")
        wr, wu = val("dram__bytes_write.sum")
        ts = to_sec(t, tu)
        rb, wb = to_bytes(rd, ru), to_bytes(wr, wu)
        gbs = (rb + wb) / ts / 1e9
        name = d[kn].replace("<unnamed>::", "").replace("void ", "")
        name = name.split("(")[0][:46]
        g = lambda n: (val(n)[0] or "-")
        print("%-46s %9.2f %9.3f %9.3f %8.1f %6.3f %7.5s %6.5s %6.5s %5s %6s %6.5s" % (
            name, ts * 1e6, rb / 1e6, wb / 1e6, gbs, gbs / peak, g("lts__t_sector_hit_rate.pct"),
            g("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
            g("sm__throughput.avg.pct_of_peak_sustained_elapsed"), g("launch__registers_per_thread"), g("launch__grid_size"),
            g("sm__cycles_elapsed.avg.per_second")))


if __name__ == "__main__":
    main()
