"""Attribute the warp-stall samples of an `ncu --import-source on` capture to CUDA source lines.
   ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > sass.csv      (per-SASS-instruction samples)
   nvdisasm -g -c the.cubin > dis.txt                                          (SASS offset -> file:line)
   python tools/ncu_lines.py sass.csv dis.txt <mangled-substring> [top]
The ncu addresses are absolute; the offset inside the function is address - (address of the first row)."""
import csv
import re
import sys
from collections import defaultdict


def main():
    sass_csv, dis, func = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    # offset -> (file, line, inline chain)
    line_of = {}
    cur = None
    on = False
    for ln in open(dis):
        if ln.startswith(".text.") and ln.rstrip().endswith(":"):
            on = func in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m and cur:
            line_of[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(sass_csv)))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    base = int(data[0][ix["Address"]], 16)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    per = defaultdict(lambda: defaultdict(int))
    tot = 0
    for r in data:
        if len(r) <= ix["# Samples"] or not r[ix["# Samples"]]:
            continue
        off = int(r[ix["Address"]], 16) - base
        key = line_of.get(off, ("?", 0))
        n = int(r[ix["# Samples"]])
        per[key]["n"] += n
        tot += n
        for h in stalls:
            per[key][h] += int(r[ix[h]] or 0)
    print("total samples %d over %d source lines" % (tot, len(per)))
    src = {}
    for key, d in sorted(per.items(), key=lambda kv: -kv[1]["n"])[:top]:
        f, l = key
        if f not in src:
            try:
                src[f] = open("monoport_b200/csrc/" + f).read().split("\n")
            except OSError:
                src[f] = []
        text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
        st = sorted(((d[h], h[6:]) for h in stalls), reverse=True)[:2]
        print("%6.2f%%  %-16s %-22s %s" % (100.0 * d["n"] / tot, "%s:%d" % (f, l), " ".join("%s=%.0f%%" % (h, 100.0 * v / d["n"]) for v, h in st if v), text))


if __name__ == "__main__":
    main()
