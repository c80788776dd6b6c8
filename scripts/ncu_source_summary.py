"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` output: per CUDA source line the stall
samples and executed instructions, top N lines, for the k-th profiled kernel in the file.
usage: ncu_source_summary.py source.csv[.gz] [kernel_index=0] [top=40]"""
import csv
import gzip
import sys


def main():
    fn = sys.argv[1]
    kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    f = gzip.open(fn, "rt") if fn.endswith(".gz") else open(fn)
    rows = csv.reader(f)
    kernel, seen_files, cur_file, hdr = -1, set(), None, None
    lines = {}          # (file, line) -> [samples, inst, text]
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1]
            if cur_file in seen_files or kernel < 0:
                if cur_file in seen_files:
                    seen_files = set()
                kernel += 1 if (kernel < 0 or True) and cur_file not in seen_files and (kernel < 0 or len(seen_files) == 0) else 0
            seen_files.add(cur_file)
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            i_samp = hdr.index("# Samples")
            i_inst = hdr.index("Instructions Executed")
            continue
        if kernel != kidx or hdr is None:
            continue
        if r[0] not in ("", "-") and r[0].isdigit():          # a CUDA source line with aggregated metrics
            try:
                s, n = int(r[i_samp]), int(r[i_inst])
            except ValueError:
                continue
            if s or n:
                key = (cur_file.split("/")[-1], int(r[0]))
                e = lines.setdefault(key, [0, 0, r[1].strip()])
                e[0] += s
                e[1] += n
    tot_s = sum(v[0] for v in lines.values()) or 1
    tot_n = sum(v[1] for v in lines.values()) or 1
    print(f"kernel #{kidx}: {tot_s} samples, {tot_n} warp instructions")
    for (fl, ln), (s, n, txt) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * s / tot_s:5.1f}% samp {100 * n / tot_n:5.1f}% inst  {fl}:{ln}  {txt[:110]}")


main()
