"""Per-source-line shares of a kernel from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` (needs
--import-source on at capture time and -lineinfo at compile time).
usage: ncu_source_lines.py report.ncu-rep [top N]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
per = collections.Counter()
samp = collections.Counter()
thr = collections.Counter()
fname = ""
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie, st, te = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or not r[0].isdigit():
        continue
    # the source text may hold quotes / commas that ncu does not escape: locate the columns from the "-","-" pair
    try:
        k0 = next(i for i in range(2, len(r) - 1) if r[i] == "-" and r[i + 1] == "-")
    except StopIteration:
        continue
    off = k0 - 2
    key = (fname, int(r[0]), ",".join(r[1:k0]))
    try:
        per[key] += float(r[ie + off] or 0)
        samp[key] += float(r[st + off] or 0)
        thr[key] += float(r[te + off] or 0)
    except (ValueError, IndexError):
        continue
tot, ts = sum(per.values()), sum(samp.values())
print("total warp-instructions %.1f M, samples %d, avg active threads %.1f" % (tot / 1e6, ts, sum(thr.values()) / max(tot, 1)))
print("| inst % | stall % | thr/inst | file:line | source |")
print("|---|---|---|---|---|")
for k, v in per.most_common(top):
    print("| %.1f | %.1f | %.0f | %s:%d | `%s` |" % (100 * v / tot, 100 * samp[k] / max(ts, 1), thr[k] / max(v, 1), k[0], k[1], k[2][:100]))
