"""Summarise the source page of an ncu report per CUDA source line (needs -lineinfo and
ncu --import-source on):

    ncu -i REPORT --page source --csv --print-source sass,cuda > src.csv
    python tools/ncu_source_summary.py src.csv [kernel-substring] [top]

For each kernel (first captured launch): the source lines that executed the most warp
instructions, with the average number of active lanes on them."""
import collections
import csv
import sys

path = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
kern = collections.OrderedDict()      # kernel name -> {(file, line, text): [warp inst, thread inst]}
cur_file = cur_fn = hdr = None
done_sections = set()
skip = False
with open(path, newline="", errors="replace") as f:
    for r in csv.reader(f):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            hdr = None
            continue
        if r[0] == "Function Name":
            cur_fn = r[1]
            key = (cur_fn, cur_file)
            skip = key in done_sections          # later launches of the same kernel repeat the sections
            done_sections.add(key)
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or skip or len(r) != len(hdr) or pat not in (cur_fn or "") or not r[0].strip():
            continue      # (rows without a line number are the SASS instructions under the line above)
        ie, te = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        try:
            n, t = int(r[ie] or 0), int(r[te] or 0)
        except ValueError:
            continue
        d = kern.setdefault(cur_fn, collections.OrderedDict())
        k = (cur_file, r[0], r[1].strip())
        a = d.setdefault(k, [0, 0, 0])
        a[0] += n
        a[1] += t
        try:
            a[2] += int(r[hdr.index("# Samples")] or 0)
        except ValueError:
            pass
for fn, d in kern.items():
    tot = sum(a[0] for a in d.values())
    tth = sum(a[1] for a in d.values())
    tsm = max(sum(a[2] for a in d.values()), 1)
    print("== %s\n   warp instructions %d, avg active lanes %.1f; columns: share of instructions, share of stall samples" % (fn[:110], tot, tth / max(tot, 1)))
    for (fl, ln, text), a in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
        if a[0] == 0:
            break
        print("   %5.1f%% %5.1f%%  lanes %4.1f  %s:%s  %s" % (100.0 * a[0] / tot, 100.0 * a[2] / tsm, a[1] / a[0], fl, ln, text[:100]))
