"""Attribute ncu stall samples of one kernel to CUDA source lines.

  python tools/ncu_lines.py <report.ncu-rep> <cubin.txt from `nvdisasm -g -c`> <mangled-kernel-substring> [min_samples] [demangled-name-substring]

The SASS page of the report and the nvdisasm listing are walked in lock-step (same instruction order);
the `//## File ..., line N` markers of the listing give every instruction its source line.
"""
import collections, csv, re, subprocess, sys

rep, listing, kern = sys.argv[1:4]
min_s = int(sys.argv[4]) if len(sys.argv) > 4 else 200
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# a report with several launches: the first launch of the kernel whose demangled name contains argv[5] (default: the first)
want = sys.argv[5] if len(sys.argv) > 5 else ""
heads = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
first = next(i for i in heads if want in rows[i][1])
end = next((i for i in heads if i > first), len(rows))
h = rows[first + 1]; ix = {k: i for i, k in enumerate(h)}
body = rows[first + 2:end]
sass = [(r[ix["Source"]].strip(), int(r[ix["# Samples"]] or 0)) for r in body if len(r) > ix["# Samples"] and r[ix["# Samples"]].isdigit()]
# listing: instructions of the kernel in order with current line
lines = open(listing).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
cur = None; instr = []
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith("//--------------------- .text") and instr: 
        if instr: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m: instr.append((m.group(1).strip(), cur))
print("sass rows", len(sass), "listing instr", len(instr), file=sys.stderr)
by = collections.Counter(); tot = 0
for (s, n), (t, ln) in zip(sass, instr):
    by[ln] += n; tot += n
src = {}
for ln, n in by.most_common():
    if n < min_s: break
    f, k = ln if ln else ("?", 0)
    if f not in src:
        try: src[f] = open("phylogenetics_b200/csrc/" + f).read().splitlines()
        except OSError: src[f] = []
    text = src[f][k - 1].strip() if 0 < k <= len(src[f]) else ""
    print("%6d %5.1f%%  %s:%d  %s" % (n, 100 * n / tot, f, k, text[:100]))
