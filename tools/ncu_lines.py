"""Per-source-line instruction counts / stall samples of one kernel from an .ncu-rep (needs -lineinfo):
    python tools/ncu_lines.py gpurun_out/x.ncu-rep adjust_rank ftk_adjust adjust_rank_kernelIiLb0ELi21 [top]
joins ncu's SASS source page (per-instruction counters) with nvdisasm -g line annotations by order."""
import csv, io, os, re, subprocess, sys, tempfile

rep, kregex, cub, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(repo, "finaletoolkit_b200", "libftk_b200.so")], cwd=tmp,
               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.startswith(cub) and f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# walk the function's section
lines, cur, infn = [], None, False
for ln in dis.splitlines():
    if ln.startswith(".text.") or re.match(r"^\s*\.section\s+\.text\.", ln):
        infn = mangled in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kregex],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[h]
ie, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
body = [r for r in rows[h + 1:] if len(r) > ie and r[ie].isdigit()]
if len(body) != len(lines):
    print(f"warning: {len(body)} SASS rows in the report vs {len(lines)} in nvdisasm", file=sys.stderr)
agg = {}
tot_i = tot_s = 0
for r, key in zip(body, lines):
    n, s = int(r[ie]), int(r[sm] or 0)
    a = agg.setdefault(key, [0, 0]); a[0] += n; a[1] += s
    tot_i += n; tot_s += s
src = {}
print(f"total warp-instructions {tot_i}, samples {tot_s}")
for key, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if key:
        path = os.path.join(repo, "finaletoolkit_b200", "csrc", key[0])
        if os.path.exists(path):
            src.setdefault(path, open(path).read().splitlines())
            text = src[path][key[1] - 1].strip()[:100] if key[1] - 1 < len(src[path]) else ""
    print(f"{100 * n / tot_i:5.1f}% inst {100 * s / max(tot_s, 1):5.1f}% smp  {key}  {text}")
