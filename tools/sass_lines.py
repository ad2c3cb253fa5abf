"""Static view of a kernel's SASS by source line (no GPU needed): python tools/sass_lines.py <mangled substring> [first_line last_line]
Without a line range: instruction count per source line, largest first.  With one: the SASS attributed to those lines of drt_render.cuh."""
import collections, os, re, subprocess, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
so = os.environ.get("DRT_CUDA_LIB") or os.path.join(HERE, "..", "daily-ray-trace_b200", "libdrt_cuda.so")
kern = sys.argv[1]
by_addr = len(sys.argv) > 3 and sys.argv[2].startswith("0x")      # an address range instead of a line range
lo, hi = (int(sys.argv[2], 0), int(sys.argv[3], 0)) if len(sys.argv) > 3 else (None, None)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
per = collections.Counter(); ops = collections.Counter(); total = 0
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"): continue
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if kern not in sass: continue
    inside, cur = False, ("?", 0)
    for ln in sass.splitlines():
        if ln.startswith(".text."):
            inside = kern in ln; continue
        if not inside: continue
        m = re.search(r'//## File "(.*?)", line (\d+)', ln)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if not m: continue
        total += 1; per[cur] += 1
        ins = m.group(2).strip()
        op = (ins.split()[1] if ins.startswith("@") else ins.split()[0])
        ops[re.sub(r"\..*", "", op)] += 1
        if lo is not None and ((by_addr and lo <= int(m.group(1), 16) <= hi) or (not by_addr and cur[0] == "drt_render.cuh" and lo <= cur[1] <= hi)):
            print(f"{cur[0][:12]:12s}{cur[1]:5d}  {m.group(1)}  {ins}")
print(f"{kern}: {total} SASS instructions")
if lo is None:
    for (f, l), c in per.most_common(40): print(f"{f[:16]:16s}{l:5d} {c:5d}")
    print(", ".join(f"{k}:{v}" for k, v in ops.most_common(30)))
