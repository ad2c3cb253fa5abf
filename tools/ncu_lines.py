"""Attribute an .ncu-rep of render_kernel to CUDA source lines.

ncu's CSV source page is SASS-only; this joins it, by instruction offset, with `nvdisasm -g` of the same cubin (line info
from -lineinfo) and prints the lines that execute the most warp-instructions.  The cubin must be the build the report
was captured from.
usage: python tools/ncu_lines.py report.ncu-rep paths_in_launch [kernel_mangled_substring] [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, paths = sys.argv[1], float(sys.argv[2])
kern = sys.argv[3] if len(sys.argv) > 3 else "render_kernelIfLi5"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
HERE = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(HERE, "..", "daily-ray-trace_b200", "libdrt_cuda.so")

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
line_of = {}   # instruction offset -> (line, function-level line for inlined code is not kept: innermost line only)
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if kern not in sass: continue
    inside, cur = False, ('?', 0)
    for ln in sass.splitlines():
        if ln.startswith(".text."):
            inside = kern in ln
            continue
        if not inside: continue
        m = re.search(r'//## File "(.*?)", line (\d+)', ln)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m: line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
ia, iaddr, ismp, it = (hdr.index(k) for k in ("Instructions Executed", "Address", "# Samples", "Avg. Threads Executed"))
base = int(data[0][iaddr], 16)
per_line = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
tot = totS = 0.0
for r in data:
    off = int(r[iaddr], 16) - base
    ln = line_of.get(off, (("?", 0), "?"))[0]
    ex, sm = float(r[ia]), float(r[ismp])
    p = per_line[ln]; p[0] += ex; p[1] += sm; p[2] += ex * float(r[it]); p[3] += 1
    tot += ex; totS += sm
srcs = {}
def text_of(f, ln):
    if f not in srcs:
        for d in ("daily-ray-trace_b200/csrc", "include"):
            q = os.path.join(HERE, "..", d, f)
            if os.path.exists(q): srcs[f] = open(q).read().splitlines(); break
        else: srcs[f] = []
    return srcs[f][ln - 1].strip()[:100] if 0 < ln <= len(srcs[f]) else "?"
groups = collections.defaultdict(float)
print(f"total {tot / paths:.1f} warp-instr/path, {len(data)} SASS instructions")
for (f, ln), (ex, sm, thr, cnt) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    text = text_of(f, ln)
    print(f"{f[:14]:14s}{ln:5d} {ex / paths:6.2f}/path {100 * ex / tot:5.1f}% smp {100 * sm / totS:5.1f}% thr {thr / max(ex, 1):4.1f} sass {cnt:4d} | {text}")
# regions of drt_render.cuh, found by their section markers so that the table follows the source
MARKS = [("small vector algebra", "vector algebra (inlined)"), ("per-path random stream", "rng"), ("K2: closest hit", "nearest_surface / hit_*"),
         ("template <typename R> struct Hit", "closest_hit / visible"), ("BSDF evaluation reduced to basis weights", "eval weights"),
         ("dielectric Fresnel at one wavelength", "fresnel"), ("K4: the six direction samplers", "samplers"),
         ("path records in shared memory", "trace_path"), ("packed f32x2 arithmetic", "packed ops"),
         ("phase 2: spectral replay", "replay"), ("film of one pixel held by a half warp", "film"), ("the kernel */", "kernel body")]
text_of("drt_render.cuh", 1)
RANGES = []
for key, name in MARKS:
    at = next((i + 1 for i, l in enumerate(srcs["drt_render.cuh"]) if key in l), None)
    if at: RANGES.append([at, 10**9, name])
RANGES.sort()
for i in range(len(RANGES) - 1): RANGES[i][1] = RANGES[i + 1][0] - 1
for (f, ln), (ex, sm, thr, cnt) in per_line.items():
    name = f
    if f == "drt_render.cuh":
        name = next((n for a, b, n in RANGES if a <= ln <= b), "other")
    groups[name] += ex
print("--- by region (line ranges of drt_render.cuh as of the capture; other files by name)")
for name, ex in sorted(groups.items(), key=lambda kv: -kv[1]):
    print(f"  {ex / paths:7.2f}/path {100 * ex / tot:5.1f}%  {name}")
