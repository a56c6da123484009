"""Join an `ncu --page source --csv` export (per-SASS-instruction counters) with nvdisasm's line table of the same
cubin: warp instructions executed and stall samples per CUDA source line.

    python tools/ncu_lines.py gpurun_out/eval3_src.csv cat.cu cat_eval_kernelILi0 [tiles]
"""
import collections, csv, os, re, subprocess, sys, tempfile

csv_path, cu, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
csrc = os.path.join(root, "constraints_as_terminations_b200", "csrc")
cubin = os.path.join(tempfile.gettempdir(), cu + ".cubin")
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", f"-I{root}/include", f"-I{csrc}",
                "-cubin", "-o", cubin, os.path.join(csrc, cu)], check=True)
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l)
cur, off2line = None, {}
for l in dis[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > idx["Instructions Executed"] and r[idx["Instructions Executed"]].isdigit()]
base = int(data[0][0], 16)
per, samp, tot = collections.Counter(), collections.Counter(), 0
for r in data:
    key = off2line.get(int(r[0], 16) - base)
    n = int(r[idx["Instructions Executed"]])
    per[key] += n
    samp[key] += int(r[idx["# Samples"]])
    tot += n
src = open(os.path.join(csrc, cu)).read().split("\n")
print(f"total warp instructions {tot}  ({tot / units:.1f} per unit), {sum(samp.values())} samples")
order = sorted(per, key=lambda k: -samp[k]) if os.environ.get("BY_SAMPLES") else [k for k, _ in per.most_common()]
for key in order[:45]:
    n = per[key]
    txt = src[key[1] - 1].strip()[:100] if key and key[0] == cu else ""
    print(f"{100 * n / tot:5.1f}% inst {n / units:8.1f}/unit {samp[key]:5d} samp  {key}  {txt}")
