"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): tcgen05.mma ->
UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, vector reductions -> REDG...F32x4, CREDUX, and the
warp-level HMMA (expected ONLY in head_mma_kernel: the 12- / 1-wide head products are mma.sync tf32 fragments, DESIGN.md §4;
the hidden-layer GEMMs must show none).   python tools/sass_summary.py > profiles/r2_sass_summary.txt   (no GPU needed)"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "constraints_as_terminations_b200", "lib", "libcatb200.so")
pat = re.compile(r"\b(UTC[A-Z0-9]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTMAPF|CREDUX|HMMA|HGMMA|REDG|ATOMG|SYNCS|UTCBAR|UTCATOMSWS)\b")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
arch = set(re.findall(r"arch = (sm_\w+)", out))
counts, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void catb200::", "").replace("catb200::", "")
        counts[name] = collections.Counter()
        continue
    if name:
        m = pat.search(line)
        if m:
            key = m.group(1)
            if key == "REDG" and "F32x4" in line:
                key = "REDG.F32x4"
            counts[name][key] += 1
cols = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "REDG.F32x4", "REDG", "CREDUX", "SYNCS", "HMMA"]
print(f"# {os.path.relpath(so, ROOT)}: cubin architectures {sorted(arch)}; cuobjdump -sass mnemonic counts per kernel")
print(f"{'kernel':58s} " + " ".join(f"{c:>10s}" for c in cols))
tot = collections.Counter()
for k, c in counts.items():
    c2 = collections.Counter()
    for key, v in c.items():
        c2["UTCHMMA" if key.startswith("UTC") and key.endswith("MMA") else key] += v
    tot.update(c2)
    print(f"{k[:58]:58s} " + " ".join(f"{c2.get(col, 0):10d}" for col in cols))
print(f"{'TOTAL':58s} " + " ".join(f"{tot.get(col, 0):10d}" for col in cols))
