"""Blackwell-specific SASS mnemonics per kernel of the in-tree library (CPU: needs cuobjdump and c++filt).
    python tools/sass_evidence.py > profiles/r02_sass_evidence.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "super_diffusion_b200", "libsuperdiff_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
want = ["UTCHMMA", "UTMALDG", "LDTM", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "UTMACCTL", "HMMA"]
out = ["# SASS evidence, round 2 (`cuobjdump -sass super_diffusion_b200/libsuperdiff_b200.so`, sm_100a; `tools/sass_evidence.py`)\n",
       "Counts of the Blackwell-specific mnemonics per kernel (B200_PROFILING.md: UTCHMMA = tcgen05.mma kind::f16, UTMALDG = TMA tensor load,",
       "LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc / dealloc, SYNCS = mbarrier ops,",
       "UTMACCTL = tensormap prefetch; HMMA = legacy mma.sync, must be 0).\n",
       "| kernel | " + " | ".join(want) + " | instructions |", "|---|" + "---|" * (len(want) + 1)]
excerpts, tot = [], collections.Counter()
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    ins = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+([^;]+);", f)
    cnt = collections.Counter()
    for i in ins:
        m = i.strip().split()
        op = m[1] if m[0].startswith("@") and len(m) > 1 else m[0]
        cnt[op.split(".")[0]] += 1
    if not any(cnt[w] for w in want[:6]):
        continue
    dem = re.sub(r"\(.*", "", subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip())
    out.append(f"| `{dem}` | " + " | ".join(str(cnt[w]) for w in want) + f" | {len(ins)} |")
    for w in want:
        tot[w] += cnt[w]
    ex = []
    for w in ("UTMALDG", "UTCHMMA", "UTCBAR", "LDTM", "UBLKCP"):
        for line in f.split("\n"):
            if re.search(r"\b" + w, line) and "/*" in line:
                ex.append(line.strip()[:150])
                break
    excerpts.append((dem, ex))
out.append("| **total** | " + " | ".join(str(tot[w]) for w in want) + " | |")
out.append("\n## First occurrence of each mnemonic per kernel\n")
for dem, ex in excerpts:
    out.append(f"### `{dem}`\n```")
    out += ex
    out.append("```")
print("\n".join(out))
