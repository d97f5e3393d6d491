"""Per-kernel SASS mnemonic counts of libctradon.so (cuobjdump -sass): which kernels use the TMA bulk-copy engine
(UBLKCP), mbarriers (SYNCS), 128-bit shared-memory loads (LDS.128), clusters / distributed shared memory, system-scope
release/acquire for the peer exchange, and that no kernel contains an atomic on the data path.
  python tools/sass_counts.py > profiles/r2_sass_counts.txt"""
import collections
import os
import re
import subprocess

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ct_pvae_b200", "libctradon.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pats = collections.OrderedDict([
    ("UBLKCP (cp.async.bulk, TMA engine)", r"\bUBLKCP"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("LDS.128", r"\bLDS(\.U)?\.128"),
    ("LDS.64", r"\bLDS(\.U)?\.64"), ("STS", r"\bSTS"), ("LDG", r"\bLDG"), ("STG", r"\bSTG"), ("FFMA", r"\bFFMA"), ("DFMA/DADD/DMUL", r"\bD(FMA|ADD|MUL)"),
    ("ATOM/RED (global)", r"\b(ATOMG|ATOM|RED)\b"), ("ATOMS (shared)", r"\bATOMS"), ("UCGABAR (barrier.cluster)", r"\bUCGABAR|\bCGABAR"),
    ("ST.E (generic store: st.shared::cluster through mapa)", r"\bST\.E"), ("MEMBAR.SYS / .STRONG.SYS", r"MEMBAR\.\w*\.?SYS|\.STRONG\.SYS"),
    ("UTMALDG/UTMASTG (tensor TMA)", r"\bUTMA(LDG|STG)"), ("UTC*MMA / HMMA (tensor cores)", r"\bUTC\w*MMA|\bHMMA"), ("NANOSLEEP", r"\bNANOSLEEP")])
kern, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        kern[name] = collections.Counter()
        continue
    if name and re.match(r"\s*/\*[0-9a-f]{4,6}\*/", line):
        kern[name]["instructions"] += 1
        for k, p in pats.items():
            if re.search(p, line):
                kern[name][k] += 1
cols = ["instructions"] + list(pats)
print("arch:", ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", out)))))
for n, c in kern.items():
    print(f"\n{n}")
    print("   " + "  ".join(f"{k}: {c[k]}" for k in cols if c[k]))
tot = collections.Counter()
for c in kern.values():
    tot.update(c)
print("\nTOTAL over", len(kern), "kernels:", "  ".join(f"{k}: {tot[k]}" for k in cols if tot[k]))
