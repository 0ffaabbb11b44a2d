"""SASS mnemonic counts per kernel (cuobjdump -sass on the objects behind libvsom_b200.so) -> profiles/r02_sass_mnemonics.txt.
Run here (no GPU needed) after a build: python tests/micro/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt"""
import glob, os, re, subprocess, sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
COLS = [("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("UTCHMMA.2CTA", r"UTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"UTMALDG[.\w]*"), ("UTMALDG.2CTA", r"UTMALDG[.\w]*\.2CTA"),
        ("UTCBAR", r"\bUTCBAR\b"), ("UTCBAR.MULTICAST", r"UTCBAR\.2CTA\.MULTICAST"), ("UCGABAR", r"UCGABAR_"), ("ELECT", r"\bELECT\b"), ("SYNCS", r"\bSYNCS"),
        ("FADD2", r"\bFADD2"), ("FMUL2", r"\bFMUL2"), ("FFMA2", r"\bFFMA2"), ("LDGSTS", r"\bLDGSTS"), ("FMNMX3", r"\bFMNMX3"), ("MUFU.RCP", r"MUFU\.RCP"),
        ("REDUX", r"\bC?REDUX"), ("HMMA(legacy)", r"\bHMMA\b"), ("instructions", r"^\s+/\*[0-9a-f]{4}\*/")]
print("# SASS mnemonic counts per kernel (cuobjdump -sass on the objects of libvsom_b200.so, round 2, tests/micro/sass_mnemonics.py)")
print("# UTCHMMA = tcgen05.mma (kind::f16; .2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG = TMA load (cp.async.bulk.tensor), UTCBAR = tcgen05.commit,")
print("# UCGABAR = cluster barrier, ELECT = elect.sync, SYNCS = mbarrier ops, FADD2/FMUL2/FFMA2 = packed f32x2, LDGSTS = cp.async, FMNMX3 = three-input min")
print("kernel | " + " | ".join(c for c, _ in COLS))
for obj in sorted(glob.glob(os.path.join(REPO, "variational-self-organizing-maps_b200", "lib", "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, lines = None, {}
    for l in out.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            lines[cur] = []
        elif cur:
            lines[cur].append(l)
    for k, ls in lines.items():
        text = "\n".join(ls)
        counts = [len(re.findall(p, text, re.M)) for _, p in COLS]
        print(f"{os.path.basename(obj)}: {k} | " + " | ".join(str(c) for c in counts))
