"""Writes profiles/r02_sass_evidence.txt: per kernel, the tcgen05 / TMA / TMEM SASS mnemonics found in the built library.
    python scripts/sass_evidence.py            # needs cuobjdump + c++filt, no GPU"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "gan_b200", "csrc", "libgan_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r'\b(UTCHMMA(?:\.2CTA)?|UTMALDG[.\w]*|UTCBAR[.\w]*|LDTM[.\w]*|UTMAPF[.\w]*|SYNCS[.\w]*|UCGABAR[.\w]*|'
                 r'RED\.E\.ADD[.\w]*|UTCATOMSWS[.\w]*)')
fn, counts = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        fn = m.group(1)
        continue
    if fn:
        for t in pat.findall(line):
            counts[fn][t] += 1
rows = []
for f, c in counts.items():
    if not any(k.startswith(('UTCHMMA', 'UTMALDG', 'LDTM')) for k in c):
        continue
    name = subprocess.run(['c++filt', f], capture_output=True, text=True).stdout.strip()
    rows.append((re.sub(r'\((?!anonymous namespace\)).*$', '', name), sorted(c.items())))
rows.sort(key=lambda r: r[0])
with open(os.path.join(ROOT, "profiles", "r02_sass_evidence.txt"), "w") as fh:
    fh.write("cuobjdump -sass gan_b200/csrc/libgan_b200.so (sm_100a): tensor-core / TMA / TMEM mnemonics per kernel\n"
             "UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2); UTMALDG = cp.async.bulk.tensor (.2CTA = pair-wide\n"
             "barrier signalling); LDTM = tcgen05.ld; UTCBAR = tcgen05.commit (.MULTICAST = multicast::cluster);\n"
             "SYNCS = mbarrier operations; UCGABAR = cluster barrier; RED.E.ADD = red.global.add\n\n")
    for n, c in rows:
        fh.write(n + "\n")
        for k, v in c:
            fh.write(f"    {k:44s} {v}\n")
print(open(os.path.join(ROOT, "profiles", "r02_sass_evidence.txt")).read())
