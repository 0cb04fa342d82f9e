"""SASS opcode census of libyad.so: static counts per kernel of the instructions that prove which hardware path a kernel uses.
usage: cuobjdump -sass yolo_ad_refine_b200/lib/libyad.so | python tools/sass_census.py > profiles/r2_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

OPS = ['UTCHMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'UTCBAR', 'SYNCS', 'ELECT', 'HMMA', 'LDGSTS']
cur, cnt = None, collections.OrderedDict()
for line in sys.stdin:
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1).split('.')[0]
        cnt[cur]['_total'] += 1
        if op in OPS:
            cnt[cur][op] += 1


def dem(n):
    try:
        return subprocess.run(['cu++filt', n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n


print("# SASS opcode census of yolo_ad_refine_b200/lib/libyad.so (cuobjdump -sass, sm_100a)")
print("# static instruction counts per kernel; UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld,")
print("# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, ELECT = elect.sync, HMMA = mma.sync (legacy tensor path), LDGSTS = cp.async")
rows = []
for k, c in cnt.items():
    if any(c[o] for o in ('UTCHMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'HMMA')):
        name = re.sub(r'\(anonymous namespace\)::', '', dem(k))
        name = re.sub(r'^void ', '', name)
        name = name[:name.rfind('>(') + 1] if '>(' in name else name.split('(')[0]
        name = re.sub(r'\((?:bool|int)\)', '', name).replace('<unnamed>::', '')
        rows.append((name, c))
rows.sort(key=lambda r: (-r[1]['UTCHMMA'], r[0]))
hdr = OPS + ['_total']
print("%-72s " % "kernel" + " ".join("%8s" % h for h in hdr))
for n, c in rows:
    print("%-72s " % n[:72] + " ".join("%8d" % c[h] for h in hdr))
