"""Per-kernel count of the Blackwell-specific SASS mnemonics in the shipped library (cuobjdump -sass):
UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UTMALDG (TMA tensor load),
UBLKCP (cp.async.bulk), SYNCS (mbarrier), HMMA (legacy mma.sync), plus the instruction total.  Writes profiles/<name>."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'danet-tensorflow_b200', 'lib', 'libdanet_sm100.so')
out = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE).stdout.decode()
keys = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTCBAR', 'UTMALDG', 'UBLKCP', 'SYNCS', 'HMMA', 'MUFU']
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], stdout=subprocess.PIPE).stdout.decode().strip()
        cur = per.setdefault(re.sub(r'\(.*', '', name), collections.Counter())
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)', line)
    if m and cur is not None:
        cur['total'] += 1
        for k in keys:
            if m.group(1).startswith(k):
                cur[k] += 1
lines = ['%-58s %7s ' % ('kernel', 'instrs') + ' '.join('%8s' % k for k in keys)]
tot = collections.Counter()
for name, c in per.items():
    tot.update(c)
    if any(c[k] for k in keys[:7]) or c['HMMA']:
        lines.append('%-58s %7d ' % (name[-58:], c['total']) + ' '.join('%8d' % c[k] for k in keys))
lines.append('%-58s %7d ' % ('ALL %d KERNELS OF THE LIBRARY' % len(per), tot['total']) + ' '.join('%8d' % tot[k] for k in keys))
text = '\n'.join(lines) + '\n'
dst = os.path.join(ROOT, 'profiles', sys.argv[1] if len(sys.argv) > 1 else 'r02_sass_summary.txt')
open(dst, 'w').write('cuobjdump -sass danet-tensorflow_b200/lib/libdanet_sm100.so (sm_100a), kernels that use tcgen05 / TMA / bulk copies / mma.sync\n' + text)
print(text)
