"""Counts the Blackwell-native SASS mnemonics per kernel of libimgcomp_b200.so (no GPU needed):
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UBLKCP = TMA tensor / bulk copies, UTCBAR = tcgen05.commit,
HMMA = legacy mma.sync (none expected).   python tools/sass_evidence.py > profiles/r1_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'imgcomp_cvpr_b200', 'libimgcomp_b200.so')
MNEMONICS = ('UTCHMMA', 'UTCQMMA', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDTM', 'STTM', 'HMMA')


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    cur, stats = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            stats[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for mn in MNEMONICS:
            if re.search(r'\b' + mn + r'[\.\s]', line):
                stats[cur][mn] += 1
    print('SASS mnemonics per kernel of %s (cuobjdump -sass, sm_100a)' % os.path.basename(LIB))
    for fn, c in stats.items():
        if any(c[k] for k in ('UTCHMMA', 'UTMALDG', 'LDTM', 'UBLKCP')):
            dem = subprocess.run(['c++filt', fn], capture_output=True, text=True).stdout.strip()
            dem = re.sub(r'ic::\(anonymous namespace\)::|ic::tc::\(anonymous namespace\)::', '', dem).split('(')[0]
            print('%-60s %s' % (dem[:60], ' '.join('%s=%d' % (k, v) for k, v in c.items() if v)))
    print('%d kernels in the library, %d with tcgen05 MMAs, %d with legacy HMMA (mma.sync)' % (
        len(stats), sum(1 for c in stats.values() if c['UTCHMMA']), sum(1 for c in stats.values() if c['HMMA'])))


if __name__ == '__main__':
    sys.exit(main())
