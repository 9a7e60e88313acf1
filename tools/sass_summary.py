"""Per-kernel count of the tcgen05 / TMA / mbarrier / cluster SASS mnemonics in the built library (no GPU needed):
    python tools/sass_summary.py > profiles/r1_sass_tensor_kernels.txt
UTCHMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor (TMA), UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc / dealloc,
SYNCS.* = mbarrier, UCGABAR = barrier.cluster; the .2CTA / .MULTICAST suffixes are the cta_group::2 / multicast forms."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'lstm_unet_b200', 'liblstm_unet_b200.so')
PAT = re.compile(r'\s(UTC[A-Z0-9]+(?:\.[A-Z0-9_]+)*|UTMALDG(?:\.[A-Z0-9_]+)*|SYNCS(?:\.[A-Z0-9_]+)*|UCGABAR_[A-Z]+|LDTM(?:\.[A-Za-z0-9_]+)*)\s')


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], check=True, capture_output=True, text=True).stdout
    names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True, text=True).stdout.split('\n')
    demangled = dict(zip(re.findall(r'Function : (\S+)', sass), names))
    cur, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        if re.search(r'/\*[0-9a-f]{4}\*/', line):
            total[cur] += 1
        m = PAT.search(line)
        if m:
            counts[cur][m.group(1)] += 1
    print('# tcgen05 / TMA / mbarrier SASS of every kernel of liblstm_unet_b200.so that has any (cuobjdump -sass, sm_100a)')
    for k in sorted(counts, key=lambda k: demangled[k]):
        if not any(op.startswith(('UTC', 'UTMALDG', 'LDTM')) for op in counts[k]):
            continue
        print('\n%s   [%d instructions]' % (demangled[k], total[k]))
        for op, c in sorted(counts[k].items()):
            print('    %-44s %d' % (op, c))


if __name__ == '__main__':
    sys.exit(main())
