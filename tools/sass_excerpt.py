"""SASS evidence of the Blackwell-native instructions (tcgen05.mma / tcgen05.ld,st / cp.async.bulk / mbarrier) in the
built library, summarised per kernel:  python tools/sass_excerpt.py > profiles/r2_sass_tcgen05.txt   (needs cuobjdump)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ('UTCHMMA', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'UTCBAR', 'SYNCS')


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'dccf_b200', 'libdccf_b200.so')
    out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
    kern, stats, lines_by = None, collections.OrderedDict(), {}
    for ln in out.splitlines():
        m = re.search(r'Function : (\S+)', ln)
        if m:
            kern = m.group(1)
            stats[kern], lines_by[kern] = collections.Counter(), []
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m and kern:
            ins = m.group(2).strip()
            op = ins.split()[1] if ins.startswith('@') else ins.split()[0]
            stats[kern][op.split('.')[0]] += 1
            if re.match(r'(UTCHMMA|UTCQMMA|LDTM|STTM|UBLKCP|UTMALDG|UTCBAR|SYNCS|UTCATOM)', op):
                lines_by[kern].append((op.split('.')[0], '        /*%s*/  %s ;' % (m.group(1), ins)))
    print('# SASS evidence of the Blackwell-native instructions in dccf_b200/libdccf_b200.so (sm_100a), generated in the build\n'
          '# container by tools/sass_excerpt.py (cuobjdump -sass).  Per kernel: instruction count, count of each tensor-core /\n'
          '# TMEM / bulk-copy / mbarrier mnemonic, and the first occurrences with their addresses.\n'
          '#   UTCHMMA  = tcgen05.mma (kind::tf32 here)          LDTM / STTM = tcgen05.ld / tcgen05.st (tensor memory)\n'
          '#   UBLKCP   = cp.async.bulk (TMA engine, 1-D bulk)    UTCBAR      = tcgen05.commit -> mbarrier\n'
          '#   SYNCS    = mbarrier init / arrive / try_wait        UTCATOMSWS  = tcgen05.alloc / dealloc\n')
    for k, c in stats.items():
        keys = [m for m in KEYS if c.get(m)]
        if not any(m in keys for m in ('UTCHMMA', 'LDTM', 'UBLKCP')):
            continue
        name = subprocess.run(['c++filt', k], capture_output=True, text=True).stdout.strip()
        print('%s\n    %d instructions; %s' % (name, sum(c.values()), ', '.join('%s x%d' % (m, c[m]) for m in keys)))
        seen = collections.Counter()
        for op, l in lines_by[k]:
            seen[op] += 1
            if seen[op] <= 3:
                print(l)
        print()


if __name__ == '__main__':
    main()
