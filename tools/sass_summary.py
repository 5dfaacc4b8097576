#!/usr/bin/env python
"""Instruction-class counts per kernel of libcae_b200.so from ``cuobjdump -sass`` (evidence that
the hot kernels are tcgen05 / TMEM / TMA code and not recompiled mma.sync):

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'cnn_autoencoder_b200', 'lib', 'libcae_b200.so')
CLASSES = [
    ('UTCHMMA  (tcgen05.mma kind::f16)', r'\bUTCHMMA'),
    ('UTCQMMA/UTCOMMA (tcgen05.mma other kinds)', r'\bUTC[QO]MMA'),
    ('UTCBAR   (tcgen05.commit)', r'\bUTCBAR'),
    ('LDTM     (tcgen05.ld, TMEM -> registers)', r'\bLDTM'),
    ('STTM     (tcgen05.st)', r'\bSTTM'),
    ('UTCATOMSWS (tcgen05.alloc/dealloc)', r'\bUTCATOMSWS'),
    ('UTMALDG  (TMA tensor load)', r'\bUTMALDG'),
    ('UTMASTG  (TMA tensor store)', r'\bUTMASTG'),
    ('UBLKCP   (bulk copy global -> shared)', r'\bUBLKCP'),
    ('SYNCS    (mbarrier)', r'\bSYNCS'),
    ('LDGSTS   (cp.async)', r'\bLDGSTS'),
    ('HMMA/IMMA (legacy mma.sync)', r'\b[HI]MMA\b'),
    ('LDG', r'\bLDG'), ('STG', r'\bSTG'), ('LDS', r'\bLDS'), ('STS', r'\bSTS'),
    ('ATOM/RED (global/shared atomics)', r'\b(ATOM|ATOMS|ATOMG|RED)\b'),
    ('SHFL/VOTE/MATCH (warp collectives)', r'\b(SHFL|VOTE|MATCH)'),
]


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = kernels.setdefault(m.group(1), [])
            continue
        if cur is not None and re.match(r'\s+/\*[0-9a-f]{4,}\*/', line):
            cur.append(line)
    demangle = subprocess.run(['c++filt'], input='\n'.join(kernels), capture_output=True, text=True).stdout.splitlines()
    print('# cuobjdump -sass %s (sm_100a), instruction-class counts per kernel' % os.path.relpath(LIB, ROOT))
    total = collections.Counter()
    for (name, lines), pretty in zip(kernels.items(), demangle):
        pretty = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', pretty)
        pretty = re.sub(r'\(.*', '', pretty)
        counts = [(label, sum(1 for l in lines if re.search(rx, l))) for label, rx in CLASSES]
        print('\n%s   [%d SASS instructions]' % (pretty, len(lines)))
        for label, n in counts:
            if n:
                print('    %-46s %6d' % (label, n))
                total[label] += n
    print('\n# library totals')
    for label, _ in CLASSES:
        if total[label]:
            print('    %-46s %6d' % (label, total[label]))


if __name__ == '__main__':
    main()
