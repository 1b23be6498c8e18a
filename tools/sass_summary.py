"""Per-kernel SASS evidence of the Blackwell-native instructions (B200_PROFILING.md table):
    python tools/sass_summary.py > profiles/r02_sass_summary.txt
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor loads/stores, UBLKCP = bulk copy,
SYNCS = mbarrier ops, LDGSTS = cp.async, HMMA = legacy mma.sync (none expected), MUFU = SFU ops."""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'parl_b200', 'csrc', 'libparl_b200.so')
PAT = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'LDGSTS', 'HMMA', 'MUFU', 'SHFL',
       'REDUX', 'ATOM']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m:
            op = m.group(1)
            kernels[cur]['_total'] += 1
            for p in PAT:
                if op.startswith(p):
                    kernels[cur][p] += 1
    demangled = subprocess.run(['c++filt'], input='\n'.join(kernels), capture_output=True, text=True).stdout.splitlines()
    print('# cuobjdump -sass parl_b200/csrc/libparl_b200.so — instruction counts per kernel (sm_100a)')
    print('# %-86s %6s %s' % ('kernel', 'SASS', ' '.join('%7s' % p for p in PAT)))
    tot = collections.Counter()
    for (k, c), name in zip(kernels.items(), demangled):
        name = re.sub(r'\(.*', '', name).replace('void ', '').replace('rl::', '')
        print('%-88s %6d %s' % (name[:88], c['_total'], ' '.join('%7d' % c[p] for p in PAT)))
        tot.update(c)
    print('%-88s %6d %s' % ('TOTAL (%d kernels)' % len(kernels), tot['_total'], ' '.join('%7d' % tot[p] for p in PAT)))


if __name__ == '__main__':
    sys.exit(main())
