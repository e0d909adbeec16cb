"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (tcgen05 / TMEM / TMA / clusters) in the shipped library:
    python tests/sass_summary.py > profiles/r02_sass_summary.txt
Reads `cuobjdump -sass ctgan_b200/libctgan_sm100.so`; not a test."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
SO = os.path.join(ROOT, 'ctgan_b200', 'libctgan_sm100.so')
WATCH = ['UTCHMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'UTCBAR', 'UTCATOMSWS', 'SYNCS', 'UCGABAR_ARV', 'UCGABAR_WAIT', 'REDG', 'ACQBULK', 'FFMA', 'HMMA']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', SO], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(['cu++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True, text=True).stdout.split('\n')
    kernels, cur, idx = [], None, 0
    for line in sass.split('\n'):
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = [names[idx] if idx < len(names) else m.group(1), collections.Counter(), collections.Counter()]
            idx += 1
            kernels.append(cur)
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)', line)
        if m and cur is not None:
            op = m.group(1)
            cur[1][op.split('.')[0]] += 1
            if op.split('.')[0] in ('UTCHMMA', 'UTMALDG', 'REDG', 'LDTM'):
                cur[2][op] += 1
    total = collections.Counter()
    print('SASS summary of %s (sm_100a), %d kernels; counts of static instructions per kernel' % (os.path.basename(SO), len(kernels)))
    print('mnemonics: UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load (cp.async.bulk.tensor), LDTM = tcgen05.ld (TMEM -> registers),')
    print('UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier arrive / try_wait, UCGABAR_* = barrier.cluster (the split-K kernel reduces through')
    print('distributed shared memory), REDG...F32x4 = red.global.add.v4.f32, UTCATOMSWS = tcgen05.alloc / dealloc; HMMA (mma.sync) = 0 everywhere.')
    print('kind::f16 vs kind::tf32 is a field of the instruction descriptor (idesc[UR]) and not visible in the mnemonic: the tf32 kernels are the *_tf32_* ones.\n')
    for name, c, full in sorted(kernels, key=lambda k: -k[1]['UTCHMMA']):
        for w in WATCH:
            total[w] += c[w]
        if not (c['UTCHMMA'] or c['UTMALDG'] or c['UCGABAR_ARV']):
            continue
        short = re.sub(r'\((?:[^()]|\([^()]*\))*\)\s*$', '', name).replace('(int)', '').replace('void ', '').replace('ctgan::', '')
        print('%-58s %s' % (short[:58], '  '.join('%s=%d' % (w, c[w]) for w in WATCH if c[w] and w not in ('FFMA', 'ATOM'))))
        variants = ['%s x%d' % (k, v) for k, v in sorted(full.items()) if k.startswith(('UTMALDG', 'REDG', 'LDTM'))]
        print('    ' + ', '.join(variants))
    print('\nwhole library: ' + '  '.join('%s=%d' % (w, total[w]) for w in WATCH))


if __name__ == '__main__':
    sys.exit(main())
