#!/usr/bin/env python
"""Per-kernel SASS evidence for profiles/ (runs in the build container, no GPU needed):
`cuobjdump -sass boosting_rcnn_b200/libbrcnn.so`, split per kernel; for the kernels named on the
command line (default: the TMA / mbarrier / cluster / packed-FMA kernels) write the opcode
histogram and every line carrying one of the mnemonics that prove the design
(UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, LDGSTS = cp.async, FFMA2 = fma.rn.f32x2,
UCGABAR / CGAERRBAR / MEMBAR.ALL.GPU... = cluster barriers, ATOM/RED = atomics).

  python tools/sass_extract.py r02            # -> profiles/r02_sass_<kernel>.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'boosting_rcnn_b200', 'libbrcnn.so')
DEFAULT = ['roi_align_fwd3_kernel', 'roi_bwd_gather5_kernel', 'roi_align_fwd_tma_kernel',
           'rpn_nms_image_kernel', 'rpn_loss_main_kernel', 'roi_bwd_prep_kernel']
KEY = re.compile(r'\b(UBLKCP|SYNCS|LDGSTS|FFMA2|UCGABAR|CGAERRBAR|ATOMS?|ATOMG|RED|REDUX|BAR|'
                 r'LDS|STS|STG|LDG|MATCH|VOTE|SHFL|BRX)\b')
PROOF = re.compile(r'\b(UBLKCP|SYNCS|LDGSTS|FFMA2|UCGABAR|CGAERRBAR|MAPA|ATOMG|RED\.)')


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r02'
    want = sys.argv[2:] or DEFAULT
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    kernels, cur = {}, None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
        elif cur is not None:
            kernels[cur].append(line)
    os.makedirs(os.path.join(ROOT, 'profiles'), exist_ok=True)
    for w in want:
        for name, lines in kernels.items():
            if w not in name:
                continue
            ops = collections.Counter()
            proof = []
            n = 0
            for l in lines:
                m = re.search(r'/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
                if not m:
                    continue
                n += 1
                ops[m.group(2).split('.')[0]] += 1
                if PROOF.search(l):
                    proof.append(re.sub(r'\s*/\* 0x[0-9a-f]+ \*/\s*$', '', l).strip())
            path = os.path.join(ROOT, 'profiles', f'{tag}_sass_{w}.txt')
            with open(path, 'w') as f:
                f.write(f'# cuobjdump -sass boosting_rcnn_b200/libbrcnn.so  (sm_100a)\n# {name}\n')
                f.write(f'# {n} SASS instructions\n\n## opcode histogram\n')
                for op, c in ops.most_common():
                    f.write(f'{c:6d}  {op}\n')
                f.write('\n## lines with TMA / mbarrier / cp.async / packed-FMA / cluster mnemonics\n')
                f.write('\n'.join(proof[:400]) + '\n')
            print(path, n, 'instrs;', {k: ops[k] for k in ('UBLKCP', 'SYNCS', 'LDGSTS', 'FFMA2') if ops[k]})
            break


if __name__ == '__main__':
    main()
