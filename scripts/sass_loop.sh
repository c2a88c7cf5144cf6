#!/bin/bash
# Compile admm_fwd.cu and print the SASS of the last VOTE.ANY-closed loop (the diagonal ADMM loop) of one kernel.
# usage: scripts/sass_loop.sh [mangled-kernel-name]
fun=${1:-_ZN2dq15admm_fwd_kernelILi8ELb0EEEvNS_9FwdParamsE}
mkdir -p /tmp/t
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v -c diffqcqp_b200/csrc/admm_fwd.cu -o /tmp/t/fwd.o 2>&1 | grep -A1 "$fun" | grep -E "registers|spill" 
cuobjdump -sass /tmp/t/fwd.o -fun "$fun" | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//' > /tmp/t/k.txt
python3 - <<'PY'
import re
L=[l.rstrip('\n') for l in open('/tmp/t/k.txt')]
# find backward branches preceded by VOTE.ANY
loops=[]
for i,l in enumerate(L):
    m=re.match(r'([0-9a-f]{4}) @P\d\s+BRA 0x([0-9a-f]+)',l)
    if m and int(m.group(2),16)<int(m.group(1),16) and 'VOTE.ANY' in L[i-1]:
        loops.append((int(m.group(2),16),int(m.group(1),16)))
print('vote-closed loops:',[(hex(a),hex(b),(b-a)//16+1) for a,b in loops])
a,b=loops[-1]
body=[l for l in L if a<=int(l[:4],16)<=b]
open('/tmp/t/loop.txt','w').write('\n'.join(body))
import collections
c=collections.Counter(re.sub(r'^@!?U?P\d\s+','',l[5:]).split()[0].split('.')[0] for l in body)
print(len(body),'instructions in loop region;',dict(c.most_common(14)))
PY
