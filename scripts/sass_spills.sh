#!/bin/bash
# usage: scripts/sass_spills.sh numpyro_b200/csrc  -> compiles the KS=7 instances of the streaming kernel (writes ks7.o / ks7.sass there) and counts
# the local-memory instructions between the first and the last HMMA of stream_engine_kernel<7,0,false> (the sweep): must print 0
cd $1 && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false --extended-lambda -Xcompiler -fPIC -Xptxas -v -DB2_INST_KS=7 -c -o ks7.o stream_instances.cu 2>&1 | grep -A2 "Compiling entry function '_ZN2b220stream_engine_kernelILi7ELi0ELb0" | grep spill
cuobjdump -sass -fun '_ZN2b220stream_engine_kernelILi7ELi0ELb0EEEvNS_12StreamParamsE' ks7.o > ks7.sass
python3 - <<PY
import re
L=[l for l in open('ks7.sass') if re.search(r'/\*[0-9a-f]{4,}\*/',l)]
ops=[]
for l in L:
    m=re.search(r'\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)',l)
    ops.append(m.group(2) if m else '')
h=[i for i,o in enumerate(ops) if o.startswith('HMMA')]
lo,hi=h[0],h[-1]
pos=[i for i in range(lo,hi+1) if ops[i].startswith(('LDL','STL'))]
print('sass',len(ops),'mma region',hi-lo,'spill instrs in region',len(pos))
PY
