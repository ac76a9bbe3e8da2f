#!/bin/bash
# Opcode histogram of the hottest loop (the last backward-branch loop containing the most FP64 ops)
# usage: tools/sass_loop.sh <mangled-name-substring> [file.so|file.cubin]
so="${2:-$(dirname "$0")/../lpm_v2_b200/liblpmgpu.so}"
cuobjdump -sass "$so" | awk -v pat="$1" '
/Function :/ { on = index($0, pat) > 0 }
on && /\/\*[0-9a-f]{4}\*\// { print }' | grep -v "^\s*/\* 0x" | sed 's#/\* 0x[0-9a-f]* \*/##' > /tmp/_fn.sass
python3 - <<'PY'
import re,collections
lines=[l.rstrip() for l in open('/tmp/_fn.sass')]
ins=[]
for l in lines:
    m=re.match(r'\s*/\*([0-9a-f]{4})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)(.*)',l)
    if m: ins.append((int(m.group(1),16),m.group(3),m.group(4)))
addr={a:i for i,(a,_,_) in enumerate(ins)}
best=None
for i,(a,op,rest) in enumerate(ins):
    if op.startswith('BRA'):
        m=re.search(r'0x([0-9a-f]+)',rest)
        if m:
            t=int(m.group(1),16)
            if t<a and t in addr:
                body=ins[addr[t]:i+1]
                inner=any(o.startswith('BRA') and (mm:=re.search(r'0x([0-9a-f]+)',r)) and int(mm.group(1),16)<aa and int(mm.group(1),16)>=t for aa,o,r in body[:-1])
                if inner: continue
                n64=sum(1 for _,o,_ in body if o.startswith(('DFMA','DMUL','DADD')))
                if best is None or n64>best[0]: best=(n64,body,t,a)
n64,body,t,a=best
h=collections.Counter(o.split('.')[0] for _,o,_ in body)
print(f"loop 0x{t:x}..0x{a:x}: {len(body)} instructions, {n64} FP64")
print(', '.join(f"{k}:{v}" for k,v in h.most_common()))
PY
