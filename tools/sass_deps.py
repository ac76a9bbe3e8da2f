"""Dependency distances between FP64 instructions in the hottest innermost loop of a kernel
(after tools/sass_loop.sh <pattern> has written /tmp/_fn.sass).  A producer->consumer distance of
1-2 FP64 instructions means the warp stalls on DFMA latency unless other warps cover it."""
import collections, re, sys
ins = []
for l in open('/tmp/_fn.sass'):
    m = re.match(r'\s*/\*([0-9a-f]{4})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)\s+(.*);', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
best = None
for i, (a, op, rest) in enumerate(ins):
    if op.startswith('BRA'):
        m = re.search(r'0x([0-9a-f]+)', rest)
        if m and int(m.group(1), 16) < a:
            t = int(m.group(1), 16)
            body = [x for x in ins if t <= x[0] <= a]
            inner = False
            for aa, o, r in body[:-1]:
                mm = re.search(r'0x([0-9a-f]+)', r)
                if o.startswith('BRA') and mm and t <= int(mm.group(1), 16) < aa:
                    inner = True
            if inner:
                continue
            n = sum(o.startswith(('DFMA', 'DMUL', 'DADD')) for _, o, _ in body)
            if best is None or n > best[0]:
                best = (n, body)
body = best[1]
last_write = {}
dist = []
k = 0
for _, o, r in body:
    ops = r.split(',')
    dst = re.match(r'\s*R(\d+)', ops[0])
    if o.startswith(('DFMA', 'DMUL', 'DADD')):
        srcs = [int(x) for x in re.findall(r'R(\d+)', ','.join(ops[1:]))]
        srcs += [s + 1 for s in srcs]
        d = min([k - last_write[s] for s in srcs if s in last_write] or [99])
        dist.append(d)
        if dst:
            last_write[int(dst.group(1))] = k
            last_write[int(dst.group(1)) + 1] = k
        k += 1
c = collections.Counter(min(d, 8) for d in dist)
print(f"{len(body)} instructions, {k} FP64; distance to the producing FP64 instruction (1 = back to back, 8 = 8 or more / none):")
print("  " + ", ".join(f"{d}: {n}" for d, n in sorted(c.items())))
