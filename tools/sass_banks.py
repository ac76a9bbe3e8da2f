"""Register-bank pressure of the FP64 instructions in a kernel's hottest innermost loop.
Model (fitted to measurements, see profiles/README.md): the kernel is bound by operand delivery, not by
the FP64 pipe alone -- time grows with the number of FRESH 64-bit register operands per pair (reads that
miss the operand reuse cache) and with bank conflicts among them, banks being (R / 2) % 2.
usage: python tools/sass_banks.py <mangled-name-substring> [file.so|file.cubin] [pairs-per-iteration]"""
import collections, os, re, subprocess, sys

def function_sass(pattern, path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    ins, on = [], False
    for l in out.splitlines():
        if "Function :" in l:
            if on and ins:
                break
            on = pattern in l
            continue
        if on:
            m = re.match(r'\s*/\*([0-9a-f]{4})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)\s+(.*);', l)
            if m:
                ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
    return ins

def hot_loop(ins, allow_inner=False):
    """The backward-branch loop with the most FP64 instructions that has no loop inside.  With
    allow_inner, loops inside are tolerated (and cut out of the body) when they hold < 25 % of the
    body's FP64 instructions -- a rarely taken slow path, e.g. the library-log() retry of the log kernels."""
    def target(x):
        mm = re.search(r'0x([0-9a-f]+)', x[2])
        return int(mm.group(1), 16) if (x[1].startswith('BRA') and mm) else None
    fp = lambda body: sum(o.startswith(('DFMA', 'DMUL', 'DADD')) for _, o, _ in body)
    best = None
    for x in ins:
        t = target(x)
        if t is None or t >= x[0]:
            continue
        body = [y for y in ins if t <= y[0] <= x[0]]
        inner = [(target(y), y[0]) for y in body[:-1] if target(y) is not None and target(y) < y[0]]
        if inner:
            if not allow_inner:
                continue
            cut = [y for y in body if any(a <= y[0] <= b for a, b in inner)]
            if fp(cut) >= 0.25 * fp(body):
                continue
            body = [y for y in body if y not in cut]
        n = fp(body)
        if best is None or n > best[0]:
            best = (n, body)
    return best[1] if best else []

def bank_stats(body, bank=lambda r: (r // 2) % 2):
    """Operand-read statistics of the FP64 instructions of a loop body (walked twice, counted on
    the second pass, because the loop is cyclic).  An operand is FRESH (read from the register
    file) unless the immediately preceding instruction carried .reuse on the same register in
    the same operand slot.  same2 / same3: FP64 instructions with two / three fresh register
    pairs in one bank."""
    st = collections.Counter()
    cache = {}
    for pas in range(2):
        for _, o, r in body:
            ops = [x.strip() for x in r.split(',')]
            isfp = o.startswith(('DFMA', 'DMUL', 'DADD'))
            newcache, fresh = {}, []
            for slot, x in enumerate(ops[1:]):
                mm = re.search(r'(?<!U)R(\d+)(\.reuse)?', x)
                if not mm:
                    continue
                reg = int(mm.group(1))
                if cache.get(slot) != reg:
                    fresh.append(reg)
                if mm.group(2):
                    newcache[slot] = reg
            cache = newcache
            if isfp and pas == 1:
                st['fp64'] += 1
                st['fresh'] += len(set(fresh))
                c = collections.Counter(bank(x) for x in set(fresh))
                if c:
                    st['same%d' % min(max(c.values()), 3)] += 1
    st['instructions'] = len(body)
    return st


def model_ms_l7(st, pairs):
    """Fitted on 24 builds of the BVE velocity kernel measured at icosTri level 7 (R^2 = 0.86):
    ms = 41.0 + 2.38 fresh + 0.81 same2 + 2.84 same3   (all per pair)."""
    return 41.0 + (2.38 * st['fresh'] + 0.81 * st['same2'] + 2.84 * st['same3']) / pairs


if __name__ == "__main__":
    pat = sys.argv[1]
    path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "lpm_v2_b200", "liblpmgpu.so")
    st = bank_stats(hot_loop(function_sass(pat, path)))
    print(dict(st))
    if len(sys.argv) > 3:
        print("model: %.2f ms at icosTri 7" % model_ms_l7(st, int(sys.argv[3])))
