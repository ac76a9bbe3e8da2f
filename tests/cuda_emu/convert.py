"""TEST INFRASTRUCTURE ONLY: rewrite the CUDA-only syntax of lpm_v2_b200/csrc/* so that g++ can compile the
library against the SIMT emulator (emu_core.h, emu_runtime.h).  Four rewrites, nothing else is touched:

  kernel<targs><<<grid, block, smem, stream>>>(args);   ->  emu_launch[_seq](grid, block, [&]() { kernel<targs>(args); });
  extern __shared__ __align__(N) unsigned char name[];   ->  unsigned char* name = ::emu_dynamic_smem();
  __shared__ T x[...];                                    ->  static T x[...];     (one CTA runs at a time)
  __noinline__                                            ->  __attribute__((noinline))

A kernel whose body (or a device function it calls, by name) synchronises or shuffles runs with one OS thread per
CUDA thread; every other kernel runs as a plain loop over its threads.

usage: python convert.py <csrc dir> <out dir>"""
import os
import re
import sys

NEEDS_THREADS = re.compile(r'__syncthreads|__shfl_|__ballot_sync|mbar_wait|block_inclusive_scan')


def kernel_bodies(text):
    """{kernel name: body text} for every __global__ function defined in `text`."""
    out = {}
    for m in re.finditer(r'__global__', text):
        # the name is the identifier before the first '(' that opens the parameter list
        i = text.index('(', m.end())
        # skip __launch_bounds__(...) if it comes first
        head = text[m.end():i]
        if '__launch_bounds__' in head:
            depth, j = 0, i
            while True:
                depth += text[j] == '('
                depth -= text[j] == ')'
                j += 1
                if depth == 0:
                    break
            i = text.index('(', j)
            head = text[j:i]
        name = re.findall(r'[A-Za-z_]\w*', head)[-1]
        depth, j = 0, i
        while True:                                  # parameter list
            depth += text[j] == '('
            depth -= text[j] == ')'
            j += 1
            if depth == 0:
                break
        k = j
        while text[k] in ' \n\t':
            k += 1
        if text[k] != '{':
            continue                                 # a declaration / explicit instantiation
        depth, e = 0, k
        while True:
            depth += text[e] == '{'
            depth -= text[e] == '}'
            e += 1
            if depth == 0:
                break
        out[name] = text[k:e]
    return out


def split_top(s):
    parts, depth, cur = [], 0, ''
    for ch in s:
        if ch in '([{':
            depth += 1
        elif ch in ')]}':
            depth -= 1
        if ch == ',' and depth == 0:
            parts.append(cur.strip())
            cur = ''
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite_launches(text, threaded):
    out, pos = '', 0
    while True:
        i = text.find('<<<', pos)
        if i < 0:
            return out + text[pos:]
        # kernel expression: identifier, optionally followed by <template args>, right before '<<<'
        j = i
        while text[j - 1] in ' \n\t':
            j -= 1
        k = j
        if text[k - 1] == '>':
            depth = 0
            while True:
                k -= 1
                depth += text[k] == '>'
                depth -= text[k] == '<'
                if depth == 0:
                    break
        while text[k - 1].isalnum() or text[k - 1] in '_:':
            k -= 1
        kexpr = text[k:j]
        name = re.match(r'[\w:]+', kexpr).group(0).split('::')[-1]
        e = text.index('>>>', i)
        cfg = split_top(text[i + 3:e])
        a0 = e + 3
        while text[a0] in ' \n\t':
            a0 += 1
        assert text[a0] == '(', (kexpr, text[a0:a0 + 20])
        depth, a1 = 0, a0
        while True:
            depth += text[a1] == '('
            depth -= text[a1] == ')'
            a1 += 1
            if depth == 0:
                break
        args = text[a0 + 1:a1 - 1]
        s1 = a1
        while text[s1] in ' \n\t':
            s1 += 1
        assert text[s1] == ';', (kexpr, text[s1:s1 + 20])
        fn = 'emu_launch' if name in threaded else 'emu_launch_seq'
        out += text[pos:k] + f'{fn}((unsigned)({cfg[0]}), (unsigned)({cfg[1]}), [&]() {{ {kexpr}({args}); }});'
        pos = s1 + 1


def convert(src_dir, out_dir):
    os.makedirs(out_dir, exist_ok=True)
    files = [f for f in sorted(os.listdir(src_dir)) if f.endswith(('.cuh', '.cu', '.inc', '.cpp', '.h'))]
    texts = {f: open(os.path.join(src_dir, f)).read() for f in files}
    bodies = {}
    for t in texts.values():
        bodies.update(kernel_bodies(t))
    threaded = {n for n, b in bodies.items() if NEEDS_THREADS.search(b)}
    for f, t in texts.items():
        t = re.sub(r'extern\s+__shared__\s+__align__\(\d+\)\s+unsigned char (\w+)\[\];',
                   r'unsigned char* \1 = ::emu_dynamic_smem();', t)
        t = re.sub(r'\b__shared__\s+', 'static ', t)
        t = t.replace('__noinline__', '__attribute__((noinline))')
        t = rewrite_launches(t, threaded)
        name = f[:-3] + '.cpp' if f.endswith('.cu') else f
        open(os.path.join(out_dir, name), 'w').write(t)
    return sorted(threaded), sorted(set(bodies) - threaded)


if __name__ == '__main__':
    th, seq = convert(sys.argv[1], sys.argv[2])
    print('threaded kernels:', ' '.join(th))
    print('sequential kernels:', ' '.join(seq))
