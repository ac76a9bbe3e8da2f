"""Compare the SASS of every kernel of two builds of liblpmgpu.so (instruction text, addresses and encodings ignored).
Used to show that additions to the translation unit left the measured kernels untouched.
usage: python tools/sass_diff.py old/liblpmgpu.so new/liblpmgpu.so"""
import hashlib
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    res, name, body = {}, None, []
    for line in out.splitlines():
        if "Function :" in line:
            if name:
                res[name] = hashlib.md5("\n".join(body).encode()).hexdigest()
            name, body = line.split("Function :")[1].strip(), []
        else:
            m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(.*?);', line)
            if m:
                body.append(m.group(1))
    if name:
        res[name] = hashlib.md5("\n".join(body).encode()).hexdigest()
    return res


if __name__ == "__main__":
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    changed = [k for k in a if k in b and a[k] != b[k]]
    print(f"kernels: {len(a)} -> {len(b)}; identical {sum(1 for k in a if k in b and a[k] == b[k])}, changed {len(changed)}, "
          f"removed {sum(1 for k in a if k not in b)}, new {sum(1 for k in b if k not in a)}")
    for k in changed:
        print(" changed:", subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()[:160])
