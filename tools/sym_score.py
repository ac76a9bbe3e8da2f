"""Score builds of the EXPERIMENTAL symmetric BVE kernel (csrc/sym_kernels.cuh) without a GPU.

Compiles sym_bve_kernel<T, 128, SB, MINB, ORDER> alone (a one-instantiation translation unit) and prints
the operand-delivery statistics of its hot loop (tools/sass_banks.py) per INTERACTION (a symmetric
pair is two interactions) with the cycle model that matches the two measured kernels
(BVE velocity: model 20.8 / measured 20.3 cycles per interaction per SM sub-partition; BVE stream
functions: 29.6 / 32.3):   cycles = max(2 * FP64 instructions, fresh + 0.5 same2 + same3).

usage: sym_score.py [T] [SB] [MINB] [order ...]        (orders default to 0)
       SYM_KERNEL=SymBveStream SYM_BLOCK=256 sym_score.py 4 4 2"""
import concurrent.futures as cf
import os, shutil, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from sass_banks import bank_stats, function_sass, hot_loop

TU = r'''
#include "sym_kernels.cuh"
using namespace lpm;
template __global__ void lpm::sym_kernel<KK, TT, BLK, SBB, LB_MIN, ORD>(const SymParams, const SymGeom, const double*, double*);
'''

def cycles(st):
    return max(2.0 * st["fp64"], st["fresh"] + 0.5 * st["same2"] + st["same3"])

def score(order, T, SB, minb, work, kern="SymBveVel", block=128):
    out = os.path.join(work, f"o{order}_{T}_{SB}_{minb}_{kern}_{block}.cubin")
    r = subprocess.run(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                        f"-I{ROOT}/include", f"-I{ROOT}/lpm_v2_b200/csrc", f"-DORD={order}", f"-DTT={T}", f"-DSBB={SB}",
                        f"-DLB_MIN={minb}", f"-DKK={kern}", f"-DBLK={block}", "-Xptxas", "-v", "-cubin", "-o", out, os.path.join(work, "tu.cu")],
                       capture_output=True, text=True)
    if r.returncode != 0:
        return order, None, r.stderr[-400:]
    lines = r.stderr.splitlines()
    at = max((i for i, l in enumerate(lines) if "Compiling entry function" in l and "sym_kernel" in l), default=0)
    regs = [l.split("Used")[1].split(",")[0].strip() for l in lines[at:at + 4] if "registers" in l]
    spill = [l.strip() for l in lines[at:at + 4] if "spill" in l]
    st = bank_stats(hot_loop(function_sass("sym_kernel", out), allow_inner=True))
    os.remove(out)
    inter = 2.0 * T * SB
    return order, {k: v / inter for k, v in st.items()} | {"cycles": cycles(st) / inter}, (regs[0] if regs else "?") + " | " + (spill[0] if spill else "")

if __name__ == "__main__":
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    SB = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    MINB = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    orders = [int(x) for x in sys.argv[4:]] or [0]
    work = tempfile.mkdtemp(prefix="sym_score_")
    open(os.path.join(work, "tu.cu"), "w").write(TU)
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        res = list(ex.map(lambda o: score(o, T, SB, MINB, work, os.environ.get('SYM_KERNEL', 'SymBveVel'),
                                          int(os.environ.get('SYM_BLOCK', '128'))), orders))
    shutil.rmtree(work, ignore_errors=True)
    for order, sc, info in sorted(res, key=lambda r: (r[1] or {}).get("cycles", 1e9)):
        if sc is None:
            print(f"ORDER {order}: compile failed: {info}")
            continue
        print(f"ORDER {order:5d}  cycles/interaction {sc['cycles']:.2f}  fp64 {sc['fp64']:.2f} fresh {sc['fresh']:.2f} "
              f"same2 {sc['same2']:.2f} same3 {sc['same3']:.2f} instr {sc['instructions']:.2f}   {info}")
