"""Search BveVelT<4, ORDER> statement orders for the fewest register-bank conflicts in the hot loop.

Compiles the BVE velocity kernel alone (a one-instantiation translation unit, ~1.5 s) for many
ORDER seeds, scores each hot loop with the bank model of tools/sass_banks.py
(fresh operand reads + bank conflicts -> predicted ms at icosTri 7) and prints the best; the GPU sweep
(tools/sweep_bve.py) then measures the short list.   usage: search_order.py [n_random] [T] [U] [min CTAs per SM]
FENCED=1 searches the kernel with a scheduling fence after every source (Fenced<>, directsum.cuh)."""
import concurrent.futures as cf
import os, random, shutil, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from sass_banks import bank_stats, function_sass, hot_loop, model_ms_l7

TU = r'''
#include "directsum.cuh"
#include "pairs.cuh"
using namespace lpm;
#ifdef FENCED
template __global__ void lpm::ds_kernel<Fenced<BveVelT<4, ORD_SEED>>, TT, 128, UU, LB_MIN>(const BveVelParams, const DsGeom, const double*, const int32_t*, double*);
#else
template __global__ void lpm::ds_kernel<BveVelT<4, ORD_SEED>, TT, 128, UU, LB_MIN>(const BveVelParams, const DsGeom, const double*, const int32_t*, double*);
#endif
'''

def score(seed, T, U, work, minb=1):
    cu = os.path.join(work, "tu.cu")
    out = os.path.join(work, f"o{seed}_{T}_{U}_{minb}.cubin")
    r = subprocess.run(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                        f"-I{ROOT}/include", f"-I{ROOT}/lpm_v2_b200/csrc", f"-DORD_SEED={seed}", f"-DTT={T}", f"-DUU={U}", f"-DLB_MIN={minb}",
                        *(["-DFENCED"] if os.environ.get("FENCED") == "1" else []), "-cubin", "-o", out, cu], capture_output=True, text=True)
    if r.returncode != 0:
        return seed, None
    st = bank_stats(hot_loop(function_sass("ds_kernel", out), allow_inner=True))
    os.remove(out)
    pairs = T * U
    return seed, {"cost": model_ms_l7(st, pairs), "same2": st["same2"] / pairs, "fresh": st["fresh"] / pairs,
                  "same3": st["same3"] / pairs, "instr": st["instructions"] / pairs, "fp64": st["fp64"] / pairs}

def decode(seed):
    return dict(pd=(seed & 31) % 24, pa=((seed >> 5) & 31) % 24, dn=(seed >> 10) & 1, an=(seed >> 11) & 3, pt=((seed >> 13) & 3) % 3)

def make_seed(pd, pa, dn, an, pt):
    return pd | (pa << 5) | (dn << 10) | (an << 11) | (pt << 13)

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    U = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    MINB = int(sys.argv[4]) if len(sys.argv) > 4 else 1      # __launch_bounds__ min CTAs per SM (register cap)
    rng = random.Random(2024)
    seeds = {0}
    while len(seeds) < n:
        seeds.add(make_seed(rng.randrange(24), rng.randrange(24), rng.randrange(2), rng.randrange(3), rng.randrange(3)))
    work = tempfile.mkdtemp(prefix="order_search_")
    open(os.path.join(work, "tu.cu"), "w").write(TU)
    res = []
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        for seed, sc in ex.map(lambda s: score(s, T, U, work, MINB), sorted(seeds)):
            if sc:
                res.append((sc["cost"], seed, sc))
    shutil.rmtree(work, ignore_errors=True)
    res.sort()
    for cost, seed, sc in res[:16]:
        print(f"ORDER {seed:6d} {decode(seed)}  model {cost:.2f} ms  fresh {sc['fresh']:.2f} same2 {sc['same2']:.2f} same3 {sc['same3']:.2f} instr {sc['instr']:.2f}")
    print("...")
    for cost, seed, sc in res[-3:]:
        print(f"ORDER {seed:6d} model {cost:.2f} ms")
