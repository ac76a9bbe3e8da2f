#!/bin/bash
# EXPERIMENT (not the product build): build/renamed/lpm_v2_b200/ = a copy of the Python package with a
# liblpmgpu.so whose default BVE velocity kernel went through tools/sass_rename.py.
# It replays nvcc's own steps (nvcc -dryrun -keep) and patches lpm_gpu.cubin between ptxas and fatbinary.
# Then, on a B200:   python tools/ab_bve.py . 7;  python tools/ab_bve.py build/renamed 7     (one gpurun call)
# and the parity tests against build/renamed before trusting it.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
KERNEL="${1:-BveVelTILi4ELi3680EEELi8ELi128ELi2ELi1}"
ITERS="${2:-40000}"
OUT="$ROOT/build/renamed"
rm -rf "$OUT"; mkdir -p "$OUT/tmp" "$OUT/lpm_v2_b200"
cd "$ROOT/lpm_v2_b200"
nvcc -dryrun -keep -keep-dir "$OUT/tmp" -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
     -Xcompiler -fPIC,-Wall,-Wno-unused-function -I../include -Icsrc -shared -o "$OUT/lpm_v2_b200/liblpmgpu.so" \
     csrc/lpm_gpu.cu csrc/mesh.cpp -ldl 2>&1 | sed 's/^#\$ //' > "$OUT/tmp/steps.sh"
# the leading VAR=value lines become the environment of the steps; `rm` steps are dropped (we keep everything)
{
  echo "set -e"
  grep -E '^[A-Za-z_]+=[^ ]*$' "$OUT/tmp/steps.sh" | sed 's/^/export /'      # single-token assignments only (CICC_PATH, PATH, ...)
  grep -vE '^[A-Za-z_]+=|^rm ' "$OUT/tmp/steps.sh" | while IFS= read -r line; do
    echo "$line"
    case "$line" in
      ptxas\ *) echo "python3 \"$ROOT/tools/sass_rename.py\" \"$OUT/tmp/lpm_gpu.cubin\" \"$KERNEL\" \"$OUT/tmp/lpm_gpu.renamed.cubin\" --iters $ITERS && cp \"$OUT/tmp/lpm_gpu.renamed.cubin\" \"$OUT/tmp/lpm_gpu.cubin\"" ;;
    esac
  done
} > "$OUT/tmp/run.sh"
bash "$OUT/tmp/run.sh"
cp "$ROOT"/lpm_v2_b200/*.py "$OUT/lpm_v2_b200/"
echo "built $OUT/lpm_v2_b200/liblpmgpu.so"
