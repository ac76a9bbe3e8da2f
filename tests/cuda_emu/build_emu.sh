#!/bin/bash
# TEST INFRASTRUCTURE ONLY: builds tests/cuda_emu/liblpmgpu_emu.so = lpm_v2_b200/csrc compiled with g++ against
# the SIMT emulator (the whole C ABI on CPU threads), and libcuda_emu.so (the kernel-level harness).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
# up to date?  (both libraries newer than every source they are made from)
newest_src=$(ls -t "$ROOT"/lpm_v2_b200/csrc/* "$ROOT"/include/*.h "$HERE"/*.h "$HERE"/*.cpp "$HERE"/*.py "$HERE"/*.sh "$HERE"/cub/device/*.cuh | head -1)
if [ -f "$HERE/liblpmgpu_emu.so" ] && [ -f "$HERE/libcuda_emu.so" ] && [ "$HERE/liblpmgpu_emu.so" -nt "$newest_src" ] && [ "$HERE/libcuda_emu.so" -nt "$newest_src" ]; then
  echo "up to date: $HERE/liblpmgpu_emu.so $HERE/libcuda_emu.so"; exit 0
fi
python3 "$HERE/convert.py" "$ROOT/lpm_v2_b200/csrc" "$HERE/_gen" > "$HERE/_gen.log"
FLAGS="-std=c++20 -O1 -fPIC -shared -pthread -fvisibility=hidden -Wl,-Bsymbolic -DLPM_CUDA_EMU=1 -I$HERE -I$HERE/_gen -I$ROOT/include"
g++ $FLAGS -o "$HERE/libcuda_emu.so" "$HERE/emu_lib.cpp" &
g++ $FLAGS -fvisibility=default -o "$HERE/liblpmgpu_emu.so" "$HERE/_gen/lpm_gpu.cpp" "$HERE/_gen/mesh.cpp" -ldl
wait
echo "built $HERE/liblpmgpu_emu.so $HERE/libcuda_emu.so"
