#!/bin/bash
# TEST INFRASTRUCTURE ONLY: builds tests/cuda_emu/liblpmgpu_emu.so = lpm_v2_b200/csrc compiled with g++ against
# the SIMT emulator (the whole C ABI on CPU threads) and
# fake_nccl/libnccl.so.2 (the nine NCCL calls over shared memory, for emulated rank-mode runs).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
# up to date?  (both libraries newer than every source they are made from)
newest_src=$(ls -t "$ROOT"/lpm_v2_b200/csrc/* "$ROOT"/include/*.h "$HERE"/*.h "$HERE"/convert.py "$HERE"/build_emu.sh "$HERE"/cub/device/*.cuh "$HERE"/fake_nccl/*.cpp | head -1)
if [ -f "$HERE/liblpmgpu_emu.so" ] && [ "$HERE/fake_nccl/libnccl.so.2" -nt "$newest_src" ] && [ "$HERE/liblpmgpu_emu.so" -nt "$newest_src" ]; then
  echo "up to date: $HERE/liblpmgpu_emu.so"; exit 0
fi
python3 "$HERE/convert.py" "$ROOT/lpm_v2_b200/csrc" "$HERE/_gen" > "$HERE/_gen.log"
FLAGS="-std=c++20 -O1 -fPIC -shared -pthread -fvisibility=hidden -Wl,-Bsymbolic -DLPM_CUDA_EMU=1 -I$HERE -I$HERE/_gen -I$ROOT/include"
g++ -std=c++17 -O1 -fPIC -shared -pthread -fvisibility=hidden -Wno-format-truncation -o "$HERE/fake_nccl/libnccl.so.2" "$HERE/fake_nccl/fake_nccl.cpp" &
g++ $FLAGS -fvisibility=default -o "$HERE/liblpmgpu_emu.so" "$HERE/_gen/lpm_gpu.cpp" -ldl
wait
echo "built $HERE/liblpmgpu_emu.so"
