#!/bin/bash
# TEST INFRASTRUCTURE ONLY: data-race check of the kernels' shared-memory protocols.
# Builds the emulated library with -fsanitize=thread and runs race_check.py on it (1 and 2 emulated devices).
# Prints the number of ThreadSanitizer reports; the full log goes to $1 (default /tmp/lpm_race_check.log).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
LOG="${1:-/tmp/lpm_race_check.log}"
bash "$HERE/build_emu.sh" > /dev/null
g++ -std=c++20 -O1 -g -fPIC -shared -pthread -fsanitize=thread -Wl,-Bsymbolic -DLPM_CUDA_EMU=1 \
    -I"$HERE" -I"$HERE/_gen" -I"$ROOT/include" -o "$HERE/liblpmgpu_emu_tsan.so" "$HERE/_gen/lpm_gpu.cpp" -ldl
TSAN="$(g++ -print-file-name=libtsan.so)"
: > "$LOG"
for nd in 1 2; do
  LD_PRELOAD="$TSAN" TSAN_OPTIONS="report_signal_unsafe=0 exitcode=0 history_size=4" LPM_EMU_DEVICES=$nd \
    LPM_GPU_LIBRARY="$HERE/liblpmgpu_emu_tsan.so" python3 "$HERE/race_check.py" >> "$LOG" 2>&1
done
echo "workloads completed: $(grep -c 'race_check workload done' "$LOG")"
echo "ThreadSanitizer reports: $(grep -c 'WARNING: ThreadSanitizer' "$LOG")"
