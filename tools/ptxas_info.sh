#!/bin/bash
# Registers / spills / shared memory of every ds_kernel instantiation (ptxas -v), demangled.
# usage: tools/ptxas_info.sh [grep-pattern]
cd "$(dirname "$0")/../lpm_v2_b200" || exit 1
make ptxas-info 2>&1 | c++filt | awk '
/Compiling entry function/ { name=$0; sub(/.*Compiling entry function ./,"",name); sub(/. for .sm_100a.*/,"",name); sub(/\(.*/,"",name) }
/bytes stack frame/ { spill=$0; sub(/^ +/,"",spill) }
/Used [0-9]+ registers/ { u=$0; sub(/.*Used /,"",u); print name " | " u " | " spill }' | grep -E "${1:-.}"
