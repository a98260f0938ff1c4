#!/bin/sh
# Development aid: replays the local traces on the host emulation of the device headers and compares with the
# compiled reference's outputs (build/traces/*_ref*.out).  Usage: tools/dev_emu/check.sh [c2]
set -e
cd "$(dirname "$0")/../.."
sh tools/dev_emu/build.sh >/dev/null
T=build/traces
run() { # trace ref data-dir
  build/dev_emu/emu_player $T/$1 --data-dir $3 --out /tmp/emu_$1.out >/dev/null
  echo "== $1"; python tools/compare_outputs.py $T/$2 /tmp/emu_$1.out
}
run kat.sglt kat_ref.out $T
run kat_ms.sglt kat_ms_ref.out $T
run c1.sglt c1_ref_st.out $T
if [ "$1" = "c2" ]; then run c2.sglt c2_ref_st.out build/cache; fi
