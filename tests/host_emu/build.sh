#!/bin/sh
# Builds the CPU-emulated twin of libchromo_b200.so for GPU-less debugging of the
# kernel logic (see cuda_emu.h).  Output: tests/host_emu/libchromo_emu.so
set -e
here=$(cd "$(dirname "$0")" && pwd)
root=$(cd "$here/../.." && pwd)
src="$root/chromo_b200/csrc"
out="$here/_build"
mkdir -p "$out"
FLAGS="-x c++ -std=c++17 -O1 -g -ffp-contract=off -fPIC -pthread -DCHROMO_HOST_EMU -D_GNU_SOURCE -Wno-unknown-pragmas -I$here"
g++ $FLAGS -c "$src/chromo_b200.cu" -o "$out/api.o" &
g++ $FLAGS -c "$src/rediscretize.cu" -o "$out/rd.o" &
g++ $FLAGS -DCB_INST_REPLAY=1 -DCB_INST_HI=0 -c "$src/mc_inst.cu" -o "$out/r12.o" &
g++ $FLAGS -DCB_INST_REPLAY=1 -DCB_INST_HI=1 -c "$src/mc_inst.cu" -o "$out/r34.o" &
g++ $FLAGS -DCB_INST_REPLAY=0 -DCB_INST_HI=0 -c "$src/mc_inst.cu" -o "$out/p12.o" &
g++ $FLAGS -DCB_INST_REPLAY=0 -DCB_INST_HI=1 -c "$src/mc_inst.cu" -o "$out/p34.o" &
g++ $FLAGS -DCB_INST_REPLAY=1 -DCB_INST_HI=0 -DCB_TWIST=1 -c "$src/mc_inst.cu" -o "$out/rt12.o" &
g++ $FLAGS -DCB_INST_REPLAY=0 -DCB_INST_HI=0 -DCB_TWIST=1 -c "$src/mc_inst.cu" -o "$out/pt12.o" &
wait
g++ -shared -pthread -o "$here/libchromo_emu.so" "$out/api.o" "$out/rd.o" "$out/r12.o" "$out/r34.o" "$out/p12.o" "$out/p34.o" "$out/rt12.o" "$out/pt12.o"
