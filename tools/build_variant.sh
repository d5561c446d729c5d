#!/bin/sh
# Build the C-ABI library from another git revision (or the working tree with extra -D flags) for
# same-box A/B timing:  tools/build_variant.sh <git-rev|WORK> <out.so> [extra nvcc flags...]
# then  python tools/move_breakdown.py --lib <out.so>
set -e
rev=$1; out=$(realpath -m "$2"); shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
mkdir -p "$tmp/chromo_b200/csrc" "$tmp/include" "$(dirname "$out")"
if [ "$rev" = WORK ]; then
  cp "$root"/chromo_b200/csrc/*.cu* "$tmp/chromo_b200/csrc/"; cp "$root/include/chromo_b200.h" "$tmp/include/"
else
  for f in $(git -C "$root" ls-tree --name-only "$rev" chromo_b200/csrc/ | grep '\.cu'); do git -C "$root" show "$rev:$f" > "$tmp/$f"; done
  git -C "$root" show "$rev:include/chromo_b200.h" > "$tmp/include/chromo_b200.h"
fi
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC $*"
cd "$tmp/chromo_b200/csrc"
nvcc $F -c chromo_b200.cu -o "$tmp/api.o" &
if [ -f rediscretize.cu ]; then nvcc $F -c rediscretize.cu -o "$tmp/rd.o" & else echo 'int cb_rd_absent;' > "$tmp/rd.cu"; nvcc $F -c "$tmp/rd.cu" -o "$tmp/rd.o" & fi
nvcc $F -DCB_INST_REPLAY=1 -DCB_INST_HI=0 -c mc_inst.cu -o "$tmp/r12.o" &
nvcc $F -DCB_INST_REPLAY=1 -DCB_INST_HI=1 -c mc_inst.cu -o "$tmp/r34.o" &
nvcc $F -DCB_INST_REPLAY=0 -DCB_INST_HI=0 -c mc_inst.cu -o "$tmp/p12.o" &
nvcc $F -DCB_INST_REPLAY=0 -DCB_INST_HI=1 -c mc_inst.cu -o "$tmp/p34.o" &
nvcc $F -DCB_INST_REPLAY=1 -DCB_INST_HI=0 -DCB_TWIST=1 -c mc_inst.cu -o "$tmp/rt12.o" &
nvcc $F -DCB_INST_REPLAY=0 -DCB_INST_HI=0 -DCB_TWIST=1 -c mc_inst.cu -o "$tmp/pt12.o" &
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$out" "$tmp/api.o" "$tmp/rd.o" "$tmp/r12.o" "$tmp/r34.o" "$tmp/p12.o" "$tmp/p34.o" "$tmp/rt12.o" "$tmp/pt12.o"
rm -rf "$tmp"
echo "built $out from $rev"
