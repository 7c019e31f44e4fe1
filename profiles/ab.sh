#!/bin/bash
# A/B of kernel builds on ONE B200 box (the only comparison that means anything: boxes differ by a few percent).
#
#   profiles/ab.sh build  name1:"-DVR_JUMP_MIN=32"  name2:"-DVR_SVO_MIN_CTAS=7 -DVR_TILE_H=4" ...     (here, CPU box)
#   gpurun -- 'profiles/ab.sh run c3 main name1 name2'                                                (GPU box)
#
# `build` compiles voxel-raycaster_b200/csrc with the extra nvcc flags into build/ab/lib_<name>.so (git-ignored, but it
# travels with gpurun); `run` times bench.py per variant through VR_CASTER_LIB (caster.py) and prints
# "<config> <variant> <ms/frame of bench.py's walk (default 2)> <ms/frame of the other walks> <frame checksum>".  "main" = the in-tree library.
set -e
cd "$(dirname "$0")/.."
cmd=$1; shift
if [ "$cmd" = build ]; then
  mkdir -p build/ab
  for v in "$@"; do
    name=${v%%:*}; flags=${v#*:}
    make -C voxel-raycaster_b200/csrc OUT="$PWD/build/ab/lib_$name.so" EXTRA="$flags" > "build/ab/$name.log" 2>&1 &
  done
  wait
  for v in "$@"; do
    name=${v%%:*}
    [ -f "build/ab/lib_$name.so" ] || { echo "build of $name failed: build/ab/$name.log"; exit 1; }
    echo "$name $(cuobjdump -res-usage build/ab/lib_$name.so 2>/dev/null | grep -A1 'vr_svo_kernelILb0ELi2ELb0' | grep -o 'REG:[0-9]* STACK:[0-9]*')"
  done
elif [ "$cmd" = run ]; then
  cfg=$1; shift
  for vv in "$@"; do
    v=${vv%%@*}; extra=""; [ "$v" != "$vv" ] && extra=${vv#*@}      # name@--bench-arg=value,--other=value
    if [ "$v" = main ]; then unset VR_CASTER_LIB; else export VR_CASTER_LIB="$PWD/build/ab/lib_$v.so"; fi
    python bench.py --config "$cfg" --steps 30 --warmup 5 --no-cpu-baseline ${extra//,/ } 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$cfg $vv', round(j['ms_per_step'],4), {k: round(x, 4) for k, x in (j['config'].get('other_walk_ms_per_frame') or {}).items()}, j['config']['frame_checksum'])"
  done
else
  echo "usage: $0 build name:flags ... | run <config> <variant> ..."; exit 2
fi
