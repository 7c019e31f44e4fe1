#!/bin/bash
# Dumps the SASS of one kernel of vr_kernels.cu (no GPU needed): profiles/sass_kernel.sh 'vr_svo_kernelILb0ELi2ELb0' [extra nvcc flags]
# Prints the instruction count, the opcode histogram and writes the listing to /tmp/sass/<pattern>.sass
set -e
pat="$1"; shift
cd "$(dirname "$0")/../voxel-raycaster_b200/csrc"
mkdir -p /tmp/sass
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false "$@" -cubin -o /tmp/sass/vr_kernels.cubin vr_kernels.cu
fn=$(cuobjdump -sass /tmp/sass/vr_kernels.cubin | grep "Function :" | grep "$pat" | head -1 | awk '{print $3}')
cuobjdump -sass -fun "$fn" /tmp/sass/vr_kernels.cubin | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's@^\s+/\*([0-9a-f]{4})\*/\s+@\1 @; s@\s*/\* 0x[0-9a-f]+ \*/@@; s@\s+;\s*$@@' > "/tmp/sass/$pat.sass"
echo "$fn: $(wc -l < /tmp/sass/$pat.sass) instructions"
awk '{op=$2; if (op ~ /^@/) op=$3; sub(/\..*/, "", op); print op}' "/tmp/sass/$pat.sass" | sort | uniq -c | sort -rn | head -25 | tr '\n' ' '; echo
