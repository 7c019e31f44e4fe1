/* vr_kernels.h -- launchers of the sm_100a kernels (internal; the public boundary is include/vr_caster.h). */
#ifndef VR_KERNELS_H
#define VR_KERNELS_H

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "vr_types.h"

/* Kernel variant selection (vr_set_option): persistent warps with warp-level pixel refill. */
typedef struct vr_launch_options {
    int persistent;            /* 0 = one thread per pixel over a 2-D grid, 1 = persistent warps */
    int refill_min;            /* idle lanes needed before a warp fetches new pixels */
    int ctas_per_sm;
    int num_sms;
    unsigned int *counter;     /* device pixel counter */
    int walk;                  /* 0 = merged in-cell walk (bit-identical to the reference on every pixel),
                                  1 = per-axis walk (identical except distance_traveled on exact-tie rays),
                                  2 = closed-form crossing times (vr_canon.h: within BASELINE.json's tolerance) */
} vr_launch_options;

/* One frame (or one row-band slab of it).  use_svo selects the 64-tree traversal kernel, otherwise the
 * dense DDA kernel.  with_aux additionally writes one vr_aux record per pixel.  *launches is bumped
 * once per kernel launched. */
cudaError_t vr_launch_raycast(const vr_frame_params &P, int use_svo, int with_aux, cudaStream_t stream,
                              unsigned long long *launches, const vr_launch_options *opt);

cudaError_t vr_launch_fill(uint32_t *dst, size_t n, uint32_t value, cudaStream_t stream,
                           unsigned long long *launches);

#endif
