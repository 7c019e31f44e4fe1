/*
 * vr_oracle.h -- CPU restatement of the reference ray caster.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity oracle for the B200 port of MitchellHansen/voxel-raycaster's per-pixel
 * ray casting hot path.  It restates, function by function, `kernels/ray_caster_kernel.cl`
 * (the OpenCL kernel) and the host-side input producers `CLCaster::create_viewport`
 * (src/CLCaster.cpp:233-299) and `Octree::GenerationRecursion` (src/map/Octree.cpp:171-323).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Nothing under voxel-raycaster_b200/ links, imports or executes it.
 *
 * PARITY PINNED AGAINST THE REFERENCE'S OWN CODE, EXECUTED.  The reference ships no tests, golden images or
 * known-answer vectors for this path and no OpenCL runtime, SFML or GL exists in the image,
 * but its sources can be compiled for the CPU from where they lie (`make -C oracle ref` -> oracle/_ref/ (the .so files)):
 *   - kernels/ray_caster_kernel.cl by g++ through ref_shim/cl_shim.h (OpenCL C vector types, operators and built-ins
 *     in C++; the only edit, by sed on the fly, is the vector-literal syntax `(typeN)(` -> `typeN(`), verbatim, with
 *     kernel:326's max_distance read from a variable, and with the three 8-entry private stacks widened to 32;
 *   - src/map/Octree.cpp + include/util.hpp with stand-ins for the three SFML headers they include.
 * tests/test_reference_kernel.py requires this restatement to reproduce that code bit for bit: every RGBA8 pixel and
 * the written/skipped mask on 16 scene/camera combinations (HEAD, shadows, reflections, biased camera, 64^3..512^3
 * terrain, transparent values, the 1024^3 bench frame on sampled rows) and 80 random scenes, all 100 000 entries of Octree::Generate's buffer and its root index, util.hpp's
 * Normalize.  Not pinned by the reference: the OpenCL built-ins below (an OpenCL runtime supplies them; cl_shim.h and
 * this file define them identically) and CLCaster::create_viewport's loop (OpenCL/GL translation unit: restated, its
 * Normalize call pinned).  Further pins: the derived known answers in tests/test_oracle.py and tests/golden/.
 *
 * OpenCL built-ins whose results are implementation-defined in the reference (it is built with
 * -cl-fast-relaxed-math, src/CLCaster.cpp:771) are pinned here to IEEE-754 binary32:
 *   normalize(v) = v / sqrtf(dot(v,v));  fast_length = sqrtf(dot);  dot = (x*x' + y*y') + z*z';
 *   pow(x, 1.0f) = x;  mix(a,b,t) = a + (b-a)*t;  max(x,y) = (x < y) ? y : x;
 *   convert_int(float) truncates toward zero;  write_imagef = rint(clamp(c*255, 0, 255)) (RTE);
 *   sin/cos of the camera angles = host sinf/cosf, passed in precomputed (they are frame-uniform).
 * No FMA contraction: compile with -ffp-contract=off -fno-fast-math.
 */
#ifndef VR_ORACLE_H
#define VR_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Everything the kernel reads (kernel:256-273), in host memory. */
typedef struct vro_scene {
    /* viewport */
    int32_t width, height;              /* viewport_resolution (kernel arg 2)                    */
    const float *ray_table;             /* viewport_matrix, float4 stride, W*H entries (arg 3)   */
    /* dense map (args 0,1) */
    const int8_t *map;                  /* char map[x + X*(y + Z*z)]                              */
    int32_t map_dim[3];
    /* camera (args 4,5) */
    float cam_dir[2];                   /* (inclination, azimuth)                                 */
    float cam_pos[3];
    float trig[4];                      /* sin(dir.x), cos(dir.x), sin(dir.y), cos(dir.y)         */
    /* lights (args 6,7): 10 floats each {r,g,b,i,x,y,z,dx,dy,dz}; only light 0 is read           */
    const float *lights;
    int32_t light_count;
    /* atlas (args 9,10,11): RGBA8 */
    const uint8_t *atlas;
    int32_t atlas_dim[2];
    int32_t tile_dim[2];
    /* octree (args 12-14) -- may be NULL: then get_oct_vox's start bias is 0 */
    const uint64_t *oct_desc;
    uint64_t oct_desc_len;
    /* settings (arg 15) */
    int64_t octdim;                     /* setting(OCTDIM)                                        */
    int64_t oct_root_index;             /* setting(OCTREE_ROOT_INDEX)                             */
    /* kernel:326 -- HEAD hard-codes 20; lifted into a parameter                                  */
    int32_t max_distance;
    /* extension beyond the reference: number of shadow lights, 1 = reference      */
    int32_t shadow_lights;
    /* widened map access for volumes too large to materialise (4096^3 = 64 GiB): when `map` is
     * NULL, voxel (x,y,z) holds 5 iff col_lo[x + X*y] <= z <= col_hi[x + X*y], else 0.  Same loop, same arithmetic;
     * only the `map[...]` load of kernel:569 is answered by the column table. */
    const int32_t *col_lo;
    const int32_t *col_hi;
    /* 0 = the reference (kernel:559: intersection_t accumulates one rounded addition per step).  1 = "Oracle-B" of the
     * parity chain: crossing times in closed form, t(k) = fma(k, delta_t, t0) -- the arithmetic of
     * the octree kernel's walk = 2.  Not part of the reference; see vr_oracle.cpp. */
    int32_t canonical_t;
    /* kernel:357 `bounce_count < 2`: lifted into a parameter like max_distance (0 = the reference's 2).  The reflection
     * limit as a setting is on the reference's own TODO list (src/main.cpp:31-33). */
    int32_t max_bounces;
    /* Counters only (the D_svo byte model of maps too large for a reference-format buffer: 4096^3 has no dense map to run
     * Octree::Generate on).  An occupancy tree with 4x4x4 children per node = two levels of the reference's 2^3 octree
     * per node: tree64[4 * i + 0..1] = 64-bit child mask (bit cx | cy << 2 | cz << 4), tree64[4 * i + 2] = index of the
     * first child node (children contiguous, ascending set-bit order), root at 0, `tree64_levels` levels, dimension
     * covered 4^levels == octdim.  When oct_desc is NULL and this is given, svo_lookup derives the path of 2^3 descriptors
     * from it (an octant of 2x2x2 slots that is empty = an empty child one level up).  Ignored by the ray cast itself. */
    const uint32_t *tree64;
    int32_t tree64_levels;
} vro_scene;

/* Per-pixel auxiliary record, 32 bytes.  The reference kernel only writes RGBA8; these expose
 * the integer state the parity tests compare bit-exactly. */
typedef struct vro_aux {
    int32_t hit[3];          /* first voxel hit by the primary ray (type 5 or 6); -1,-1,-1 if none */
    uint8_t face;            /* bits0-2 face_mask x/y/z at first hit; bits 3-5 voxel_step<0 per axis */
    uint8_t status;          /* VRO_ST_* terminal event                                             */
    uint8_t flags;           /* VRO_FL_*                                                            */
    uint8_t hit_type;        /* voxel value at first hit (5 / 6), 0 if none                         */
    uint32_t steps_first;    /* distance_traveled when the first hit happened                       */
    uint32_t steps_total;    /* distance_traveled at termination                                    */
    uint32_t pad[2];
} vro_aux;

enum {
    VRO_ST_SKIP_PRIMARY = 0,   /* kernel:293  zero component in the rotated ray: pixel not written   */
    VRO_ST_OOB          = 1,   /* kernel:563  ray left the map                                       */
    VRO_ST_MAXDIST      = 2,   /* kernel:357  distance_traveled reached max_distance                 */
    VRO_ST_SHADOW_HIT   = 3,   /* kernel:706  shadow ray (or any ray while shadow_ray) hit 5/6       */
    VRO_ST_SKIP_REDIRECT= 4,   /* kernel:671/694 zero component after redirect: pixel not written    */
    VRO_ST_BOUNCES      = 5    /* kernel:357  bounce_count reached 2                                 */
};
enum {
    VRO_FL_LIT       = 1,      /* a type-5 primary hit happened (shadow ray spawned)                  */
    VRO_FL_REFLECTED = 2,      /* at least one type-6 reflection                                     */
    VRO_FL_TIE       = 4,      /* some DDA step moved along more than one axis (exact t tie)         */
    VRO_FL_ATLAS_CLAMP = 8,    /* atlas texel coordinate fell outside the image and was clamped      */
    VRO_FL_FRAC0     = 16,     /* a camera position component is an exact integer                    */
    VRO_FL_NEAR      = 32,     /* (canonical_t == 0 only) some step would be taken along other axes under closed-form
                                * crossing times -- two crossing times closer than the accumulated rounding: the ray
                                * passes a voxel edge within float noise -- AND one of the two voxels entered is set or
                                * outside the map, so the difference can show (BASELINE.json: "degenerate" ray)       */
    VRO_FL_NEAR_AIR  = 64      /* same, between empty voxels: the two paths meet again one step later (statistics)   */
};

/* Whole-frame counters (SURVEY 8d byte model). */
typedef struct vro_counters {
    uint64_t pixels;
    uint64_t pixels_written;
    uint64_t primary_rays;     /* pixels that entered the loop                                       */
    uint64_t shadow_rays;      /* shadow redirects                                                   */
    uint64_t reflect_rays;
    uint64_t dda_steps;        /* S_dense: loop iterations that loaded map[]                         */
    uint64_t texel_fetches;    /* T                                                                  */
    uint64_t svo_desc_fetches; /* D_svo: canonical descent fetches (only if counted)                 */
    uint64_t svo_cell_changes;
    uint64_t tie_pixels;
} vro_counters;

/* create_viewport's ray table (src/CLCaster.cpp:233-299).  out: W*H float4. */
void vro_make_ray_table(int width, int height, float *out);

/* raycaster kernel, dense branch (kernel:256-724 with setting(OCTENABLED) != 0), rows y0, y0+row_stride,
 * ... < y1 (row_stride > 1 is the bounded sample used for CPU-baseline timing and byte counting).
 * rgba must be pre-filled by the caller (skipped pixels keep their contents, kernel:293).
 * aux / counters may be NULL.  count_svo != 0 additionally walks the reference-format octree to
 * count canonical descriptor fetches (needs oct_desc).  Returns 0 on success. */
int vro_raycast(const vro_scene *s, int y0, int y1, int row_stride, uint8_t *rgba, vro_aux *aux,
                vro_counters *counters, int count_svo, int num_threads);

/* get_oct_vox (kernel:140-251) at one position: writes found, sub_oct_pos[3], resolution, scale. */
void vro_get_oct_vox(const uint64_t *desc, int64_t root_index, int64_t octdim, const int32_t pos[3],
                     int32_t *found, int32_t sub_oct_pos[3], int32_t *resolution, int32_t *scale);

/* Octree::Generate + GenerationRecursion (src/map/Octree.cpp:13-43,171-323) into a caller buffer of
 * buffer_size entries (reference: 100000), written from the top down.  Returns the root index, or
 * -1 if the buffer is too small (the reference has no such check).  *used = descriptors written. */
int64_t vro_octree_generate(const int8_t *data, int dim, uint64_t *buffer, uint64_t buffer_size,
                            uint64_t *used);

int vro_num_procs(void);

#ifdef __cplusplus
}
#endif
#endif
