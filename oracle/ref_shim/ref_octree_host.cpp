/*
 * ref_octree_host.cpp -- TEST INFRASTRUCTURE.  C entry points around the REFERENCE's own host code, compiled from where
 * it lies (src/map/Octree.cpp, include/util.hpp) with stand-ins for the three SFML headers it includes
 * (oracle/ref_shim/SFML/*): Octree::Generate / GetVoxel / Validate (src/map/Octree.cpp:13, :45, :329) and util.hpp's
 * Normalize (:64), which CLCaster::create_viewport applies to every ray of the viewport table (src/CLCaster.cpp:258).
 * Built by `make -C oracle ref` into oracle/_ref/libref_octree.so; pins oracle/vr_oracle.cpp's restatements
 * (tests/test_reference_kernel.py).  Never loaded by the product.
 */
#include <cstdint>
#include <cstring>

#include "map/Octree.h"

extern "C" {

void *ref_octree_new(void) { return new Octree(); }

void ref_octree_free(void *o) {
    Octree *t = static_cast<Octree *>(o);
    delete[] t->descriptor_buffer;
    delete[] t->attachment_lookup;
    delete[] t->attachment_buffer;
    delete t;
}

int ref_octree_buffer_size(void) { return Octree::buffer_size; }

/* Octree::Generate (writes raw_output.txt / raw_data.txt into the current directory, Octree.cpp:33-41: the caller
 * runs it in a scratch directory).  out: Octree::buffer_size descriptors. */
void ref_octree_generate(void *o, char *data, int dim, uint64_t *out, uint64_t *root_index, uint64_t *buffer_position) {
    Octree *t = static_cast<Octree *>(o);
    t->Generate(data, sf::Vector3i(dim, dim, dim));
    memcpy(out, t->descriptor_buffer, sizeof(uint64_t) * Octree::buffer_size);
    *root_index = t->root_index;
    *buffer_position = t->descriptor_buffer_position;
}

/* Octree::GetVoxel: found flag, cell origin, stack depth */
int ref_octree_get_voxel(void *o, int x, int y, int z, int *oct_pos3, int *stack_position) {
    OctState s = static_cast<Octree *>(o)->GetVoxel(sf::Vector3i(x, y, z));
    oct_pos3[0] = s.oct_pos.x; oct_pos3[1] = s.oct_pos.y; oct_pos3[2] = s.oct_pos.z;
    *stack_position = s.parent_stack_position;
    return s.found;
}

int ref_octree_validate(void *o, char *data, int dim) { return static_cast<Octree *>(o)->Validate(data, sf::Vector3i(dim, dim, dim)) ? 1 : 0; }

/* util.hpp:64 */
void ref_normalize(const float *in3, float *out3) {
    const sf::Vector3f r = Normalize(sf::Vector3f(in3[0], in3[1], in3[2]));
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
}
