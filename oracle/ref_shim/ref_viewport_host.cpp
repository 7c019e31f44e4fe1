/*
 * ref_viewport_host.cpp -- TEST INFRASTRUCTURE.  The ray-table loop of the REFERENCE's CLCaster::create_viewport
 * (src/CLCaster.cpp:244-275), compiled from where it lies: `make -C oracle ref` cuts the statements between
 * "viewport_matrix = new sf::Vector4f[...]" and the create_buffer("viewport_matrix") call out of src/CLCaster.cpp with sed
 * into a temporary file (VR_REF_VIEWPORT, outside the repo) which is included below as the body of a function that only
 * supplies the names the loop uses (width, height, view_res, viewport_matrix).  The rest of CLCaster.cpp (OpenCL, OpenGL,
 * SFML window code) cannot be compiled here and is not needed for the table.  Normalize comes from the reference's
 * include/util.hpp, sf::Vector4f from its include/Vector4.hpp, sf::Vector2/3 from the stand-ins under ref_shim/SFML.
 * Built into oracle/_ref/libref_viewport.so; pins oracle/vr_oracle.cpp: vro_make_ray_table and the library's
 * vr_create_viewport (tests/test_reference_kernel.py).  Never loaded by the product.
 */
#include <cmath>
#include <cstdint>
#include <cstring>

#include "util.hpp"

extern "C" void ref_create_viewport_table(int width, int height, float *out4) {
    sf::Vector2i view_res(width, height);
    sf::Vector4f *viewport_matrix = nullptr;
#include VR_REF_VIEWPORT
    for (long i = 0; i < (long)width * height; i++) {
        out4[4 * i + 0] = viewport_matrix[i].x;
        out4[4 * i + 1] = viewport_matrix[i].y;
        out4[4 * i + 2] = viewport_matrix[i].z;
        out4[4 * i + 3] = viewport_matrix[i].w;
    }
    delete[] viewport_matrix;
}
