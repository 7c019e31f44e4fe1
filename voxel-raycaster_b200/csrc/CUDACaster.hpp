/*
 * CUDACaster.hpp -- C++ facade over the C ABI (include/vr_caster.h) with the reference's class shape.
 *
 * A maintainer of MitchellHansen/voxel-raycaster swaps `CLCaster` (ref include/CLCaster.h:93-329) for this
 * class: same method names, same call order (ref src/Application.cpp:27-88,151-159), same
 * bool-returning convention.  SFML/GL types are replaced by raw pointers so the header has no dependency
 * beyond the C ABI; INTEGRATION.md shows the three-line adapters for sf::Texture / Camera / Map.
 * Header-only; link with libvrcaster.so.
 */
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/vr_caster.h"

class CUDACaster {
public:
    CUDACaster() = default;
    ~CUDACaster() { vr_destroy(ctx_); }
    CUDACaster(const CUDACaster &) = delete;
    CUDACaster &operator=(const CUDACaster &) = delete;

    /* ref CLCaster::init (include/CLCaster.h:110) */
    bool init(int device = -1) { return vr_init(&ctx_, device, VR_INIT_HEADLESS) != 0; }

    /* ref :114-115 */
    bool create_viewport(int width, int height, float v_fov, float h_fov) {
        width_ = width;
        height_ = height;
        return vr_create_viewport(ctx_, width, height, v_fov, h_fov) != 0;
    }
    bool release_viewport() { return vr_release_viewport(ctx_) != 0; }

    /* ref :119 -- `packed` is LightController's std::vector<PackedData>::data(): 10 floats per light, aliased */
    bool assign_lights(const float *packed, int count) { return vr_assign_lights(ctx_, packed, count) != 0; }

    /* ref :123-124 -- ArrayMap::getDataPtr() / getDimensions() */
    bool assign_map(const char *voxels, int nx, int ny, int nz) {
        return vr_assign_map(ctx_, reinterpret_cast<const int8_t *>(voxels), nx, ny, nz) != 0;
    }
    bool release_map() { return vr_release_map(ctx_) != 0; }

    /* ref :127-128 -- Octree::descriptor_buffer / buffer_size / root_index */
    bool assign_octree(const uint64_t *descriptor_buffer, const uint32_t *attachment_lookup, const uint64_t *attachment_buffer,
                       uint64_t buffer_size, uint64_t root_index) {
        return vr_assign_octree(ctx_, descriptor_buffer, attachment_lookup, attachment_buffer, buffer_size, root_index) != 0;
    }
    bool release_octree() { return vr_release_octree(ctx_) != 0; }

    /* ref :131-132 -- Camera::get_direction_pointer() / get_position_pointer(), aliased */
    bool assign_camera(const float *direction, const float *position) { return vr_assign_camera(ctx_, direction, position) != 0; }
    bool release_camera() { return vr_release_camera(ctx_) != 0; }

    /* ref :136 -- sf::Image::getPixelsPtr() of the sprite sheet + tile size */
    bool create_texture_atlas(const uint8_t *rgba, int width, int height, int tile_w, int tile_h) {
        return vr_create_texture_atlas(ctx_, rgba, width, height, tile_w, tile_h) != 0;
    }

    /* ref :139, :142, :169 */
    bool validate() { return vr_validate(ctx_) != 0; }
    bool compute() { return vr_compute(ctx_) != 0; }
    bool debug_quick_recompile() { return vr_debug_quick_recompile(ctx_) != 0; }

    /* ref :145 draw(sf::RenderWindow*) -- headless: hands the frame to the caller (RGBA8, width*height*4) */
    bool draw(std::vector<uint8_t> &frame) {
        frame.resize(static_cast<size_t>(width_) * height_ * 4);
        return vr_read_framebuffer(ctx_, frame.data(), frame.size()) != 0;
    }

    /* draw(sf::RenderWindow*) with CL/GL sharing (src/CLCaster.cpp:330-332, :840-842): register the sprite's texture once
     * (sf::Texture::getNativeHandle()), then draw_gl() after every compute() copies the frame into it on the device */
    bool register_gl_texture(unsigned texture, unsigned target = 0x0DE1 /* GL_TEXTURE_2D */) { return vr_gl_register_texture(ctx_, texture, target) != 0; }
    bool draw_gl() { return vr_gl_draw(ctx_) != 0; }

    /* ref :148-151 */
    bool load_config() { return vr_load_config(ctx_, nullptr) != 0; }
    void save_config() { vr_save_config(ctx_, nullptr); }

    /* ref :154-155 */
    void set_define(const std::string &name, const std::string &value) { vr_set_define(ctx_, name.c_str(), value.c_str()); }
    void remove_define(const std::string &name) { vr_remove_define(ctx_, name.c_str()); }

    /* ref :157-164 */
    bool create_settings_buffer() { return vr_create_settings_buffer(ctx_) != 0; }
    bool release_settings_buffer() { return vr_release_settings_buffer(ctx_) != 0; }
    template <typename T>
    bool add_to_settings_buffer(const std::string &setting_name, const std::string &define_accessor_name, T value) {
        return vr_add_to_settings_buffer(ctx_, setting_name.c_str(), define_accessor_name.c_str(), static_cast<int64_t>(value)) != 0;
    }
    bool overwrite_setting(const std::string &settings_name, int64_t *value) {
        return vr_overwrite_setting(ctx_, settings_name.c_str(), value) != 0;
    }
    bool remove_from_settings_buffer(const std::string &setting_name) {
        return vr_remove_from_settings_buffer(ctx_, setting_name.c_str()) != 0;
    }

    /* ---- beyond the reference: one CUDACaster per process and GPU, frames split over the GPUs of the node
     * (include/vr_caster.h: vr_mgpu_*).  Where Application::game_loop calls compute() + draw() (ref src/Application.cpp:
     * 151-159), a multi-GPU loop calls frame() on every rank and frame_wait() / frame_release() on rank 0. */
    bool set_option(const std::string &name, int64_t value) { return vr_set_option(ctx_, name.c_str(), value) != 0; }
    bool mgpu_init(const std::string &session, int world, int rank, bool host_frame = false) {
        return vr_mgpu_init(ctx_, session.c_str(), world, rank, host_frame ? VR_MGPU_HOST_FRAME : 0u) != 0;
    }
    bool mgpu_broadcast_octree() { return vr_mgpu_broadcast_octree(ctx_) != 0; }
    bool mgpu_frame(uint64_t *frame_no) { return vr_mgpu_frame(ctx_, frame_no) != 0; }
    bool mgpu_frame_wait(uint64_t frame_no, const uint8_t **rgba) { return vr_mgpu_frame_wait(ctx_, frame_no, rgba) != 0; }
    bool mgpu_frame_release(uint64_t frame_no) { return vr_mgpu_frame_release(ctx_, frame_no) != 0; }
    bool mgpu_barrier() { return vr_mgpu_barrier(ctx_) != 0; }
    bool mgpu_shutdown() { return vr_mgpu_shutdown(ctx_) != 0; }

    const char *last_error() const { return vr_last_error(ctx_); }
    vr_ctx *handle() { return ctx_; }

private:
    vr_ctx *ctx_ = nullptr;
    int width_ = 0, height_ = 0;
};
