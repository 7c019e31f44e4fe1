"""The C++ facade (voxel-raycaster_b200/csrc/CUDACaster.hpp) must compile against the C ABI header and link
against libvrcaster.so with the method set of the reference's CLCaster (include/CLCaster.h:110-164)."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

SRC = r'''
#include "CUDACaster.hpp"
#include <cstdio>
#include <cstring>
#include <unistd.h>
int main() {
    CUDACaster c;
    bool ok = c.init(0);                       // false without a GPU: must not crash, must report why
    std::printf("init=%d err=%s\n", (int)ok, ok ? "" : "no device");
    if (!ok) return 0;
    float dir[2] = {2.0f, 1.0f}, pos[3] = {2.5f, 2.5f, 2.5f}, light[10] = {1,1,1,1, 5,5,5, 0,0,-1};
    static char vox[16 * 16 * 16];
    for (int i = 0; i < 16 * 16 * 4; i++) vox[i] = 5;
    static unsigned char atlas[256 * 256 * 4];
    int64_t md = 48;
    ok = c.add_to_settings_buffer("octree_dimensions", "OCTDIM", 16) && c.add_to_settings_buffer("using_octree", "OCTENABLED", 0)
      && c.add_to_settings_buffer("max_distance", "MAX_DISTANCE", 20) && c.overwrite_setting("max_distance", &md)
      && c.assign_map(vox, 16, 16, 16) && c.assign_camera(dir, pos) && c.create_viewport(64, 48, 56.25f, 90.f)
      && c.assign_lights(light, 1) && c.create_texture_atlas(atlas, 256, 256, 16, 16) && c.validate() && c.compute();
    std::vector<uint8_t> frame;
    ok = ok && c.draw(frame) && frame.size() == 64u * 48u * 4u;
    std::printf("frame=%d\n", (int)ok);
    if (!ok) return 1;
    // the multi-GPU frame loop of the library with a world of one: same frame, in shared host memory
    char session[64];
    std::snprintf(session, sizeof(session), "facade_%d", (int)getpid());
    ok = c.mgpu_init(session, 1, 0, true) && c.mgpu_broadcast_octree() && c.mgpu_barrier();
    for (int i = 0; ok && i < 6; i++) {
        uint64_t k = 0;
        const uint8_t *rgba = nullptr;
        ok = c.mgpu_frame(&k) && k == (uint64_t)i && c.mgpu_frame_wait(k, &rgba) && rgba != nullptr
          && std::memcmp(rgba, frame.data(), frame.size()) == 0 && c.mgpu_frame_release(k);
    }
    ok = ok && c.mgpu_shutdown();
    std::printf("mgpu=%d %s\n", (int)ok, ok ? "" : c.last_error());
    return ok ? 0 : 1;
}
'''


def test_facade_compiles_and_links(pkg, tmp_path):
    src = tmp_path / "facade.cpp"
    src.write_text(SRC)
    exe = tmp_path / "facade"
    lib_dir = ROOT / "voxel-raycaster_b200"
    r = subprocess.run(["/usr/bin/g++", "-std=c++14", "-Wall", "-I", str(lib_dir / "csrc"), str(src), "-o", str(exe),
                        "-L", str(lib_dir), "-lvrcaster", f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "init=" in r.stdout


import pytest  # noqa: E402


@pytest.mark.gpu
def test_facade_renders_and_schedules_on_gpu(pkg, tmp_path):
    """The same C++ program on a GPU: one frame through CUDACaster (compute + draw) and six frames through the library's
    multi-GPU frame loop (vr_mgpu_*, world of one, frame in shared host memory) that must equal it byte for byte."""
    src = tmp_path / "facade.cpp"
    src.write_text(SRC)
    exe = tmp_path / "facade"
    lib_dir = ROOT / "voxel-raycaster_b200"
    r = subprocess.run(["/usr/bin/g++", "-std=c++14", "-Wall", "-I", str(lib_dir / "csrc"), str(src), "-o", str(exe),
                        "-L", str(lib_dir), "-lvrcaster", f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "init=1" in r.stdout and "frame=1" in r.stdout and "mgpu=1" in r.stdout, r.stdout + r.stderr
