/*
 * vr_mgpu.cu -- multi-GPU frame scheduler behind the C ABI (include/vr_caster.h: vr_mgpu_*).
 *
 * What CLCaster::run_kernel (reference src/CLCaster.cpp:946-987: acquire, NDRange, clFinish, release) is to one OpenCL
 * device, this is to the GPUs of one node: ONE process per GPU, each with its own vr_ctx; a frame is split into the
 * 32x4-pixel CTA tiles of the ray kernel, tile (tx, ty) belongs to rank (tx + ty) mod world.
 *
 *   bootstrap   a POSIX shared-memory segment named after the caller's session string carries everything the ranks
 *               exchange on the host: the NCCL unique id, the CUDA-IPC handles of the root's frame buffers, the per-rank
 *               frame counters.  No other channel (MPI, sockets, torch.distributed) is needed.
 *   scene       vr_mgpu_broadcast_octree: rank 0's 64-tree goes to every rank with ncclBroadcast over NVLink ("the
 *               octree is broadcast once"); camera / lights / settings stay per-rank host pointers as in vr_compute.
 *   frame       device frames (default): every rank's ray kernel stores its tiles IN PLACE into the frame that lives on
 *               the root GPU (mapped through CUDA IPC: the stores travel over NVLink, there is no slab and no gather
 *               kernel); host frames (VR_MGPU_HOST_FRAME): every rank renders row bands into a local slab and copies
 *               them with one strided copy-engine transfer into a frame in shared pinned host memory -- all PCIe links
 *               in parallel, no rank ever holds the whole frame.
 *   completion  after its kernel (and copy) a rank's stream executes a one-thread kernel that stores the frame number
 *               with release semantics at system scope into its slot of the shared segment (page-locked and mapped by
 *               every process).  The root's CPU polls those slots -- no collective, no spinning kernel, every wait has
 *               a timeout on the CPU.  `ring` frame buffers: a rank renders frame k once the root has released frame
 *               k - ring.
 * NCCL is loaded with dlopen (libnccl.so.2: the copy a host framework already loaded, else the system one), so that
 * libvrcaster.so has no link-time dependency on it.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <errno.h>
#include <fcntl.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "../../include/vr_caster.h"
#include "vr_ctx.h"

#define VR_MGPU_MAX_RANKS 16
#define VR_MGPU_RING 4
#define VR_MGPU_MAGIC 0x5652364D47505532ull      /* "VR6MGPU2" */

/* host-visible state shared by the ranks (head of the shared segment) */
struct vr_mgpu_shared {
    volatile unsigned long long magic;           /* written last by rank 0: the fields below are valid */
    unsigned long long world, frame_bytes, host_frames;
    unsigned long long created_s;                /* CLOCK_REALTIME seconds when rank 0 created the segment (stale-segment guard) */
    unsigned char nccl_id[128];
    unsigned char frame_handle[VR_MGPU_RING][64];        /* cudaIpcMemHandle_t of the root's device frames */
    volatile unsigned long long joined[VR_MGPU_MAX_RANKS];        /* bootstrap barrier: generation reached per rank */
    volatile unsigned long long done[VR_MGPU_MAX_RANKS];          /* frames this rank has finished (stored by the device) */
    volatile unsigned long long released;                         /* frames the root has released for reuse */
    volatile unsigned long long failed;
};

struct vr_mgpu {
    int world = 1, rank = 0;
    unsigned flags = 0;
    char shm_name[96] = {0};
    size_t shm_bytes = 0, frames_offset = 0;
    vr_mgpu_shared *sh = nullptr;                /* host mapping */
    unsigned long long *d_done = nullptr;        /* device pointer of sh->done[rank] (mapped pinned memory) */
    bool registered = false;
    size_t frame_bytes = 0;
    int width = 0, height = 0, padded_rows = 0;
    uint8_t *frames[VR_MGPU_RING] = {};          /* device mode: root cudaMalloc / peers IPC-mapped */
    uint8_t *slabs[VR_MGPU_RING] = {};           /* host mode: this rank's bands, one slab per ring slot */
    unsigned long long issued = 0, generation = 0;
    /* consecutive frames are launched on two alternating streams: a 1/world share of a frame is only a few waves of CTAs
     * whose run times differ by an order of magnitude (sky vs horizon), and the first CTAs of frame k + 1 fill the SMs
     * the tail of frame k leaves idle.  Host frames: the copy of frame k runs on a third stream under frame k + 1. */
    cudaStream_t stream2 = nullptr, copy = nullptr;
    cudaEvent_t ev_rendered[VR_MGPU_RING] = {}, ev_ready = nullptr, ev_tail = nullptr;
    /* NCCL, loaded at run time */
    void *lib = nullptr;
    ncclComm_t comm = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    /* cuStreamWriteValue64 (driver API, resolved at run time): the completion counter is written by the stream itself,
     * after everything before it in the stream, with a system-wide memory barrier -- no kernel launch for the signal */
    CUresult (*WriteValue64)(CUstream, CUdeviceptr, cuuint64_t, unsigned int) = nullptr;
};

namespace {

constexpr int kBandRows = 8;                     /* host frames: interleaved bands of 8 rows (a multiple of the 4-row CTA tile) */

double now_s() {
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

/* polls a host-visible counter until it reaches `value`; false after `timeout_s` or when a rank reported failure */
bool wait_ge(const volatile unsigned long long *p, unsigned long long value, const vr_mgpu_shared *sh, double timeout_s) {
    const double t0 = now_s();
    for (unsigned spin = 0;; spin++) {
        if (__atomic_load_n(p, __ATOMIC_ACQUIRE) >= value) return true;
        if (sh && __atomic_load_n(&sh->failed, __ATOMIC_ACQUIRE)) return false;
        if ((spin & 1023u) == 1023u) {
            if (now_s() - t0 > timeout_s) return false;
            if (spin > (1u << 16)) usleep(50);
        }
    }
}

/* all ranks reach generation `gen` */
bool barrier(vr_mgpu *m, double timeout_s) {
    const unsigned long long gen = ++m->generation;
    __atomic_store_n(&m->sh->joined[m->rank], gen, __ATOMIC_RELEASE);
    for (int r = 0; r < m->world; r++)
        if (!wait_ge(&m->sh->joined[r], gen, m->sh, timeout_s)) return false;
    return true;
}

__global__ void vr_mgpu_signal(unsigned long long *slot, unsigned long long value) {
    /* everything this stream did before (the ray kernel's stores into a peer's frame over NVLink, the band copy into
     * host memory) is visible system-wide before the counter is */
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(slot), "l"(value) : "memory");
}

bool load_nccl(vr_ctx *c, vr_mgpu *m) {
    /* the copy the process already uses, if any (a host framework such as PyTorch ships its own NCCL under the same
     * soname: loading another one first would break that framework's later import); VR_NCCL_LIB overrides */
    const char *env = getenv("VR_NCCL_LIB");
    if (env && env[0]) m->lib = dlopen(env, RTLD_NOW | RTLD_LOCAL);
    if (!m->lib) m->lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!m->lib) m->lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!m->lib) m->lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!m->lib) return vr_i_fail(c, "mgpu: cannot load NCCL (%s)", dlerror()) != 0;
    *(void **)&m->GetUniqueId = dlsym(m->lib, "ncclGetUniqueId");
    *(void **)&m->CommInitRank = dlsym(m->lib, "ncclCommInitRank");
    *(void **)&m->Broadcast = dlsym(m->lib, "ncclBroadcast");
    *(void **)&m->CommDestroy = dlsym(m->lib, "ncclCommDestroy");
    *(void **)&m->GetErrorString = dlsym(m->lib, "ncclGetErrorString");
    if (!m->GetUniqueId || !m->CommInitRank || !m->Broadcast || !m->CommDestroy || !m->GetErrorString)
        return vr_i_fail(c, "mgpu: NCCL library lacks a required symbol") != 0;
    return true;
}

#define VR_NCCL(c, m, call)                                                                              \
    do {                                                                                                 \
        ncclResult_t r__ = (call);                                                                       \
        if (r__ != ncclSuccess) return vr_i_fail((c), "%s failed: %s", #call, (m)->GetErrorString(r__)); \
    } while (0)
#define VR_CU(c, call)                                                                                  \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) return vr_i_fail((c), "%s failed: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

/* the scheduler's second stream waits for what the context's stream has been given so far (uploads, the broadcast) */
bool catch_up(vr_ctx *c, vr_mgpu *m) {
    return cudaEventRecord(m->ev_ready, c->stream) == cudaSuccess && cudaStreamWaitEvent(m->stream2, m->ev_ready, 0) == cudaSuccess;
}

void teardown(vr_ctx *c) {
    vr_mgpu *m = c->mgpu;
    if (!m) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (m->comm && m->CommDestroy) m->CommDestroy(m->comm);
    for (int i = 0; i < VR_MGPU_RING; i++) {
        if (!m->frames[i]) continue;
        if (m->flags & VR_MGPU_HOST_FRAME) continue;                       /* part of the shared segment */
        if (m->rank == 0) cudaFree(m->frames[i]); else cudaIpcCloseMemHandle(m->frames[i]);
    }
    for (int i = 0; i < VR_MGPU_RING; i++) {
        if (m->slabs[i]) cudaFree(m->slabs[i]);
        if (m->ev_rendered[i]) cudaEventDestroy(m->ev_rendered[i]);
    }
    if (m->stream2) { cudaStreamSynchronize(m->stream2); cudaStreamDestroy(m->stream2); }
    if (m->copy) { cudaStreamSynchronize(m->copy); cudaStreamDestroy(m->copy); }
    if (m->ev_ready) cudaEventDestroy(m->ev_ready);
    if (m->ev_tail) cudaEventDestroy(m->ev_tail);
    if (m->registered) cudaHostUnregister((void *)m->sh);
    if (m->sh) munmap((void *)m->sh, m->shm_bytes);
    if (m->rank == 0 && m->shm_name[0]) shm_unlink(m->shm_name);
    /* the NCCL library stays loaded: a host framework may be using the same copy */
    c->tile_world = 1; c->tile_rank = 0;
    c->band_rows = 1; c->band_stride = 1; c->band_first = 0;
    delete m;
    c->mgpu = nullptr;
}

}  // namespace

int vr_mgpu_init(vr_ctx *c, const char *session, int world, int rank, unsigned flags) {
    if (!c || !session || !session[0]) return 0;
    if (world < 1 || world > VR_MGPU_MAX_RANKS || rank < 0 || rank >= world) return vr_i_fail(c, "mgpu_init: bad world / rank");
    if (c->width <= 0 || c->height <= 0) return vr_i_fail(c, "mgpu_init: create_viewport first (the frame size must be known)");
    if (strlen(session) > 60) return vr_i_fail(c, "mgpu_init: session name too long");
    if (c->mgpu) teardown(c);
    cudaSetDevice(c->device);
    vr_mgpu *m = new vr_mgpu();
    c->mgpu = m;
    m->world = world; m->rank = rank; m->flags = flags;
    m->width = c->width; m->height = c->height;
    const bool host = (flags & VR_MGPU_HOST_FRAME) != 0;
    /* frame rows padded so that every rank owns whole bands (host frames) */
    const int group = kBandRows * world;
    m->padded_rows = host ? ((c->height + group - 1) / group) * group : c->height;
    m->frame_bytes = (size_t)m->padded_rows * c->width * 4;
    m->frames_offset = (sizeof(vr_mgpu_shared) + 4095) & ~(size_t)4095;
    m->shm_bytes = m->frames_offset + (host ? VR_MGPU_RING * m->frame_bytes : 0);
    snprintf(m->shm_name, sizeof(m->shm_name), "/vrcaster_%s", session);
    if (!load_nccl(c, m)) { teardown(c); return 0; }

    /* ---- the shared segment: rank 0 creates and sizes it, the others wait for it.  A segment left behind by a crashed
     * run under the same session name is recognised by its age (rank 0 unlinks and recreates the name) and skipped. */
    void *map = MAP_FAILED;
    if (rank == 0) {
        shm_unlink(m->shm_name);
        int fd = shm_open(m->shm_name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)m->shm_bytes) != 0) {
            if (fd >= 0) close(fd);
            vr_i_fail(c, "mgpu_init: cannot create shared segment %s (%s)", m->shm_name, strerror(errno));
            teardown(c);
            return 0;
        }
        map = mmap(nullptr, m->shm_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
    } else {
        const double t0 = now_s();
        const unsigned long long entered = (unsigned long long)time(nullptr);
        for (;;) {
            int fd = shm_open(m->shm_name, O_RDWR, 0600);
            struct stat st;
            if (fd >= 0 && fstat(fd, &st) == 0 && (size_t)st.st_size >= m->shm_bytes) {
                map = mmap(nullptr, m->shm_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
                close(fd);
                if (map != MAP_FAILED) {
                    const vr_mgpu_shared *probe = static_cast<const vr_mgpu_shared *>(map);
                    /* wait for rank 0 to finish writing this incarnation; an old one is complete but too old */
                    const bool fresh = wait_ge(&probe->magic, VR_MGPU_MAGIC, nullptr, 2.0) && probe->magic == VR_MGPU_MAGIC &&
                                       probe->created_s + 120 >= entered;
                    if (fresh) break;
                    munmap(map, m->shm_bytes);
                    map = MAP_FAILED;
                }
            } else if (fd >= 0) {
                close(fd);
            }
            if (now_s() - t0 > 120.0) { vr_i_fail(c, "mgpu_init: timed out waiting for rank 0's segment %s", m->shm_name); teardown(c); return 0; }
            usleep(2000);
        }
    }
    if (map == MAP_FAILED) { vr_i_fail(c, "mgpu_init: mmap failed (%s)", strerror(errno)); teardown(c); return 0; }
    m->sh = static_cast<vr_mgpu_shared *>(map);
    if (cudaHostRegister(map, m->shm_bytes, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) {
        vr_i_fail(c, "mgpu_init: cudaHostRegister of the shared segment failed: %s", cudaGetErrorString(cudaGetLastError()));
        teardown(c);
        return 0;
    }
    m->registered = true;
    void *dmap = nullptr;
    if (cudaHostGetDevicePointer(&dmap, map, 0) != cudaSuccess) { vr_i_fail(c, "mgpu_init: no device pointer for the shared segment"); teardown(c); return 0; }
    m->d_done = reinterpret_cast<unsigned long long *>(static_cast<char *>(dmap) + offsetof(vr_mgpu_shared, done)) + rank;

    /* ---- rank 0 publishes the NCCL id and the frame buffers */
    if (rank == 0) {
        ncclUniqueId id;
        if (m->GetUniqueId(&id) != ncclSuccess) { vr_i_fail(c, "mgpu_init: ncclGetUniqueId failed"); teardown(c); return 0; }
        memcpy(m->sh->nccl_id, &id, sizeof(id) < 128 ? sizeof(id) : 128);
        m->sh->world = (unsigned long long)world;
        m->sh->frame_bytes = m->frame_bytes;
        m->sh->host_frames = host ? 1 : 0;
        m->sh->created_s = (unsigned long long)time(nullptr);
        if (!host) {
            for (int i = 0; i < VR_MGPU_RING; i++) {
                cudaIpcMemHandle_t h;
                if (cudaMalloc(&m->frames[i], m->frame_bytes) != cudaSuccess || cudaIpcGetMemHandle(&h, m->frames[i]) != cudaSuccess) {
                    vr_i_fail(c, "mgpu_init: frame buffer allocation / IPC export failed: %s", cudaGetErrorString(cudaGetLastError()));
                    teardown(c);
                    return 0;
                }
                memcpy(m->sh->frame_handle[i], &h, sizeof(h));
            }
        }
        __atomic_store_n(&m->sh->magic, VR_MGPU_MAGIC, __ATOMIC_RELEASE);
    } else {
        if (!wait_ge(&m->sh->magic, VR_MGPU_MAGIC, nullptr, 60.0) || m->sh->magic != VR_MGPU_MAGIC ||
            m->sh->world != (unsigned long long)world || m->sh->frame_bytes != m->frame_bytes || m->sh->host_frames != (host ? 1ull : 0ull)) {
            vr_i_fail(c, "mgpu_init: rank 0's segment does not match (world / viewport / flags differ?)");
            teardown(c);
            return 0;
        }
        if (!host) {
            for (int i = 0; i < VR_MGPU_RING; i++) {
                cudaIpcMemHandle_t h;
                memcpy(&h, m->sh->frame_handle[i], sizeof(h));
                if (cudaIpcOpenMemHandle((void **)&m->frames[i], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    vr_i_fail(c, "mgpu_init: cannot map the root's frame buffer: %s", cudaGetErrorString(cudaGetLastError()));
                    teardown(c);
                    return 0;
                }
            }
        }
    }
    if (host) {
        char *base = static_cast<char *>(dmap) + m->frames_offset;
        for (int i = 0; i < VR_MGPU_RING; i++) m->frames[i] = reinterpret_cast<uint8_t *>(base + (size_t)i * m->frame_bytes);
        const size_t slab_rows = (size_t)(m->padded_rows / world);
        for (int i = 0; i < VR_MGPU_RING; i++)
            if (cudaMalloc(&m->slabs[i], slab_rows * c->width * 4) != cudaSuccess) { vr_i_fail(c, "mgpu_init: slab allocation failed"); teardown(c); return 0; }
    }
    {
        bool ok = cudaStreamCreateWithFlags(&m->stream2, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&m->copy, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaEventCreateWithFlags(&m->ev_ready, cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&m->ev_tail, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < VR_MGPU_RING && ok; i++) ok = cudaEventCreateWithFlags(&m->ev_rendered[i], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { vr_i_fail(c, "mgpu_init: stream / event creation failed"); teardown(c); return 0; }
    }
    if (!getenv("VR_MGPU_SIGNAL_KERNEL")) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue64", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            *(void **)&m->WriteValue64 = fn;
        else
            cudaGetLastError();
    }
    /* ---- NCCL communicator (scene broadcast) */
    {
        ncclUniqueId id;
        memcpy(&id, m->sh->nccl_id, sizeof(id) < 128 ? sizeof(id) : 128);
        const ncclResult_t r = m->CommInitRank(&m->comm, world, id, rank);
        if (r != ncclSuccess) { vr_i_fail(c, "mgpu_init: ncclCommInitRank failed: %s", m->GetErrorString(r)); m->comm = nullptr; teardown(c); return 0; }
    }
    /* ---- this rank's share of a frame */
    if (host) {
        c->tile_world = 1; c->tile_rank = 0;
        c->band_rows = kBandRows; c->band_stride = world; c->band_first = rank;
    } else {
        c->band_rows = 1; c->band_stride = 1; c->band_first = 0;
        c->tile_world = world; c->tile_rank = rank;
    }
    if (!catch_up(c, m)) { vr_i_fail(c, "mgpu_init: stream set-up failed"); teardown(c); return 0; }
    if (!barrier(m, 120.0)) { vr_i_fail(c, "mgpu_init: a rank did not arrive"); teardown(c); return 0; }
    return 1;
}

int vr_mgpu_broadcast_octree(vr_ctx *c) {
    if (!c || !c->mgpu) return c ? vr_i_fail(c, "mgpu_broadcast_octree: call mgpu_init first") : 0;
    vr_mgpu *m = c->mgpu;
    cudaSetDevice(c->device);
    if (m->rank == 0 && !vr_i_ensure_tree(c)) return 0;
    /* sizes first (through a small device buffer), then the two arrays */
    unsigned long long meta_h[4] = {0, 0, 0, 0}, *meta_d = nullptr;
    if (m->rank == 0) { meta_h[0] = c->n_nodes; meta_h[1] = c->n_leaf_types; meta_h[2] = (unsigned long long)c->levels; meta_h[3] = (unsigned long long)c->tree_dim; }
    VR_CU(c, cudaMalloc(&meta_d, sizeof(meta_h)));
    VR_CU(c, cudaMemcpyAsync(meta_d, meta_h, sizeof(meta_h), cudaMemcpyHostToDevice, c->stream));
    VR_NCCL(c, m, m->Broadcast(meta_d, meta_d, sizeof(meta_h), ncclUint8, 0, m->comm, c->stream));
    VR_CU(c, cudaMemcpyAsync(meta_h, meta_d, sizeof(meta_h), cudaMemcpyDeviceToHost, c->stream));
    VR_CU(c, cudaStreamSynchronize(c->stream));
    cudaFree(meta_d);
    if (!meta_h[0] || !meta_h[1] || meta_h[2] < 1 || meta_h[2] > VR_MAX_LEVELS) return vr_i_fail(c, "mgpu_broadcast_octree: rank 0 has no octree");
    if (m->rank != 0) {
        vr_i_free_tree(c);
        VR_CU(c, cudaMalloc(&c->d_nodes, meta_h[0] * sizeof(vr_node)));
        VR_CU(c, cudaMalloc(&c->d_leaf_types, meta_h[1]));
    }
    VR_NCCL(c, m, m->Broadcast(c->d_nodes, c->d_nodes, meta_h[0] * sizeof(vr_node), ncclUint8, 0, m->comm, c->stream));
    VR_NCCL(c, m, m->Broadcast(c->d_leaf_types, c->d_leaf_types, meta_h[1], ncclUint8, 0, m->comm, c->stream));
    VR_CU(c, cudaStreamSynchronize(c->stream));
    if (m->rank != 0) {
        c->n_nodes = meta_h[0];
        c->n_leaf_types = meta_h[1];
        c->solid_voxels = meta_h[1];
        c->levels = (int)meta_h[2];
        c->tree_dim = (int)meta_h[3];
        c->tree_valid = true;
        c->tree_from_map = true;
    }
    if (!catch_up(c, m)) return vr_i_fail(c, "mgpu_broadcast_octree: stream set-up failed");
    return 1;
}

int vr_mgpu_frame(vr_ctx *c, uint64_t *frame_no) {
    if (!c || !c->mgpu) return c ? vr_i_fail(c, "mgpu_frame: call mgpu_init first") : 0;
    vr_mgpu *m = c->mgpu;
    cudaSetDevice(c->device);
    const unsigned long long k = m->issued;
    const int slot = (int)(k % VR_MGPU_RING);
    /* the buffer of frame k was last used by frame k - ring: wait until the root has released that one */
    if (k >= VR_MGPU_RING && !wait_ge(&m->sh->released, k - VR_MGPU_RING + 1, m->sh, 30.0))
        return vr_i_fail(c, "mgpu_frame: timed out waiting for the root to release frame %llu", k - VR_MGPU_RING);
    /* odd frames go to the second stream (ordered after the scene set-up by catch_up at init / broadcast) */
    cudaStream_t main = c->stream, s = (k & 1) ? m->stream2 : main;
    c->stream = s;
    const bool host = (m->flags & VR_MGPU_HOST_FRAME) != 0;
    const int ok = vr_i_launch_frame(c, host ? m->slabs[slot] : m->frames[slot], false);
    c->stream = main;
    if (!ok) { __atomic_store_n(&m->sh->failed, 1ull, __ATOMIC_RELEASE); return 0; }
    cudaStream_t tail = s;
    if (host) {
        /* this rank's bands -> frame order in shared pinned host memory: one strided copy over this GPU's PCIe link, on
         * the copy stream so that the next frame renders underneath it */
        const size_t band_bytes = (size_t)kBandRows * m->width * 4;
        const size_t nbands = (size_t)(m->padded_rows / (kBandRows * m->world));
        uint8_t *host_frame = reinterpret_cast<uint8_t *>(m->sh) + m->frames_offset + (size_t)slot * m->frame_bytes;
        VR_CU(c, cudaEventRecord(m->ev_rendered[slot], s));
        VR_CU(c, cudaStreamWaitEvent(m->copy, m->ev_rendered[slot], 0));
        VR_CU(c, cudaMemcpy2DAsync(host_frame + (size_t)m->rank * band_bytes, band_bytes * m->world, m->slabs[slot], band_bytes, band_bytes, nbands,
                                   cudaMemcpyDeviceToHost, m->copy));
        tail = m->copy;
    }
    if (m->WriteValue64 && m->WriteValue64((CUstream)tail, (CUdeviceptr)(uintptr_t)m->d_done, (cuuint64_t)(k + 1), CU_STREAM_WRITE_VALUE_DEFAULT) == CUDA_SUCCESS) {
        /* done: a stream memory operation, ordered after the kernel (and the copy) like a kernel would be */
    } else {
        m->WriteValue64 = nullptr;                       /* not supported here: a one-thread kernel does the same */
        vr_mgpu_signal<<<1, 1, 0, tail>>>(m->d_done, k + 1);
        c->launches++;
        VR_CU(c, cudaGetLastError());
    }
    m->issued = k + 1;
    if (frame_no) *frame_no = k;
    return 1;
}

int vr_mgpu_flush(vr_ctx *c) {
    if (!c || !c->mgpu) return 0;
    vr_mgpu *m = c->mgpu;
    cudaSetDevice(c->device);
    /* the context's stream waits for everything the scheduler enqueued on its own streams */
    VR_CU(c, cudaEventRecord(m->ev_tail, m->stream2));
    VR_CU(c, cudaStreamWaitEvent(c->stream, m->ev_tail, 0));
    VR_CU(c, cudaEventRecord(m->ev_tail, m->copy));
    VR_CU(c, cudaStreamWaitEvent(c->stream, m->ev_tail, 0));
    return 1;
}

int vr_mgpu_frame_wait(vr_ctx *c, uint64_t frame_no, const uint8_t **rgba) {
    if (!c || !c->mgpu) return c ? vr_i_fail(c, "mgpu_frame_wait: call mgpu_init first") : 0;
    vr_mgpu *m = c->mgpu;
    if (frame_no >= m->issued) return vr_i_fail(c, "mgpu_frame_wait: frame %llu was not issued", (unsigned long long)frame_no);
    if (m->rank != 0) {
        /* a non-root rank has nothing to collect: its part is done when its own counter says so */
        if (!wait_ge(&m->sh->done[m->rank], frame_no + 1, m->sh, 30.0)) return vr_i_fail(c, "mgpu_frame_wait: timed out");
        if (rgba) *rgba = nullptr;
        return 1;
    }
    for (int r = 0; r < m->world; r++)
        if (!wait_ge(&m->sh->done[r], frame_no + 1, m->sh, 30.0))
            return vr_i_fail(c, "mgpu_frame_wait: timed out waiting for rank %d (frame %llu)", r, (unsigned long long)frame_no);
    if (rgba) {
        /* device frames: the root's own buffer; host frames: the host address of the shared frame */
        if (m->flags & VR_MGPU_HOST_FRAME)
            *rgba = reinterpret_cast<const uint8_t *>(m->sh) + m->frames_offset + (size_t)(frame_no % VR_MGPU_RING) * m->frame_bytes;
        else
            *rgba = m->frames[frame_no % VR_MGPU_RING];
    }
    return 1;
}

int vr_mgpu_frame_release(vr_ctx *c, uint64_t frame_no) {
    if (!c || !c->mgpu) return 0;
    vr_mgpu *m = c->mgpu;
    if (m->rank != 0) return 1;
    if (frame_no + 1 > m->sh->released) __atomic_store_n(&m->sh->released, (unsigned long long)frame_no + 1, __ATOMIC_RELEASE);
    return 1;
}

int vr_mgpu_barrier(vr_ctx *c) {
    if (!c || !c->mgpu) return c ? vr_i_fail(c, "mgpu_barrier: call mgpu_init first") : 0;
    /* the bootstrap barrier of the shared segment (one generation counter per rank, spun on by the CPU): the ranks leave
     * it within a cache-line transfer of each other.  Nothing is enqueued and no stream is waited for. */
    if (!barrier(c->mgpu, 30.0)) return vr_i_fail(c, "mgpu_barrier: a rank did not arrive within 30 s");
    return 1;
}

int vr_mgpu_shutdown(vr_ctx *c) {
    if (!c || !c->mgpu) return 0;
    vr_mgpu *m = c->mgpu;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(m->stream2);
    cudaStreamSynchronize(m->copy);
    barrier(m, 30.0);                            /* nobody unmaps the root's frames while a peer may still write them */
    teardown(c);
    return 1;
}

void vr_i_mgpu_destroy(vr_ctx *c) { teardown(c); }
