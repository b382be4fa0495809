// capi_multi.inl — N GPUs behind the C ABI (included at the end of capi.cu): rtw_render_multi (one process, N devices; fused
// peer-memory reduce + resolve, or NCCL), rtw_comm_* / rtw_render_rank (one process per GPU, NCCL), rtw_scene_sync.
// The reference has no distributed path (its only parallelism is rayon over pixels, shared/src/camera.rs:353); the seam that
// grows an `n_gpus` is Camera::render as called from bin/src/main.rs:82-86.
#include <dlfcn.h>

namespace {

// ---- NCCL, loaded on demand -------------------------------------------------------------------------------------------
// Only the handful of entry points used here, declared locally (stable since NCCL 2.0) so that neither the build nor a
// single-GPU user of the library depends on NCCL being installed.
struct NcclId { char internal[128]; };
typedef void* NcclComm;
enum { kNcclUint8 = 1, kNcclUint64 = 5, kNcclSum = 0 };
struct Nccl {
    void* handle = nullptr;
    std::string error;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
    int (*CommInitAll)(NcclComm*, int, const int*) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok() const { return handle != nullptr; }
};
Nccl load_nccl() {
    Nccl n;
    std::vector<std::string> names;
    if (const char* e = std::getenv("RTW_NCCL_LIBRARY")) names.push_back(e);
    names.push_back("libnccl.so.2");
    names.push_back("libnccl.so");
    for (const std::string& name : names) {
        n.handle = dlopen(name.c_str(), RTLD_NOW | RTLD_GLOBAL);
        if (n.handle) break;
        n.error = dlerror();
    }
    if (!n.handle) return n;
    bool all = true;
    auto sym = [&](const char* name) { void* p = dlsym(n.handle, name); if (!p) { all = false; n.error = std::string("missing symbol ") + name; } return p; };
    n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(sym("ncclGetUniqueId"));
    n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(sym("ncclCommInitRank"));
    n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(sym("ncclCommInitAll"));
    n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
    n.Reduce = reinterpret_cast<decltype(n.Reduce)>(sym("ncclReduce"));
    n.Send = reinterpret_cast<decltype(n.Send)>(sym("ncclSend"));
    n.Recv = reinterpret_cast<decltype(n.Recv)>(sym("ncclRecv"));
    n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
    n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
    n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
    if (!all) { dlclose(n.handle); n.handle = nullptr; }
    return n;
}
Nccl& nccl() { static Nccl n = load_nccl(); return n; }
int need_nccl() {
    if (nccl().ok()) return RTW_OK;
    return fail(RTW_E_UNSUPPORTED, "NCCL is not available (" + nccl().error + "); set RTW_NCCL_LIBRARY or use RTW_COLLECTIVE_PEER");
}
#define NC(expr)                                                                                                    \
    do {                                                                                                            \
        int r_ = (expr);                                                                                            \
        if (r_ != 0) return fail(RTW_E_CUDA, std::string(#expr) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r_) : "NCCL error")); \
    } while (0)

// packed accumulator block: [3 * slots] u64 radiance sums, then [slots] u32 poison words (padded to a whole u64)
size_t block_words(size_t slots) { return 3 * slots + (slots + 1) / 2; }
void sample_share(uint32_t spp, uint32_t rank, uint32_t world, uint32_t* begin, uint32_t* count) {
    uint32_t b = (uint32_t)((uint64_t)spp * rank / world), e = (uint32_t)((uint64_t)spp * (rank + 1) / world);
    *begin = b; *count = e - b;
}
bool fixed_point_renderer(const rtw_scene* s, const rtw_opts* o) {
    return o->precision == RTW_F32 && !(o->flags & RTW_FLAG_LANE_PER_PIXEL);
    (void)s;
}

struct DeviceGuard {
    int saved = -1;
    DeviceGuard() { if (cudaGetDevice(&saved) != cudaSuccess) { cudaGetLastError(); saved = -1; } }
    ~DeviceGuard() { if (saved >= 0) cudaSetDevice(saved); }
};

}  // namespace

// one replica per device of rtw_render_multi; replica 0 is the scene handle itself
struct MultiReplica {
    int device = 0;
    rtw_scene* scene = nullptr;          // clone on `device` (replica 0: the caller's handle, not owned)
    bool owned = false;
    cudaStream_t stream = nullptr;
    cudaEvent_t rendered = nullptr, resolved = nullptr;
    DevBuf<unsigned long long> block;    // sample partition: this device's packed accumulators
    DevBuf<unsigned char> tiles;         // tile partition: this device's tile buffer
    NcclComm comm = nullptr;
};
struct rtw_comm { NcclComm comm = nullptr; int rank = 0, world = 1, device = 0; };

namespace {

void multi_release(rtw_scene* s) {
    DeviceGuard guard;
    for (MultiReplica* r : s->replicas) {
        cudaSetDevice(r->device);
        if (r->comm && nccl().ok()) nccl().CommDestroy(r->comm);
        r->block.release(); r->tiles.release();
        if (r->rendered) cudaEventDestroy(r->rendered);
        if (r->resolved) cudaEventDestroy(r->resolved);
        if (r->stream) cudaStreamDestroy(r->stream);
        if (r->owned && r->scene) rtw_scene_destroy(r->scene);
        delete r;
    }
    s->replicas.clear();
}

// the same scene on another device, from the host copy every handle keeps
int clone_scene(const rtw_scene* src, int device, rtw_scene** out) {
    DeviceGuard guard;
    CU(cudaSetDevice(device));
    const int saved_builder = g_bvh_builder;
    g_bvh_builder = src->bvh_builder;                    // same tree on every device (Hittable::hit does not depend on it; the timing does)
    int rc;
    if (src->general) {
        const GeneralDesc& g = src->gdesc;
        rtw_scene_desc d{};
        d.spheres = g.spheres.data(); d.n_spheres = g.spheres.size(); d.planes = g.planes.data(); d.n_planes = g.planes.size();
        d.quads = g.quads.data(); d.n_quads = g.quads.size(); d.cuboids = g.cuboids.data(); d.n_cuboids = g.cuboids.size();
        d.transforms = g.transforms.data(); d.n_transforms = g.transforms.size(); d.materials = g.materials.data(); d.n_materials = g.materials.size();
        d.textures = g.textures.data(); d.n_textures = g.textures.size(); d.perlins = g.perlins.data(); d.n_perlins = g.perlins.size();
        d.world = g.world.data(); d.n_world = g.world.size(); d.lights = g.lights.data(); d.n_lights = g.lights.size();
        d.lights_is_bvh = g.lights_is_bvh ? 1u : 0u;
        rc = rtw_scene_create_general(&d, out);
    } else {
        rc = rtw_scene_create(src->spheres.data(), src->sphere_material.data(), src->spheres.size(), src->planes.data(), src->plane_material.data(),
                              src->planes.size(), src->materials.data(), src->materials.size(), src->lights.data(), src->lights.size(), out);
    }
    g_bvh_builder = saved_builder;
    return rc;
}

// replicas on exactly `devices` (rebuilt when the device list changes)
int ensure_replicas(rtw_scene* s, int n, const int* devices) {
    bool same = (int)s->replicas.size() == n;
    for (int i = 0; same && i < n; ++i) same = s->replicas[i]->device == devices[i];
    if (same) return RTW_OK;
    multi_release(s);
    DeviceGuard guard;
    for (int i = 0; i < n; ++i) {
        MultiReplica* r = new MultiReplica();
        r->device = devices[i];
        s->replicas.push_back(r);
        if (devices[i] == s->device && i == 0) r->scene = s;
        else {
            int rc = clone_scene(s, devices[i], &r->scene);
            if (rc) { multi_release(s); return rc; }
            r->owned = true;
        }
        cudaError_t e = cudaSetDevice(devices[i]);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->rendered, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->resolved, cudaEventDisableTiming);
        if (e != cudaSuccess) { multi_release(s); cudaGetLastError(); return fail(RTW_E_CUDA, cudaGetErrorString(e)); }
    }
    return RTW_OK;
}

// peer access between every ordered pair; *ok = false when some pair cannot
int enable_peer_access(int n, const int* devices, bool* ok) {
    *ok = true;
    DeviceGuard guard;
    for (int a = 0; a < n && *ok; ++a)
        for (int b = 0; b < n; ++b) {
            if (a == b) continue;
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, devices[a], devices[b]));
            if (!can) { *ok = false; break; }
        }
    if (!*ok) return RTW_OK;
    for (int a = 0; a < n; ++a) {
        CU(cudaSetDevice(devices[a]));
        for (int b = 0; b < n; ++b) {
            if (a == b) continue;
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); continue; }
            if (e != cudaSuccess) { cudaGetLastError(); return fail(RTW_E_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); }
        }
    }
    return RTW_OK;
}

int collect_stats(rtw_scene* s, rtw_stats* stats, double total_ms) {
    // counters of every replica (after the streams have been synchronised); kernel time = the slowest device's
    DeviceGuard guard;
    std::memset(stats, 0, sizeof(*stats));
    double kernel_ms = 0.;
    uint32_t launches = 0;
    for (MultiReplica* r : s->replicas) {
        CU(cudaSetDevice(r->device));
        DeviceCounters c;
        CU(cudaMemcpy(&c, r->scene->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
        rtw_stats one{};
        read_stats(c, &one);
        stats->paths += one.paths; stats->rays += one.rays; stats->node_visits += one.node_visits; stats->sphere_tests += one.sphere_tests;
        stats->light_tests += one.light_tests; stats->lambertian += one.lambertian; stats->metal += one.metal; stats->dielectric += one.dielectric;
        stats->absorbed += one.absorbed; stats->missed += one.missed; stats->depth_out += one.depth_out;
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, r->scene->ev[0], r->scene->ev[1]));
        kernel_ms = std::max(kernel_ms, (double)ms);
        launches += r->scene->last_launches;
    }
    stats->kernel_ms = kernel_ms; stats->total_ms = total_ms; stats->launches = launches + (uint32_t)s->replicas.size();
    return RTW_OK;
}

}  // namespace

extern "C" {

int rtw_scene_sync(rtw_scene* s, double* kernel_ms) {
    if (!s) return fail(RTW_E_INVALID, "scene is NULL");
    DeviceGuard guard;
    CU(cudaSetDevice(s->device));
    CU(cudaDeviceSynchronize());
    if (kernel_ms) {
        float ms = 0.f;
        cudaError_t e = cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]);
        if (e != cudaSuccess) { cudaGetLastError(); ms = 0.f; }       // no render call yet
        *kernel_ms = ms;
    }
    return check_reference_panic(s);
}

int rtw_render_multi(rtw_scene* s, const rtw_camera* cam, const rtw_opts* o, int n_gpus, const int* devices, uint32_t collective,
                     double* rgb_sum, uint8_t* rgb8, rtw_stats* stats) {
    if (!s) return fail(RTW_E_INVALID, "scene is NULL");
    int rc = check_camera(cam); if (rc) return rc;
    rc = check_opts(o); if (rc) return rc;
    if (collective > RTW_COLLECTIVE_NCCL) return fail(RTW_E_INVALID, "collective");
    int ndev = rtw_device_count();
    if (ndev < 0) return ndev;
    if (n_gpus < 1 || n_gpus > ndev || n_gpus > kMaxPeers) return fail(RTW_E_INVALID, "n_gpus: " + std::to_string(n_gpus) + " requested, " + std::to_string(ndev) + " CUDA device(s) visible");
    std::vector<int> devs(n_gpus);
    for (int i = 0; i < n_gpus; ++i) {
        devs[i] = devices ? devices[i] : (i == 0 ? s->device : (i <= s->device ? i - 1 : i));      // the scene's own device first
        if (devs[i] < 0 || devs[i] >= ndev) return fail(RTW_E_INVALID, "device ordinal out of range");
        for (int k = 0; k < i; ++k) if (devs[k] == devs[i]) return fail(RTW_E_INVALID, "duplicate device");
    }
    if (n_gpus == 1 && devs[0] == s->device) return rtw_render(s, cam, o, rgb_sum, rgb8, stats);
    if (devs[0] != s->device) return fail(RTW_E_INVALID, "devices[0] must be the device the scene was created on");
    DeviceGuard guard;
    rc = ensure_replicas(s, n_gpus, devs.data()); if (rc) return rc;
    const uint32_t w = cam->image_width, h = cam->image_height, spp = cam->samples_per_pixel, N = (uint32_t)n_gpus;
    const size_t npx = (size_t)w * h;
    const bool samples = fixed_point_renderer(s, o) && N <= 15 && spp >= N;
    MultiReplica* root = s->replicas[0];
    CU(cudaSetDevice(root->device));
    if (rgb_sum) CU(s->d_rgb_sum.reserve(npx * 3));
    if (rgb8) CU(s->d_rgb8.reserve(npx * 3));
    double* d_sum = rgb_sum ? s->d_rgb_sum.p : nullptr;
    uint8_t* d_8 = rgb8 ? s->d_rgb8.p : nullptr;
    CU(cudaEventRecord(s->ev[2], root->stream));
    if (samples) {
        bool peer = collective != RTW_COLLECTIVE_NCCL;
        if (peer) {
            bool ok = false;
            rc = enable_peer_access(n_gpus, devs.data(), &ok); if (rc) return rc;
            if (!ok && collective == RTW_COLLECTIVE_PEER) return fail(RTW_E_UNSUPPORTED, "RTW_COLLECTIVE_PEER: no peer access between some pair of the devices");
            peer = ok;
        }
        if (!peer) {
            rc = need_nccl(); if (rc) return rc;
            if (!s->replicas[0]->comm) {
                std::vector<NcclComm> comms(n_gpus, nullptr);
                NC(nccl().CommInitAll(comms.data(), n_gpus, devs.data()));
                for (int i = 0; i < n_gpus; ++i) s->replicas[i]->comm = comms[i];
            }
        }
        const size_t slots = rtw_accum_slots(w, h), words = block_words(slots);
        for (uint32_t g = 0; g < N; ++g) {
            MultiReplica* r = s->replicas[g];
            CU(cudaSetDevice(r->device));
            CU(r->block.reserve(words));
            SampleRange sr; sr.own_rank = g; sr.own_world = N;        // its pixels, or its samples of every pixel (render_device_impl)
            rc = render_device_impl(r->scene, cam, o, 0, 1, nullptr, r->stream, nullptr, sr, r->block.p, reinterpret_cast<uint32_t*>(r->block.p + 3 * slots));
            if (rc) return rc;
            CU(cudaEventRecord(r->rendered, r->stream));
        }
        if (peer) {
            PeerBlocks B{};
            B.n = n_gpus;
            for (uint32_t g = 0; g < N; ++g) { B.accum[g] = s->replicas[g]->block.p; B.poison[g] = reinterpret_cast<const uint32_t*>(s->replicas[g]->block.p + 3 * slots); }
            for (uint32_t g = 0; g < N; ++g) {
                MultiReplica* r = s->replicas[g];
                CU(cudaSetDevice(r->device));
                for (uint32_t k = 0; k < N; ++k) if (k != g) CU(cudaStreamWaitEvent(r->stream, s->replicas[k]->rendered, 0));
                // whole tiles per GPU: slot ranges cut at multiples of 256
                const uint32_t tiles = (uint32_t)(slots / 256);
                const uint32_t q0 = (uint32_t)((uint64_t)tiles * g / N) * 256u, q1 = (uint32_t)((uint64_t)tiles * (g + 1) / N) * 256u;
                CU(launch_peer_reduce_resolve_f32(B, q0, q1, w, h, spp, d_sum, d_8, r->stream));
                CU(cudaEventRecord(r->resolved, r->stream));
            }
            CU(cudaSetDevice(root->device));
            for (uint32_t k = 1; k < N; ++k) CU(cudaStreamWaitEvent(root->stream, s->replicas[k]->resolved, 0));
        } else {
            NC(nccl().GroupStart());
            for (uint32_t g = 0; g < N; ++g) {
                MultiReplica* r = s->replicas[g];
                CU(cudaSetDevice(r->device));
                NC(nccl().Reduce(r->block.p, r->block.p, words, kNcclUint64, kNcclSum, 0, r->comm, r->stream));
            }
            NC(nccl().GroupEnd());
            CU(cudaSetDevice(root->device));
            CU(launch_resolve_accum_f32(root->block.p, reinterpret_cast<const uint32_t*>(root->block.p + 3 * slots), w, h, spp, d_sum, d_8, root->stream));
        }
    } else {
        // tile partition: device g renders tile slots g, g + N, ...; the buffers are copied to the root and untiled there
        const size_t elem = o->precision == RTW_F32 ? sizeof(float) : sizeof(double);
        const size_t bytes = (size_t)rtw_tiles_per_rank(w, h, N) * kTileW * kTileH * 3 * elem;
        CU(cudaSetDevice(root->device));
        CU(s->d_gather.reserve(bytes * N));
        for (uint32_t g = 0; g < N; ++g) {
            MultiReplica* r = s->replicas[g];
            CU(cudaSetDevice(r->device));
            CU(r->tiles.reserve(bytes));
            rc = render_device_impl(r->scene, cam, o, g, N, r->tiles.p, r->stream, nullptr, SampleRange(), nullptr, nullptr);
            if (rc) return rc;
            CU(cudaMemcpyPeerAsync(s->d_gather.p + bytes * g, root->device, r->tiles.p, r->device, bytes, r->stream));
            CU(cudaEventRecord(r->rendered, r->stream));
        }
        CU(cudaSetDevice(root->device));
        for (uint32_t k = 1; k < N; ++k) CU(cudaStreamWaitEvent(root->stream, s->replicas[k]->rendered, 0));
        rc = rtw_untile_resolve_device(s->d_gather.p, o->precision, w, h, N, spp, d_sum, d_8, root->stream);
        if (rc) return rc;
    }
    CU(cudaSetDevice(root->device));
    if (rgb_sum) CU(cudaMemcpyAsync(rgb_sum, d_sum, npx * 3 * sizeof(double), cudaMemcpyDeviceToHost, root->stream));
    if (rgb8) CU(cudaMemcpyAsync(rgb8, d_8, npx * 3, cudaMemcpyDeviceToHost, root->stream));
    CU(cudaEventRecord(s->ev[3], root->stream));
    for (MultiReplica* r : s->replicas) { CU(cudaSetDevice(r->device)); CU(cudaStreamSynchronize(r->stream)); }
    for (MultiReplica* r : s->replicas) { rc = check_reference_panic(r->scene); if (rc) return rc; }
    if (stats) {
        CU(cudaSetDevice(root->device));
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, s->ev[2], s->ev[3]));
        rc = collect_stats(s, stats, t); if (rc) return rc;
    }
    return RTW_OK;
}

// ---- one process per GPU ------------------------------------------------------------------------------------------------
int rtw_comm_unique_id(uint8_t id[RTW_COMM_ID_BYTES]) {
    if (!id) return fail(RTW_E_INVALID, "id is NULL");
    int rc = need_nccl(); if (rc) return rc;
    NcclId u;
    NC(nccl().GetUniqueId(&u));
    std::memcpy(id, u.internal, RTW_COMM_ID_BYTES);
    return RTW_OK;
}
int rtw_comm_init_rank(const uint8_t id[RTW_COMM_ID_BYTES], int rank, int world, rtw_comm** out) {
    if (!id || !out) return fail(RTW_E_INVALID, "NULL argument");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail(RTW_E_INVALID, "rank/world");
    int rc = need_nccl(); if (rc) return rc;
    int ndev = rtw_device_count();
    if (ndev <= 0) return ndev < 0 ? ndev : fail(RTW_E_NO_DEVICE, "no CUDA device: this backend has no CPU fallback");
    rtw_comm* c = new rtw_comm();
    c->rank = rank; c->world = world;
    cudaError_t e = cudaGetDevice(&c->device);
    if (e != cudaSuccess) { delete c; cudaGetLastError(); return fail(RTW_E_CUDA, cudaGetErrorString(e)); }
    NcclId u;
    std::memcpy(u.internal, id, RTW_COMM_ID_BYTES);
    int r = nccl().CommInitRank(&c->comm, world, u, rank);
    if (r != 0) { delete c; return fail(RTW_E_CUDA, std::string("ncclCommInitRank: ") + nccl().GetErrorString(r)); }
    *out = c;
    return RTW_OK;
}
void rtw_comm_destroy(rtw_comm* c) {
    if (!c) return;
    if (c->comm && nccl().ok()) nccl().CommDestroy(c->comm);
    delete c;
}
int rtw_comm_rank(const rtw_comm* c) { return c ? c->rank : fail(RTW_E_INVALID, "comm is NULL"); }
int rtw_comm_world(const rtw_comm* c) { return c ? c->world : fail(RTW_E_INVALID, "comm is NULL"); }

int rtw_render_rank_device(rtw_scene* s, const rtw_camera* cam, const rtw_opts* o, rtw_comm* c, double* d_rgb_sum, uint8_t* d_rgb8,
                           void* stream, rtw_stats* stats) {
    if (!s || !c) return fail(RTW_E_INVALID, "NULL argument");
    int rc = check_camera(cam); if (rc) return rc;
    rc = check_opts(o); if (rc) return rc;
    if (c->device != s->device) return fail(RTW_E_INVALID, "the communicator and the scene live on different devices");
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t w = cam->image_width, h = cam->image_height, spp = cam->samples_per_pixel, N = (uint32_t)c->world, rank = (uint32_t)c->rank;
    CU(cudaSetDevice(s->device));
    const bool samples = fixed_point_renderer(s, o) && N <= 15 && spp >= N;
    if (samples) {
        const size_t slots = rtw_accum_slots(w, h), words = block_words(slots);
        CU(s->d_block.reserve(words));
        uint32_t* poison = reinterpret_cast<uint32_t*>(s->d_block.p + 3 * slots);
        SampleRange sr; sr.own_rank = rank; sr.own_world = N;         // its pixels, or its samples of every pixel (render_device_impl)
        rc = render_device_impl(s, cam, o, 0, 1, nullptr, st, nullptr, sr, s->d_block.p, poison);
        if (rc) return rc;
        if (N > 1) NC(nccl().Reduce(s->d_block.p, s->d_block.p, words, kNcclUint64, kNcclSum, 0, c->comm, st));
        if (rank == 0) CU(launch_resolve_accum_f32(s->d_block.p, poison, w, h, spp, d_rgb_sum, d_rgb8, st));
    } else {
        const size_t elem = o->precision == RTW_F32 ? sizeof(float) : sizeof(double);
        const size_t bytes = (size_t)rtw_tiles_per_rank(w, h, N) * kTileW * kTileH * 3 * elem;
        CU(s->d_tiles_rank.reserve(bytes));
        if (rank == 0) CU(s->d_gather.reserve(bytes * N));
        rc = render_device_impl(s, cam, o, rank, N, s->d_tiles_rank.p, st, nullptr, SampleRange(), nullptr, nullptr);
        if (rc) return rc;
        if (rank == 0) {
            CU(cudaMemcpyAsync(s->d_gather.p, s->d_tiles_rank.p, bytes, cudaMemcpyDeviceToDevice, st));
            if (N > 1) {
                NC(nccl().GroupStart());
                for (uint32_t r = 1; r < N; ++r) NC(nccl().Recv(s->d_gather.p + bytes * r, bytes, kNcclUint8, (int)r, c->comm, st));
                NC(nccl().GroupEnd());
            }
            rc = rtw_untile_resolve_device(s->d_gather.p, o->precision, w, h, N, spp, d_rgb_sum, d_rgb8, st);
            if (rc) return rc;
        } else {
            NC(nccl().Send(s->d_tiles_rank.p, bytes, kNcclUint8, 0, c->comm, st));
        }
    }
    if (stats) {
        DeviceCounters cn;
        CU(cudaMemcpyAsync(&cn, s->d_counters, sizeof(cn), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        rc = check_reference_panic(s); if (rc) return rc;
        std::memset(stats, 0, sizeof(*stats));
        read_stats(cn, stats);
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]));
        stats->kernel_ms = ms; stats->total_ms = ms; stats->launches = s->last_launches + (rank == 0 ? 1u : 0u);
    }
    return RTW_OK;
}

int rtw_render_rank(rtw_scene* s, const rtw_camera* cam, const rtw_opts* o, rtw_comm* c, double* rgb_sum, uint8_t* rgb8, rtw_stats* stats) {
    if (!s || !c) return fail(RTW_E_INVALID, "NULL argument");
    int rc = check_camera(cam); if (rc) return rc;
    const size_t npx = (size_t)cam->image_width * cam->image_height;
    const bool root = c->rank == 0;
    CU(cudaSetDevice(s->device));
    if (root && rgb_sum) CU(s->d_rgb_sum.reserve(npx * 3));
    if (root && rgb8) CU(s->d_rgb8.reserve(npx * 3));
    CU(cudaEventRecord(s->ev[2], 0));
    rtw_stats local{};
    rc = rtw_render_rank_device(s, cam, o, c, root && rgb_sum ? s->d_rgb_sum.p : nullptr, root && rgb8 ? s->d_rgb8.p : nullptr, nullptr, &local);
    if (rc) return rc;
    if (root && rgb_sum) CU(cudaMemcpy(rgb_sum, s->d_rgb_sum.p, npx * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (root && rgb8) CU(cudaMemcpy(rgb8, s->d_rgb8.p, npx * 3, cudaMemcpyDeviceToHost));
    CU(cudaEventRecord(s->ev[3], 0));
    CU(cudaEventSynchronize(s->ev[3]));
    if (stats) {
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, s->ev[2], s->ev[3]));
        *stats = local;
        stats->total_ms = t;
    }
    return RTW_OK;
}

}  // extern "C"
