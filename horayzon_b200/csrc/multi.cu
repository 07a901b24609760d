// multi.cu -- multi-GPU host tier: one process, one host thread per GPU.
//
// The reference parallelises horizon_gridded_comp over rows of the inner domain with TBB
// (horizon_comp.cpp:739-744) and the shadow loop over rows as well (shadow_comp.cpp:390-394); no cell
// reads another cell's result.  hzb_horizon_gridded_multi deals the 4-row blocks of the inner domain out
// to the GPUs in turn (block b to shard b % n: every GPU gets the same mix of cheap rim rows and
// expensive centre rows), every GPU builds the BVH of the replicated DEM and computes its blocks into a
// packed send buffer, and the shards are joined
//   * device_gather == 0: by each GPU copying its own blocks to their places in the caller's host array
//     over its own PCIe link (the fastest way to a HOST result -- no GPU needs the whole array), or
//   * device_gather != 0: by ONE in-place ncclAllGather of the packed shards over NVLink / NVSwitch
//     (single process, ncclCommInitAll), after which every GPU holds the whole array; GPU 0 puts the
//     blocks in domain order and returns them.  This is the device-side product the resident tier and
//     bench.py use (there through torch.distributed, one process per GPU).
// NCCL is looked up at run time (dlopen of libnccl.so.2 -- the copy PyTorch has already loaded when the
// caller is a Python process), so the library itself links against nothing but the CUDA runtime.
#include "hzb_common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace hzb {
namespace {

struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) { err = std::string("NCCL not found: ") + dlerror(); return false; }
#define HZB_SYM(field, name) field = (decltype(field))dlsym(lib, name); if (!field) { err = std::string("NCCL symbol missing: ") + name; lib = nullptr; return false; }
        HZB_SYM(CommInitAll, "ncclCommInitAll") HZB_SYM(CommDestroy, "ncclCommDestroy") HZB_SYM(AllGather, "ncclAllGather")
        HZB_SYM(GroupStart, "ncclGroupStart") HZB_SYM(GroupEnd, "ncclGroupEnd") HZB_SYM(GetErrorString, "ncclGetErrorString")
#undef HZB_SYM
        return true;
    }
};
Nccl& nccl() { static Nccl n; return n; }

// packed shards [n][per_rows][row_elems] -> domain order [rows][row_elems] (block j of shard r is block j * n + r)
__global__ void k_unpack_blocks(const float4* __restrict__ gathered, float4* __restrict__ out, int rows, long long row_vec,
                                int n, int per_rows) {
    const long long total = (long long)rows * row_vec;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(g / row_vec);
        const long long e = g - (long long)row * row_vec;
        const int blk = row >> 2, r = blk % n, j = blk / n;
        out[g] = gathered[((long long)r * per_rows + j * 4 + (row & 3)) * row_vec + e];
    }
}
__global__ void k_unpack_blocks_f1(const float* __restrict__ gathered, float* __restrict__ out, int rows, long long row_elems,
                                   int n, int per_rows) {
    const long long total = (long long)rows * row_elems;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(g / row_elems);
        const long long e = g - (long long)row * row_elems;
        const int blk = row >> 2, r = blk % n, j = blk / n;
        out[g] = gathered[((long long)r * per_rows + j * 4 + (row & 3)) * row_elems + e];
    }
}

struct ShardResult { int rc = 0; std::string err; hzb_stats st{}; double t_scene = 0, t_kernel = 0; };

}  // namespace
}  // namespace hzb

using namespace hzb;

extern "C" {

int hzb_set_device(int device) {
    if (cudaSetDevice(device) != cudaSuccess) { set_error(std::string("cudaSetDevice(") + std::to_string(device) + ") failed"); cudaGetLastError(); return 1; }
    return 0;
}

// Multi-GPU twin of hzb_horizon_gridded (same leading arguments; horizon_comp.h:8-20).  n_devices <= 0: all
// visible devices.  n_shards <= 0: one shard per device; more shards than devices are dealt round-robin (the tests
// run two shards on one GPU).  vec_tilt / svf_buffer: optional fused sky view factor (both NULL: none).
int hzb_horizon_gridded_multi(const float* vert_grid, int dem_dim_0, int dem_dim_1, const float* vec_norm, const float* vec_north,
                              int offset_0, int offset_1, float* hori_buffer, int dim_in_0, int dim_in_1, int azim_num,
                              float dist_search, float hori_acc, const char* ray_algorithm, const char* geom_type,
                              const float* vert_simp, int num_vert_simp, const int32_t* tri_ind_simp, int num_tri_simp,
                              float elev_ang_low_lim, const uint8_t* mask, float hori_fill, float ray_org_elev,
                              const float* vec_tilt, float* svf_buffer, int n_devices, int n_shards, int device_gather) {
    const double t_start = now_s();
    const int avail = hzb_device_count();
    if (avail <= 0) { set_error("no CUDA device available (libhorayzon_b200 has no CPU fallback)"); return 1; }
    if (parse_algorithm(ray_algorithm) < 0) { set_error("invalid input argument for ray_algorithm"); return 1; }
    if (parse_geom_type(geom_type) < 0) { set_error("invalid input argument for geom_type"); return 1; }
    if (!vert_grid || !vec_norm || !vec_north || !hori_buffer || !mask) { set_error("null pointer argument"); return 1; }
    if ((vec_tilt == nullptr) != (svf_buffer == nullptr)) { set_error("vec_tilt and svf_buffer go together"); return 1; }
    if (dim_in_0 < 0 || dim_in_1 < 0 || azim_num < 1 || (svf_buffer && azim_num < 2)) { set_error("invalid dimensions"); return 1; }
    const int ndev = n_devices <= 0 ? avail : std::min(n_devices, avail);
    const int n = n_shards <= 0 ? ndev : n_shards;
    if (device_gather && n != ndev) { set_error("the NCCL all-gather needs one shard per device"); return 1; }
    if (dim_in_0 == 0 || dim_in_1 == 0) return 0;
    int dev0 = 0; cudaGetDevice(&dev0);

    const size_t row_elems = (size_t)dim_in_1 * (size_t)azim_num;
    const int per_rows = hzb_shard_rows(dim_in_0, 0, n);          // shard 0 holds the most blocks
    std::vector<ShardResult> res((size_t)n);
    std::vector<hzb_scene*> scenes((size_t)n, nullptr);
    std::vector<float*> d_gather((size_t)n, nullptr), d_in((size_t)n * 2, nullptr), d_svf((size_t)n, nullptr), d_tilt((size_t)n, nullptr), d_azim((size_t)n, nullptr);
    std::vector<uint8_t*> d_mask((size_t)n, nullptr);
    std::vector<cudaStream_t> streams((size_t)n, nullptr);
    std::vector<float> az((size_t)azim_num);
    for (int i = 0; i < azim_num; ++i) az[i] = (float)((2 * M_PI) / azim_num * i);    // horizon.pyx:190-195
    const size_t nc = (size_t)dim_in_0 * dim_in_1;

    const bool out_pinned = host_is_pinned(hori_buffer);
    auto fail = [&](int r, const std::string& m) { res[r].rc = 1; res[r].err = m; };
    // ---- phase 1 (thread per shard): scene, inputs, kernels into the packed shard buffer
    auto phase1 = [&](int r) {
        const int dev = r % ndev;
        if (cudaSetDevice(dev) != cudaSuccess) { fail(r, "cudaSetDevice failed"); return; }
        const double t0 = now_s();
        scenes[r] = hzb_scene_create(vert_grid, dem_dim_0, dem_dim_1, vert_simp, num_vert_simp, tri_ind_simp, num_tri_simp, dev);
        if (!scenes[r]) { fail(r, hzb_last_error()); return; }
        res[r].t_scene = now_s() - t0;
        if (cudaStreamCreateWithFlags(&streams[r], cudaStreamNonBlocking) != cudaSuccess) { fail(r, "stream creation failed"); return; }
        auto up = [&](const void* h, size_t bytes) -> void* {
            void* d = pool_alloc(bytes);
            if (d && cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, streams[r]) != cudaSuccess) { pool_free(d); d = nullptr; }
            return d;
        };
        d_in[2 * r] = (float*)up(vec_norm, nc * 12); d_in[2 * r + 1] = (float*)up(vec_north, nc * 12); d_mask[r] = (uint8_t*)up(mask, nc);
        // the all-gather receives every shard: [n][per_rows][row_elems]; otherwise only this shard's rows are needed
        const size_t buf_rows = device_gather ? (size_t)n * per_rows : (size_t)per_rows;
        d_gather[r] = (float*)pool_alloc(buf_rows * row_elems * sizeof(float));
        if (!d_in[2 * r] || !d_in[2 * r + 1] || !d_mask[r] || !d_gather[r]) { fail(r, std::string("device allocation / upload failed: ") + hzb_last_error()); return; }
        float* mine = d_gather[r] + (device_gather ? (size_t)r * per_rows * row_elems : 0);
        const double t1 = now_s();
        if (hzb_horizon_gridded_dev_sharded(scenes[r], d_in[2 * r], d_in[2 * r + 1], d_mask[r], offset_0, offset_1, dim_in_0, dim_in_1, azim_num,
                                            dist_search, hori_acc, ray_algorithm, elev_ang_low_lim, hori_fill, ray_org_elev, mine, r, n, 1,
                                            streams[r])) { fail(r, hzb_last_error()); return; }
        const int my_rows = hzb_shard_rows(dim_in_0, r, n);
        if (svf_buffer && my_rows > 0) {
            // per-shard tilt vectors in packed order (host gather), integral on the packed shard
            std::vector<float> tl((size_t)my_rows * dim_in_1 * 3, 0.f);
            for (int lr = 0; lr < my_rows; ++lr) {
                const int row = ((lr >> 2) * n + r) * 4 + (lr & 3);
                if (row < dim_in_0) memcpy(&tl[(size_t)lr * dim_in_1 * 3], vec_tilt + (size_t)row * dim_in_1 * 3, (size_t)dim_in_1 * 12);
                else for (int j = 0; j < dim_in_1; ++j) tl[((size_t)lr * dim_in_1 + j) * 3 + 2] = 1.f;
            }
            d_tilt[r] = (float*)pool_alloc(tl.size() * 4); d_azim[r] = (float*)pool_alloc((size_t)azim_num * 4);
            d_svf[r] = (float*)pool_alloc((size_t)my_rows * dim_in_1 * 4);
            if (!d_tilt[r] || !d_azim[r] || !d_svf[r]) { fail(r, "device allocation failed"); return; }
            if (cudaMemcpyAsync(d_tilt[r], tl.data(), tl.size() * 4, cudaMemcpyHostToDevice, streams[r]) != cudaSuccess ||
                cudaMemcpyAsync(d_azim[r], az.data(), (size_t)azim_num * 4, cudaMemcpyHostToDevice, streams[r]) != cudaSuccess ||
                cudaStreamSynchronize(streams[r]) != cudaSuccess) { fail(r, "tilt upload failed"); return; }   // tl dies at scope end
            // padding rows of the last block hold no horizon values: integrate whole rows of real cells only
            const int real_rows = [&] { int c = 0; for (int lr = 0; lr < my_rows; ++lr) if (((lr >> 2) * n + r) * 4 + (lr & 3) < dim_in_0) c = lr + 1; return c; }();
            if (launch_svf(0, d_azim[r], mine, d_tilt[r], (long long)real_rows * dim_in_1, azim_num, d_svf[r], streams[r])) { fail(r, hzb_last_error()); return; }
        }
        if (cudaStreamSynchronize(streams[r]) != cudaSuccess) { fail(r, std::string("horizon kernel failed: ") + cudaGetErrorString(cudaGetLastError())); return; }
        res[r].t_kernel = now_s() - t1;
    };
    // ---- phase 2: join the shards
    auto copy_out = [&](int r) {      // each shard's blocks straight to their places in the host arrays
        if (cudaSetDevice(r % ndev) != cudaSuccess) { fail(r, "cudaSetDevice failed"); return; }
        const int my_rows = hzb_shard_rows(dim_in_0, r, n);
        const float* mine = d_gather[r] + (device_gather ? (size_t)r * per_rows * row_elems : 0);
        for (int j = 0; j * 4 < my_rows; ++j) {
            const int row0 = (j * n + r) * 4;
            const int nrows = std::min(4, dim_in_0 - row0);
            if (nrows <= 0) break;
            float* dst = hori_buffer + (size_t)row0 * row_elems;
            const size_t bytes = (size_t)nrows * row_elems * sizeof(float);
            if (out_pinned) {   // page-locked output (hzb_host_alloc): plain DMA, every GPU over its own PCIe link at once
                if (cudaMemcpyAsync(dst, mine + (size_t)j * 4 * row_elems, bytes, cudaMemcpyDeviceToHost, streams[r]) != cudaSuccess) { fail(r, "copy failed"); return; }
            } else if (staged_d2h(dst, mine + (size_t)j * 4 * row_elems, bytes, streams[r])) {   // pageable: the process-wide staging engine
                fail(r, hzb_last_error()); return;
            }
            if (svf_buffer && cudaMemcpyAsync(svf_buffer + (size_t)row0 * dim_in_1, d_svf[r] + (size_t)j * 4 * dim_in_1, (size_t)nrows * dim_in_1 * 4,
                                              cudaMemcpyDeviceToHost, streams[r]) != cudaSuccess) { fail(r, "svf copy failed"); return; }
        }
        if (cudaStreamSynchronize(streams[r]) != cudaSuccess) fail(r, "copy failed");
    };

    auto run_all = [&](auto&& fn) {
        std::vector<std::thread> th;
        for (int r = 1; r < n; ++r) th.emplace_back(fn, r);
        fn(0);
        for (auto& t : th) t.join();
    };
    run_all(phase1);
    bool ok = true;
    for (int r = 0; r < n; ++r) ok = ok && res[r].rc == 0;
    std::string gather_err;
    if (ok && device_gather) {
        // one in-place all-gather: send = recv + rank * count
        Nccl& N = nccl();
        static std::mutex mu; std::lock_guard<std::mutex> l(mu);
        std::vector<ncclComm_t> comms((size_t)n); std::vector<int> devs((size_t)n);
        for (int r = 0; r < n; ++r) devs[r] = r;
        ncclResult_t e = ncclSuccess;
        if (!N.load(gather_err)) ok = false;
        else if ((e = N.CommInitAll(comms.data(), n, devs.data())) != ncclSuccess) { gather_err = std::string("ncclCommInitAll: ") + N.GetErrorString(e); ok = false; }
        else {
            const size_t count = (size_t)per_rows * row_elems;
            N.GroupStart();
            for (int r = 0; r < n; ++r) {
                cudaSetDevice(r);
                e = N.AllGather(d_gather[r] + (size_t)r * count, d_gather[r], count, ncclFloat, comms[r], streams[r]);
                if (e != ncclSuccess) break;
            }
            const ncclResult_t e2 = N.GroupEnd();
            if (e == ncclSuccess) e = e2;
            for (int r = 0; r < n; ++r) { cudaSetDevice(r); if (cudaStreamSynchronize(streams[r]) != cudaSuccess && e == ncclSuccess) e = ncclUnhandledCudaError; }
            for (int r = 0; r < n; ++r) N.CommDestroy(comms[r]);
            if (e != ncclSuccess) { gather_err = std::string("ncclAllGather: ") + N.GetErrorString(e); ok = false; }
        }
        if (ok) {   // GPU 0 holds every shard: domain order on the device, one copy to the host
            cudaSetDevice(0);
            float* d_full = (float*)pool_alloc((size_t)dim_in_0 * row_elems * sizeof(float));
            if (!d_full) { gather_err = hzb_last_error(); ok = false; }
            else {
                const int grid = sm_count() * 8;
                if (row_elems % 4 == 0) k_unpack_blocks<<<grid, 256, 0, streams[0]>>>((const float4*)d_gather[0], (float4*)d_full, dim_in_0, (long long)(row_elems / 4), n, per_rows);
                else k_unpack_blocks_f1<<<grid, 256, 0, streams[0]>>>(d_gather[0], d_full, dim_in_0, (long long)row_elems, n, per_rows);
                if (cudaStreamSynchronize(streams[0]) != cudaSuccess || staged_d2h(hori_buffer, d_full, (size_t)dim_in_0 * row_elems * sizeof(float), streams[0])) {
                    gather_err = "unpack / copy of the gathered array failed"; ok = false;
                }
                pool_free(d_full);
            }
            if (ok && svf_buffer) {   // the SVF shards are small: each GPU returns its own
                std::vector<std::thread> th;
                auto svf_only = [&](int r) {
                    cudaSetDevice(r % ndev);
                    const int my_rows = hzb_shard_rows(dim_in_0, r, n);
                    for (int j = 0; j * 4 < my_rows; ++j) {
                        const int row0 = (j * n + r) * 4, nrows = std::min(4, dim_in_0 - row0);
                        if (nrows <= 0) break;
                        cudaMemcpyAsync(svf_buffer + (size_t)row0 * dim_in_1, d_svf[r] + (size_t)j * 4 * dim_in_1, (size_t)nrows * dim_in_1 * 4, cudaMemcpyDeviceToHost, streams[r]);
                    }
                    if (cudaStreamSynchronize(streams[r]) != cudaSuccess) fail(r, "svf copy failed");
                };
                for (int r = 1; r < n; ++r) th.emplace_back(svf_only, r);
                svf_only(0);
                for (auto& t : th) t.join();
            }
        }
    } else if (ok) {
        run_all(copy_out);
    }
    for (int r = 0; r < n; ++r) ok = ok && res[r].rc == 0;

    // ---- statistics (summed over the shards) and clean-up
    hzb_stats total{};
    std::string first_err = gather_err;
    for (int r = 0; r < n; ++r) {
        cudaSetDevice(r % ndev);
        if (scenes[r]) {
            hzb_stats st{};
            if (hzb_scene_stats(scenes[r], &st) == 0) {
                total.rays += st.rays; total.node_visits += st.node_visits; total.prim_tests += st.prim_tests; total.units += st.units;
                total.fallback_packets += st.fallback_packets;
                total.segment_tasks += st.segment_tasks; total.segment_redos += st.segment_redos;
                total.num_prims = st.num_prims; total.num_nodes = st.num_nodes; total.bvh_bytes = st.bvh_bytes;
                total.t_h2d = std::max(total.t_h2d, st.t_h2d); total.t_build = std::max(total.t_build, st.t_build);
            } else if (ok) { ok = false; first_err = hzb_last_error(); }
        }
        total.t_trace = std::max(total.t_trace, res[r].t_kernel);
        if (res[r].rc && first_err.empty()) first_err = "shard " + std::to_string(r) + ": " + res[r].err;
        pool_free(d_in[2 * r]); pool_free(d_in[2 * r + 1]); pool_free(d_mask[r]); pool_free(d_gather[r]);
        pool_free(d_svf[r]); pool_free(d_tilt[r]); pool_free(d_azim[r]);
        if (streams[r]) cudaStreamDestroy(streams[r]);
        if (scenes[r]) hzb_scene_destroy(scenes[r]);
    }
    cudaSetDevice(dev0);
    total.t_total = now_s() - t_start;
    set_last_stats(total);
    if (!ok) { set_error(first_err.empty() ? "multi-GPU horizon failed" : first_err); return 1; }
    return 0;
}

}  // extern "C"
