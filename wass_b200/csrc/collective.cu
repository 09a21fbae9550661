// The one cross-frame data flow of the reference, as a collective inside the C ABI.
//
// The reference runs one wass_stereo process per frame; the only thing that crosses frames is 4 doubles per frame:
// plane.txt -> output/planes.txt (cli/wasscli/wasscli.py:341-343) -> np.nanmean over the rows
// (gridding/wassgridsurface/wassgridsurface.py:672-678).  With one rank per GPU and frames sharded over ranks that is ONE
// all-reduce(sum) over 5 doubles -- sums of a, b, c, d over the frames whose plane is not NaN, and their count -- on NCCL
// over NVLink / NVSwitch; 40 bytes, latency-bound.
//
// NCCL is bound at run time (dlopen libnccl.so.2), so libwassgpu.so has no link-time dependency on it: inside a process
// that already carries a libnccl (torch's bundled copy in bench.py and the tests) the loader hands back that one, in the
// C++ hosts it is the system library.  The communicator is created by the caller (wsg_nccl_comm_create below, or its own
// ncclCommInitRank) and passed as an opaque pointer, as SURVEY 8b sketches.
#include "handle.cuh"
#include <cmath>
#include <dlfcn.h>
#include <mutex>
#include <nccl.h>

namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};

NcclApi& nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) { api.err = std::string("libnccl.so.2 not found: ") + dlerror(); return; }
        auto sym = [&](const char* n) { void* p = dlsym(api.lib, n); if (!p) api.err = std::string("missing NCCL symbol ") + n; return p; };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    });
    return api;
}

thread_local std::string g_coll_err;
int fail(const std::string& what) { g_coll_err = what; return WSG_ERR_CUDA; }
}  // namespace

extern "C" {

const char* wsg_collective_last_error(void) { return g_coll_err.c_str(); }

int wsg_nccl_unique_id(unsigned char id[WSG_NCCL_UNIQUE_ID_BYTES])
{
    static_assert(sizeof(ncclUniqueId) == WSG_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
    if (!id) return WSG_ERR_INVALID_ARG;
    NcclApi& n = nccl();
    if (!n.err.empty()) return fail(n.err);
    ncclUniqueId u;
    const ncclResult_t r = n.GetUniqueId(&u);
    if (r != ncclSuccess) return fail(std::string("ncclGetUniqueId: ") + n.GetErrorString(r));
    memcpy(id, &u, sizeof(u));
    return WSG_OK;
}

int wsg_nccl_comm_create(int device, int nranks, int rank, const unsigned char id[WSG_NCCL_UNIQUE_ID_BYTES], void** comm)
{
    if (!id || !comm || nranks <= 0 || rank < 0 || rank >= nranks) return WSG_ERR_INVALID_ARG;
    NcclApi& n = nccl();
    if (!n.err.empty()) return fail(n.err);
    if (cudaSetDevice(device) != cudaSuccess) return fail("cudaSetDevice failed");
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t c = nullptr;
    const ncclResult_t r = n.CommInitRank(&c, nranks, u, rank);
    if (r != ncclSuccess) return fail(std::string("ncclCommInitRank: ") + n.GetErrorString(r));
    // NCCL connects the ranks lazily, inside the first collective (~2 s on an NVSwitch box): pay that here, once, so that
    // the plane reduction at the end of a sequence costs what 40 bytes over NVLink cost
    double* d = nullptr;
    if (cudaMalloc(&d, 16 * sizeof(double)) == cudaSuccess) {
        cudaMemset(d, 0, 16 * sizeof(double));
        const ncclResult_t w = n.AllReduce(d, d + 8, 5, ncclDouble, ncclSum, c, (cudaStream_t)0);
        cudaStreamSynchronize((cudaStream_t)0);
        cudaFree(d);
        if (w != ncclSuccess) { n.CommDestroy(c); return fail(std::string("ncclAllReduce (warm-up): ") + n.GetErrorString(w)); }
    }
    *comm = c;
    return WSG_OK;
}

void wsg_nccl_comm_destroy(void* comm)
{
    NcclApi& n = nccl();
    if (comm && n.err.empty()) n.CommDestroy((ncclComm_t)comm);
}

// acc[5]: this rank's NaN-aware sums (wsg_plane_mean_accumulate).  One ncclAllReduce(sum) over 5 doubles on the handle's
// stream; mean[4] = the sequence mean plane (NaN when no rank had a valid plane), *frames = planes that entered it.
int wsg_plane_allreduce(wsg_handle* h, void* nccl_comm, const double acc[5], double mean[4], long long* frames)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (!nccl_comm || !acc || !mean) { h->err = "null communicator or buffer"; return WSG_ERR_INVALID_ARG; }
    NcclApi& n = nccl();
    if (!n.err.empty()) { h->err = n.err; return WSG_ERR_CUDA; }
    CK(h, cudaSetDevice(h->device));
    int rc;
    if ((rc = ensure(h, h->m_small, 4096))) return rc;
    double* d = (double*)h->m_small.p;
    CK(h, cudaMemcpyAsync(d, acc, 5 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    const ncclResult_t r = n.AllReduce(d, d + 8, 5, ncclDouble, ncclSum, (ncclComm_t)nccl_comm, h->stream);
    if (r != ncclSuccess) { h->err = std::string("ncclAllReduce: ") + n.GetErrorString(r); return WSG_ERR_CUDA; }
    double tot[5];
    CK(h, cudaMemcpyAsync(tot, d + 8, 5 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    wsg_plane_mean_finish(tot, mean);
    if (frames) *frames = (long long)tot[4];
    return WSG_OK;
}

// Every rank's per-frame planes, in rank order, on every rank: what a driver needs to write planes.txt in FRAME order
// (the reference's file is in completion order, wasscli.py:343).  planes: n_local x 4 doubles; all: nranks x n_local x 4.
int wsg_plane_allgather(wsg_handle* h, void* nccl_comm, int nranks, const double* planes, int n_local, double* all)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (!nccl_comm || !planes || !all || n_local <= 0 || nranks <= 0) { h->err = "bad argument"; return WSG_ERR_INVALID_ARG; }
    NcclApi& n = nccl();
    if (!n.err.empty()) { h->err = n.err; return WSG_ERR_CUDA; }
    CK(h, cudaSetDevice(h->device));
    const size_t cnt = (size_t)n_local * 4;
    int rc;
    if ((rc = ensure(h, h->m_scratch, (size_t)(nranks + 1) * cnt * sizeof(double)))) return rc;
    double* d = (double*)h->m_scratch.p;
    CK(h, cudaMemcpyAsync(d, planes, cnt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    const ncclResult_t r = n.AllGather(d, d + cnt, cnt, ncclDouble, (ncclComm_t)nccl_comm, h->stream);
    if (r != ncclSuccess) { h->err = std::string("ncclAllGather: ") + n.GetErrorString(r); return WSG_ERR_CUDA; }
    CK(h, cudaMemcpyAsync(all, d + cnt, (size_t)nranks * cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return WSG_OK;
}

}  // extern "C"
