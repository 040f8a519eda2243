// One-level vertical halo exchange of level-sharded fields over NCCL: the only
// collective of the hot path.  Vertical levels couple only through FD2(., dz, axis=0)
// (grids/cartesian_grid.py:250-251,275,321, grids/cylindrical_grid.py:388-389,419,466), so a rank that owns
// a contiguous block of levels needs ONE plane of each z-differentiated field from each neighbour.
//
// The library does not link against NCCL: it binds, at run time, to the libnccl.so.2 that is already loaded
// in the process (the one torch.distributed brought in -- two NCCL copies in one process must be avoided) or
// to the path handed to mt_nccl_load().  Only the stable C entry points below are used; planes travel as
// bytes (ncclChar), so no NCCL enum beyond 0 is relied on.
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace {

struct NcclUniqueId {
    char internal[128];
};
typedef void *NcclComm;
typedef int NcclResult;   // ncclSuccess == 0

struct Api {
    void *handle = nullptr;
    NcclResult (*GetUniqueId)(NcclUniqueId *) = nullptr;
    NcclResult (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    NcclResult (*CommDestroy)(NcclComm) = nullptr;
    NcclResult (*GroupStart)() = nullptr;
    NcclResult (*GroupEnd)() = nullptr;
    NcclResult (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    NcclResult (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(NcclResult) = nullptr;
    NcclResult (*GetVersion)(int *) = nullptr;
    bool ok() const { return handle && GetUniqueId && CommInitRank && CommDestroy && GroupStart && GroupEnd && Send && Recv; }
};

Api g_api;
std::mutex g_mu;

bool bind(void *h)
{
    Api a;
    a.handle = h;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
    a.GroupStart = (decltype(a.GroupStart))dlsym(h, "ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))dlsym(h, "ncclGroupEnd");
    a.Send = (decltype(a.Send))dlsym(h, "ncclSend");
    a.Recv = (decltype(a.Recv))dlsym(h, "ncclRecv");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
    a.GetVersion = (decltype(a.GetVersion))dlsym(h, "ncclGetVersion");
    if (!a.ok()) return false;
    g_api = a;
    return true;
}

// already bound, or the copy that is resident in the process (RTLD_NOLOAD: never loads a second NCCL)
bool ensure_api()
{
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_api.ok()) return true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (h && bind(h)) return true;
    return false;
}

int nccl_fail(NcclResult r, const char *what)
{
    mt::set_error("NCCL error %d (%s) in %s", (int)r, g_api.GetErrorString ? g_api.GetErrorString(r) : "?", what);
    return MT_ECUDA;
}

#define MT_NCCL(call)                                   \
    do {                                                \
        NcclResult _r = (call);                         \
        if (_r != 0) return nccl_fail(_r, #call);       \
    } while (0)

}  // namespace

extern "C" {

int mt_nccl_load(const char *path)
{
    if (ensure_api()) return MT_OK;
    MT_REQUIRE(path && *path, "mt_nccl_load: no libnccl.so.2 is loaded in this process and no path was given");
    std::lock_guard<std::mutex> lock(g_mu);
    void *h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    MT_REQUIRE(h, "mt_nccl_load: dlopen(%s) failed: %s", path, dlerror());
    MT_REQUIRE(bind(h), "mt_nccl_load: %s does not export the NCCL point-to-point API", path);
    return MT_OK;
}

int mt_nccl_version(void)
{
    if (!ensure_api() || !g_api.GetVersion) return 0;
    int v = 0;
    return g_api.GetVersion(&v) == 0 ? v : 0;
}

int mt_nccl_unique_id(void *id128)
{
    MT_REQUIRE(id128, "mt_nccl_unique_id: null buffer");
    MT_REQUIRE(ensure_api(), "mt_nccl_unique_id: NCCL is not loaded (call mt_nccl_load first)");
    MT_NCCL(g_api.GetUniqueId(reinterpret_cast<NcclUniqueId *>(id128)));
    return MT_OK;
}

int mt_nccl_comm_init(void **comm, int nranks, int rank, const void *id128)
{
    MT_REQUIRE(comm && id128 && nranks >= 1 && rank >= 0 && rank < nranks, "mt_nccl_comm_init: bad arguments");
    MT_REQUIRE(ensure_api(), "mt_nccl_comm_init: NCCL is not loaded (call mt_nccl_load first)");
    NcclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NcclComm c = nullptr;
    MT_NCCL(g_api.CommInitRank(&c, nranks, id, rank));
    *comm = c;
    return MT_OK;
}

int mt_nccl_comm_destroy(void *comm)
{
    if (!comm) return MT_OK;
    MT_REQUIRE(ensure_api(), "mt_nccl_comm_destroy: NCCL is not loaded");
    MT_NCCL(g_api.CommDestroy(comm));
    return MT_OK;
}

int mt_halo_exchange(void *comm, double *const *fields, int nfields, int64_t plane_elems, int64_t nz_local,
                     int halo_lo, int halo_hi, int rank_lo, int rank_hi, mt_stream_t stream)
{
    MT_REQUIRE(comm && fields && nfields >= 1 && nfields <= 64 && plane_elems > 0, "mt_halo_exchange: bad arguments");
    MT_REQUIRE((halo_lo == 0 || halo_lo == 1) && (halo_hi == 0 || halo_hi == 1) && nz_local > halo_lo + halo_hi,
               "mt_halo_exchange: halo_lo / halo_hi must be 0 or 1 and leave at least one owned level");
    MT_REQUIRE((!halo_lo || rank_lo >= 0) && (!halo_hi || rank_hi >= 0), "mt_halo_exchange: neighbour rank missing");
    MT_REQUIRE(ensure_api(), "mt_halo_exchange: NCCL is not loaded");
    if (!halo_lo && !halo_hi) return MT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t bytes = (size_t)plane_elems * sizeof(double);
    MT_NCCL(g_api.GroupStart());
    NcclResult r = 0;
    for (int f = 0; f < nfields && r == 0; ++f) {
        double *base = fields[f];
        if (!base) {
            r = -1;
            break;
        }
        if (halo_lo) {   // my first owned level goes down, the neighbour's last owned level comes up
            r = g_api.Send(base + (int64_t)halo_lo * plane_elems, bytes, 0 /* ncclChar */, rank_lo, comm, s);
            if (r == 0) r = g_api.Recv(base, bytes, 0, rank_lo, comm, s);
        }
        if (halo_hi && r == 0) {
            r = g_api.Send(base + (nz_local - 1 - halo_hi) * plane_elems, bytes, 0, rank_hi, comm, s);
            if (r == 0) r = g_api.Recv(base + (nz_local - 1) * plane_elems, bytes, 0, rank_hi, comm, s);
        }
    }
    const NcclResult rg = g_api.GroupEnd();
    if (r == -1) {
        mt::set_error("mt_halo_exchange: null field pointer");
        return MT_EINVAL;
    }
    if (r != 0) return nccl_fail(r, "ncclSend / ncclRecv");
    if (rg != 0) return nccl_fail(rg, "ncclGroupEnd");
    return MT_OK;
}

}  // extern "C"
