// slmgs_comm.cu -- the ONE collective of the sharded batch path behind the C ABI:
// an NCCL all-gather of the final near-field phases, plus the small all-reduce the pixel-sharded compressed
// hologram needs.  NCCL is loaded with dlopen (libnccl.so.2: the copy already in the process if there is one),
// so libslmgs.so has no link-time dependency on it and no torch / torch.distributed is involved.
// The unique id is created here (rank 0) and distributed by the caller (slmsuite_b200/comm.py: a TCP rendezvous on
// MASTER_ADDR / MASTER_PORT); every rank then calls slmgs_comm_create.
#include "../../include/slmgs.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>

#ifndef SLMGS_EMULATE
#include <cuda_runtime.h>
#include <dlfcn.h>

extern "C" void* slmgs_phase_device_ptr(slmgs_ctx* c);
extern "C" void* slmgs_stream(slmgs_ctx* c);
extern "C" int slmgs_sync(slmgs_ctx* c);

namespace {

typedef struct { char internal[128]; } nccl_uid;   // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* nccl_comm;
enum { NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };

struct NcclApi {
    void* handle;
    int (*GetUniqueId)(nccl_uid*);
    int (*CommInitRank)(nccl_comm*, int, nccl_uid, int);
    int (*CommDestroy)(nccl_comm);
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm, cudaStream_t);
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t);
    const char* (*GetErrorString)(int);
    int (*GetVersion)(int*);
};
NcclApi g_nccl = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
std::string g_comm_error;

int comm_fail(int code, const std::string& msg) {
    g_comm_error = msg;
    return code;
}

int load_nccl() {
    if (g_nccl.handle) return 0;
    const char* env = getenv("SLMGS_NCCL_LIB");
    void* h = nullptr;
    if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);  // the copy the process already uses
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return comm_fail(SLMGS_ERR_NCCL, std::string("cannot load libnccl.so.2 (set SLMGS_NCCL_LIB): ") + dlerror());
#define SYM(field, name)                                                             \
    *(void**)(&g_nccl.field) = dlsym(h, name);                                       \
    if (!g_nccl.field) return comm_fail(SLMGS_ERR_NCCL, std::string("libnccl has no ") + name);
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllGather, "ncclAllGather")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GetErrorString, "ncclGetErrorString")
    SYM(GetVersion, "ncclGetVersion")
#undef SYM
    g_nccl.handle = h;
    return 0;
}

int nccl_check(int r, const char* what) {
    if (r == 0) return 0;
    return comm_fail(SLMGS_ERR_NCCL, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error"));
}

}  // namespace

struct slmgs_comm {
    nccl_comm comm;
    int rank, world, device;
    float* gather;        // device buffer of the last all-gather (grow-only)
    size_t gather_bytes;
    float* pad;           // device staging of a short last shard
    size_t pad_bytes;
};

extern "C" const char* slmgs_comm_last_error(void) { return g_comm_error.c_str(); }

extern "C" int slmgs_comm_nccl_version(void) {
    if (load_nccl()) return -1;
    int v = 0;
    g_nccl.GetVersion(&v);
    return v;
}

extern "C" int slmgs_comm_unique_id(unsigned char* out128) {
    if (!out128) return comm_fail(SLMGS_ERR_INVALID, "out is NULL");
    int e = load_nccl();
    if (e) return e;
    nccl_uid id;
    if ((e = nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId"))) return e;
    memcpy(out128, &id, 128);
    return SLMGS_OK;
}

extern "C" int slmgs_comm_create(slmgs_comm** out, const unsigned char* id128, int rank, int world, int device) {
    if (!out || !id128 || world < 1 || rank < 0 || rank >= world) return comm_fail(SLMGS_ERR_INVALID, "bad communicator arguments");
    *out = nullptr;
    int e = load_nccl();
    if (e) return e;
    if (cudaSetDevice(device) != cudaSuccess) return comm_fail(SLMGS_ERR_CUDA, "cudaSetDevice failed");
    nccl_uid id;
    memcpy(&id, id128, 128);
    slmgs_comm* c = new slmgs_comm();
    c->comm = nullptr; c->rank = rank; c->world = world; c->device = device;
    c->gather = nullptr; c->gather_bytes = 0; c->pad = nullptr; c->pad_bytes = 0;
    if ((e = nccl_check(g_nccl.CommInitRank(&c->comm, world, id, rank), "ncclCommInitRank"))) {
        delete c;
        return e;
    }
    *out = c;
    return SLMGS_OK;
}

extern "C" int slmgs_comm_destroy(slmgs_comm* c) {
    if (!c) return SLMGS_OK;
    cudaSetDevice(c->device);
    if (c->gather) cudaFree(c->gather);
    if (c->pad) cudaFree(c->pad);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    delete c;
    return SLMGS_OK;
}

static int reserve(float** p, size_t* have, size_t bytes) {
    if (bytes <= *have) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *have = 0;
    if (cudaMalloc((void**)p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return comm_fail(SLMGS_ERR_OOM, "device allocation for the all-gather failed");
    }
    *have = bytes;
    return 0;
}

// All-gather of the final phases: every rank contributes `per_rank` holograms of `elems` floats
// (ctx holds n_local <= per_rank of them: a short or empty last shard is zero-padded; ctx may be NULL when
// n_local == 0) and receives all world * per_rank of them.  The collective runs on the context's stream behind the
// loop's kernels; `out_host` (may be NULL) receives the gathered array, `*out_dev` (may be NULL) its device address
// (owned by the communicator, valid until the next call).  `ms` (may be NULL): device time of the collective alone.
extern "C" int slmgs_allgather_phase(slmgs_ctx* ctx, slmgs_comm* c, int n_local, int per_rank, long long elems,
                                     float* out_host, void** out_dev, float* ms) {
    if (!c || per_rank < 1 || n_local < 0 || n_local > per_rank || elems < 1) return comm_fail(SLMGS_ERR_INVALID, "bad all-gather arguments");
    if (n_local > 0 && !ctx) return comm_fail(SLMGS_ERR_INVALID, "a context is needed to contribute holograms");
    if (cudaSetDevice(c->device) != cudaSuccess) return comm_fail(SLMGS_ERR_CUDA, "cudaSetDevice failed");
    const size_t shard = (size_t)per_rank * (size_t)elems * sizeof(float);
    int e;
    if ((e = reserve(&c->gather, &c->gather_bytes, shard * c->world))) return e;
    cudaStream_t stream = ctx ? (cudaStream_t)slmgs_stream(ctx) : (cudaStream_t)0;
    const float* src = ctx ? (const float*)slmgs_phase_device_ptr(ctx) : nullptr;
    if (n_local < per_rank) {
        if ((e = reserve(&c->pad, &c->pad_bytes, shard))) return e;
        cudaMemsetAsync(c->pad, 0, shard, stream);
        if (n_local > 0)
            cudaMemcpyAsync(c->pad, src, (size_t)n_local * elems * sizeof(float), cudaMemcpyDeviceToDevice, stream);
        src = c->pad;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ms) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, stream);
    }
    if ((e = nccl_check(g_nccl.AllGather(src, c->gather, (size_t)per_rank * elems, NCCL_FLOAT32, c->comm, stream), "ncclAllGather")))
        return e;
    if (ms) cudaEventRecord(e1, stream);
    if (out_host) {
        if (cudaMemcpyAsync(out_host, c->gather, shard * c->world, cudaMemcpyDeviceToHost, stream) != cudaSuccess)
            return comm_fail(SLMGS_ERR_CUDA, "download of the gathered phases failed");
    }
    if (cudaStreamSynchronize(stream) != cudaSuccess) return comm_fail(SLMGS_ERR_CUDA, cudaGetErrorString(cudaGetLastError()));
    if (ms) {
        cudaEventElapsedTime(ms, e0, e1);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    if (out_dev) *out_dev = c->gather;
    return SLMGS_OK;
}

// In-place sum over ranks of `count` float64 values at a device address, ordered on `stream` (no host sync)
extern "C" int slmgs_comm_allreduce_f64(slmgs_comm* c, void* dev_ptr, long long count, void* stream) {
    if (!c || !dev_ptr || count < 1) return comm_fail(SLMGS_ERR_INVALID, "bad all-reduce arguments");
    if (cudaSetDevice(c->device) != cudaSuccess) return comm_fail(SLMGS_ERR_CUDA, "cudaSetDevice failed");
    return nccl_check(g_nccl.AllReduce(dev_ptr, dev_ptr, (size_t)count, NCCL_FLOAT64, NCCL_SUM, c->comm, (cudaStream_t)stream),
                      "ncclAllReduce");
}

#else  // ---- host emulation build: no NCCL; the Python layer uses its TCP communicator instead -------------------

struct slmgs_comm { int unused; };
static const char* kNoNccl = "the host-emulation build has no NCCL";
extern "C" const char* slmgs_comm_last_error(void) { return kNoNccl; }
extern "C" int slmgs_comm_nccl_version(void) { return -1; }
extern "C" int slmgs_comm_unique_id(unsigned char*) { return SLMGS_ERR_NCCL; }
extern "C" int slmgs_comm_create(slmgs_comm** out, const unsigned char*, int, int, int) {
    if (out) *out = nullptr;
    return SLMGS_ERR_NCCL;
}
extern "C" int slmgs_comm_destroy(slmgs_comm*) { return SLMGS_OK; }
extern "C" int slmgs_allgather_phase(slmgs_ctx*, slmgs_comm*, int, int, long long, float*, void**, float*) { return SLMGS_ERR_NCCL; }
extern "C" int slmgs_comm_allreduce_f64(slmgs_comm*, void*, long long, void*) { return SLMGS_ERR_NCCL; }
#endif
