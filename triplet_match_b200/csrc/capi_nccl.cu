// capi_nccl.cu — the one collective of the path: best-pose argmax over ranks, and the int64 sum
// all-reduce of the scene-sharded ICP.  NCCL is resolved at run time with dlopen so the library loads in
// processes that never go multi-GPU.
#include "capi_internal.cuh"

typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess_ = 0 };
enum { ncclUint8_ = 1, ncclUint32_ = 3, ncclInt64_ = 4, ncclUint64_ = 5 };  // ncclDataType_t
enum { ncclSum_ = 0, ncclMax_ = 2 };       // ncclRedOp_t
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mutex;
static int nccl_load() {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl.lib) return TM_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names)
        if ((lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!lib) return fail(TM_ERR_NCCL, std::string("dlopen(libnccl.so.2) failed: ") + dlerror());
    g_nccl.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(lib, "ncclCommDestroy");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(
        lib, "ncclAllReduce");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(lib, "ncclAllGather");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce || !g_nccl.AllGather)
        return fail(TM_ERR_NCCL, "libnccl is missing required symbols");
    g_nccl.lib = lib;
    return TM_OK;
}
#define NC(call)                                                                             \
    do {                                                                                     \
        int r_ = (call);                                                                     \
        if (r_ != 0)                                                                         \
            return fail(TM_ERR_NCCL, std::string(#call) + ": " +                            \
                                         (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?")); \
    } while (0)


extern "C" {

int tm_nccl_unique_id(uint8_t out[128]) {
    REQUIRE(out, "null out");
    TRY(nccl_load());
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(out, id.internal, 128);
    return TM_OK;
}
int tm_comm_create(tm_ctx* c, const uint8_t idb[128], int rank, int world, tm_comm** out) {
    REQUIRE(c && idb && out && world > 0 && rank >= 0 && rank < world, "tm_comm_create: bad argument");
    TRY(nccl_load());
    TRY(bind(c));
    ncclUniqueId id;
    memcpy(id.internal, idb, 128);
    tm_comm* cm = new tm_comm{c, nullptr, rank, world, DevBuf()};
    int r = g_nccl.CommInitRank(&cm->comm, world, id, rank);
    if (r != 0) {
        delete cm;
        return fail(TM_ERR_NCCL, std::string("ncclCommInitRank: ") +
                                     (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    }
    *out = cm;
    return TM_OK;
}
void tm_comm_destroy(tm_comm* cm) {
    if (!cm) return;
    cudaSetDevice(cm->ctx->device);
    if (cm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(cm->comm);
    cm->stage.release();
    delete cm;
}
// batched form for several queries over the same scene (BASELINE configs[3]: 16 models x one
// scene): ONE max all-reduce over the n packed keys and ONE sum all-reduce over the n x 72-byte
// (score, pose) records instead of 2n collectives
int tm_queries_allreduce_best(tm_query** qs, uint32_t n, tm_comm* cm) {
    REQUIRE(cm && (n == 0 || qs), "tm_queries_allreduce_best: null argument");
    if (!n) return TM_OK;
    tm_ctx* c = cm->ctx;
    for (uint32_t i = 0; i < n; ++i)
        REQUIRE(qs[i] && qs[i]->ran && qs[i]->s->ctx == c, "tm_queries_allreduce_best: bad query");
    TRY(bind(c));
    TRY(cm->stage.ensure((size_t)n * 8 + (size_t)n * 72));
    unsigned long long* keys = cm->stage.as<unsigned long long>();
    uint8_t* recs = reinterpret_cast<uint8_t*>(keys + n);
    for (uint32_t i = 0; i < n; ++i)
        CU(cudaMemcpyAsync(keys + i, &qs[i]->out.as<QueryOut>()->best, 8, cudaMemcpyDeviceToDevice, c->stream));
    NC(g_nccl.AllReduce(keys, keys, n, ncclUint64_, ncclMax_, cm->comm, c->stream));
    for (uint32_t i = 0; i < n; ++i) {
        tm_query* q = qs[i];
        QueryOut* out = q->out.as<QueryOut>();
        CU(cudaMemcpyAsync(&out->best, keys + i, 8, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemsetAsync(out->best_T16, 0, 64, c->stream));
        CU(cudaMemsetAsync(&out->best_score, 0, 8, c->stream));
        TRY(finalize_best(q));
        CU(cudaMemcpyAsync(recs + 72 * (size_t)i, &out->best_score, 72, cudaMemcpyDeviceToDevice, c->stream));
    }
    CU(cudaGetLastError());
    // A broadcast from a rank that is only known on the device, written as a byte-wise SUM: the global key is
    // unique, so exactly one rank (the owner of the winning hypothesis) exports a record and every other rank's
    // 72 bytes are zero (memsets above) — each byte position therefore has at most one non-zero addend and the
    // uint8 sum reproduces the owner's bytes exactly, without wrap-around.  Not a sum of floats.
    NC(g_nccl.AllReduce(recs, recs, (size_t)n * 72, ncclUint8_, ncclSum_, cm->comm, c->stream));
    for (uint32_t i = 0; i < n; ++i)
        CU(cudaMemcpyAsync(&qs[i]->out.as<QueryOut>()->best_score, recs + 72 * (size_t)i, 72,
                           cudaMemcpyDeviceToDevice, c->stream));
    return TM_OK;
}
}  // extern "C"

// every rank contributes `bytes` bytes; recv holds world x bytes in rank order
int comm_allgather_bytes(tm_comm* cm, const void* send, void* recv, size_t bytes, cudaStream_t st) {
    NC(g_nccl.AllGather(send, recv, bytes, ncclUint8_, cm->comm, st));
    return TM_OK;
}
int comm_allreduce_max_u32(tm_comm* cm, void* buf, size_t count, cudaStream_t st) {
    NC(g_nccl.AllReduce(buf, buf, count, ncclUint32_, ncclMax_, cm->comm, st));
    return TM_OK;
}
int comm_allreduce_sum_i64(tm_comm* cm, void* buf, size_t count, cudaStream_t st) {
    NC(g_nccl.AllReduce(buf, buf, count, ncclInt64_, ncclSum_, cm->comm, st));
    return TM_OK;
}

extern "C" {

int tm_query_allreduce_best(tm_query* q, tm_comm* cm) {
    REQUIRE(q && cm && q->ran, "tm_query_allreduce_best: bad argument");
    tm_ctx* c = q->s->ctx;
    REQUIRE(c == cm->ctx, "communicator belongs to another context");
    TRY(bind(c));
    QueryOut* out = q->out.as<QueryOut>();
    // max over ranks of (inliers << 32 | ~global id): 8 bytes, latency-bound
    NC(g_nccl.AllReduce(&out->best, &out->best, 1, ncclUint64_, ncclMax_, cm->comm, c->stream));
    // the owner re-exports the winning pose; everybody else contributes zeros
    CU(cudaMemsetAsync(out->best_T16, 0, 64, c->stream));
    CU(cudaMemsetAsync(&out->best_score, 0, 8, c->stream));
    TRY(finalize_best(q));
    CU(cudaGetLastError());
    // best_score (8 B) and best_T16 (64 B) are adjacent in QueryOut.  Byte-wise sum = broadcast from the one rank
    // that owns the (unique) winning key: all other ranks contribute zero bytes (see tm_queries_allreduce_best).
    NC(g_nccl.AllReduce(&out->best_score, &out->best_score, 72, ncclUint8_, ncclSum_, cm->comm,
                        c->stream));
    return TM_OK;
}

}  // extern "C"
