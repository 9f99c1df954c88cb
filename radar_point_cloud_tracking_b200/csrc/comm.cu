// Collective layer of the time-sharded path (SURVEY.md section 8 e / 8 b: rb_comm_init): NCCL, called by the library
// itself on the caller's stream. One process per GPU; a context holds ONE communicator, and every block slot of a rank
// (a context + a stream each, sharded.py) has its own, so the collectives of interleaved blocks do not queue behind each
// other inside one communicator.
//
// NCCL is not linked: the process already holds one copy (the one PyTorch loaded, nvidia/nccl/lib/libnccl.so.2) and this
// library binds to it at run time - dlopen with RTLD_NOLOAD first, the path in RB_NCCL_LIBRARY or the soname otherwise.
// Only the handful of entry points below are used; their signatures are the public NCCL 2.x API (nccl.h).
#include <dlfcn.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

namespace {

typedef struct { char internal[128]; } nccl_unique_id;          // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* nccl_comm;                                         // ncclComm_t
enum { NCCL_SUM = 0 };                                           // ncclRedOp_t
enum { NCCL_UINT8 = 1, NCCL_INT32 = 2, NCCL_INT64 = 4, NCCL_FLOAT64 = 8 };      // ncclDataType_t

struct NcclApi {
    int (*GetUniqueId)(nccl_unique_id*);
    int (*CommInitRank)(nccl_comm*, int, nccl_unique_id, int);
    int (*CommDestroy)(nccl_comm);
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm, cudaStream_t);
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t);
    int (*Send)(const void*, size_t, int, int, nccl_comm, cudaStream_t);
    int (*Recv)(void*, size_t, int, int, nccl_comm, cudaStream_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    const char* (*GetErrorString)(int);
    void* handle = nullptr;
    bool ok = false;
};

NcclApi g_nccl;
std::mutex g_nccl_mu;

int load_nccl() {
    std::lock_guard<std::mutex> lock(g_nccl_mu);
    if (g_nccl.ok) return RB_OK;
    const char* names[] = {getenv("RB_NCCL_LIBRARY"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names)
        if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL))) break;       // the copy the process already holds
    if (!h)
        for (const char* n : names)
            if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) {
        rb_set_error("rb_comm: cannot load NCCL (%s); set RB_NCCL_LIBRARY to libnccl.so.2", dlerror());
        return RB_ERR_NCCL;
    }
    g_nccl.handle = h;
#define RB_SYM(field, name)                                                        \
    *(void**)(&g_nccl.field) = dlsym(h, name);                                     \
    if (!g_nccl.field) { rb_set_error("rb_comm: NCCL has no symbol %s", name); return RB_ERR_NCCL; }
    RB_SYM(GetUniqueId, "ncclGetUniqueId")
    RB_SYM(CommInitRank, "ncclCommInitRank")
    RB_SYM(CommDestroy, "ncclCommDestroy")
    RB_SYM(AllGather, "ncclAllGather")
    RB_SYM(AllReduce, "ncclAllReduce")
    RB_SYM(Send, "ncclSend")
    RB_SYM(Recv, "ncclRecv")
    RB_SYM(GroupStart, "ncclGroupStart")
    RB_SYM(GroupEnd, "ncclGroupEnd")
    RB_SYM(GetErrorString, "ncclGetErrorString")
#undef RB_SYM
    g_nccl.ok = true;
    return RB_OK;
}

#define RB_NCCL(call)                                                                                  \
    do {                                                                                               \
        int _r = (call);                                                                               \
        if (_r != 0) {                                                                                 \
            rb_set_error("%s:%d: %s -> NCCL error %d (%s)", __FILE__, __LINE__, #call, _r, g_nccl.GetErrorString(_r)); \
            return RB_ERR_NCCL;                                                                        \
        }                                                                                              \
    } while (0)

}  // namespace

struct rb_comm {
    nccl_comm comm = nullptr;
    int rank = 0, world = 1;
};

void rb_comm_free(rb_ctx* ctx) {
    if (!ctx->comm) return;
    if (ctx->comm->comm && g_nccl.ok) g_nccl.CommDestroy(ctx->comm->comm);
    delete ctx->comm;
    ctx->comm = nullptr;
}

extern "C" int rb_comm_unique_id(uint8_t* out128) {
    RB_REQUIRE(out128, "NULL argument");
    RB_TRY(load_nccl());
    nccl_unique_id id;
    RB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, id.internal, sizeof id.internal);
    return RB_OK;
}

extern "C" int rb_comm_init(rb_ctx* ctx, const uint8_t* unique_id128, int rank, int world) {
    RB_REQUIRE(ctx && unique_id128, "NULL argument");
    RB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank / world");
    RB_TRY(load_nccl());
    rb_comm_free(ctx);
    int prev = -1;
    RB_CUDA(cudaGetDevice(&prev));
    RB_CUDA(cudaSetDevice(ctx->device));
    nccl_unique_id id;
    memcpy(id.internal, unique_id128, sizeof id.internal);
    rb_comm* c = new rb_comm();
    c->rank = rank;
    c->world = world;
    const int r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
    if (r != 0) {
        rb_set_error("rb_comm_init: ncclCommInitRank -> %d (%s)", r, g_nccl.GetErrorString(r));
        delete c;
        return RB_ERR_NCCL;
    }
    ctx->comm = c;
    return RB_OK;
}

extern "C" int rb_comm_destroy(rb_ctx* ctx) {
    RB_REQUIRE(ctx, "ctx is NULL");
    rb_comm_free(ctx);
    return RB_OK;
}

extern "C" int rb_comm_info(rb_ctx* ctx, int* rank, int* world) {
    RB_REQUIRE(ctx && ctx->comm, "no communicator (rb_comm_init)");
    if (rank) *rank = ctx->comm->rank;
    if (world) *world = ctx->comm->world;
    return RB_OK;
}

extern "C" int rb_comm_all_gather(rb_ctx* ctx, const void* send, void* recv, int64_t bytes_per_rank, void* stream) {
    RB_REQUIRE(ctx && ctx->comm, "no communicator (rb_comm_init)");
    RB_REQUIRE(bytes_per_rank >= 0 && (bytes_per_rank == 0 || (send && recv)), "bad arguments");
    if (bytes_per_rank == 0) return RB_OK;
    RB_NCCL(g_nccl.AllGather(send, recv, (size_t)bytes_per_rank, NCCL_UINT8, ctx->comm->comm, (cudaStream_t)stream));
    return RB_OK;
}

// in place: buf[count] of int32 (dtype 0) or float64 (dtype 1), summed over the ranks
extern "C" int rb_comm_all_reduce_sum(rb_ctx* ctx, void* buf, int64_t count, int dtype, void* stream) {
    RB_REQUIRE(ctx && ctx->comm, "no communicator (rb_comm_init)");
    RB_REQUIRE(count >= 0 && (count == 0 || buf) && (dtype == 0 || dtype == 1), "bad arguments");
    if (count == 0) return RB_OK;
    RB_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, dtype == 0 ? NCCL_INT32 : NCCL_FLOAT64, NCCL_SUM, ctx->comm->comm, (cudaStream_t)stream));
    return RB_OK;
}

// The land grids of a block in one call: int32 counts and float64 intensity sums (sums of integers: exact in any order).
extern "C" int rb_comm_all_reduce_grids(rb_ctx* ctx, int32_t* count, double* isum, int64_t cells, void* stream) {
    RB_REQUIRE(ctx && ctx->comm, "no communicator (rb_comm_init)");
    RB_REQUIRE(cells >= 0 && (cells == 0 || (count && isum)), "bad arguments");
    if (cells == 0 || ctx->comm->world == 1) return RB_OK;
    RB_NCCL(g_nccl.GroupStart());
    const int r1 = g_nccl.AllReduce(count, count, (size_t)cells, NCCL_INT32, NCCL_SUM, ctx->comm->comm, (cudaStream_t)stream);
    const int r2 = g_nccl.AllReduce(isum, isum, (size_t)cells, NCCL_FLOAT64, NCCL_SUM, ctx->comm->comm, (cudaStream_t)stream);
    RB_NCCL(g_nccl.GroupEnd());
    RB_NCCL(r1);
    RB_NCCL(r2);
    return RB_OK;
}

// Neighbour exchange of the time shards: up to `n_parts` arrays go to the left neighbour (rank - 1) and as many to the
// right one (rank + 1), and the matching arrays come back from them, all in ONE NCCL group. Sizes in bytes; a rank without
// that neighbour (or a zero size) skips the part. Pointers are device pointers.
extern "C" int rb_comm_exchange(rb_ctx* ctx, int n_parts, const void* const* to_left, const int64_t* to_left_bytes,
                                const void* const* to_right, const int64_t* to_right_bytes, void* const* from_left,
                                const int64_t* from_left_bytes, void* const* from_right, const int64_t* from_right_bytes, void* stream_) {
    RB_REQUIRE(ctx && ctx->comm, "no communicator (rb_comm_init)");
    RB_REQUIRE(n_parts >= 0 && n_parts <= 8, "bad n_parts");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int rank = ctx->comm->rank, world = ctx->comm->world;
    nccl_comm comm = ctx->comm->comm;
    int rc = 0;
    RB_NCCL(g_nccl.GroupStart());
    for (int k = 0; k < n_parts && rc == 0; ++k) {
        if (rank > 0) {
            if (to_left_bytes[k] > 0) rc = g_nccl.Send(to_left[k], (size_t)to_left_bytes[k], NCCL_UINT8, rank - 1, comm, stream);
            if (rc == 0 && from_left_bytes[k] > 0) rc = g_nccl.Recv(from_left[k], (size_t)from_left_bytes[k], NCCL_UINT8, rank - 1, comm, stream);
        }
        if (rc == 0 && rank < world - 1) {
            if (to_right_bytes[k] > 0) rc = g_nccl.Send(to_right[k], (size_t)to_right_bytes[k], NCCL_UINT8, rank + 1, comm, stream);
            if (rc == 0 && from_right_bytes[k] > 0) rc = g_nccl.Recv(from_right[k], (size_t)from_right_bytes[k], NCCL_UINT8, rank + 1, comm, stream);
        }
    }
    const int re = g_nccl.GroupEnd();
    RB_NCCL(rc);
    RB_NCCL(re);
    return RB_OK;
}
