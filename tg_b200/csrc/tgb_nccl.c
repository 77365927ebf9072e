/*
 * tgb_nccl.c -- the one real exchange step of the path: merging per-GPU visibility buffers with
 * ncclAllReduce(ncclUint64, ncclMin) over NVLink. The reference has no
 * multi-GPU path at all (one VkDevice, tgvk_core.c:4382-4392); the 64-bit min it resolves hits with
 * (visibility.frag:206) is associative and commutative, so the all-reduce is bit-identical to a
 * single-GPU frame over the union of the shards.
 *
 * NCCL is bound with dlopen so that libtgb200.so loads on hosts without it and so that, inside a
 * process that already loaded torch's bundled libnccl.so.2, the very same library instance is used.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "tgb_internal.h"

typedef struct { char internal[128]; } tgb_nccl_unique_id;
typedef void* tgb_nccl_comm;

/* values from nccl.h (ncclDataType_t / ncclRedOp_t), stable across NCCL 2.x */
enum { TGB_NCCL_UINT8 = 1, TGB_NCCL_UINT32 = 3, TGB_NCCL_UINT64 = 5 };
enum { TGB_NCCL_SUM = 0, TGB_NCCL_MAX = 2, TGB_NCCL_MIN = 3 };

static struct
{
    void* p_lib;
    int (*GetUniqueId)(tgb_nccl_unique_id*);
    int (*CommInitRank)(tgb_nccl_comm*, int, tgb_nccl_unique_id, int);
    int (*CommDestroy)(tgb_nccl_comm);
    int (*AllReduce)(const void*, void*, size_t, int, int, tgb_nccl_comm, void*);
    int (*AllGather)(const void*, void*, size_t, int, tgb_nccl_comm, void*);
    int (*ReduceScatter)(const void*, void*, size_t, int, int, tgb_nccl_comm, void*);
    const char* (*GetErrorString)(int);
} tgb__nccl;

static b32 tgb__nccl_load(void)
{
    if (tgb__nccl.p_lib) return TG_TRUE;
    const char* p_names[] = { getenv("TGB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
    for (int i = 0; i < 3 && !tgb__nccl.p_lib; i++)
    {
        if (p_names[i] && p_names[i][0]) tgb__nccl.p_lib = dlopen(p_names[i], RTLD_NOW | RTLD_GLOBAL);
    }
    if (!tgb__nccl.p_lib) { tgb_set_error("NCCL not found (dlopen libnccl.so.2): %s", dlerror()); return TG_FALSE; }
    *(void**)&tgb__nccl.GetUniqueId    = dlsym(tgb__nccl.p_lib, "ncclGetUniqueId");
    *(void**)&tgb__nccl.CommInitRank   = dlsym(tgb__nccl.p_lib, "ncclCommInitRank");
    *(void**)&tgb__nccl.CommDestroy    = dlsym(tgb__nccl.p_lib, "ncclCommDestroy");
    *(void**)&tgb__nccl.AllReduce      = dlsym(tgb__nccl.p_lib, "ncclAllReduce");
    *(void**)&tgb__nccl.AllGather      = dlsym(tgb__nccl.p_lib, "ncclAllGather");
    *(void**)&tgb__nccl.ReduceScatter  = dlsym(tgb__nccl.p_lib, "ncclReduceScatter");
    *(void**)&tgb__nccl.GetErrorString = dlsym(tgb__nccl.p_lib, "ncclGetErrorString");
    if (!tgb__nccl.GetUniqueId || !tgb__nccl.CommInitRank || !tgb__nccl.CommDestroy || !tgb__nccl.AllReduce || !tgb__nccl.AllGather || !tgb__nccl.ReduceScatter ||
        !tgb__nccl.GetErrorString)
    {
        tgb_set_error("NCCL library lacks a required symbol");
        dlclose(tgb__nccl.p_lib);
        memset(&tgb__nccl, 0, sizeof(tgb__nccl));
        return TG_FALSE;
    }
    return TG_TRUE;
}

#define TGB_NCCL(call) do { int r__ = (call); if (r__ != 0) { tgb_set_error("%s -> %s", #call, tgb__nccl.GetErrorString(r__)); return 0; } } while (0)

b32 tgbn_unique_id(u8* p_out_128)
{
    memset(p_out_128, 0, 128);
    if (!tgb__nccl_load()) return TG_FALSE;
    tgb_nccl_unique_id id;
    TGB_NCCL(tgb__nccl.GetUniqueId(&id));
    memcpy(p_out_128, &id, 128);
    return TG_TRUE;
}

void* tgbn_init(const u8* p_unique_id_128, u32 rank, u32 n_ranks)
{
    if (!tgb__nccl_load()) return NULL;
    tgb_nccl_unique_id id;
    memcpy(&id, p_unique_id_128, 128);
    tgb_nccl_comm comm = NULL;
    TGB_NCCL(tgb__nccl.CommInitRank(&comm, (int)n_ranks, id, (int)rank));
    return comm;
}

void tgbn_destroy(void* p_comm)
{
    if (p_comm && tgb__nccl.p_lib) tgb__nccl.CommDestroy((tgb_nccl_comm)p_comm);
}

b32 tgbn_allreduce_min_u64(void* p_comm, void* p_device_buffer, u64 count, void* p_stream)
{
    if (!p_comm || !tgb__nccl.p_lib) { tgb_set_error("tgbn_allreduce_min_u64: no communicator"); return TG_FALSE; }
    TGB_NCCL(tgb__nccl.AllReduce(p_device_buffer, p_device_buffer, (size_t)count, TGB_NCCL_UINT64, TGB_NCCL_MIN, (tgb_nccl_comm)p_comm, p_stream));
    return TG_TRUE;
}

/* per-node arrival counts of the sharded SVO build: sum over ranks, in place */
b32 tgbn_allreduce_sum_u32(void* p_comm, void* p_device_buffer, u64 count, void* p_stream)
{
    if (!p_comm || !tgb__nccl.p_lib) { tgb_set_error("tgbn_allreduce_sum_u32: no communicator"); return TG_FALSE; }
    TGB_NCCL(tgb__nccl.AllReduce(p_device_buffer, p_device_buffer, (size_t)count, TGB_NCCL_UINT32, TGB_NCCL_SUM, (tgb_nccl_comm)p_comm, p_stream));
    return TG_TRUE;
}

/* every rank contributes n_bytes, receives n_ranks * n_bytes in rank order (object records, partial SVO leaves, radiance tiles) */
b32 tgbn_allgather_bytes(void* p_comm, const void* p_send, void* p_recv, u64 n_bytes, void* p_stream)
{
    if (!p_comm || !tgb__nccl.p_lib) { tgb_set_error("tgbn_allgather_bytes: no communicator"); return TG_FALSE; }
    TGB_NCCL(tgb__nccl.AllGather(p_send, p_recv, (size_t)n_bytes, TGB_NCCL_UINT8, (tgb_nccl_comm)p_comm, p_stream));
    return TG_TRUE;
}

/* owner-resolved material words: element-wise max over ranks, rank r receives elements [r * count_per_rank, (r + 1) * count_per_rank) */
b32 tgbn_reducescatter_max_u64(void* p_comm, const void* p_send, void* p_recv, u64 count_per_rank, void* p_stream)
{
    if (!p_comm || !tgb__nccl.p_lib) { tgb_set_error("tgbn_reducescatter_max_u64: no communicator"); return TG_FALSE; }
    TGB_NCCL(tgb__nccl.ReduceScatter(p_send, p_recv, (size_t)count_per_rank, TGB_NCCL_UINT64, TGB_NCCL_MAX, (tgb_nccl_comm)p_comm, p_stream));
    return TG_TRUE;
}
