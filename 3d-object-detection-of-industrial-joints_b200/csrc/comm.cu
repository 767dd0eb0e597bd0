// comm.cu — the path's one exchange step, inside the library: an NCCL communicator owned by the context, a
// broadcast of one scene to every rank and the gather of the per-rank correspondence lists.
//
// The reference's callers are single-process C++ programs (SHOT.cpp:204, 6Dpose.cpp:216); with this file they reach
// the multi-GPU path through the C ABI alone (b200_comm_*, b200_gather_correspondences,
// b200_register_scene_shot_sharded) — one context per GPU, driven by one host thread or process each.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already in the process if a host framework brought
// its own, else the system one), so the library loads — and every single-GPU entry point works — on a machine
// without NCCL.  Collectives are stream-ordered on the context's stream; message sizes are small (the scene is
// 16 MB, a correspondence list ~0.3 MB) and ride NVLink / NVSwitch.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>
#include <string>

#include "common.cuh"

namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      api.error = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "");
      return;
    }
    auto sym = [&](const char *s) {
      void *p = dlsym(api.handle, s);
      if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + s;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return &api;
}

int nccl_fail(b200_ctx *ctx, ncclResult_t r, const char *what) {
  NcclApi *a = nccl_api();
  std::string msg = std::string("NCCL error in ") + what + ": " + (a->GetErrorString ? a->GetErrorString(r) : "?");
  return ctx->fail(B200_ERR_CUDA, msg.c_str());
}

#define B200_NCCL(ctx, expr)                                 \
  do {                                                       \
    ncclResult_t r__ = (expr);                               \
    if (r__ != ncclSuccess) return nccl_fail(ctx, r__, #expr); \
  } while (0)

// rank r's list (count[r] records at in + r * cap) -> out, ranks in order; total -> *n_out
__global__ void concat_lists_kernel(const b200_corr *__restrict__ in, const int *__restrict__ counts, int world, int cap,
                                    b200_corr *__restrict__ out, int out_cap, int *__restrict__ n_out) {
  int base = 0;
  for (int r = 0; r < world; ++r) {
    const int c = min(max(counts[r], 0), cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c; i += gridDim.x * blockDim.x)
      if (base + i < out_cap) out[base + i] = in[(size_t)r * cap + i];
    base += c;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && n_out) *n_out = min(base, out_cap);
}

// index_match of a slab's correspondences: position within the slab -> position within the scene's keypoints
__global__ void offset_scene_index_kernel(b200_corr *__restrict__ c, const int *__restrict__ n, int cap, int offset) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < min(*n, cap)) c[i].index_match += offset;
}

}  // namespace

int comm_unique_id(void *id128, size_t bytes, std::string *err) {
  NcclApi *a = nccl_api();
  if (!a->error.empty() || !a->GetUniqueId) {
    *err = a->error.empty() ? "NCCL unavailable" : a->error;
    return B200_ERR_NODEVICE;
  }
  if (!id128 || bytes < sizeof(ncclUniqueId)) {
    *err = "comm_unique_id: buffer of at least 128 bytes required";
    return B200_ERR_INVALID;
  }
  ncclUniqueId id;
  ncclResult_t r = a->GetUniqueId(&id);
  if (r != ncclSuccess) {
    *err = std::string("ncclGetUniqueId: ") + a->GetErrorString(r);
    return B200_ERR_CUDA;
  }
  memcpy(id128, &id, sizeof(id));
  return B200_OK;
}

int comm_init(b200_ctx *ctx, const void *id128, int rank, int world) {
  NcclApi *a = nccl_api();
  if (!a->error.empty()) return ctx->fail(B200_ERR_NODEVICE, a->error.c_str());
  if (!id128 || world < 1 || rank < 0 || rank >= world) return ctx->fail(B200_ERR_INVALID, "comm_init: bad arguments");
  if (ctx->nccl_comm) return ctx->fail(B200_ERR_INVALID, "comm_init: the context already has a communicator");
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  B200_NCCL(ctx, a->CommInitRank(&comm, world, id, rank));
  ctx->nccl_comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  return B200_OK;
}

int comm_destroy(b200_ctx *ctx) {
  if (!ctx->nccl_comm) return B200_OK;
  NcclApi *a = nccl_api();
  cudaStreamSynchronize(ctx->stream);
  a->CommDestroy(static_cast<ncclComm_t>(ctx->nccl_comm));
  ctx->nccl_comm = nullptr;
  ctx->comm_rank = 0;
  ctx->comm_world = 1;
  return B200_OK;
}

int comm_broadcast(b200_ctx *ctx, void *d_buf, size_t bytes, int root) {
  if (ctx->comm_world == 1 || bytes == 0) return B200_OK;
  NcclApi *a = nccl_api();
  B200_NCCL(ctx, a->Broadcast(d_buf, d_buf, bytes, ncclChar, root, static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream));
  return B200_OK;
}

int comm_allgather(b200_ctx *ctx, const void *d_send, void *d_recv, size_t bytes_per_rank) {
  if (bytes_per_rank == 0) return B200_OK;
  if (ctx->comm_world == 1) {
    if (d_send != d_recv)
      B200_CUDA(ctx, cudaMemcpyAsync(d_recv, d_send, bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
    return B200_OK;
  }
  NcclApi *a = nccl_api();
  B200_NCCL(ctx, a->AllGather(d_send, d_recv, bytes_per_rank, ncclChar, static_cast<ncclComm_t>(ctx->nccl_comm),
                              ctx->stream));
  return B200_OK;
}

// Every rank contributes *d_count (<= cap) records; afterwards every rank holds all lists, d_gathered[r * cap ...] with
// d_counts[r] valid records each — the padded all-gather of SURVEY.md 8(e): one collective for the counts, one for
// the records (latency-bound: tens of KB to ~1 MB per rank).
int dev_gather_correspondences(b200_ctx *ctx, const b200_corr *d_corrs, const int *d_count, int cap,
                               b200_corr *d_gathered, int *d_counts) {
  if (cap < 0) return ctx->fail(B200_ERR_INVALID, "gather_correspondences: negative capacity");
  B200_TRY(comm_allgather(ctx, d_count, d_counts, sizeof(int)));
  B200_TRY(comm_allgather(ctx, d_corrs, d_gathered, (size_t)cap * sizeof(b200_corr)));
  return B200_OK;
}

int dev_concat_lists(b200_ctx *ctx, const b200_corr *d_gathered, const int *d_counts, int world, int cap,
                     b200_corr *d_out, int out_cap, int *d_n_out) {
  concat_lists_kernel<<<std::max(1, std::min(ctx->sm_count, ceil_div(cap, 256))), 256, 0, ctx->stream>>>(
      d_gathered, d_counts, world, cap, d_out, out_cap, d_n_out);
  B200_LAUNCHED(ctx);
  return B200_OK;
}

int dev_offset_scene_index(b200_ctx *ctx, b200_corr *d_corrs, const int *d_n, int cap, int offset) {
  if (cap <= 0 || offset == 0) return B200_OK;
  offset_scene_index_kernel<<<ceil_div(cap, 256), 256, 0, ctx->stream>>>(d_corrs, d_n, cap, offset);
  B200_LAUNCHED(ctx);
  return B200_OK;
}
