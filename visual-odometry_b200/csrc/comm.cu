// comm.cu — the multi-GPU entry points of the C ABI (BASELINE config 4): one process,
// one vo_nn handle per GPU, the map replicated, a query batch sharded in contiguous blocks, and the
// int32 match indices of all shards gathered with ONE ncclAllGather over NVLink — the only
// collective on the path (each query's answer depends on the replicated map alone).
//
// What it replaces: nothing in the reference (which is single-threaded, src/apps/vo_complete.cpp:12-49
// answers one query at a time); it is the non-Python way to run the sharded sweep that bench.py
// drives through torch.distributed (visual-odometry_b200/sharding.py).
//
// NCCL is bound at run time (dlopen "libnccl.so.2"), so libvo_b200.so has no link-time dependency
// on it and single-GPU users never load it; inside a process that already has NCCL (torch) the
// loader hands back that copy.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <vector>

#include "nn.cuh"

namespace vo {
namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

const NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (lib) {
      api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
      api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
      api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(lib, "ncclAllGather"));
      api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(dlsym(lib, "ncclBroadcast"));
      api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(lib, "ncclGroupStart"));
      api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
      api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
      if (api.CommInitAll && api.CommDestroy && api.AllGather && api.Broadcast && api.GroupStart &&
          api.GroupEnd && api.GetErrorString)
        api.lib = lib;
    }
  }
  return api.lib ? &api : nullptr;
}

#define VO_NCCL(api, expr)                                                                  \
  do {                                                                                      \
    ncclResult_t _r = (expr);                                                               \
    if (_r != ncclSuccess) {                                                                \
      ::vo::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, (api)->GetErrorString(_r)); \
      return VO_ERR_CUDA;                                                                   \
    }                                                                                       \
  } while (0)

}  // namespace
}  // namespace vo

using namespace vo;

struct vo_comm_s {
  int n = 0;
  std::vector<ncclComm_t> comms;
  std::vector<vo_nn_t> nn;          // one handle per GPU (device d = rank d)
  std::vector<DevBuf> send, recv;   // per GPU: its shard's indices / everybody's
  bool have_map = false;
};

extern "C" {

int vo_comm_init_all(vo_comm_t* out, int n_gpus) {
  VO_REQUIRE(out != nullptr, VO_ERR_ARG, "null handle pointer");
  int ndev = 0;
  VO_CUDA(cudaGetDeviceCount(&ndev));
  VO_REQUIRE(n_gpus >= 1 && n_gpus <= ndev, VO_ERR_ARG, "n_gpus must be between 1 and the device count");
  const NcclApi* api = nullptr;
  if (n_gpus > 1) {
    api = nccl_api();
    VO_REQUIRE(api != nullptr, VO_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
  }
  vo_comm_s* c = new vo_comm_s();
  c->n = n_gpus;
  c->nn.assign((size_t)n_gpus, nullptr);
  c->send.resize((size_t)n_gpus);
  c->recv.resize((size_t)n_gpus);
  for (int d = 0; d < n_gpus; ++d) {
    const int rc = vo_nn_create(&c->nn[(size_t)d], d);
    if (rc) {
      vo_comm_destroy(c);
      return rc;
    }
  }
  if (n_gpus > 1) {
    c->comms.assign((size_t)n_gpus, nullptr);
    std::vector<int> devs((size_t)n_gpus);
    for (int d = 0; d < n_gpus; ++d) devs[(size_t)d] = d;
    const ncclResult_t r = api->CommInitAll(c->comms.data(), n_gpus, devs.data());
    if (r != ncclSuccess) {
      set_error("ncclCommInitAll(%d) -> %s", n_gpus, api->GetErrorString(r));
      c->comms.clear();
      vo_comm_destroy(c);
      return VO_ERR_CUDA;
    }
  }
  *out = c;
  return VO_OK;
}

int vo_comm_destroy(vo_comm_t c) {
  if (!c) return VO_OK;
  const NcclApi* api = c->comms.empty() ? nullptr : nccl_api();
  for (size_t d = 0; d < c->nn.size(); ++d) {
    if (c->nn[d]) {
      DeviceGuard g((int)d);
      cudaStreamSynchronize(c->nn[d]->stream);
      c->send[d].release();
      c->recv[d].release();
    }
  }
  if (api)
    for (ncclComm_t cm : c->comms)
      if (cm) api->CommDestroy(cm);
  for (vo_nn_t h : c->nn) vo_nn_destroy(h);
  delete c;
  return VO_OK;
}

int vo_comm_size(vo_comm_t c) { return c ? c->n : VO_ERR_ARG; }

// Replicates the map: GPU 0 takes the host rows through the staging ring and re-packs them (as
// vo_nn_set_map); the PACKED buffers (FP32 rows, f16 tiles, max norm) then travel to the other GPUs
// with ncclBroadcast over NVLink instead of N-1 more trips across PCIe.
int vo_nn_set_map_replicated(vo_comm_t c, const float* rows_host, int64_t n_rows, int row_stride,
                             int skip_cols) {
  VO_REQUIRE(c != nullptr, VO_ERR_ARG, "null communicator");
  c->have_map = false;
  int rc = vo_nn_set_map(c->nn[0], rows_host, n_rows, row_stride, skip_cols);
  if (rc) return rc;
  vo_nn_s* h0 = c->nn[0];
  if (c->n > 1 && (!h0->fast || n_rows == 0)) {
    // general dimension: no packed form to broadcast, every GPU takes the host rows itself
    for (int d = 1; d < c->n; ++d)
      if ((rc = vo_nn_set_map(c->nn[(size_t)d], rows_host, n_rows, row_stride, skip_cols))) return rc;
  } else if (c->n > 1) {
    const NcclApi* api = nccl_api();
    const size_t packed_bytes = (size_t)h0->n_tiles * NN_TM * NN_ROW_BYTES;
    const size_t tiles16_bytes = h0->tc_ready ? (size_t)h0->n_tiles16 * 8192 : 0;
    for (int d = 1; d < c->n; ++d) {
      vo_nn_s* h = c->nn[(size_t)d];
      DeviceGuard g(d);
      // the same bookkeeping as set_map, with the buffers filled by the broadcast below
      h->have_mm_max = false;
      h->n_rows = n_rows;
      h->row_stride = row_stride;
      h->skip = skip_cols;
      h->dim = h0->dim;
      h->fast = true;
      h->n_tiles = h0->n_tiles;
      h->n_tiles16 = h0->n_tiles16;
      h->tc_ready = h0->tc_ready;
      if ((rc = h->packed.reserve(packed_bytes))) return rc;
      if ((rc = h->scalars.reserve(64))) return rc;
      if (tiles16_bytes && (rc = h->tiles16.reserve(tiles16_bytes))) return rc;
      if (tiles16_bytes && (rc = h->tc_stats.reserve(128))) return rc;
      h->rows_dev = h->packed.as<float>();
      h->map_stride = NN_ROW_BYTES / (int)sizeof(float);
      h->map_skip = 0;
    }
    VO_NCCL(api, api->GroupStart());
    for (int d = 0; d < c->n; ++d) {
      vo_nn_s* h = c->nn[(size_t)d];
      VO_NCCL(api, api->Broadcast(h->packed.p, h->packed.p, packed_bytes, ncclUint8, 0, c->comms[(size_t)d], h->stream));
    }
    VO_NCCL(api, api->GroupEnd());
    if (tiles16_bytes) {
      VO_NCCL(api, api->GroupStart());
      for (int d = 0; d < c->n; ++d) {
        vo_nn_s* h = c->nn[(size_t)d];
        VO_NCCL(api, api->Broadcast(h->tiles16.p, h->tiles16.p, tiles16_bytes, ncclUint8, 0, c->comms[(size_t)d], h->stream));
      }
      VO_NCCL(api, api->GroupEnd());
    }
    VO_NCCL(api, api->GroupStart());
    for (int d = 0; d < c->n; ++d) {
      vo_nn_s* h = c->nn[(size_t)d];
      VO_NCCL(api, api->Broadcast(h->scalars.p, h->scalars.p, 64, ncclUint8, 0, c->comms[(size_t)d], h->stream));
    }
    VO_NCCL(api, api->GroupEnd());
    for (int d = 0; d < c->n; ++d) {
      DeviceGuard g(d);
      VO_CUDA(cudaStreamSynchronize(c->nn[(size_t)d]->stream));
    }
  }
  c->have_map = true;
  return VO_OK;
}

// bruteForceBestMatch for a batch, sharded: GPU g answers queries [g*Q/G, (g+1)*Q/G); one
// ncclAllGather(int32) leaves all Q indices on every GPU; GPU 0's copy goes back to the host.
int vo_nn_best_match_sharded(vo_comm_t c, const float* queries_host, int64_t n_queries, int query_stride,
                             float norm, int32_t* best_idx_host) {
  VO_REQUIRE(c != nullptr, VO_ERR_ARG, "null communicator");
  VO_REQUIRE(c->have_map, VO_ERR_STATE, "vo_nn_set_map_replicated not called");
  VO_REQUIRE(n_queries >= 0 && (n_queries == 0 || (queries_host && best_idx_host)), VO_ERR_ARG, "bad queries");
  if (n_queries == 0) return VO_OK;
  const int G = c->n;
  if (G == 1) return vo_nn_best_match(c->nn[0], queries_host, n_queries, query_stride, norm, best_idx_host, nullptr);
  const NcclApi* api = nccl_api();
  const int64_t per = (n_queries + G - 1) / G;  // padded shard length: the collective wants equal counts
  int rc;
  for (int d = 0; d < G; ++d) {
    vo_nn_s* h = c->nn[(size_t)d];
    DeviceGuard g(d);
    const int64_t lo = (int64_t)d * n_queries / G, hi = (int64_t)(d + 1) * n_queries / G, nq = hi - lo;
    if ((rc = c->send[(size_t)d].reserve((size_t)per * sizeof(int32_t)))) return rc;
    if ((rc = c->recv[(size_t)d].reserve((size_t)per * G * sizeof(int32_t)))) return rc;
    if ((rc = h->q_stage.reserve((size_t)std::max<int64_t>(nq, 1) * query_stride * sizeof(float)))) return rc;
    VO_CUDA(cudaMemsetAsync(c->send[(size_t)d].p, 0xFF, (size_t)per * sizeof(int32_t), h->stream));  // padding = -1
    if (nq > 0) {
      rc = stage_h2d(d, h->q_stage.p, queries_host + lo * (int64_t)query_stride,
                     (size_t)nq * query_stride * sizeof(float), h->stream);
      if (rc) return rc;
      rc = vo_nn_best_match_device(h, h->q_stage.as<float>(), nq, query_stride, norm,
                                   c->send[(size_t)d].as<int32_t>(), nullptr);
      if (rc) return rc;
    }
  }
  VO_NCCL(api, api->GroupStart());
  for (int d = 0; d < G; ++d)
    VO_NCCL(api, api->AllGather(c->send[(size_t)d].p, c->recv[(size_t)d].p, (size_t)per, ncclInt32,
                                c->comms[(size_t)d], c->nn[(size_t)d]->stream));
  VO_NCCL(api, api->GroupEnd());
  // un-pad GPU 0's gathered copy into the caller's array
  {
    DeviceGuard g(0);
    cudaStream_t s0 = c->nn[0]->stream;
    for (int d = 0; d < G; ++d) {
      const int64_t lo = (int64_t)d * n_queries / G, hi = (int64_t)(d + 1) * n_queries / G;
      if (hi > lo)
        VO_CUDA(cudaMemcpyAsync(best_idx_host + lo, c->recv[0].as<int32_t>() + (int64_t)d * per,
                                (size_t)(hi - lo) * sizeof(int32_t), cudaMemcpyDeviceToHost, s0));
    }
    VO_CUDA(cudaStreamSynchronize(s0));
  }
  for (int d = 1; d < G; ++d) {
    DeviceGuard g(d);
    VO_CUDA(cudaStreamSynchronize(c->nn[(size_t)d]->stream));
  }
  return VO_OK;
}

}  // extern "C"
