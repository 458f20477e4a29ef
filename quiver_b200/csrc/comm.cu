// comm.cu — several GPUs of one box behind the C ABI (include/quiver_gpu.h: qg_comm_*, qg_group_*).
//
// The exchange of a row-sharded search is tiny (q * k * 8 bytes per rank) and happens once per batch
// after the scan, so it is a plain NCCL all-gather over NVLink / NVSwitch followed by the merge kernel —
// there is no GEMM-then-collective to fuse tile by tile. NCCL is loaded with dlopen on
// first use, so libquivergpu.so itself carries no NCCL dependency; a process that already loaded a
// libnccl.so.2 (PyTorch ships one) shares that copy.
//
// Reference shape this replaces: one Go process, one index, BatchSearch fanning a batch out to
// goroutines (pkg/hybrid/hybrid_index.go:677-811, pkg/core/db.go:707-845).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/quiver_gpu.h"
#include "common.cuh"

namespace qg {

// ---- NCCL entry points, resolved at run time ----------------------------------------------------
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};
static NcclApi g_nccl;
static std::once_flag g_nccl_once;

static int nccl_load() {
  std::call_once(g_nccl_once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) {
      g_nccl.error = std::string("NCCL is not available (dlopen libnccl.so.2): ") + dlerror();
      return;
    }
    auto sym = [&](const char* s) -> void* {
      void* p = dlsym(g_nccl.handle, s);
      if (!p && g_nccl.error.empty()) g_nccl.error = std::string("libnccl lacks ") + s;
      return p;
    };
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
    g_nccl.CommInitAll = reinterpret_cast<decltype(g_nccl.CommInitAll)>(sym("ncclCommInitAll"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
    g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(sym("ncclAllGather"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
  });
  if (!g_nccl.error.empty()) return fail(QG_ERR_UNSUPPORTED, g_nccl.error);
  return 0;
}

#define QG_NCCL_OK(expr)                                                                              \
  do {                                                                                                \
    ncclResult_t _r = (expr);                                                                         \
    if (_r != ncclSuccess) return ::qg::fail(QG_ERR_CUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(_r)); \
  } while (0)

struct DBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(need, 256));
    if (e != cudaSuccess) return fail(QG_ERR_OOM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    bytes = std::max<size_t>(need, 256);
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};

// Result block of one rank in the query-split layout: rows (int64) | distances (float32) | counts (int32)
// for `per` queries — rows first so that every part is aligned.
__host__ __device__ inline size_t block_bytes(int per, int k) {
  return ((size_t)per * ((size_t)k * 12 + 4) + 15) & ~(size_t)15;  // whole 16-byte units: the next rank's rows stay aligned
}

// gathered blocks (rank-major) -> [q x k] outputs; the last blocks may be partial or empty
__global__ void __launch_bounds__(256) scatter_blocks_kernel(const unsigned char* __restrict__ blocks, int world, int per,
                                                             int q, int k, float* __restrict__ out_dist,
                                                             long long* __restrict__ out_row, int* __restrict__ out_count) {
  const size_t bb = block_bytes(per, k);
  const long long total = (long long)q * k;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int qi = (int)(i / k), j = (int)(i - (long long)qi * k);
    const int r = qi / per, l = qi - r * per;
    const unsigned char* b = blocks + (size_t)r * bb;
    out_row[i] = reinterpret_cast<const long long*>(b)[(size_t)l * k + j];
    out_dist[i] = reinterpret_cast<const float*>(b + (size_t)per * k * 8)[(size_t)l * k + j];
    if (j == 0) out_count[qi] = reinterpret_cast<const int*>(b + (size_t)per * k * 12)[l];
  }
}

}  // namespace qg

using namespace qg;

struct qg_comm {
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0, device = 0;
  DBuf keys, gathered, block, blocks;
};

static int comm_alloc(qg_comm** out, ncclComm_t comm, int world, int rank, int device) {
  qg_comm* c = new qg_comm();
  c->comm = comm;
  c->world = world;
  c->rank = rank;
  c->device = device;
  *out = c;
  return 0;
}

// shared by the collective entry points and the group workers
static int comm_search_rows(qg_comm* c, qg_index* shard, const void* d_queries, int q, int dim, int k, qg_filter* filter,
                            int64_t row_base, void* d_out_dist, void* d_out_row, void* d_out_count, cudaStream_t st,
                            bool merge_here) {
  QG_CUDA_OK(cudaSetDevice(c->device));
  const size_t nkeys = (size_t)q * k;
  if (int rc = c->keys.ensure(nkeys * 8)) return rc;
  if (int rc = c->gathered.ensure(nkeys * 8 * c->world)) return rc;
  if (int rc = qg_search_shard_keys_device(shard, d_queries, q, dim, k, filter, row_base, c->keys.p, st)) return rc;
  if (c->world > 1) {
    QG_NCCL_OK(g_nccl.AllGather(c->keys.p, c->gathered.p, nkeys, ncclUint64, c->comm, st));
  } else {
    QG_CUDA_OK(cudaMemcpyAsync(c->gathered.p, c->keys.p, nkeys * 8, cudaMemcpyDeviceToDevice, st));
  }
  if (!merge_here) return 0;
  return qg_merge_shard_keys_device(c->device, c->gathered.p, c->world, q, k, d_out_dist, d_out_row, d_out_count, st);
}

extern "C" {

int qg_comm_unique_id(void* id_out) {
  if (!id_out) return fail(QG_ERR_INVALID, "id_out is null");
  if (int rc = nccl_load()) return rc;
  static_assert(sizeof(ncclUniqueId) <= QG_COMM_ID_BYTES, "NCCL unique id larger than QG_COMM_ID_BYTES");
  ncclUniqueId id;
  QG_NCCL_OK(g_nccl.GetUniqueId(&id));
  std::memset(id_out, 0, QG_COMM_ID_BYTES);
  std::memcpy(id_out, &id, sizeof(id));
  return 0;
}

int qg_comm_create_rank(const void* id, int world, int rank, int device, qg_comm** out) {
  if (!id || !out) return fail(QG_ERR_INVALID, "null argument");
  *out = nullptr;
  if (world <= 0 || rank < 0 || rank >= world) return fail(QG_ERR_INVALID, "bad world / rank");
  if (int rc = nccl_load()) return rc;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
    return fail(QG_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
  if (device < 0 || device >= n) return fail(QG_ERR_INVALID, "device ordinal out of range");
  QG_CUDA_OK(cudaSetDevice(device));
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  ncclComm_t comm = nullptr;
  QG_NCCL_OK(g_nccl.CommInitRank(&comm, world, uid, rank));
  return comm_alloc(out, comm, world, rank, device);
}

int qg_comm_destroy(qg_comm* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  c->keys.release();
  c->gathered.release();
  c->block.release();
  c->blocks.release();
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  delete c;
  return 0;
}

int qg_comm_world(const qg_comm* c) { return c ? c->world : 0; }
int qg_comm_rank(const qg_comm* c) { return c ? c->rank : -1; }

int qg_comm_search_rows_device(qg_comm* c, qg_index* shard, const void* d_queries, int q, int dim, int k,
                               qg_filter* filter, int64_t row_base, void* d_out_dist, void* d_out_row,
                               void* d_out_count, void* stream) {
  if (!c || !shard) return fail(QG_ERR_INVALID, "null handle");
  if (q < 0) return fail(QG_ERR_INVALID, "negative query count");
  if (q == 0) return 0;
  if (!d_out_dist || !d_out_row || !d_out_count) return fail(QG_ERR_INVALID, "null device buffer");
  return comm_search_rows(c, shard, d_queries, q, dim, k, filter, row_base, d_out_dist, d_out_row, d_out_count,
                          (cudaStream_t)stream, true);
}

int qg_comm_search_queries_device(qg_comm* c, qg_index* replica, const void* d_queries, int q, int dim, int k,
                                  qg_filter* filter, int gather, void* d_out_dist, void* d_out_row, void* d_out_count,
                                  void* stream) {
  if (!c || !replica) return fail(QG_ERR_INVALID, "null handle");
  if (q < 0) return fail(QG_ERR_INVALID, "negative query count");
  if (q == 0) return 0;
  if (k <= 0) return fail(QG_ERR_K, "k must be positive");
  if (!d_queries || !d_out_dist || !d_out_row || !d_out_count) return fail(QG_ERR_INVALID, "null device buffer");
  QG_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int per = (q + c->world - 1) / c->world;
  const int q0 = std::min(q, c->rank * per), q1 = std::min(q, q0 + per);
  const float* qp = static_cast<const float*>(d_queries) + (size_t)q0 * dim;
  if (!gather || c->world == 1) {
    if (q1 > q0)
      return qg_search_batch_device(replica, qp, q1 - q0, dim, k, filter, nullptr,
                                    static_cast<float*>(d_out_dist) + (size_t)q0 * k, nullptr,
                                    static_cast<long long*>(d_out_row) + (size_t)q0 * k,
                                    static_cast<int*>(d_out_count) + q0, st);
    return 0;
  }
  const size_t bb = block_bytes(per, k);
  if (int rc = c->block.ensure(bb)) return rc;
  if (int rc = c->blocks.ensure(bb * c->world)) return rc;
  unsigned char* b = static_cast<unsigned char*>(c->block.p);
  if (q1 > q0) {
    if (int rc = qg_search_batch_device(replica, qp, q1 - q0, dim, k, filter, nullptr, b + (size_t)per * k * 8, nullptr, b,
                                        b + (size_t)per * k * 12, st))
      return rc;
  }
  QG_NCCL_OK(g_nccl.AllGather(c->block.p, c->blocks.p, bb, ncclUint8, c->comm, st));
  const long long total = (long long)q * k;
  const int grid = (int)std::min<long long>(148 * 4, (total + 255) / 256);
  scatter_blocks_kernel<<<grid, 256, 0, st>>>(static_cast<const unsigned char*>(c->blocks.p), c->world, per, q, k,
                                              static_cast<float*>(d_out_dist), static_cast<long long*>(d_out_row),
                                              static_cast<int*>(d_out_count));
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"

// ================================================================================================
// One host process driving all GPUs: one worker thread, stream, index and communicator per device.
// ================================================================================================
struct GroupJob {
  int kind = 0;  // 1 upload host rows, 2 upload synthetic, 3 search, 4 quit
  const float* rows = nullptr;
  int64_t n = 0;
  int synth_kind = 0;
  uint64_t seed = 0;
  const float* queries = nullptr;
  int q = 0, dim = 0, k = 0;
  float* out_dist = nullptr;
  int64_t* out_row = nullptr;
  int* out_count = nullptr;
};

struct GroupWorker {
  int device = 0, rank = 0;
  qg_index* index = nullptr;
  qg_comm* comm = nullptr;
  cudaStream_t stream = nullptr;
  int64_t row_base = 0;
  std::thread thread;
  // per-call device / pinned staging
  DBuf d_q, d_dist, d_row, d_cnt;
  void* h_q = nullptr;
  void* h_out = nullptr;
  size_t h_q_bytes = 0, h_out_bytes = 0;
  int rc = 0;
  std::string error;
};

struct qg_group {
  int n = 0, dim = 0, metric = 0;
  int layout = QG_LAYOUT_ROWS;
  int64_t rows = 0;
  bool loaded = false;
  std::vector<std::unique_ptr<GroupWorker>> workers;
  std::mutex mu;               // one call at a time drives the group
  std::mutex job_mu;
  std::condition_variable job_cv, done_cv;
  GroupJob job;
  uint64_t generation = 0;     // bumped for every job
  int pending = 0;             // workers still busy with the current job
};

static int pin_ensure(void** p, size_t* have, size_t need) {
  if (need <= *have) return 0;
  if (*p) cudaFreeHost(*p);
  *p = nullptr;
  *have = 0;
  cudaError_t e = cudaHostAlloc(p, std::max<size_t>(need, 4096), cudaHostAllocDefault);
  if (e != cudaSuccess) return fail(QG_ERR_OOM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
  *have = std::max<size_t>(need, 4096);
  return 0;
}

static int worker_search(qg_group* g, GroupWorker* w, const GroupJob& j) {
  QG_CUDA_OK(cudaSetDevice(w->device));
  const int q = j.q, k = j.k, dim = j.dim;
  const bool by_queries = g->layout == QG_LAYOUT_QUERIES;
  const int per = (q + g->n - 1) / g->n;
  const int q0 = by_queries ? std::min(q, w->rank * per) : 0;
  const int q1 = by_queries ? std::min(q, q0 + per) : q;
  const int nq = q1 - q0;
  // queries: this worker's block (replicas) or the whole batch (row shards), through pinned staging
  const size_t qbytes = (size_t)nq * dim * 4;
  if (int rc = pin_ensure(&w->h_q, &w->h_q_bytes, qbytes)) return rc;
  if (int rc = w->d_q.ensure(qbytes)) return rc;
  if (nq > 0) {
    std::memcpy(w->h_q, j.queries + (size_t)q0 * dim, qbytes);
    QG_CUDA_OK(cudaMemcpyAsync(w->d_q.p, w->h_q, qbytes, cudaMemcpyHostToDevice, w->stream));
  }
  const size_t on = (size_t)std::max(nq, 1) * k;
  if (int rc = w->d_dist.ensure(on * 4)) return rc;
  if (int rc = w->d_row.ensure(on * 8)) return rc;
  if (int rc = w->d_cnt.ensure((size_t)std::max(nq, 1) * 4)) return rc;
  const bool emit = by_queries ? nq > 0 : w->rank == 0;  // who copies results out
  if (by_queries) {
    // replicas: no collective; a query the device API could not certify is rare and the host-buffer entry
    // point handles it, so the block goes through qg_search_batch's device path and is checked below
    if (nq > 0) {
      if (int rc = qg_search_batch_device(w->index, w->d_q.p, nq, dim, k, nullptr, nullptr, w->d_dist.p, nullptr,
                                          w->d_row.p, w->d_cnt.p, w->stream))
        return rc;
    }
  } else {
    if (int rc = comm_search_rows(w->comm, w->index, w->d_q.p, q, dim, k, nullptr, w->row_base, w->d_dist.p, w->d_row.p,
                                  w->d_cnt.p, w->stream, /*merge_here=*/w->rank == 0))
      return rc;
  }
  if (emit) {
    const size_t ob = on * 12 + (size_t)nq * 4;
    if (int rc = pin_ensure(&w->h_out, &w->h_out_bytes, ob)) return rc;
    char* ho = static_cast<char*>(w->h_out);
    QG_CUDA_OK(cudaMemcpyAsync(ho, w->d_row.p, on * 8, cudaMemcpyDeviceToHost, w->stream));
    QG_CUDA_OK(cudaMemcpyAsync(ho + on * 8, w->d_dist.p, on * 4, cudaMemcpyDeviceToHost, w->stream));
    QG_CUDA_OK(cudaMemcpyAsync(ho + on * 12, w->d_cnt.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, w->stream));
  }
  QG_CUDA_OK(cudaStreamSynchronize(w->stream));
  if (emit && nq > 0) {
    const char* ho = static_cast<const char*>(w->h_out);
    const int* cnt = reinterpret_cast<const int*>(ho + on * 12);
    std::memcpy(j.out_row + (size_t)q0 * k, ho, (size_t)nq * k * 8);
    std::memcpy(j.out_dist + (size_t)q0 * k, ho + on * 8, (size_t)nq * k * 4);
    std::memcpy(j.out_count + q0, cnt, (size_t)nq * 4);
    if (by_queries) {
      // uncertified queries of the asynchronous device API (count -1): repeat through the host entry point
      for (int i = 0; i < nq; ++i) {
        if (cnt[i] >= 0) continue;
        if (int rc = qg_search_batch(w->index, j.queries + (size_t)(q0 + i) * dim, 1, dim, k, nullptr, nullptr,
                                     j.out_dist + (size_t)(q0 + i) * k, nullptr, j.out_row + (size_t)(q0 + i) * k,
                                     j.out_count + q0 + i))
          return rc;
      }
    }
  }
  return 0;
}

static int worker_upload(qg_group* g, GroupWorker* w, const GroupJob& j) {
  QG_CUDA_OK(cudaSetDevice(w->device));
  int64_t row0 = 0, n = j.n;
  if (g->layout == QG_LAYOUT_ROWS) {
    const int64_t per = (j.n + g->n - 1) / g->n;
    row0 = std::min<int64_t>(j.n, (int64_t)w->rank * per);
    n = std::min<int64_t>(j.n, row0 + per) - row0;
  }
  w->row_base = row0;
  if (n <= 0) return 0;
  if (j.kind == 1) return qg_index_upload(w->index, j.rows + (size_t)row0 * g->dim, n, nullptr);
  return qg_index_upload_synthetic(w->index, j.synth_kind, j.seed, row0, n, nullptr);
}

static void worker_main(qg_group* g, GroupWorker* w) {
  uint64_t seen = 0;
  for (;;) {
    GroupJob j;
    {
      std::unique_lock<std::mutex> lk(g->job_mu);
      g->job_cv.wait(lk, [&] { return g->generation != seen; });
      seen = g->generation;
      j = g->job;
    }
    int rc = 0;
    if (j.kind == 1 || j.kind == 2) rc = worker_upload(g, w, j);
    else if (j.kind == 3) rc = worker_search(g, w, j);
    w->rc = rc;
    w->error = rc ? qg_last_error() : "";
    {
      std::lock_guard<std::mutex> lk(g->job_mu);
      if (--g->pending == 0) g->done_cv.notify_all();
    }
    if (j.kind == 4) return;
  }
}

// hand a job to every worker and wait; the first failure becomes the caller's error
static int group_run(qg_group* g, const GroupJob& j) {
  {
    std::lock_guard<std::mutex> lk(g->job_mu);
    g->job = j;
    g->pending = g->n;
    ++g->generation;
  }
  g->job_cv.notify_all();
  {
    std::unique_lock<std::mutex> lk(g->job_mu);
    g->done_cv.wait(lk, [&] { return g->pending == 0; });
  }
  for (auto& w : g->workers)
    if (w->rc) return fail(w->rc, "device " + std::to_string(w->device) + ": " + w->error);
  return 0;
}

extern "C" {

int qg_group_create(const int* devices, int n_devices, int dim, int metric, const qg_config* cfg, qg_group** out) {
  if (!out) return fail(QG_ERR_INVALID, "out is null");
  *out = nullptr;
  if (!devices || n_devices <= 0 || n_devices > 64) return fail(QG_ERR_INVALID, "bad device list");
  if (n_devices > 1)
    if (int rc = nccl_load()) return rc;
  std::unique_ptr<qg_group> g(new qg_group());
  g->n = n_devices;
  g->dim = dim;
  g->metric = metric;
  std::vector<ncclComm_t> comms((size_t)n_devices, nullptr);
  if (n_devices > 1) QG_NCCL_OK(g_nccl.CommInitAll(comms.data(), n_devices, devices));
  for (int i = 0; i < n_devices; ++i) {
    std::unique_ptr<GroupWorker> w(new GroupWorker());
    w->device = devices[i];
    w->rank = i;
    qg_config c{};
    if (cfg) c = *cfg;
    c.device = devices[i];
    c.reserve_rows = 0;
    if (int rc = qg_index_create(&w->index, dim, metric, &c)) return rc;
    QG_CUDA_OK(cudaSetDevice(devices[i]));
    QG_CUDA_OK(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
    if (int rc = comm_alloc(&w->comm, comms[(size_t)i], n_devices, i, devices[i])) return rc;
    g->workers.push_back(std::move(w));
  }
  for (auto& w : g->workers) w->thread = std::thread(worker_main, g.get(), w.get());
  *out = g.release();
  return 0;
}

int qg_group_destroy(qg_group* g) {
  if (!g) return 0;
  {
    std::lock_guard<std::mutex> lk(g->mu);
    GroupJob j;
    j.kind = 4;
    group_run(g, j);
  }
  for (auto& w : g->workers) {
    if (w->thread.joinable()) w->thread.join();
    cudaSetDevice(w->device);
    cudaStreamSynchronize(w->stream);
    w->d_q.release(); w->d_dist.release(); w->d_row.release(); w->d_cnt.release();
    if (w->h_q) cudaFreeHost(w->h_q);
    if (w->h_out) cudaFreeHost(w->h_out);
    qg_comm_destroy(w->comm);
    qg_index_destroy(w->index);
    cudaStreamDestroy(w->stream);
  }
  delete g;
  return 0;
}

int qg_group_devices(const qg_group* g) { return g ? g->n : 0; }
qg_index* qg_group_index(qg_group* g, int i) { return (g && i >= 0 && i < g->n) ? g->workers[(size_t)i]->index : nullptr; }
int64_t qg_group_row_base(const qg_group* g, int i) {
  return (g && i >= 0 && i < g->n) ? g->workers[(size_t)i]->row_base : -1;
}
int64_t qg_group_rows(const qg_group* g) { return g ? g->rows : 0; }
int qg_group_layout(const qg_group* g) { return g ? g->layout : -1; }

static int group_pick_layout(qg_group* g, int64_t n, int layout) {
  if (layout == QG_LAYOUT_ROWS || layout == QG_LAYOUT_QUERIES) return layout;
  // replicate a corpus whose fp32 rows + bf16 copy (6 bytes per element) stay under a quarter of one GPU
  size_t free_b = 0, total_b = 0;
  cudaSetDevice(g->workers[0]->device);
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) total_b = 180ull << 30;
  return ((double)n * g->dim * 6.0 <= 0.25 * (double)total_b) ? QG_LAYOUT_QUERIES : QG_LAYOUT_ROWS;
}

int qg_group_upload(qg_group* g, const float* rows, int64_t n, int layout) {
  if (!g) return fail(QG_ERR_INVALID, "group handle is null");
  if (n < 0 || (n > 0 && !rows)) return fail(QG_ERR_INVALID, "bad rows");
  std::lock_guard<std::mutex> lk(g->mu);
  if (g->loaded) return fail(QG_ERR_UNSUPPORTED, "a group is filled by one upload call");
  g->layout = group_pick_layout(g, n, layout);
  GroupJob j;
  j.kind = 1;
  j.rows = rows;
  j.n = n;
  if (int rc = group_run(g, j)) return rc;
  g->rows = n;
  g->loaded = true;
  return 0;
}

int qg_group_upload_synthetic(qg_group* g, int kind, uint64_t seed, int64_t n, int layout) {
  if (!g) return fail(QG_ERR_INVALID, "group handle is null");
  if (n < 0) return fail(QG_ERR_INVALID, "bad rows");
  std::lock_guard<std::mutex> lk(g->mu);
  if (g->loaded) return fail(QG_ERR_UNSUPPORTED, "a group is filled by one upload call");
  g->layout = group_pick_layout(g, n, layout);
  GroupJob j;
  j.kind = 2;
  j.synth_kind = kind;
  j.seed = seed;
  j.n = n;
  if (int rc = group_run(g, j)) return rc;
  g->rows = n;
  g->loaded = true;
  return 0;
}

int qg_group_search_batch(qg_group* g, const float* queries, int q, int dim, int k, float* out_dist, int64_t* out_row,
                          int* out_count) {
  if (!g) return fail(QG_ERR_INVALID, "group handle is null");
  if (q < 0) return fail(QG_ERR_INVALID, "negative query count");
  if (q == 0) return 0;
  if (!out_count) return fail(QG_ERR_INVALID, "out_count is null");
  std::lock_guard<std::mutex> lk(g->mu);
  // error order of exact.go:96-106: empty index -> no results, no error; then dimension; then k
  if (g->rows == 0) {
    for (int i = 0; i < q; ++i) out_count[i] = 0;
    for (long long i = 0; k > 0 && i < (long long)q * k; ++i) {
      if (out_dist) out_dist[i] = __builtin_huge_valf();
      if (out_row) out_row[i] = -1;
    }
    return 0;
  }
  if (dim != g->dim)
    return fail(QG_ERR_DIM, "query dimension mismatch: expected " + std::to_string(g->dim) + ", got " + std::to_string(dim));
  if (k <= 0) return fail(QG_ERR_K, "k must be positive");
  if (!queries || !out_dist || !out_row) return fail(QG_ERR_INVALID, "null buffer");
  GroupJob j;
  j.kind = 3;
  j.queries = queries;
  j.q = q;
  j.dim = dim;
  j.k = k;
  j.out_dist = out_dist;
  j.out_row = out_row;
  j.out_count = out_count;
  return group_run(g, j);
}

}  // extern "C"
