// api.cu — the C ABI of libquivergpu.so (include/quiver_gpu.h): index lifecycle, upload through
// pinned staging buffers, facet columns / filters, search orchestration.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/quiver_gpu.h"
#include "common.cuh"
#include "compact.cuh"
#include "exact.cuh"
#include "finalize.cuh"
#include "hnsw.cuh"
#include "misc.cuh"
#include "scan.cuh"
#include "synth.cuh"
#include "tc_scan.cuh"

namespace qg {

static thread_local std::string g_error;
static thread_local qg_scan_stats g_my_stats{};  // stats of THIS thread's last search_enqueue (qg_search_batch reads them back)
void set_error(const std::string& msg) { g_error = msg; }
int fail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("QG_PDL");
    return e == nullptr || std::atoi(e) != 0;
  }();
  return on;
}

// ---- per-device one-time state ----------------------------------------------------------------
struct DeviceState {
  bool attrs_done = false;
  int sm_count = 0;
};
static std::mutex g_dev_mu;
static DeviceState g_dev[64];

static int ensure_device(int device) {
  if (device < 0 || device >= 64) return fail(QG_ERR_INVALID, "device ordinal out of range");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(QG_ERR_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") +
                                 cudaGetErrorString(e));
  if (device >= n) return fail(QG_ERR_INVALID, "device ordinal >= device count");
  QG_CUDA_OK(cudaSetDevice(device));
  std::lock_guard<std::mutex> lk(g_dev_mu);
  DeviceState& ds = g_dev[device];
  if (!ds.attrs_done) {
    cudaDeviceProp prop;
    QG_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
      return fail(QG_ERR_CUDA, std::string("libquivergpu is built for sm_100a only; device is ") + prop.name);
    ds.sm_count = prop.multiProcessorCount;
    if (int rc = scan_set_attributes()) return rc;
    if (int rc = finalize_set_attributes()) return rc;
    if (int rc = tc_set_attributes()) return rc;
    ds.attrs_done = true;
  }
  return 0;
}

// ---- growable device buffer ---------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    size_t want = std::max(need, (size_t)256);
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      p = nullptr;
      return fail(QG_ERR_OOM, std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e));
    }
    bytes = want;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};

struct PinBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr;
    bytes = 0;
    size_t want = std::max(need, (size_t)4096);
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
    if (e != cudaSuccess) {
      p = nullptr;
      return fail(QG_ERR_OOM, std::string("cudaHostAlloc(") + std::to_string(want) + "): " + cudaGetErrorString(e));
    }
    bytes = want;
    return 0;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    bytes = 0;
  }
};

// Per-call scratch: acquired from the index's pool so concurrent searches never share one.
struct Workspace {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;  // recorded when an async (device-API) call finished using this workspace
  bool busy_async = false;
  cudaStream_t last_stream = nullptr;  // caller stream of the async call that parked this workspace
  // profiling: event pairs (begin, end); kind 0 = scan, 1 = finalize
  std::vector<cudaEvent_t> prof_ev;
  std::vector<int> prof_kind;
  size_t prof_used = 0;  // events in use (2 per bracket)
  int prof_begin(int kind, cudaStream_t st) {
    if (prof_used + 2 > prof_ev.size()) {
      cudaEvent_t a = nullptr, b = nullptr;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return -1;
      prof_ev.push_back(a);
      prof_ev.push_back(b);
      prof_kind.push_back(kind);
    }
    prof_kind[prof_used / 2] = kind;
    cudaEventRecord(prof_ev[prof_used], st);
    return 0;
  }
  void prof_end(cudaStream_t st) {
    cudaEventRecord(prof_ev[prof_used + 1], st);
    prof_used += 2;
  }
  // host-buffer searches stage their queries chunk by chunk on a copy stream, one event per chunk
  cudaStream_t copy_stream = nullptr, d2h_stream = nullptr;
  std::vector<cudaEvent_t> chunk_ev;  // per chunk: queries staged, chunk computed, results on the host
  DevBuf qpad, negpad, partial, mask, counters;
  DevBuf xchg;              // dense flat scan: threshold exchange words (scan.cuh), zeroed once
  uint32_t xchg_epoch = 0;  // bumped per scan launch; words of older launches do not count
  DevBuf tc_sample, tc_tau, tc_cand, tc_cnt, tc_bias, tc_apack;
  DevBuf d_q, d_neg, d_dist, d_negdist, d_row, d_count, d_rows32, d_rows64, d_fetch;
  PinBuf h_in, h_out;
  ExhaustiveWork ex;
  void destroy() {
    qpad.release(); negpad.release(); partial.release(); mask.release(); counters.release(); xchg.release();
    tc_sample.release(); tc_tau.release(); tc_cand.release(); tc_cnt.release(); tc_bias.release(); tc_apack.release();
    d_q.release(); d_neg.release(); d_dist.release(); d_negdist.release(); d_row.release(); d_count.release();
    d_rows32.release(); d_rows64.release(); d_fetch.release();
    h_in.release(); h_out.release();
    exhaustive_free(ex);
    for (cudaEvent_t e : prof_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : chunk_ev) cudaEventDestroy(e);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (d2h_stream) cudaStreamDestroy(d2h_stream);
    prof_ev.clear();
    if (done) cudaEventDestroy(done);
    if (stream) cudaStreamDestroy(stream);
  }
};

struct FacetColumn {
  bool set = false;
  long long n = 0;
  DevBuf kind, num, scode, fcode;
  DevBuf arr_off, arr_code;  // optional CSR of array elements (qg_facets_set_array_column)
  long long arr_rows = -1;   // rows covered by arr_off, -1 = none
  bool synthetic = false;    // created by prepare_columns for a field nobody set (every row MISSING)
};

}  // namespace qg

using namespace qg;

struct qg_filter {
  qg_index* owner = nullptr;
  int device = 0;  // the owner's device, for qg_filter_destroy (which may run after the index is gone)
  std::vector<qg_pred> preds;
  std::vector<qg_clause> clauses;
  std::vector<int32_t> iset;
  std::vector<double> fset;
  DevBuf d_preds, d_clauses, d_iset, d_fset;
  // cache (guarded by mu)
  std::mutex mu;
  long long raw_rows = -1;
  uint64_t raw_facet_epoch = ~0ull;
  DevBuf raw_mask;
  long long raw_matches = 0;
  uint64_t comb_live_epoch = ~0ull;
  DevBuf comb_mask, gather_list;
  long long comb_matches = 0;
  bool gather_valid = false;
  // Dense copy of the passing rows for batched searches (build_filter_view): fp32 rows, bf16 copy, norms, in
  // ascending row order, plus the new -> old row list. Valid for (view_live_epoch, view_facet_epoch, view_rows).
  DevBuf v_vec, v_vec16, v_inv, v_n2, v_ub, v_new2old, v_map, v_cnt, v_woff, v_tmp;
  long long view_n = -1, view_rows = -1;
  uint64_t view_live_epoch = ~0ull, view_facet_epoch = ~0ull;
};

struct qg_index {
  int device = 0;
  int dim = 0, dp = 0;
  int metric = 0, arith = 0;
  int margin = 16;
  int sm_count = 148;
  long long cap = 0, n_rows = 0, n_live = 0;
  float* vec = nullptr;
  float* inv_norm = nullptr;
  float* norm2 = nullptr;      // |x|^2, +inf beyond n_rows (tensor-core row term, L2)
  float* unit_bias = nullptr;  // 1.0, +inf beyond n_rows (tensor-core row term, dot / cosine)
  void* vec16 = nullptr;       // bf16 copy of vec, [cap x dp16]: the tensor-core stream (dim <= 512), else nullptr
  int dp16 = 0;
  bool use_bf16 = false;
  uint32_t* live = nullptr;
  float* max_norm2 = nullptr;  // device scalar
  uint64_t live_epoch = 0, facet_epoch = 0;
  std::vector<FacetColumn> cols;
  DevBuf col_table;  // FacetColDev[]
  bool col_table_dirty = true;
  // Facet columns and their device pointer table: qg_facets_set_* and the (re)evaluation of a predicate mask
  // (prepare_columns + the eval launch + its sync) exclude each other, so concurrent filtered searches — which
  // only hold the caller's SHARED lock — never see a column half rebuilt.
  std::mutex cols_mu;
  std::mutex stats_mu;  // idx->stats is introspection shared by concurrent searches
  std::mutex ws_mu;
  std::vector<std::unique_ptr<Workspace>> ws_free;
  std::vector<std::unique_ptr<Workspace>> ws_async;  // in flight on caller streams
  PinBuf stage[2];
  cudaStream_t up_stream = nullptr;
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  qg_scan_stats stats{};
  bool profiling = false;
  int tc_min_q = 8;             // query batches at least this large use the tensor-core regime
  long long tc_min_rows = 8192;  // below this the sampled threshold admits too few rows (rank * rows / 2048) and the flat scan is cheap
};

namespace qg {

// Number of corpus tiles the sample stage scores: chosen so that about G rows per query fall under
// the tc_sample_rank(k)-th smallest sampled score (G ~ 256 for k = 10).
static long long tc_sample_tiles(const TcPlan& plan, long long n_rows, int k) {
  const long long n_tiles = (n_rows + plan.tile_rows - 1) / plan.tile_rows;
  const long long G = std::max<long long>(256, 6ll * (k + 24));
  long long n_sample = ((long long)tc_sample_rank(k) * n_rows + G * plan.tile_rows - 1) / (G * plan.tile_rows);
  // the sampled fraction is rank / G (~3 %) of the corpus whatever its size
  return std::max<long long>(16, std::min<long long>(n_sample, std::min<long long>(n_tiles, 1 << 17)));
}

static int scan_mode_of(int metric) {
  switch (metric) {
    case METRIC_L2:
    case METRIC_SQL2: return MODE_L2;
    case METRIC_L1: return MODE_L1;
    default: return MODE_DOT;
  }
}

// `st_hint`: a workspace parked by an earlier async call on the SAME stream can be reused at once —
// the new work is ordered behind the old one on that stream.
static Workspace* ws_acquire(qg_index* idx, cudaStream_t st_hint = nullptr, bool have_hint = false) {
  std::lock_guard<std::mutex> lk(idx->ws_mu);
  if (have_hint) {
    for (size_t i = 0; i < idx->ws_async.size(); ++i) {
      if (idx->ws_async[i]->last_stream == st_hint) {
        Workspace* w = idx->ws_async[i].release();
        idx->ws_async.erase(idx->ws_async.begin() + i);
        w->busy_async = false;
        return w;
      }
    }
  }
  // recycle async workspaces whose work has completed
  for (size_t i = 0; i < idx->ws_async.size();) {
    if (cudaEventQuery(idx->ws_async[i]->done) == cudaSuccess) {
      idx->ws_async[i]->busy_async = false;
      idx->ws_free.push_back(std::move(idx->ws_async[i]));
      idx->ws_async.erase(idx->ws_async.begin() + i);
    } else {
      ++i;
    }
  }
  if (!idx->ws_free.empty()) {
    Workspace* w = idx->ws_free.back().release();
    idx->ws_free.pop_back();
    return w;
  }
  Workspace* w = new Workspace();
  if (cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&w->done, cudaEventDisableTiming) != cudaSuccess) {
    w->destroy();
    delete w;
    return nullptr;
  }
  return w;
}

static void ws_release(qg_index* idx, Workspace* w) {
  std::lock_guard<std::mutex> lk(idx->ws_mu);
  idx->ws_free.emplace_back(w);
}

static void ws_release_async(qg_index* idx, Workspace* w, cudaStream_t st) {
  cudaEventRecord(w->done, st);
  w->busy_async = true;
  w->last_stream = st;
  std::lock_guard<std::mutex> lk(idx->ws_mu);
  idx->ws_async.emplace_back(w);
}

static int grow(qg_index* idx, long long need_rows) {
  if (need_rows <= idx->cap) return 0;
  long long ncap = std::max<long long>(idx->cap * 2, 1024);
  ncap = std::max(ncap, need_rows);
  ncap = (ncap + 1023) & ~1023ll;  // whole mask words, 16-byte aligned columns
  float* nvec = nullptr;
  float* ninv = nullptr;
  float* nn2 = nullptr;
  float* nub = nullptr;
  void* nv16 = nullptr;
  uint32_t* nlive = nullptr;
  cudaError_t e = cudaMalloc(&nvec, (size_t)ncap * idx->dp * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&ninv, (size_t)ncap * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&nn2, (size_t)ncap * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&nub, (size_t)ncap * sizeof(float));
  if (e == cudaSuccess && idx->use_bf16) e = cudaMalloc(&nv16, (size_t)ncap * idx->dp16 * 2);
  if (e == cudaSuccess) e = cudaMalloc(&nlive, (size_t)(ncap / 32) * sizeof(uint32_t));
  if (e != cudaSuccess) {
    if (nvec) cudaFree(nvec);
    if (ninv) cudaFree(ninv);
    if (nn2) cudaFree(nn2);
    if (nub) cudaFree(nub);
    if (nv16) cudaFree(nv16);
    if (nlive) cudaFree(nlive);
    return fail(QG_ERR_OOM, std::string("device allocation for ") + std::to_string(ncap) +
                                " rows failed: " + cudaGetErrorString(e));
  }
  QG_CUDA_OK(cudaMemsetAsync(nlive, 0, (size_t)(ncap / 32) * sizeof(uint32_t), idx->up_stream));
  if (int rc = launch_fill_f32(nn2, ncap, INFINITY, idx->up_stream)) return rc;
  if (int rc = launch_fill_f32(nub, ncap, INFINITY, idx->up_stream)) return rc;
  if (idx->n_rows > 0) {
    QG_CUDA_OK(cudaMemcpyAsync(nvec, idx->vec, (size_t)idx->n_rows * idx->dp * sizeof(float),
                               cudaMemcpyDeviceToDevice, idx->up_stream));
    QG_CUDA_OK(cudaMemcpyAsync(ninv, idx->inv_norm, (size_t)idx->n_rows * sizeof(float), cudaMemcpyDeviceToDevice,
                               idx->up_stream));
    QG_CUDA_OK(cudaMemcpyAsync(nn2, idx->norm2, (size_t)idx->n_rows * sizeof(float), cudaMemcpyDeviceToDevice,
                               idx->up_stream));
    QG_CUDA_OK(cudaMemcpyAsync(nub, idx->unit_bias, (size_t)idx->n_rows * sizeof(float), cudaMemcpyDeviceToDevice,
                               idx->up_stream));
    if (nv16)
      QG_CUDA_OK(cudaMemcpyAsync(nv16, idx->vec16, (size_t)idx->n_rows * idx->dp16 * 2, cudaMemcpyDeviceToDevice,
                                 idx->up_stream));
    QG_CUDA_OK(cudaMemcpyAsync(nlive, idx->live, (size_t)((idx->n_rows + 31) / 32) * sizeof(uint32_t),
                               cudaMemcpyDeviceToDevice, idx->up_stream));
  }
  QG_CUDA_OK(cudaStreamSynchronize(idx->up_stream));
  if (idx->vec) cudaFree(idx->vec);
  if (idx->inv_norm) cudaFree(idx->inv_norm);
  if (idx->norm2) cudaFree(idx->norm2);
  if (idx->unit_bias) cudaFree(idx->unit_bias);
  if (idx->vec16) cudaFree(idx->vec16);
  if (idx->live) cudaFree(idx->live);
  idx->vec16 = nv16;
  idx->vec = nvec;
  idx->inv_norm = ninv;
  idx->norm2 = nn2;
  idx->unit_bias = nub;
  idx->live = nlive;
  idx->cap = ncap;
  return 0;
}

// Post-copy bookkeeping shared by the three upload flavours.
static int finish_append(qg_index* idx, long long n, int64_t* first_row) {
  const long long row0 = idx->n_rows;
  if (int rc = launch_row_norms(idx->vec, row0, n, idx->dp, idx->dim, idx->inv_norm, idx->norm2, idx->unit_bias,
                                idx->max_norm2, idx->up_stream))
    return rc;
  if (idx->use_bf16) {
    if (int rc = launch_tc_to_bf16(idx->vec, row0, n, idx->dp, idx->dim, idx->dp16,
                                   tc_extra_cols(idx->dim, scan_mode_of(idx->metric) == MODE_L2) ? idx->norm2 : nullptr,
                                   idx->metric == METRIC_COSINE ? idx->inv_norm : nullptr, idx->vec16, idx->up_stream))
      return rc;
  }
  if (int rc = launch_set_live(idx->live, row0, n, idx->up_stream)) return rc;
  QG_CUDA_OK(cudaStreamSynchronize(idx->up_stream));
  idx->n_rows += n;
  idx->n_live += n;
  idx->live_epoch++;
  if (first_row) *first_row = row0;
  return 0;
}

static int check_index(const qg_index* idx) {
  if (!idx) return fail(QG_ERR_INVALID, "index handle is null");
  QG_CUDA_OK(cudaSetDevice(idx->device));
  return 0;
}

}  // namespace qg

// ================================================================================================
extern "C" {

int qg_abi_version(void) { return QG_ABI_VERSION; }
const char* qg_last_error(void) { return g_error.c_str(); }

int qg_device_count(int* out_count) {
  if (!out_count) return fail(QG_ERR_INVALID, "out_count is null");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *out_count = 0;
    return fail(QG_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  }
  *out_count = n;
  return 0;
}

int qg_device_info(int device, char* name_buf, size_t name_len, int* out_sm_count, int* out_cc_major,
                   int* out_cc_minor) {
  cudaDeviceProp prop;
  QG_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (name_buf && name_len) {
    std::strncpy(name_buf, prop.name, name_len - 1);
    name_buf[name_len - 1] = 0;
  }
  if (out_sm_count) *out_sm_count = prop.multiProcessorCount;
  if (out_cc_major) *out_cc_major = prop.major;
  if (out_cc_minor) *out_cc_minor = prop.minor;
  return 0;
}

int qg_index_create(qg_index** out, int dim, int metric, const qg_config* cfg) {
  if (!out) return fail(QG_ERR_INVALID, "out is null");
  *out = nullptr;
  if (dim <= 0) return fail(QG_ERR_DIM, "dimension must be positive");
  if (metric < 0 || metric > 4) return fail(QG_ERR_INVALID, "unknown metric");
  const int device = cfg ? cfg->device : 0;
  if (cfg && (cfg->flags & ~QG_FLAG_NO_BF16_COPY) != 0) return fail(QG_ERR_INVALID, "unknown bit in qg_config.flags");
  if (int rc = ensure_device(device)) return rc;
  std::unique_ptr<qg_index> idx(new qg_index());
  idx->device = device;
  idx->dim = dim;
  idx->dp = (dim + 3) & ~3;
  idx->metric = metric;
  idx->arith = cfg ? cfg->arith : 0;
  if (idx->arith != ARITH_VECTORTYPES && idx->arith != ARITH_HNSW_F32) return fail(QG_ERR_INVALID, "unknown arith");
  if (idx->arith == ARITH_HNSW_F32 && (metric == METRIC_SQL2 || metric == METRIC_L1))
    return fail(QG_ERR_INVALID, "hnsw float32 arithmetic exists for cosine / l2 / dot only");
  idx->margin = (cfg && cfg->select_margin > 0) ? cfg->select_margin : 16;
  if (const char* e = std::getenv("QG_TC_MIN_Q")) idx->tc_min_q = std::max(1, std::atoi(e));
  if (const char* e = std::getenv("QG_TC_MIN_ROWS")) idx->tc_min_rows = std::max(128, std::atoi(e));
  // bf16 copy for the tensor-core stream (+50 % memory): every distance that is RETURNED is still
  // recomputed from the fp32 rows, the copy only feeds candidate selection. QG_TC_BF16=0 disables it.
  idx->dp16 = tc_dp16(dim, scan_mode_of(metric) == MODE_L2);
  idx->use_bf16 = dim <= 512 && metric != QG_L1 && !(cfg && (cfg->flags & QG_FLAG_NO_BF16_COPY));
  if (const char* e = std::getenv("QG_TC_BF16")) idx->use_bf16 = idx->use_bf16 && std::atoi(e) != 0;
  // With the bf16 copy a tensor-core pass streams half the bytes of the flat fp32 scan, so it wins from the
  // first query on (1M x 128, whole call: 72 us for one query and 75 us for 2..8, against 100 us for one query
  // and 134 / 220 us for 2 / 4 queries on the flat scan). The flat scan remains the regime of indexes without
  // the copy, of filtered scans over row lists, and the re-run of a query the copy could not certify.
  if (idx->use_bf16 && std::getenv("QG_TC_MIN_Q") == nullptr) idx->tc_min_q = 1;
  {
    std::lock_guard<std::mutex> lk(g_dev_mu);
    idx->sm_count = g_dev[device].sm_count;
  }
  QG_CUDA_OK(cudaStreamCreateWithFlags(&idx->up_stream, cudaStreamNonBlocking));
  QG_CUDA_OK(cudaEventCreateWithFlags(&idx->stage_ev[0], cudaEventDisableTiming));
  QG_CUDA_OK(cudaEventCreateWithFlags(&idx->stage_ev[1], cudaEventDisableTiming));
  QG_CUDA_OK(cudaMalloc(&idx->max_norm2, sizeof(float)));
  QG_CUDA_OK(cudaMemset(idx->max_norm2, 0, sizeof(float)));
  if (cfg && cfg->reserve_rows > 0) {
    if (int rc = grow(idx.get(), cfg->reserve_rows)) return rc;
  }
  *out = idx.release();
  return 0;
}

int qg_index_destroy(qg_index* idx) {
  if (!idx) return 0;
  cudaSetDevice(idx->device);
  cudaDeviceSynchronize();
  for (auto& w : idx->ws_free) w->destroy();
  for (auto& w : idx->ws_async) w->destroy();
  for (auto& c : idx->cols) {
    c.kind.release(); c.num.release(); c.scode.release(); c.fcode.release();
    c.arr_off.release(); c.arr_code.release();
  }
  idx->col_table.release();
  idx->stage[0].release();
  idx->stage[1].release();
  if (idx->vec) cudaFree(idx->vec);
  if (idx->inv_norm) cudaFree(idx->inv_norm);
  if (idx->norm2) cudaFree(idx->norm2);
  if (idx->unit_bias) cudaFree(idx->unit_bias);
  if (idx->vec16) cudaFree(idx->vec16);
  if (idx->live) cudaFree(idx->live);
  if (idx->max_norm2) cudaFree(idx->max_norm2);
  if (idx->stage_ev[0]) cudaEventDestroy(idx->stage_ev[0]);
  if (idx->stage_ev[1]) cudaEventDestroy(idx->stage_ev[1]);
  if (idx->up_stream) cudaStreamDestroy(idx->up_stream);
  delete idx;
  return 0;
}

int qg_index_upload(qg_index* idx, const float* rows, int64_t n, int64_t* first_row) {
  if (int rc = check_index(idx)) return rc;
  if (n < 0 || (n > 0 && !rows)) return fail(QG_ERR_INVALID, "rows is null or n < 0");
  if (idx->n_rows + n > 0xFFFFFFFEll) return fail(QG_ERR_RANGE, "an index holds at most 2^32-2 rows");
  if (n == 0) {
    if (first_row) *first_row = idx->n_rows;
    return 0;
  }
  if (int rc = grow(idx, idx->n_rows + n)) return rc;
  const int d = idx->dim, dp = idx->dp;
  float* dst0 = idx->vec + (size_t)idx->n_rows * dp;
  if (dp != d) QG_CUDA_OK(cudaMemsetAsync(dst0, 0, (size_t)n * dp * sizeof(float), idx->up_stream));
  // pinned double-buffered staging: host memcpy of chunk i+1 overlaps the H2D DMA of chunk i
  const size_t row_bytes = (size_t)d * sizeof(float);
  const long long chunk_rows = std::max<long long>(1, (long long)((32u << 20) / row_bytes));
  for (int b = 0; b < 2; ++b)
    if (int rc = idx->stage[b].ensure((size_t)std::min<long long>(chunk_rows, n) * row_bytes)) return rc;
  int buf = 0;
  for (long long off = 0; off < n; off += chunk_rows, buf ^= 1) {
    const long long m = std::min(chunk_rows, n - off);
    QG_CUDA_OK(cudaEventSynchronize(idx->stage_ev[buf]));  // previous DMA out of this buffer is done
    std::memcpy(idx->stage[buf].p, rows + (size_t)off * d, (size_t)m * row_bytes);
    if (dp == d) {
      QG_CUDA_OK(cudaMemcpyAsync(dst0 + (size_t)off * dp, idx->stage[buf].p, (size_t)m * row_bytes,
                                 cudaMemcpyHostToDevice, idx->up_stream));
    } else {
      QG_CUDA_OK(cudaMemcpy2DAsync(dst0 + (size_t)off * dp, (size_t)dp * sizeof(float), idx->stage[buf].p, row_bytes,
                                   row_bytes, (size_t)m, cudaMemcpyHostToDevice, idx->up_stream));
    }
    QG_CUDA_OK(cudaEventRecord(idx->stage_ev[buf], idx->up_stream));
  }
  return finish_append(idx, n, first_row);
}

int qg_index_upload_device(qg_index* idx, const void* d_rows, int64_t n, int64_t* first_row) {
  if (int rc = check_index(idx)) return rc;
  if (n < 0 || (n > 0 && !d_rows)) return fail(QG_ERR_INVALID, "d_rows is null or n < 0");
  if (idx->n_rows + n > 0xFFFFFFFEll) return fail(QG_ERR_RANGE, "an index holds at most 2^32-2 rows");
  if (n == 0) {
    if (first_row) *first_row = idx->n_rows;
    return 0;
  }
  if (int rc = grow(idx, idx->n_rows + n)) return rc;
  const int d = idx->dim, dp = idx->dp;
  float* dst0 = idx->vec + (size_t)idx->n_rows * dp;
  if (dp == d) {
    QG_CUDA_OK(cudaMemcpyAsync(dst0, d_rows, (size_t)n * d * sizeof(float), cudaMemcpyDeviceToDevice, idx->up_stream));
  } else {
    QG_CUDA_OK(cudaMemsetAsync(dst0, 0, (size_t)n * dp * sizeof(float), idx->up_stream));
    QG_CUDA_OK(cudaMemcpy2DAsync(dst0, (size_t)dp * sizeof(float), d_rows, (size_t)d * sizeof(float),
                                 (size_t)d * sizeof(float), (size_t)n, cudaMemcpyDeviceToDevice, idx->up_stream));
  }
  return finish_append(idx, n, first_row);
}

int qg_index_upload_synthetic(qg_index* idx, int kind, uint64_t seed, int64_t global_row0, int64_t n,
                              int64_t* first_row) {
  if (int rc = check_index(idx)) return rc;
  if (n < 0) return fail(QG_ERR_INVALID, "n < 0");
  if (idx->n_rows + n > 0xFFFFFFFEll) return fail(QG_ERR_RANGE, "an index holds at most 2^32-2 rows");
  if (n == 0) {
    if (first_row) *first_row = idx->n_rows;
    return 0;
  }
  if (int rc = grow(idx, idx->n_rows + n)) return rc;
  if (int rc = launch_synth_fill(idx->vec, idx->n_rows, n, idx->dp, idx->dim, kind, seed, global_row0,
                                 idx->up_stream))
    return rc;
  return finish_append(idx, n, first_row);
}

int qg_index_tombstone(qg_index* idx, const int64_t* rows, int64_t n) {
  if (int rc = check_index(idx)) return rc;
  if (n < 0 || (n > 0 && !rows)) return fail(QG_ERR_INVALID, "rows is null or n < 0");
  if (n == 0) return 0;
  for (int64_t i = 0; i < n; ++i)
    if (rows[i] < 0 || rows[i] >= idx->n_rows) return fail(QG_ERR_RANGE, "tombstone: row index out of range");
  DevBuf d_rows, d_cnt;
  if (int rc = d_rows.ensure((size_t)n * 8)) return rc;
  if (int rc = d_cnt.ensure(8)) {
    d_rows.release();
    return rc;
  }
  int rc = 0;
  unsigned long long cleared = 0;
  do {
    cudaError_t e = cudaMemcpyAsync(d_rows.p, rows, (size_t)n * 8, cudaMemcpyHostToDevice, idx->up_stream);
    if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, cudaGetErrorString(e)); break; }
    rc = launch_tombstone(idx->live, (const long long*)d_rows.p, n, idx->n_rows, (unsigned long long*)d_cnt.p,
                          idx->up_stream);
    if (rc) break;
    e = cudaMemcpyAsync(&cleared, d_cnt.p, 8, cudaMemcpyDeviceToHost, idx->up_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(idx->up_stream);
    if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, cudaGetErrorString(e)); break; }
  } while (0);
  d_rows.release();
  d_cnt.release();
  if (rc) return rc;
  idx->n_live -= (long long)cleared;
  idx->live_epoch++;
  return 0;
}

// Tombstone compaction. Out of place: the map old row -> new row comes from a prefix sum over the live
// mask, one warp per old row then moves the row's slice of every array; the index switches to the new
// arrays only after everything succeeded.
int qg_index_compact(qg_index* idx, int64_t* old_to_new, int64_t* out_rows) {
  if (int rc = check_index(idx)) return rc;
  QG_CUDA_OK(cudaDeviceSynchronize());  // nothing may still read the arrays that are about to be replaced
  // QG_COMPACT_TIMING=1: wall-clock stage times on stderr (development aid)
  static const bool timing = std::getenv("QG_COMPACT_TIMING") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaDeviceSynchronize();
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[qg_index_compact] %-28s %8.3f ms\n", what,
                 std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  const long long n_old = idx->n_rows, n_new = idx->n_live;
  if (out_rows) *out_rows = n_new;
  if (n_old == n_new) {
    if (old_to_new)
      for (long long r = 0; r < n_old; ++r) old_to_new[r] = r;
    return 0;
  }
  cudaStream_t st = idx->up_stream;
  const long long n_words = (n_old + 31) / 32;
  long long ncap = (std::max<long long>(n_new, 1024) + 1023) & ~1023ll;

  struct NewCol {
    DevBuf kind, num, scode, fcode, arr_off, arr_code, arr_cnt;
    long long n = 0;
    bool arr = false;
    void drop() {
      kind.release(); num.release(); scode.release(); fcode.release();
      arr_off.release(); arr_code.release(); arr_cnt.release();
    }
  };
  DevBuf cnt, woff, tmp, map, map64, nvec, ninv, nn2, nub, nv16, nlive;
  std::vector<NewCol> ncols(idx->cols.size());
  std::unique_ptr<uint32_t[]> h_map;  // uninitialised on purpose: filled by the D2H copy
  auto cleanup = [&](int rc) {
    cnt.release(); woff.release(); tmp.release(); map.release(); map64.release();
    nvec.release(); ninv.release(); nn2.release(); nub.release(); nv16.release(); nlive.release();
    for (auto& c : ncols) c.drop();
    return rc;
  };
  auto cuda_rc = [&](cudaError_t e) {
    return e == cudaSuccess ? 0 : fail(e == cudaErrorMemoryAllocation ? QG_ERR_OOM : QG_ERR_CUDA, cudaGetErrorString(e));
  };
  int rc = 0;
  // ---- old row -> new row ---------------------------------------------------------------------------
  if ((rc = cnt.ensure((size_t)n_words * 4)) || (rc = woff.ensure((size_t)(n_words + 1) * 4)) ||
      (rc = tmp.ensure((size_t)scan_tmp_words(std::max(n_words, ncap)) * 4)) || (rc = map.ensure((size_t)n_old * 4)))
    return cleanup(rc);
  if ((rc = launch_word_popcount(idx->live, n_words, (uint32_t*)cnt.p, st)) ||
      (rc = launch_exclusive_scan_u32((const uint32_t*)cnt.p, (uint32_t*)woff.p, n_words, (uint32_t*)tmp.p, st)) ||
      (rc = launch_compact_map(idx->live, (const uint32_t*)woff.p, n_old, (uint32_t*)map.p, st)))
    return cleanup(rc);
  uint32_t total = 0;
  if ((rc = cuda_rc(cudaMemcpyAsync(&total, (uint32_t*)woff.p + n_words, 4, cudaMemcpyDeviceToHost, st))) ||
      (rc = cuda_rc(cudaStreamSynchronize(st))))
    return cleanup(rc);
  if ((long long)total != n_new)
    return cleanup(fail(QG_ERR_CUDA, "compact: live mask holds " + std::to_string(total) + " rows, expected " +
                                             std::to_string(n_new)));
  lap("scan + map");
  bool any_col = false;
  for (const FacetColumn& c : idx->cols) any_col = any_col || (c.set && !c.synthetic);
  if (old_to_new) {
    // widened on the device and copied straight into the caller's buffer: one pass over it on the host side
    // (a 10M-row map cost 78 ms as a uint32 download plus a widening loop, page faults included)
    if ((rc = map64.ensure((size_t)n_old * 8)) ||
        (rc = launch_widen_map((const uint32_t*)map.p, n_old, (long long*)map64.p, st)) ||
        (rc = cuda_rc(cudaMemcpyAsync(old_to_new, map64.p, (size_t)n_old * 8, cudaMemcpyDeviceToHost, st))) ||
        (rc = cuda_rc(cudaStreamSynchronize(st))))
      return cleanup(rc);
    map64.release();
  } else if (any_col) {
    h_map.reset(new uint32_t[(size_t)n_old]);
    if ((rc = cuda_rc(cudaMemcpyAsync(h_map.get(), map.p, (size_t)n_old * 4, cudaMemcpyDeviceToHost, st))) ||
        (rc = cuda_rc(cudaStreamSynchronize(st))))
      return cleanup(rc);
  }
  auto new_row = [&](long long r) -> long long {
    if (old_to_new) return (long long)old_to_new[r];
    return h_map[(size_t)r] == 0xFFFFFFFFu ? -1ll : (long long)h_map[(size_t)r];
  };
  lap("map D2H");
  // ---- vectors, bf16 copy, norms, live mask -------------------------------------------------------------
  if ((rc = nvec.ensure((size_t)ncap * idx->dp * 4)) || (rc = ninv.ensure((size_t)ncap * 4)) ||
      (rc = nn2.ensure((size_t)ncap * 4)) || (rc = nub.ensure((size_t)ncap * 4)) ||
      (idx->use_bf16 && (rc = nv16.ensure((size_t)ncap * idx->dp16 * 2))) || (rc = nlive.ensure((size_t)(ncap / 32) * 4)))
    return cleanup(rc);
  lap("cudaMalloc new arrays");
  if ((rc = cuda_rc(cudaMemsetAsync(nlive.p, 0, (size_t)(ncap / 32) * 4, st))) ||
      (rc = launch_fill_f32((float*)nn2.p, ncap, INFINITY, st)) || (rc = launch_fill_f32((float*)nub.p, ncap, INFINITY, st)))
    return cleanup(rc);
  {
    CompactRowsArgs a{};
    a.map = (const uint32_t*)map.p;
    a.n_rows = n_old;
    a.vec = idx->vec; a.vec_out = (float*)nvec.p; a.dp = idx->dp;
    a.vec16 = idx->use_bf16 ? idx->vec16 : nullptr; a.vec16_out = nv16.p; a.dp16 = idx->dp16;
    a.inv_norm = idx->inv_norm; a.inv_norm_out = (float*)ninv.p;
    a.norm2 = idx->norm2; a.norm2_out = (float*)nn2.p;
    a.unit_bias = idx->unit_bias; a.unit_bias_out = (float*)nub.p;
    if ((rc = launch_compact_rows(a, idx->sm_count, st)) || (rc = launch_set_live((uint32_t*)nlive.p, 0, n_new, st)))
      return cleanup(rc);
  }
  lap("fills + row move");
  // ---- facet columns ----------------------------------------------------------------------------------
  for (size_t i = 0; i < idx->cols.size(); ++i) {
    const FacetColumn& c = idx->cols[i];
    if (!c.set || c.synthetic) continue;
    NewCol& nc = ncols[i];
    // rows the column covers after the move: the live ones among its first c.n rows
    nc.n = 0;
    for (long long r = c.n - 1; r >= 0; --r)
      if (new_row(r) >= 0) { nc.n = new_row(r) + 1; break; }
    nc.arr = c.arr_rows >= 0;
    if ((rc = nc.kind.ensure((size_t)ncap)) || (rc = nc.num.ensure((size_t)ncap * 8)) ||
        (rc = nc.scode.ensure((size_t)ncap * 4)) || (rc = nc.fcode.ensure((size_t)ncap * 4)))
      return cleanup(rc);
    if ((rc = cuda_rc(cudaMemsetAsync(nc.kind.p, QG_KIND_NOROW, nc.kind.bytes, st))) ||
        (rc = cuda_rc(cudaMemsetAsync(nc.num.p, 0, nc.num.bytes, st))) ||
        (rc = cuda_rc(cudaMemsetAsync(nc.scode.p, 0xff, nc.scode.bytes, st))) ||
        (rc = cuda_rc(cudaMemsetAsync(nc.fcode.p, 0xff, nc.fcode.bytes, st))))
      return cleanup(rc);
    if (nc.arr) {
      if ((rc = nc.arr_cnt.ensure((size_t)ncap * 4)) || (rc = nc.arr_off.ensure((size_t)(ncap + 1) * 4)))
        return cleanup(rc);
      if ((rc = cuda_rc(cudaMemsetAsync(nc.arr_cnt.p, 0, (size_t)ncap * 4, st)))) return cleanup(rc);
    }
    CompactColArgs a{};
    a.map = (const uint32_t*)map.p;
    a.n = c.n;
    a.kind = (const uint8_t*)c.kind.p; a.kind_out = (uint8_t*)nc.kind.p;
    a.num = (const double*)c.num.p; a.num_out = (double*)nc.num.p;
    a.scode = (const int32_t*)c.scode.p; a.scode_out = (int32_t*)nc.scode.p;
    a.fcode = (const int32_t*)c.fcode.p; a.fcode_out = (int32_t*)nc.fcode.p;
    a.arr_off = nc.arr ? (const int32_t*)c.arr_off.p : nullptr;
    a.arr_cnt_out = nc.arr ? (uint32_t*)nc.arr_cnt.p : nullptr;
    if ((rc = launch_compact_column(a, st))) return cleanup(rc);
    if (nc.arr) {
      // new offsets over all ncap rows (rows past nc.n own no elements, so their offsets equal the total:
      // the padding rule of qg_facets_set_array_column), then the element lists themselves
      if ((rc = launch_exclusive_scan_u32((const uint32_t*)nc.arr_cnt.p, (uint32_t*)nc.arr_off.p, ncap, (uint32_t*)tmp.p,
                                          st)))
        return cleanup(rc);
      uint32_t n_elems = 0;
      if ((rc = cuda_rc(cudaMemcpyAsync(&n_elems, (uint32_t*)nc.arr_off.p + ncap, 4, cudaMemcpyDeviceToHost, st))) ||
          (rc = cuda_rc(cudaStreamSynchronize(st))))
        return cleanup(rc);
      if ((rc = nc.arr_code.ensure(std::max<size_t>(1, n_elems) * 4))) return cleanup(rc);
      if ((rc = launch_compact_elems((const uint32_t*)map.p, c.n, (const int32_t*)c.arr_off.p,
                                     (const int32_t*)c.arr_code.p, (const uint32_t*)nc.arr_off.p, (int32_t*)nc.arr_code.p,
                                     idx->sm_count, st)))
        return cleanup(rc);
    }
  }
  if ((rc = cuda_rc(cudaStreamSynchronize(st)))) return cleanup(rc);
  lap("facet columns");
  // ---- switch over ------------------------------------------------------------------------------------
  cudaFree(idx->vec); cudaFree(idx->inv_norm); cudaFree(idx->norm2); cudaFree(idx->unit_bias); cudaFree(idx->live);
  if (idx->vec16) cudaFree(idx->vec16);
  idx->vec = (float*)nvec.p; nvec.p = nullptr; nvec.bytes = 0;
  idx->inv_norm = (float*)ninv.p; ninv.p = nullptr; ninv.bytes = 0;
  idx->norm2 = (float*)nn2.p; nn2.p = nullptr; nn2.bytes = 0;
  idx->unit_bias = (float*)nub.p; nub.p = nullptr; nub.bytes = 0;
  idx->live = (uint32_t*)nlive.p; nlive.p = nullptr; nlive.bytes = 0;
  idx->vec16 = nv16.p; nv16.p = nullptr; nv16.bytes = 0;
  for (size_t i = 0; i < idx->cols.size(); ++i) {
    FacetColumn& c = idx->cols[i];
    if (!c.set) continue;
    c.kind.release(); c.num.release(); c.scode.release(); c.fcode.release();
    c.arr_off.release(); c.arr_code.release();
    if (c.synthetic) {  // recreated on demand by prepare_columns
      c = FacetColumn();
      continue;
    }
    NewCol& nc = ncols[i];
    c.kind = nc.kind; c.num = nc.num; c.scode = nc.scode; c.fcode = nc.fcode;
    nc.kind = DevBuf(); nc.num = DevBuf(); nc.scode = DevBuf(); nc.fcode = DevBuf();
    c.n = nc.n;
    if (nc.arr) {
      c.arr_off = nc.arr_off; c.arr_code = nc.arr_code;
      nc.arr_off = DevBuf(); nc.arr_code = DevBuf();
      c.arr_rows = ncap;
    } else {
      c.arr_rows = -1;
    }
  }
  idx->cap = ncap;
  idx->n_rows = n_new;
  idx->live_epoch++;
  idx->facet_epoch++;
  idx->col_table_dirty = true;
  lap("cudaFree old arrays");
  rc = cleanup(0);
  lap("free scratch");
  return rc;
}

int64_t qg_index_size(const qg_index* idx) { return idx ? idx->n_live : 0; }
int64_t qg_index_rows(const qg_index* idx) { return idx ? idx->n_rows : 0; }
int qg_index_dim(const qg_index* idx) { return idx ? idx->dim : 0; }
int qg_index_metric(const qg_index* idx) { return idx ? idx->metric : -1; }

int qg_index_fetch(qg_index* idx, const int64_t* rows, int64_t n, float* out) {
  if (int rc = check_index(idx)) return rc;
  if (n < 0 || (n > 0 && (!rows || !out))) return fail(QG_ERR_INVALID, "fetch: null buffer or n < 0");
  if (n == 0) return 0;
  for (int64_t i = 0; i < n; ++i)
    if (rows[i] < 0 || rows[i] >= idx->n_rows) return fail(QG_ERR_RANGE, "fetch: row index out of range");
  Workspace* w = ws_acquire(idx);
  if (!w) return fail(QG_ERR_CUDA, "could not create a stream");
  int rc = 0;
  do {
    if ((rc = w->d_rows64.ensure((size_t)n * 8))) break;
    if ((rc = w->d_fetch.ensure((size_t)n * idx->dim * 4))) break;
    cudaError_t e = cudaMemcpyAsync(w->d_rows64.p, rows, (size_t)n * 8, cudaMemcpyHostToDevice, w->stream);
    if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, cudaGetErrorString(e)); break; }
    if ((rc = launch_fetch_rows(idx->vec, idx->dp, idx->dim, idx->n_rows, (const long long*)w->d_rows64.p, n,
                                (float*)w->d_fetch.p, w->stream)))
      break;
    e = cudaMemcpyAsync(out, w->d_fetch.p, (size_t)n * idx->dim * 4, cudaMemcpyDeviceToHost, w->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(w->stream);
    if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, cudaGetErrorString(e)); break; }
  } while (0);
  ws_release(idx, w);
  return rc;
}

// ---- facet columns and filters ----------------------------------------------------------------------
int qg_facets_set_column(qg_index* idx, int field, const uint8_t* kind, const double* num, const int32_t* scode,
                         const int32_t* fcode, int64_t n) {
  if (int rc = check_index(idx)) return rc;
  if (field < 0 || field >= 4096) return fail(QG_ERR_INVALID, "field index out of range");
  if (n < 0 || n > idx->n_rows) return fail(QG_ERR_RANGE, "column longer than the index");
  if (n > 0 && (!kind || !num || !scode || !fcode)) return fail(QG_ERR_INVALID, "null column buffer");
  std::lock_guard<std::mutex> cols_lock(idx->cols_mu);
  if ((size_t)field >= idx->cols.size()) idx->cols.resize(field + 1);
  FacetColumn& c = idx->cols[field];
  const size_t cap = (size_t)std::max<long long>(idx->cap, 1);
  if (int rc = c.kind.ensure(cap)) return rc;
  if (int rc = c.num.ensure(cap * 8)) return rc;
  if (int rc = c.scode.ensure(cap * 4)) return rc;
  if (int rc = c.fcode.ensure(cap * 4)) return rc;
  // rows beyond n have no metadata entry
  QG_CUDA_OK(cudaMemsetAsync(c.kind.p, QG_KIND_NOROW, c.kind.bytes, idx->up_stream));
  if (n > 0) {
    QG_CUDA_OK(cudaMemcpyAsync(c.kind.p, kind, (size_t)n, cudaMemcpyHostToDevice, idx->up_stream));
    QG_CUDA_OK(cudaMemcpyAsync(c.num.p, num, (size_t)n * 8, cudaMemcpyHostToDevice, idx->up_stream));
    QG_CUDA_OK(cudaMemcpyAsync(c.scode.p, scode, (size_t)n * 4, cudaMemcpyHostToDevice, idx->up_stream));
    QG_CUDA_OK(cudaMemcpyAsync(c.fcode.p, fcode, (size_t)n * 4, cudaMemcpyHostToDevice, idx->up_stream));
  }
  QG_CUDA_OK(cudaStreamSynchronize(idx->up_stream));
  c.set = true;
  c.synthetic = false;
  c.n = n;
  c.arr_rows = -1;  // a new column invalidates the element lists of the old one
  idx->facet_epoch++;
  idx->col_table_dirty = true;
  return 0;
}

int qg_facets_set_array_column(qg_index* idx, int field, const int32_t* offsets, const int32_t* elem_codes, int64_t n,
                               int64_t n_elems) {
  if (int rc = check_index(idx)) return rc;
  std::lock_guard<std::mutex> cols_lock(idx->cols_mu);
  if (field < 0 || (size_t)field >= idx->cols.size() || !idx->cols[field].set)
    return fail(QG_ERR_INVALID, "set the column with qg_facets_set_column first");
  FacetColumn& c = idx->cols[field];
  if (n != c.n) return fail(QG_ERR_RANGE, "array column must cover the same rows as the column");
  if (n_elems < 0 || !offsets || (n_elems > 0 && !elem_codes)) return fail(QG_ERR_INVALID, "null array column buffer");
  if (offsets[0] != 0 || offsets[n] != n_elems) return fail(QG_ERR_INVALID, "offsets must run from 0 to n_elems");
  for (int64_t i = 0; i < n; ++i)
    if (offsets[i + 1] < offsets[i]) return fail(QG_ERR_INVALID, "offsets must be non-decreasing");
  // rows appended later own no elements: the offsets are padded with n_elems up to the capacity
  const size_t cap = (size_t)std::max<long long>(idx->cap, 1);
  std::vector<int32_t> off(cap + 1, (int32_t)n_elems);
  std::copy(offsets, offsets + n + 1, off.begin());
  if (int rc = c.arr_off.ensure((cap + 1) * 4)) return rc;
  if (int rc = c.arr_code.ensure(std::max<size_t>(1, (size_t)n_elems) * 4)) return rc;
  QG_CUDA_OK(cudaMemcpy(c.arr_off.p, off.data(), (cap + 1) * 4, cudaMemcpyHostToDevice));
  if (n_elems > 0) QG_CUDA_OK(cudaMemcpy(c.arr_code.p, elem_codes, (size_t)n_elems * 4, cudaMemcpyHostToDevice));
  c.arr_rows = (long long)cap;
  idx->facet_epoch++;
  idx->col_table_dirty = true;
  return 0;
}

int qg_filter_compile(qg_index* idx, const qg_pred* preds, int n_preds, const qg_clause* clauses, int n_clauses,
                      const int32_t* iset, int n_iset, const double* fset, int n_fset, qg_filter** out) {
  if (int rc = check_index(idx)) return rc;
  if (!out) return fail(QG_ERR_INVALID, "out is null");
  *out = nullptr;
  if (n_preds < 0 || n_clauses < 0 || n_iset < 0 || n_fset < 0) return fail(QG_ERR_INVALID, "negative count");
  if ((n_preds && !preds) || (n_clauses && !clauses) || (n_iset && !iset) || (n_fset && !fset))
    return fail(QG_ERR_INVALID, "null program buffer");
  for (int i = 0; i < n_preds; ++i) {
    if (preds[i].first_clause < 0 || preds[i].n_clauses < 0 || preds[i].first_clause + preds[i].n_clauses > n_clauses)
      return fail(QG_ERR_INVALID, "predicate clause range out of bounds");
  }
  for (int i = 0; i < n_clauses; ++i) {
    const qg_clause& c = clauses[i];
    if (c.op < QG_OP_FALSE || c.op > QG_OP_WHOLE_EQ) return fail(QG_ERR_UNSUPPORTED, "unknown clause op");
    if (c.op >= QG_OP_KIND_IN) {
      if (c.field < 0 || c.field >= 4096) return fail(QG_ERR_INVALID, "clause field out of range");
    }
    if (c.op == QG_OP_SCODE_IN || c.op == QG_OP_FCODE_IN || c.op == QG_OP_ELEM_IN) {
      if (c.ia < 0 || c.ic < 0 || c.ia + c.ic > n_iset) return fail(QG_ERR_INVALID, "clause int-set out of bounds");
    }
    if (c.op == QG_OP_NUM_IN_TOL || c.op == QG_OP_NUM_IN) {
      if (c.ia < 0 || c.ic < 0 || c.ia + c.ic > n_fset) return fail(QG_ERR_INVALID, "clause float-set out of bounds");
    }
  }
  std::unique_ptr<qg_filter> f(new qg_filter());
  f->owner = idx;
  f->device = idx->device;
  f->preds.assign(preds, preds + n_preds);
  f->clauses.assign(clauses, clauses + n_clauses);
  f->iset.assign(iset, iset + n_iset);
  f->fset.assign(fset, fset + n_fset);
  if (int rc = f->d_preds.ensure(std::max<size_t>(1, n_preds) * sizeof(qg_pred))) return rc;
  if (int rc = f->d_clauses.ensure(std::max<size_t>(1, n_clauses) * sizeof(qg_clause))) return rc;
  if (int rc = f->d_iset.ensure(std::max<size_t>(1, n_iset) * 4)) return rc;
  if (int rc = f->d_fset.ensure(std::max<size_t>(1, n_fset) * 8)) return rc;
  if (n_preds) QG_CUDA_OK(cudaMemcpy(f->d_preds.p, preds, n_preds * sizeof(qg_pred), cudaMemcpyHostToDevice));
  if (n_clauses)
    QG_CUDA_OK(cudaMemcpy(f->d_clauses.p, clauses, n_clauses * sizeof(qg_clause), cudaMemcpyHostToDevice));
  if (n_iset) QG_CUDA_OK(cudaMemcpy(f->d_iset.p, iset, (size_t)n_iset * 4, cudaMemcpyHostToDevice));
  if (n_fset) QG_CUDA_OK(cudaMemcpy(f->d_fset.p, fset, (size_t)n_fset * 8, cudaMemcpyHostToDevice));
  *out = f.release();
  return 0;
}

int qg_filter_destroy(qg_filter* f) {
  if (!f) return 0;
  // the owner may already be destroyed (bindings release handles in any order): never look at it here
  if (cudaSetDevice(f->device) != cudaSuccess) (void)cudaGetLastError();
  f->d_preds.release(); f->d_clauses.release(); f->d_iset.release(); f->d_fset.release();
  f->raw_mask.release(); f->comb_mask.release(); f->gather_list.release();
  f->v_vec.release(); f->v_vec16.release(); f->v_inv.release(); f->v_n2.release(); f->v_ub.release();
  f->v_new2old.release(); f->v_map.release(); f->v_cnt.release(); f->v_woff.release(); f->v_tmp.release();
  delete f;
  return 0;
}

}  // extern "C"

namespace qg {

// Make every column referenced by the filter exist and cover all rows; refresh the pointer table.
static int prepare_columns(qg_index* idx, const qg_filter* f) {
  int max_field = -1;
  for (const qg_clause& c : f->clauses)
    if (c.op >= QG_OP_KIND_IN) max_field = std::max(max_field, c.field);
  if (max_field >= (int)idx->cols.size()) {
    idx->cols.resize(max_field + 1);
    idx->col_table_dirty = true;
  }
  const size_t cap = (size_t)std::max<long long>(idx->cap, 1);
  for (const qg_clause& c : f->clauses) {
    if (c.op < QG_OP_KIND_IN) continue;
    FacetColumn& col = idx->cols[c.field];
    if (!col.set || col.kind.bytes < cap) {
      // unknown field, or rows were appended after the column was set: those rows have no value
      DevBuf nk, nn, ns, nf;
      if (int rc = nk.ensure(cap)) return rc;
      if (int rc = nn.ensure(cap * 8)) return rc;
      if (int rc = ns.ensure(cap * 4)) return rc;
      if (int rc = nf.ensure(cap * 4)) return rc;
      QG_CUDA_OK(cudaMemset(nk.p, col.set ? QG_KIND_NOROW : QG_KIND_MISSING, nk.bytes));
      QG_CUDA_OK(cudaMemset(nn.p, 0, nn.bytes));
      QG_CUDA_OK(cudaMemset(ns.p, 0xff, ns.bytes));
      QG_CUDA_OK(cudaMemset(nf.p, 0xff, nf.bytes));
      if (col.set && col.n > 0) {
        QG_CUDA_OK(cudaMemcpy(nk.p, col.kind.p, (size_t)col.n, cudaMemcpyDeviceToDevice));
        QG_CUDA_OK(cudaMemcpy(nn.p, col.num.p, (size_t)col.n * 8, cudaMemcpyDeviceToDevice));
        QG_CUDA_OK(cudaMemcpy(ns.p, col.scode.p, (size_t)col.n * 4, cudaMemcpyDeviceToDevice));
        QG_CUDA_OK(cudaMemcpy(nf.p, col.fcode.p, (size_t)col.n * 4, cudaMemcpyDeviceToDevice));
      }
      col.kind.release(); col.num.release(); col.scode.release(); col.fcode.release();
      col.kind = nk; col.num = nn; col.scode = ns; col.fcode = nf;
      if (!col.set) col.synthetic = true;
      col.set = true;
      idx->col_table_dirty = true;
    }
  }
  if (idx->col_table_dirty) {
    std::vector<FacetColDev> tab(std::max<size_t>(1, idx->cols.size()));
    for (size_t i = 0; i < idx->cols.size(); ++i) {
      tab[i].kind = (const uint8_t*)idx->cols[i].kind.p;
      tab[i].num = (const double*)idx->cols[i].num.p;
      tab[i].scode = (const int32_t*)idx->cols[i].scode.p;
      tab[i].fcode = (const int32_t*)idx->cols[i].fcode.p;
      // only rows below the column's n can be arrays, and the element lists cover those
      const bool arr_ok = idx->cols[i].arr_rows >= 0;
      tab[i].arr_off = arr_ok ? (const int32_t*)idx->cols[i].arr_off.p : nullptr;
      tab[i].arr_code = arr_ok ? (const int32_t*)idx->cols[i].arr_code.p : nullptr;
    }
    if (int rc = idx->col_table.ensure(tab.size() * sizeof(FacetColDev))) return rc;
    QG_CUDA_OK(cudaMemcpy(idx->col_table.p, tab.data(), tab.size() * sizeof(FacetColDev), cudaMemcpyHostToDevice));
    idx->col_table_dirty = false;
  }
  return 0;
}

// Evaluate (or reuse) the raw predicate mask, then the mask combined with the live bits and, when
// the selectivity is low, the compacted row list for the gather scan.
static int filter_refresh(qg_index* idx, qg_filter* f, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(f->mu);
  const long long n = idx->n_rows;
  const size_t words = (size_t)((n + 31) / 32) + 1;
  DevBuf cnt;
  // QG_FILTER_TIMING=1: stage times of a mask (re)build on stderr (development aid)
  static const bool timing = [] { const char* e = std::getenv("QG_FILTER_TIMING"); return e && std::atoi(e) != 0; }();
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(st);
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "filter_refresh: %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  if (f->raw_rows != n || f->raw_facet_epoch != idx->facet_epoch) {
    // columns may be (re)built by another search's host layer: excluded until this mask has been evaluated
    std::lock_guard<std::mutex> cols_lock(idx->cols_mu);
    if (int rc = prepare_columns(idx, f)) return rc;
    lap("prepare_columns");
    if (int rc = f->raw_mask.ensure(words * 4)) return rc;
    if (int rc = cnt.ensure(8)) return rc;
    lap("allocate raw mask + counter");
    FilterProgDev prog{(const qg_pred*)f->d_preds.p, (int)f->preds.size(), (const qg_clause*)f->d_clauses.p,
                       (const int32_t*)f->d_iset.p, (const double*)f->d_fset.p};
    int rc = launch_filter_eval((const FacetColDev*)idx->col_table.p, prog, n, (uint32_t*)f->raw_mask.p,
                                (unsigned long long*)cnt.p, st);
    unsigned long long m = 0;
    if (!rc) {
      cudaError_t e = cudaMemcpyAsync(&m, cnt.p, 8, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) rc = fail(QG_ERR_CUDA, cudaGetErrorString(e));
    }
    if (rc) {
      cnt.release();
      return rc;
    }
    lap("filter_eval_kernel + count");
    f->raw_matches = (long long)m;
    f->raw_rows = n;
    f->raw_facet_epoch = idx->facet_epoch;
    f->comb_live_epoch = ~0ull;
  }
  if (f->comb_live_epoch != idx->live_epoch) {
    int rc = 0;
    unsigned long long m = 0;
    do {
      if ((rc = f->comb_mask.ensure(words * 4))) break;
      if ((rc = cnt.ensure(8))) break;
      if ((rc = launch_mask_and((const uint32_t*)f->raw_mask.p, idx->live, n, (uint32_t*)f->comb_mask.p,
                                (unsigned long long*)cnt.p, st)))
        break;
      cudaError_t e = cudaMemcpyAsync(&m, cnt.p, 8, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, cudaGetErrorString(e)); break; }
      lap("live AND + count (+ allocation)");
      f->comb_matches = (long long)m;
      f->gather_valid = false;
      // gather list when at most a quarter of the rows pass: the scan then reads only those rows
      if (f->comb_matches > 0 && f->comb_matches * 4 <= n) {
        if ((rc = f->gather_list.ensure((size_t)f->comb_matches * 4 + 64))) break;
        if ((rc = launch_mask_compact((const uint32_t*)f->comb_mask.p, n, (uint32_t*)f->gather_list.p,
                                      (unsigned long long*)cnt.p, st)))
          break;
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, cudaGetErrorString(e)); break; }
        f->gather_valid = true;
        lap("row list (allocation + compaction)");
      }
      f->comb_live_epoch = idx->live_epoch;
    } while (0);
    if (rc) {
      cnt.release();
      return rc;
    }
  }
  cnt.release();
  return 0;
}

// new2old[map[r]] = r for every passing row r
__global__ void __launch_bounds__(256) invert_map_kernel(const uint32_t* __restrict__ map, long long n,
                                                         uint32_t* __restrict__ new2old) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const uint32_t m = map[r];
    if (m != 0xFFFFFFFFu) new2old[m] = (uint32_t)r;
  }
}
// results of a search over a filter view carry view rows: back to index rows (+ the shard's row base)
__global__ void __launch_bounds__(256) translate_rows_kernel(long long* __restrict__ rows, uint64_t* __restrict__ keys,
                                                             long long n, const uint32_t* __restrict__ new2old,
                                                             long long row_base) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (rows != nullptr) {
      const long long r = rows[i];
      if (r >= 0) rows[i] = (long long)new2old[r] + row_base;
    }
    if (keys != nullptr) {
      const uint64_t k = keys[i];
      if (k != KEY_NONE) keys[i] = (k & 0xFFFFFFFF00000000ull) | (uint64_t)(uint32_t)((long long)new2old[(uint32_t)k] + row_base);
    }
  }
}

// Dense copy of the rows that pass `f` (filter AND live), in ascending row order — so ties between equal
// distances resolve exactly as over the whole index. A batch that would need many gather passes of the flat
// scan (each re-reading the passing rows through row-sized gathers) instead pays one ordered gather (every
// passing row read once, written once: the row move of qg_index_compact with the filter's mask in place of the
// live mask) and then runs the tensor-core scan over a dense, unmasked corpus. C3 (10M x 768, 10 % pass, 32
// queries): 8 gather passes = 9.4 ms before; copy ~1 ms + one dense pass ~0.5 ms. The copy is cached in the
// filter handle and reused until rows, tombstones or facet columns change. Caller holds f->mu.
static int build_filter_view(qg_index* idx, qg_filter* f, cudaStream_t st) {
  if (f->view_n >= 0 && f->view_rows == idx->n_rows && f->view_live_epoch == idx->live_epoch &&
      f->view_facet_epoch == idx->facet_epoch)
    return 0;
  f->view_n = -1;
  const long long n_old = idx->n_rows, n_new = f->comb_matches;
  const long long n_words = (n_old + 31) / 32;
  const long long ncap = (std::max<long long>(n_new, 1024) + 1023) & ~1023ll;
  int rc = 0;
  if ((rc = f->v_cnt.ensure((size_t)n_words * 4)) || (rc = f->v_woff.ensure((size_t)(n_words + 1) * 4)) ||
      (rc = f->v_tmp.ensure((size_t)scan_tmp_words(std::max(n_words, ncap)) * 4)) ||
      (rc = f->v_map.ensure((size_t)n_old * 4)) || (rc = f->v_new2old.ensure((size_t)ncap * 4)) ||
      (rc = f->v_vec.ensure((size_t)ncap * idx->dp * 4)) || (rc = f->v_inv.ensure((size_t)ncap * 4)) ||
      (rc = f->v_n2.ensure((size_t)ncap * 4)) || (rc = f->v_ub.ensure((size_t)ncap * 4)) ||
      (idx->use_bf16 && (rc = f->v_vec16.ensure((size_t)ncap * idx->dp16 * 2))))
    return rc;
  const uint32_t* mask = (const uint32_t*)f->comb_mask.p;
  if ((rc = launch_word_popcount(mask, n_words, (uint32_t*)f->v_cnt.p, st)) ||
      (rc = launch_exclusive_scan_u32((const uint32_t*)f->v_cnt.p, (uint32_t*)f->v_woff.p, n_words, (uint32_t*)f->v_tmp.p, st)) ||
      (rc = launch_compact_map(mask, (const uint32_t*)f->v_woff.p, n_old, (uint32_t*)f->v_map.p, st)) ||
      (rc = launch_fill_f32((float*)f->v_n2.p, ncap, INFINITY, st)) ||
      (rc = launch_fill_f32((float*)f->v_ub.p, ncap, INFINITY, st)))
    return rc;
  QG_CUDA_OK(cudaMemsetAsync(f->v_inv.p, 0, (size_t)ncap * 4, st));
  CompactRowsArgs a{};
  a.map = (const uint32_t*)f->v_map.p;
  a.n_rows = n_old;
  a.vec = idx->vec; a.vec_out = (float*)f->v_vec.p; a.dp = idx->dp;
  a.vec16 = idx->use_bf16 ? idx->vec16 : nullptr; a.vec16_out = f->v_vec16.p; a.dp16 = idx->dp16;
  a.inv_norm = idx->inv_norm; a.inv_norm_out = (float*)f->v_inv.p;
  a.norm2 = idx->norm2; a.norm2_out = (float*)f->v_n2.p;
  a.unit_bias = idx->unit_bias; a.unit_bias_out = (float*)f->v_ub.p;
  if ((rc = launch_compact_rows(a, idx->sm_count, st))) return rc;
  {
    long long blocks = std::min<long long>((n_old + 255) / 256, 148 * 8);
    invert_map_kernel<<<(int)std::max<long long>(blocks, 1), 256, 0, st>>>((const uint32_t*)f->v_map.p, n_old,
                                                                          (uint32_t*)f->v_new2old.p);
    QG_CUDA_OK(cudaGetLastError());
  }
  f->view_n = n_new;
  f->view_rows = idx->n_rows;
  f->view_live_epoch = idx->live_epoch;
  f->view_facet_epoch = idx->facet_epoch;
  return 0;
}

__global__ void fill_empty_kernel(float* dist, float* negdist, long long* row, int* count, uint64_t* keys,
                                  long long n, int nq) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (dist) dist[i] = __int_as_float(0x7f800000);
    if (negdist) negdist[i] = __int_as_float(0x7f800000);
    if (row) row[i] = -1;
    if (keys) keys[i] = KEY_NONE;
    if (count && i < nq) count[i] = 0;
  }
}

struct SearchArgs {
  const float* d_queries;  // [q x dim] contiguous, device
  int q, k;
  qg_filter* filter;
  const float* d_neg;  // [q x dim] or nullptr
  float* d_dist;
  float* d_negdist;
  long long* d_row;
  int* d_count;
  uint64_t* d_keys;  // shard mode
  long long row_base;
  bool no_tc = false;  // the flat re-run of a query the tensor-core regime could not certify
};

static void publish_stats(qg_index* idx, const qg_scan_stats& st) {
  g_my_stats = st;
  std::lock_guard<std::mutex> lk(idx->stats_mu);
  idx->stats = st;
}

// Enqueue the whole search on `st` using workspace `w`. Validation already done.
static int search_enqueue(qg_index* idx, Workspace* w, const SearchArgs& a, cudaStream_t st) {
  const int q = a.q, k = a.k, d = idx->dim, dp = idx->dp;
  qg_scan_stats stats{};
  const uint32_t* mask = nullptr;
  const uint32_t* gather = nullptr;
  long long n_pass = idx->n_live;
  long long n_items = idx->n_rows;
  if (a.filter) {
    if (int rc = filter_refresh(idx, a.filter, st)) return rc;
    mask = (const uint32_t*)a.filter->comb_mask.p;
    n_pass = a.filter->comb_matches;
    if (a.filter->gather_valid) {
      gather = (const uint32_t*)a.filter->gather_list.p;
      n_items = n_pass;
    }
  } else if (idx->n_live < idx->n_rows) {
    mask = idx->live;
  }

  const long long out_n = (long long)q * k;
  if (n_pass == 0) {
    fill_empty_kernel<<<64, 256, 0, st>>>(a.d_dist, a.d_negdist, a.d_row, a.d_count, a.d_keys,
                                          std::max<long long>(out_n, q), q);
    QG_CUDA_OK(cudaGetLastError());
    publish_stats(idx, stats);
    return 0;
  }

  // padded queries / negatives
  const float* qpad = a.d_queries;
  const float* negpad = a.d_neg;
  if (dp != d) {
    if (int rc = w->qpad.ensure((size_t)q * dp * 4)) return rc;
    QG_CUDA_OK(cudaMemsetAsync(w->qpad.p, 0, (size_t)q * dp * 4, st));
    QG_CUDA_OK(cudaMemcpy2DAsync(w->qpad.p, (size_t)dp * 4, a.d_queries, (size_t)d * 4, (size_t)d * 4, (size_t)q,
                                 cudaMemcpyDeviceToDevice, st));
    qpad = (const float*)w->qpad.p;
    if (a.d_neg) {
      if (int rc = w->negpad.ensure((size_t)q * dp * 4)) return rc;
      QG_CUDA_OK(cudaMemsetAsync(w->negpad.p, 0, (size_t)q * dp * 4, st));
      QG_CUDA_OK(cudaMemcpy2DAsync(w->negpad.p, (size_t)dp * 4, a.d_neg, (size_t)d * 4, (size_t)d * 4, (size_t)q,
                                   cudaMemcpyDeviceToDevice, st));
      negpad = (const float*)w->negpad.p;
    }
  }

  const long long keff = std::min<long long>(k, n_pass);
  int kp = 32;
  while (kp < keff + idx->margin && kp < 2048) kp <<= 1;

  if (kp > 1024) {
    // exhaustive path, one query at a time (the reference's own algorithm, exact.go:114-129)
    for (int i = 0; i < q; ++i) {
      if (int rc = exhaustive_search(w->ex, idx->vec, idx->n_rows, dp, d, mask, qpad + (size_t)i * dp,
                                     negpad ? negpad + (size_t)i * dp : nullptr, idx->metric, idx->arith, k,
                                     a.d_dist ? a.d_dist + (size_t)i * k : nullptr,
                                     a.d_negdist ? a.d_negdist + (size_t)i * k : nullptr,
                                     a.d_row ? a.d_row + (size_t)i * k : nullptr, a.d_count ? a.d_count + i : nullptr,
                                     a.d_keys ? a.d_keys + (size_t)i * k : nullptr, a.row_base, st))
        return rc;
      stats.kernel_launches += 3;
    }
    stats.path = 0;
    stats.passes = q;
    stats.queries_per_pass = 1;
    stats.rows_scanned = idx->n_rows;
    stats.bytes_algorithmic = n_pass * (long long)d * 4;
    publish_stats(idx, stats);
    return 0;
  }

  const int mode = scan_mode_of(idx->metric);

  // ---- tensor-core regime: large query batches over a dense (possibly masked) corpus ----------------
  TcPlan plan{};
  const bool tc_possible = !a.no_tc && q >= idx->tc_min_q && mode != MODE_L1 && kp <= 128 && idx->n_rows >= idx->tc_min_rows &&
                           n_pass >= idx->tc_min_rows && tc_available() == 0 &&
                           tc_plan(dp, d + tc_extra_cols(d, mode == MODE_L2), q, idx->use_bf16, &plan) == 0;
  // the arrays the tensor-core scan and its finalize read: the index's own, or the filter's dense view
  struct RowArrays {
    const float* vec;
    const float* inv_norm;
    const float* norm2;
    const float* unit_bias;
    const void* vec16;
    long long n_rows;
  } R{idx->vec, idx->inv_norm, idx->norm2, idx->unit_bias, idx->vec16, idx->n_rows};
  const uint32_t* view_new2old = nullptr;
  if (tc_possible && gather != nullptr) {
    // A selective filter has a compacted row list. Three ways to serve the batch, costed in row reads:
    //  - flat gather scan: only the matching rows, but at most max_qb queries per pass and through row-sized
    //    gathers (counted at 1.2x: 5.9 TB/s measured on 10M x 768 against 6.8 TB/s for a dense stream);
    //  - masked tensor-core scan: every row of the index once per plan.n_cols queries;
    //  - dense VIEW of the matching rows (build_filter_view: read + write once, cached in the filter) and the
    //    tensor-core scan over it.
    const int flat_qb = scan_fast_supported(dp) ? scan_fast_max_qb(dp) : 8;
    const double tc_passes = (double)((q + plan.n_cols - 1) / plan.n_cols);
    const double flat_cost = 1.2 * (double)((q + flat_qb - 1) / flat_qb) * (double)n_pass;
    const double tc_cost = tc_passes * (double)idx->n_rows;
    const bool view_cached = a.filter->view_n >= 0 && a.filter->view_rows == idx->n_rows &&
                             a.filter->view_live_epoch == idx->live_epoch && a.filter->view_facet_epoch == idx->facet_epoch;
    const double view_cost = (view_cached ? 0.0 : 2.5 * (double)n_pass) + tc_passes * (double)n_pass;
    static const bool view_on = [] { const char* e = std::getenv("QG_FILTER_VIEW"); return e == nullptr || std::atoi(e) != 0; }();
    if (view_on && view_cost < flat_cost && view_cost < tc_cost && n_pass >= idx->tc_min_rows) {
      {
        std::lock_guard<std::mutex> flk(a.filter->mu);
        if (int rc = build_filter_view(idx, a.filter, st)) return rc;
      }
      R = RowArrays{(const float*)a.filter->v_vec.p, (const float*)a.filter->v_inv.p, (const float*)a.filter->v_n2.p,
                    (const float*)a.filter->v_ub.p, idx->use_bf16 ? a.filter->v_vec16.p : nullptr, n_pass};
      view_new2old = (const uint32_t*)a.filter->v_new2old.p;
      gather = nullptr;
      mask = nullptr;
      n_items = n_pass;
    } else if (tc_cost < flat_cost) {
      gather = nullptr;
      n_items = idx->n_rows;
    }
  }
  if (tc_possible && !gather) {
    const long long n_sample = tc_sample_tiles(plan, R.n_rows, k);
    // TS variant: the sample + threshold stage runs once for all passes of the search
    const int n_pass_tc = (q + plan.n_cols - 1) / plan.n_cols;
    const bool hoist = plan.variant == 1 && n_pass_tc > 1;
    const size_t q_slots = hoist ? (size_t)n_pass_tc * plan.n_cols : (size_t)plan.n_cols;
    if (int rc = w->tc_sample.ensure(q_slots * n_sample * plan.sample_vals * 4)) return rc;
    if (int rc = w->tc_tau.ensure(std::max<size_t>(q_slots, TC_MAX_COLS) * 4)) return rc;
    // hoisted searches finalize TC_PASS_GROUP passes per launch: candidate lists, counters and tile-claim
    // counters exist once per pass of a group
    const int group = hoist ? TC_PASS_GROUP : 1;
    const size_t cnt_slots = (size_t)group * TC_MAX_COLS;  // work counters follow at [cnt_slots ..)
    if (int rc = w->tc_cand.ensure((size_t)group * plan.n_cols * TC_CAND_CAP * 8)) return rc;
    if (int rc = w->tc_cnt.ensure((cnt_slots + 8) * 4)) return rc;
    // raw scan: bf16 stream without a mask — the norm columns of the bf16 rows make the MMA output the score
    const bool tc_raw = plan.variant == 1 && plan.bf16 && mask == nullptr && R.n_rows >= plan.tile_rows &&
                        tc_raw_supported(d, mode == MODE_L2);
    if (plan.variant == 1) {
      // all queries of the search in tensor-memory order, one block per pass
      if (int rc = w->tc_apack.ensure(tc_pack_bytes(plan, q))) return rc;
      if (idx->profiling) w->prof_begin(2, st);
      if (int rc = launch_tc_pack(plan, qpad, q, dp, d, tc_raw ? tc_extra_cols(d, mode == MODE_L2) : 0,
                                  idx->metric == METRIC_COSINE, w->tc_apack.p, st))
        return rc;
      if (idx->profiling) w->prof_end(st);
      stats.kernel_launches++;
    }
    // TS variant: additive row term with the mask folded in (+inf = excluded)
    const float* tc_bias = mode == MODE_L2 ? R.norm2 : R.unit_bias;
    if (plan.variant == 1 && mask != nullptr) {
      const long long n_pad = (R.n_rows + 127) & ~127ll;
      if (int rc = w->tc_bias.ensure((size_t)n_pad * 4)) return rc;
      if (int rc = launch_tc_bias(mask, mode == MODE_L2 ? R.norm2 : nullptr, R.n_rows, n_pad,
                                  (float*)w->tc_bias.p, st))
        return rc;
      stats.kernel_launches++;
      tc_bias = (const float*)w->tc_bias.p;
    }
    FinalizeCandParams cp{};
    cp.cand = (const uint64_t*)w->tc_cand.p;
    cp.cand_cnt = (const int*)w->tc_cnt.p;
    cp.cap = TC_CAND_CAP;
    cp.tau = (const float*)w->tc_tau.p;
    cp.kp = kp;
    // |tensor-core dot - exact dot| <= tc_gamma * |q||x|: both operands truncated to tf32 (2^-10 each) or
    // rounded to bf16 (2^-9 each), plus the fp32 accumulation
    cp.tc_gamma = (plan.bf16 ? 1.02 / 256.0 : 1.02 / 512.0) + (double)d / 4194304.0;
    cp.tc_norm_gamma = (tc_raw && mode == MODE_L2) ? (double)(d + 16) / 4194304.0 : 0.0;
    FinalizeParams& fb = cp.base;
    fb.vec = R.vec;
    fb.dp = dp;
    fb.d = d;
    fb.metric = idx->metric;
    fb.arith = idx->arith;
    fb.mode = mode;
    fb.cosine = idx->metric == METRIC_COSINE;
    fb.k = k;
    fb.gamma = (float)((d + 16) * 5.9604645e-8);
    fb.max_norm2 = idx->max_norm2;
    fb.row_base = view_new2old ? 0 : a.row_base;  // view rows are translated (and based) after the finalize
    // profiling: kind 2 = sample + threshold kernels, kind 0 = main scan kernel
    TcStageHook hook{[](void* ctx, int stage, int begin, cudaStream_t s) {
                       Workspace* ws = static_cast<Workspace*>(ctx);
                       if (begin) ws->prof_begin(stage == 0 ? 2 : 0, s);
                       else ws->prof_end(s);
                     },
                     w};
    if (hoist) {
      cp.reset_cnt = (int*)w->tc_cnt.p;
      cp.reset_work = (int*)w->tc_cnt.p + cnt_slots;
      cp.n_reset_work = group;
    }
    int pass_no = 0;  // pass index within the search
    for (int p0 = hoist ? -1 : 0; p0 < q; p0 = p0 < 0 ? 0 : p0 + plan.n_cols) {
      // p0 == -1: the sample + threshold stage of every pass at once (hoisted out of the loop)
      const bool all = p0 < 0;
      const int q0 = all ? 0 : p0;
      const int nq = all ? q : std::min(plan.n_cols, q - p0);
      TcArgs ta{};
      ta.sample_only = all ? 1 : 0;
      ta.presampled = hoist ? 1 : 0;
      ta.vec = R.vec;
      ta.vec16 = R.vec16;
      ta.dp16 = idx->dp16;
      ta.n_rows = R.n_rows;
      ta.dp = dp;
      ta.row_norm2 = R.norm2;
      ta.inv_norm = idx->metric == METRIC_COSINE ? R.inv_norm : nullptr;
      ta.mask = mask;
      ta.bias = tc_bias;
      ta.queries = qpad + (size_t)q0 * dp;
      ta.apack = plan.variant == 1 ? (const char*)w->tc_apack.p + tc_pack_bytes(plan, q0) : nullptr;
      ta.raw = tc_raw;
      const int slot = all ? 0 : pass_no % group;  // this pass's lists inside its group
      ta.work_counter = (int*)w->tc_cnt.p + cnt_slots + slot;
      ta.n_cnt = group * plan.n_cols;
      ta.nq = nq;
      ta.mode = mode;
      ta.cosine = fb.cosine;
      ta.sample = (uint32_t*)w->tc_sample.p;
      ta.n_sample = (int)n_sample;
      ta.sample_rank = tc_sample_rank(k);
      ta.tau = (float*)w->tc_tau.p + (hoist ? q0 : 0);
      ta.cand = (uint64_t*)w->tc_cand.p + (size_t)slot * plan.n_cols * TC_CAND_CAP;
      ta.cand_cnt = (int*)w->tc_cnt.p + (size_t)slot * plan.n_cols;
      const int rc = launch_tc_pass(plan, ta, idx->sm_count, st, &stats.kernel_launches,
                                    idx->profiling ? &hook : nullptr);
      if (rc) return rc;
      if (all) continue;
      stats.passes++;
      ++pass_no;
      // finalize once per group of passes (a partial pass can only be the last one of the search, so the
      // queries of a group are contiguous: the lists of its passes sit one after the other)
      const bool group_done = pass_no % group == 0 || p0 + plan.n_cols >= q;
      if (!group_done) continue;
      const int g0 = (pass_no - 1) / group * group * plan.n_cols;  // first query of the group
      const int gq = p0 + nq - g0;                                  // queries in the group
      cp.tau = (const float*)w->tc_tau.p + (hoist ? g0 : 0);
      fb.queries = qpad + (size_t)g0 * dp;
      fb.negatives = negpad ? negpad + (size_t)g0 * dp : nullptr;
      fb.out_dist = a.d_dist ? a.d_dist + (size_t)g0 * k : nullptr;
      fb.out_negdist = a.d_negdist ? a.d_negdist + (size_t)g0 * k : nullptr;
      fb.out_row = a.d_row ? a.d_row + (size_t)g0 * k : nullptr;
      fb.out_count = a.d_count ? a.d_count + g0 : nullptr;
      fb.out_keys = a.d_keys ? a.d_keys + (size_t)g0 * k : nullptr;
      const bool prof2 = idx->profiling && w->prof_begin(1, st) == 0;
      const int frc = launch_finalize_cand(cp, gq, st);
      if (prof2) w->prof_end(st);
      if (frc) return frc;
      stats.kernel_launches++;
    }
    if (view_new2old != nullptr) {
      const long long total = (long long)q * k;
      const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 4);
      translate_rows_kernel<<<blocks, 256, 0, st>>>(a.d_row, a.d_keys, total, view_new2old, a.row_base);
      QG_CUDA_OK(cudaGetLastError());
      stats.kernel_launches++;
    }
    stats.path = 3;
    stats.queries_per_pass = plan.n_cols;
    stats.rows_scanned = R.n_rows;
    stats.bytes_algorithmic = R.n_rows * (long long)(plan.bf16 ? idx->dp16 * 2 : d * 4) +
                              (tc_raw ? 0 : R.n_rows * 4) + (mask ? R.n_rows / 8 : 0);
    stats.reserved = plan.bf16;
    publish_stats(idx, stats);
    return 0;
  }

  const bool fast = scan_fast_supported(dp) && mode != MODE_L1;
  // scans over the fast dimensions take the dense kernel; QG_SCAN_DENSE=0 keeps the round-1 fast kernel
  static const bool dense_on = [] { const char* e = std::getenv("QG_SCAN_DENSE"); return e == nullptr || std::atoi(e) != 0; }();
  // row lists (gather) take it too unless QG_SCAN_DENSE_GATHER=0 (then the round-1 fast kernel)
  static const bool dense_gather_on = [] { const char* e = std::getenv("QG_SCAN_DENSE_GATHER"); return e == nullptr || std::atoi(e) != 0; }();
  const bool dense = fast && dense_on && (gather == nullptr || dense_gather_on);
  int tile_rows, nw = SCAN_NW, stages = 4, max_qb;
  if (fast) {
    tile_rows = scan_fast_tile_rows(dp);
    max_qb = scan_fast_max_qb(dp);
  } else {
    const int row_bytes = dp * 4;
    tile_rows = std::max(1, std::min(32, 4096 / row_bytes));
    const size_t tile_bytes = (size_t)tile_rows * row_bytes;
    // ring budget ~160 KB: shrink stages, then warps, for very wide rows
    while (nw > 1 && (size_t)nw * 2 * tile_bytes > 160 * 1024) nw >>= 1;
    stages = (int)std::min<size_t>(4, (160 * 1024) / ((size_t)nw * tile_bytes));
    if (stages < 2) return fail(QG_ERR_UNSUPPORTED, "dimension too large for the scan kernels (max 16384)");
    max_qb = 8;
    while (max_qb > 1 && (size_t)max_qb * dp * 4 > 32 * 1024) max_qb >>= 1;
  }
  int qb = 1;
  if (dense) {
    // ring and warps depend on the query block: the largest block (<= the batch) whose pools fit
    int want = 1;
    while (want < max_qb && want < q) want <<= 1;
    size_t smem = 0;
    for (qb = want; qb >= 1; qb >>= 1) {
      if (scan_dense_geometry(dp, qb, kp, &tile_rows, &nw, &smem) == 0 && smem <= 227 * 1024) break;
    }
    if (qb < 1) return fail(QG_ERR_UNSUPPORTED, "scan: candidate pools do not fit shared memory");
  } else {
    // pools must fit next to the ring
    const size_t ring = fast ? (size_t)0 : (size_t)nw * stages * tile_rows * dp * 4;
    const size_t ring_fast = fast ? (size_t)scan_fast_ring_bytes(dp) : 0;
    const size_t avail = 225 * 1024 - (fast ? ring_fast : ring) - (fast ? 0 : (size_t)max_qb * dp * 4);
    while (max_qb > 1 && (size_t)max_qb * pool_slots(kp) * 8 > avail) max_qb >>= 1;
    while (qb < max_qb && qb < q) qb <<= 1;
  }

  const long long n_tiles = (n_items + tile_rows - 1) / tile_rows;
  int nb = (int)std::min<long long>(idx->sm_count, (n_tiles + nw - 1) / nw);
  nb = std::max(nb, 1);

  // finalize in chunks so the partial buffer stays bounded
  int qchunk = std::max(qb, std::min(q, std::max(8, 16384 / kp)));
  qchunk = (qchunk / qb) * qb;
  if (int rc = w->partial.ensure((size_t)qchunk * nb * kp * 8)) return rc;

  ScanParams sp{};
  sp.vec = idx->vec;
  sp.inv_norm = idx->metric == METRIC_COSINE ? idx->inv_norm : nullptr;
  sp.mask = mask;
  sp.gather = gather;
  sp.n_items = n_items;
  sp.kp = kp;
  sp.cosine = idx->metric == METRIC_COSINE;
  sp.dp = dp;
  sp.tile_rows = tile_rows;
  sp.stages = stages;
  // pacing load of the dense kernel (scan.cuh); QG_SCAN_PACE=0 turns it off
  static const bool scan_pace = [] { const char* e = std::getenv("QG_SCAN_PACE"); return e == nullptr || std::atoi(e) != 0; }();
  sp.pace = (scan_pace && dense && sp.inv_norm == nullptr) ? idx->norm2 : nullptr;
  static const int scan_dbg = [] { const char* e = std::getenv("QG_SCAN_DBG"); return e ? std::atoi(e) : 0; }();
  sp.dbg = scan_dbg;
  // QG_SCAN_TRACE=1: per-CTA time stamps of the dense kernel, summarised on stderr after every launch (diagnostic)
  static const bool scan_trace = [] { const char* e = std::getenv("QG_SCAN_TRACE"); return e && std::atoi(e) != 0; }();
  static unsigned long long* trace_buf = nullptr;
  if (scan_trace && trace_buf == nullptr) cudaMalloc(&trace_buf, 1024 * 32 * sizeof(unsigned long long));
  sp.trace = scan_trace ? trace_buf : nullptr;
  // threshold exchange between the CTAs of a dense launch (QG_SCAN_XCHG=0 turns it off)
  static const bool xchg_on = [] { const char* e = std::getenv("QG_SCAN_XCHG"); return e == nullptr || std::atoi(e) != 0; }();
  constexpr size_t kXchgBytes = (size_t)XCHG_MAX_QB * XCHG_MAX_CTAS * 16 * sizeof(unsigned long long);
  const bool use_xchg = dense && xchg_on && kp <= XCHG_MAX_KP && qb <= XCHG_MAX_QB && nb <= XCHG_MAX_CTAS && nw == 16;
  if (use_xchg && w->xchg.p == nullptr) {
    if (int rc = w->xchg.ensure(kXchgBytes)) return rc;
    QG_CUDA_OK(cudaMemsetAsync(w->xchg.p, 0, kXchgBytes, st));
  }

  FinalizeParams fp{};
  fp.nb = nb;
  fp.kp = kp;
  fp.vec = idx->vec;
  fp.dp = dp;
  fp.d = d;
  fp.metric = idx->metric;
  fp.arith = idx->arith;
  fp.mode = mode;
  fp.cosine = sp.cosine;
  fp.k = k;
  fp.gamma = (float)((d + 16) * 5.9604645e-8);
  fp.max_norm2 = idx->max_norm2;
  fp.row_base = a.row_base;

  for (int q0 = 0; q0 < q; q0 += qchunk) {
    const int qn = std::min(qchunk, q - q0);
    for (int p0 = 0; p0 < qn; p0 += qb) {
      sp.queries = qpad + (size_t)(q0 + p0) * dp;
      sp.nq = std::min(qb, qn - p0);
      sp.partial = (uint64_t*)w->partial.p + (size_t)p0 * nb * kp;
      if (sp.trace) {
        std::vector<unsigned long long> init((size_t)nb * 32, 0ull);
        for (int c = 0; c < nb; ++c) init[(size_t)c * 32 + 5] = ~0ull;
        cudaMemcpyAsync(sp.trace, init.data(), init.size() * 8, cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);
      }
      if (use_xchg) {
        if (++w->xchg_epoch == 0) w->xchg_epoch = 1;
        sp.xchg = (unsigned long long*)w->xchg.p;
        sp.epoch = w->xchg_epoch;
      }
      const bool prof = idx->profiling && w->prof_begin(0, st) == 0;
      int rc = dense ? launch_scan_dense(dp, qb, mode, sp, nb, st)
                     : fast ? launch_scan_fast(dp, qb, mode, sp, nb, st) : launch_scan_generic(qb, mode, sp, nb, nw, st);
      if (prof) w->prof_end(st);
      if (rc) return rc;
      if (sp.trace && dense) {
        std::vector<unsigned long long> tr((size_t)nb * 32);
        cudaStreamSynchronize(st);
        cudaMemcpy(tr.data(), sp.trace, tr.size() * 8, cudaMemcpyDeviceToHost);
        unsigned long long t0 = ~0ull, t_end = 0, pro = 0, first = 0, loop_last = 0, loop_first = ~0ull, start_last = 0;
        for (int c = 0; c < nb; ++c) {
          const unsigned long long* r = &tr[(size_t)c * 32];
          t0 = std::min(t0, r[0]);
          start_last = std::max(start_last, r[0]);
          t_end = std::max(t_end, r[4]);
          pro = std::max(pro, r[1] - r[0]);
          first = std::max(first, r[2] - r[1]);
          loop_last = std::max(loop_last, r[3]);
          loop_first = std::min(loop_first, r[5]);
        }
        {
          float ext_lo = INFINITY, ext_hi = -INFINITY, tau_lo = INFINITY, tau_hi = -INFINITY;
          int pr_max = 0, cnt_max = 0;
          long long cnt_sum = 0;
          for (int c = 0; c < nb; ++c) {
            const unsigned long long a6 = tr[(size_t)c * 32 + 6], a7 = tr[(size_t)c * 32 + 7];
            float e, t;
            const uint32_t eb = (uint32_t)a6, tb = (uint32_t)a7;
            std::memcpy(&e, &eb, 4);
            std::memcpy(&t, &tb, 4);
            ext_lo = std::min(ext_lo, e); ext_hi = std::max(ext_hi, e);
            tau_lo = std::min(tau_lo, t); tau_hi = std::max(tau_hi, t);
            pr_max = std::max(pr_max, (int)(a6 >> 32));
            cnt_max = std::max(cnt_max, (int)(a7 >> 32));
            cnt_sum += (int)(a7 >> 32);
          }
          double intra = 0;
          unsigned long long cta_first = ~0ull;
          for (int c = 0; c < nb; ++c) {
            intra += (double)(tr[(size_t)c * 32 + 3] - tr[(size_t)c * 32 + 5]);
            cta_first = std::min(cta_first, tr[(size_t)c * 32 + 3]);
          }
          for (int c = 0; c < nb; c += 37) {
            const unsigned long long* r = &tr[(size_t)c * 32];
            std::fprintf(stderr, "scan trace: CTA %3d warps (tiles @ us):", c);
            for (int wi = 0; wi < 16; ++wi)
              if (r[16 + wi]) std::fprintf(stderr, " %llu@%.0f", r[16 + wi] >> 48, (double)((r[16 + wi] & 0xffffffffffffull) - (t0 & 0xffffffffffffull)) * 1e-3);
            std::fprintf(stderr, "\n");
            std::fprintf(stderr, "scan trace: CTA %3d fetches (tile, pool, seen):", c);
            for (int f = 0; f < 4; ++f)
              std::fprintf(stderr, " (%llu, %llu, %llu)", r[8 + f] >> 48, (r[8 + f] >> 24) & 0xffffff, r[8 + f] & 0xffffff);
            std::fprintf(stderr, " threshold at %.1f us, prunes %llu, pool %llu\n", r[12] ? (r[12] - t0) * 1e-3 : -1.0, r[6] >> 32, r[7] >> 32);
          }
          std::fprintf(stderr, "scan trace: first..last warp within a CTA %.1f us on average | first CTA done at %.1f us\n",
                       intra / nb * 1e-3, (cta_first - t0) * 1e-3);
          std::fprintf(stderr, "scan trace: tau_ext %g..%g | own tau %g..%g | prunes max %d | pool at the end max %d, mean %.0f\n",
                       ext_lo, ext_hi, tau_lo, tau_hi, pr_max, cnt_max, (double)cnt_sum / nb);
        }
        std::fprintf(stderr,
                     "scan trace: span %.1f us | CTA starts spread %.1f | prologue max %.1f | first tile max %.1f | "
                     "first warp done at %.1f, last at %.1f | finish after last warp %.1f\n",
                     (t_end - t0) * 1e-3, (start_last - t0) * 1e-3, pro * 1e-3, first * 1e-3, (loop_first - t0) * 1e-3,
                     (loop_last - t0) * 1e-3, (t_end - loop_last) * 1e-3);
      }
      stats.kernel_launches++;
      stats.passes++;
    }
    fp.partial = (const uint64_t*)w->partial.p;
    fp.queries = qpad + (size_t)q0 * dp;
    fp.negatives = negpad ? negpad + (size_t)q0 * dp : nullptr;
    fp.out_dist = a.d_dist ? a.d_dist + (size_t)q0 * k : nullptr;
    fp.out_negdist = a.d_negdist ? a.d_negdist + (size_t)q0 * k : nullptr;
    fp.out_row = a.d_row ? a.d_row + (size_t)q0 * k : nullptr;
    fp.out_count = a.d_count ? a.d_count + q0 : nullptr;
    fp.out_keys = a.d_keys ? a.d_keys + (size_t)q0 * k : nullptr;
    const bool prof = idx->profiling && w->prof_begin(1, st) == 0;
    const int frc = launch_finalize(fp, qn, st);
    if (prof) w->prof_end(st);
    if (frc) return frc;
    stats.kernel_launches++;
  }
  stats.path = gather ? 2 : 1;
  stats.queries_per_pass = qb;
  stats.rows_scanned = n_items;
  stats.bytes_algorithmic = n_items * (long long)d * 4 + (mask && !gather ? idx->n_rows / 8 : 0) +
                            ((sp.inv_norm || sp.pace) ? n_items * 4 : 0) + (gather ? n_items * 4 : 0);
  publish_stats(idx, stats);
  return 0;
}

// Common argument checks in the reference's order (exact.go:96-106). Returns 1 when the index is
// empty (zero results, no error), 0 to proceed, <0 never; errors are returned via *err.
static int validate_search(qg_index* idx, int q, int dim, int k, int* err) {
  *err = 0;
  if (q < 0) {
    *err = fail(QG_ERR_INVALID, "negative query count");
    return 0;
  }
  if (idx->n_live == 0) return 1;
  if (dim != idx->dim) {
    *err = fail(QG_ERR_DIM, "query dimension mismatch: expected " + std::to_string(idx->dim) + ", got " +
                                std::to_string(dim));
    return 0;
  }
  if (k <= 0) {
    *err = fail(QG_ERR_K, "k must be positive");
    return 0;
  }
  return 0;
}

}  // namespace qg

extern "C" {

int qg_filter_eval(qg_index* idx, qg_filter* f, uint64_t* mask_out, int64_t* out_matches) {
  if (int rc = check_index(idx)) return rc;
  if (!f || f->owner != idx) return fail(QG_ERR_INVALID, "filter does not belong to this index");
  Workspace* w = ws_acquire(idx);
  if (!w) return fail(QG_ERR_CUDA, "could not create a stream");
  int rc = filter_refresh(idx, f, w->stream);
  if (!rc && mask_out) {
    const size_t words64 = (size_t)((idx->n_rows + 63) / 64);
    const size_t bytes32 = (size_t)((idx->n_rows + 31) / 32) * 4;
    std::memset(mask_out, 0, words64 * 8);
    cudaError_t e = cudaMemcpy(mask_out, f->raw_mask.p, bytes32, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(QG_ERR_CUDA, cudaGetErrorString(e));
  }
  if (!rc && out_matches) *out_matches = f->raw_matches;
  ws_release(idx, w);
  return rc;
}

int qg_search_batch_device(qg_index* idx, const void* d_queries, int q, int dim, int k, qg_filter* filter,
                           const void* d_negatives, void* d_out_dist, void* d_out_negdist, void* d_out_row,
                           void* d_out_count, void* stream) {
  if (int rc = check_index(idx)) return rc;
  int err = 0;
  const int empty = validate_search(idx, q, dim, k, &err);
  if (err) return err;
  if (q == 0) return 0;
  if (filter && filter->owner != idx) return fail(QG_ERR_INVALID, "filter does not belong to this index");
  cudaStream_t st = (cudaStream_t)stream;
  if (empty) {
    if (k > 0) {
      fill_empty_kernel<<<64, 256, 0, st>>>((float*)d_out_dist, (float*)d_out_negdist, (long long*)d_out_row,
                                            (int*)d_out_count, nullptr, std::max<long long>((long long)q * k, q), q);
    } else {
      fill_empty_kernel<<<1, 256, 0, st>>>(nullptr, nullptr, nullptr, (int*)d_out_count, nullptr, q, q);
    }
    QG_CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (!d_queries || !d_out_dist || !d_out_row || !d_out_count) return fail(QG_ERR_INVALID, "null device buffer");
  if (d_negatives && !d_out_negdist) return fail(QG_ERR_INVALID, "negatives given without out_negdist");
  Workspace* w = ws_acquire(idx, st, true);
  if (!w) return fail(QG_ERR_CUDA, "could not create a stream");
  SearchArgs a{(const float*)d_queries, q, k, filter, (const float*)d_negatives, (float*)d_out_dist,
               d_negatives ? (float*)d_out_negdist : nullptr, (long long*)d_out_row, (int*)d_out_count, nullptr, 0};
  const int rc = search_enqueue(idx, w, a, st);
  ws_release_async(idx, w, st);
  return rc;
}

int qg_search_shard_keys_device(qg_index* idx, const void* d_queries, int q, int dim, int k, qg_filter* filter,
                                int64_t row_base, void* d_out_keys, void* stream) {
  if (int rc = check_index(idx)) return rc;
  if (q < 0) return fail(QG_ERR_INVALID, "negative query count");
  if (k <= 0) return fail(QG_ERR_K, "k must be positive");
  if (dim != idx->dim)
    return fail(QG_ERR_DIM, "query dimension mismatch: expected " + std::to_string(idx->dim) + ", got " +
                                std::to_string(dim));
  if (q == 0) return 0;
  if (!d_queries || !d_out_keys) return fail(QG_ERR_INVALID, "null device buffer");
  if (filter && filter->owner != idx) return fail(QG_ERR_INVALID, "filter does not belong to this index");
  if (row_base < 0 || row_base + idx->n_rows > 0xFFFFFFFFll)
    return fail(QG_ERR_RANGE, "global rows must fit 32 bits in the packed shard keys");
  cudaStream_t st = (cudaStream_t)stream;
  if (idx->n_live == 0) {
    fill_empty_kernel<<<64, 256, 0, st>>>(nullptr, nullptr, nullptr, nullptr, (uint64_t*)d_out_keys,
                                          (long long)q * k, q);
    QG_CUDA_OK(cudaGetLastError());
    return 0;
  }
  Workspace* w = ws_acquire(idx, st, true);
  if (!w) return fail(QG_ERR_CUDA, "could not create a stream");
  int rc = w->d_count.ensure((size_t)q * 4);
  if (!rc) {
    SearchArgs a{(const float*)d_queries, q, k, filter, nullptr, nullptr, nullptr, nullptr, (int*)w->d_count.p,
                 (uint64_t*)d_out_keys, row_base};
    rc = search_enqueue(idx, w, a, st);
  }
  // The keys are merged with other shards' keys without any count travelling along, so a query whose
  // selection could not be certified is repeated here by the exhaustive path (rare; this is why the call
  // synchronises the stream once before it returns).
  if (!rc) {
    std::vector<int> h_cnt((size_t)q);
    cudaError_t e = cudaMemcpyAsync(h_cnt.data(), w->d_count.p, (size_t)q * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fail(QG_ERR_CUDA, std::string("shard search: ") + cudaGetErrorString(e));
    const uint32_t* mask = filter ? (const uint32_t*)filter->comb_mask.p
                                  : (idx->n_live < idx->n_rows ? idx->live : nullptr);
    const size_t rowb = (size_t)idx->dp * 4, srcb = (size_t)idx->dim * 4;
    for (int i = 0; i < q && !rc; ++i) {
      if (h_cnt[(size_t)i] >= 0) continue;
      if ((rc = w->qpad.ensure(rowb))) break;
      e = cudaMemsetAsync(w->qpad.p, 0, rowb, st);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(w->qpad.p, (const float*)d_queries + (size_t)i * idx->dim, srcb, cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, cudaGetErrorString(e)); break; }
      rc = exhaustive_search(w->ex, idx->vec, idx->n_rows, idx->dp, idx->dim, mask, (const float*)w->qpad.p, nullptr,
                             idx->metric, idx->arith, k, nullptr, nullptr, nullptr, (int*)w->d_count.p + i,
                             (uint64_t*)d_out_keys + (size_t)i * k, row_base, st);
      {
        std::lock_guard<std::mutex> slk(idx->stats_mu);
        idx->stats.escalations++;
      }
    }
  }
  ws_release_async(idx, w, st);
  return rc;
}

int qg_merge_shard_keys_device(int device, const void* d_keys_gathered, int world, int q, int k, void* d_out_dist,
                               void* d_out_row, void* d_out_count, void* stream) {
  if (world <= 0 || q < 0 || k <= 0) return fail(QG_ERR_INVALID, "merge: bad sizes");
  if (!d_keys_gathered || !d_out_dist || !d_out_row || !d_out_count) return fail(QG_ERR_INVALID, "null device buffer");
  QG_CUDA_OK(cudaSetDevice(device));
  return launch_merge_shards((const uint64_t*)d_keys_gathered, world, q, k, (float*)d_out_dist,
                             (long long*)d_out_row, (int*)d_out_count, (cudaStream_t)stream);
}

int qg_search_batch(qg_index* idx, const float* queries, int q, int dim, int k, qg_filter* filter,
                    const float* negatives, float* out_dist, float* out_negdist, int64_t* out_row, int* out_count) {
  if (int rc = check_index(idx)) return rc;
  int err = 0;
  const int empty = validate_search(idx, q, dim, k, &err);
  if (err) return err;
  if (q == 0) return 0;
  if (!out_count) return fail(QG_ERR_INVALID, "out_count is null");
  if (empty) {
    for (int i = 0; i < q; ++i) out_count[i] = 0;
    if (k > 0) {
      for (long long i = 0; i < (long long)q * k; ++i) {
        if (out_dist) out_dist[i] = INFINITY;
        if (out_row) out_row[i] = -1;
        if (out_negdist) out_negdist[i] = INFINITY;
      }
    }
    return 0;
  }
  if (!queries || !out_dist || !out_row) return fail(QG_ERR_INVALID, "null buffer");
  if (negatives && !out_negdist) return fail(QG_ERR_INVALID, "negatives given without out_negdist");
  if (filter && filter->owner != idx) return fail(QG_ERR_INVALID, "filter does not belong to this index");
  Workspace* w = ws_acquire(idx);
  if (!w) return fail(QG_ERR_CUDA, "could not create a stream");
  const size_t qbytes = (size_t)q * dim * 4;
  const size_t obytes = (size_t)q * k;
  int rc = 0;
  do {
    if ((rc = w->h_in.ensure(qbytes * (negatives ? 2 : 1)))) break;
    if ((rc = w->h_out.ensure(obytes * (4 + 4 + 8) + (size_t)q * 4))) break;
    if ((rc = w->d_q.ensure(qbytes))) break;
    if (negatives && (rc = w->d_neg.ensure(qbytes))) break;
    if ((rc = w->d_dist.ensure(obytes * 4))) break;
    if ((rc = w->d_negdist.ensure(obytes * 4))) break;
    if ((rc = w->d_row.ensure(obytes * 8))) break;
    if ((rc = w->d_count.ensure((size_t)q * 4))) break;
    cudaStream_t st = w->stream;
    // Large batches are staged and enqueued in chunks: the host copy into pinned memory and the H2D copy
    // of chunk c + 1 (copy stream) run while the device scans chunk c.
    // The first chunk is one pass of 256 queries, so that the device starts ~15 us after the call instead of
    // after the 1 MB host copy + DMA of a full chunk; the others are groups of 2 048.
    constexpr int CHUNK = 2048, FIRST = 256;
    std::vector<int> chunk_q0;
    if (q > CHUNK + FIRST) {
      chunk_q0.push_back(0);
      for (int q0 = FIRST; q0 < q; q0 += CHUNK) chunk_q0.push_back(q0);
    } else {
      for (int q0 = 0; q0 < q; q0 += CHUNK) chunk_q0.push_back(q0);
    }
    const int n_chunks = (int)chunk_q0.size();
    if (n_chunks > 1 && !w->copy_stream &&
        (cudaStreamCreateWithFlags(&w->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
         cudaStreamCreateWithFlags(&w->d2h_stream, cudaStreamNonBlocking) != cudaSuccess)) {
      rc = fail(QG_ERR_CUDA, "could not create the copy streams");
      break;
    }
    while (n_chunks > 1 && (int)w->chunk_ev.size() < 3 * n_chunks) {
      cudaEvent_t ev = nullptr;
      if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) break;
      w->chunk_ev.push_back(ev);
    }
    if (n_chunks > 1 && (int)w->chunk_ev.size() < 3 * n_chunks) { rc = fail(QG_ERR_CUDA, "could not create events"); break; }
    char* const ho = (char*)w->h_out.p;
    float* const h_dist = (float*)ho;
    float* const h_neg = (float*)(ho + obytes * 4);
    long long* const h_row = (long long*)(ho + obytes * 8);
    int* const h_cnt = (int*)(ho + obytes * 16);
    cudaError_t e = cudaSuccess;
    qg_scan_stats total{};
    for (int c = 0; c < n_chunks && !rc; ++c) {
      const int q0 = chunk_q0[(size_t)c], qc = (c + 1 < n_chunks ? chunk_q0[(size_t)c + 1] : q) - q0;
      const size_t off = (size_t)q0 * dim * 4, cb = (size_t)qc * dim * 4;
      cudaStream_t cs = n_chunks > 1 ? w->copy_stream : st;
      std::memcpy((char*)w->h_in.p + off, (const char*)queries + off, cb);
      e = cudaMemcpyAsync((char*)w->d_q.p + off, (char*)w->h_in.p + off, cb, cudaMemcpyHostToDevice, cs);
      if (e == cudaSuccess && negatives) {
        std::memcpy((char*)w->h_in.p + qbytes + off, (const char*)negatives + off, cb);
        e = cudaMemcpyAsync((char*)w->d_neg.p + off, (char*)w->h_in.p + qbytes + off, cb, cudaMemcpyHostToDevice, cs);
      }
      if (e == cudaSuccess && n_chunks > 1) {
        e = cudaEventRecord(w->chunk_ev[c], cs);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, w->chunk_ev[c], 0);
      }
      if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, cudaGetErrorString(e)); break; }
      SearchArgs a{(const float*)w->d_q.p + (size_t)q0 * dim, qc, k, filter,
                   negatives ? (const float*)w->d_neg.p + (size_t)q0 * dim : nullptr,
                   (float*)w->d_dist.p + (size_t)q0 * k, negatives ? (float*)w->d_negdist.p + (size_t)q0 * k : nullptr,
                   (long long*)w->d_row.p + (size_t)q0 * k, (int*)w->d_count.p + q0, nullptr, 0};
      if ((rc = search_enqueue(idx, w, a, st))) break;
      if (c == 0) {
        total = g_my_stats;
      } else {
        total.passes += g_my_stats.passes;
        total.kernel_launches += g_my_stats.kernel_launches;
      }
      if (n_chunks > 1) {
        // this chunk's results travel to the host (second copy stream) while the next chunks are scanned
        cudaEvent_t done = w->chunk_ev[n_chunks + c], landed = w->chunk_ev[2 * n_chunks + c];
        cudaStream_t ds = w->d2h_stream;
        const size_t o0 = (size_t)q0 * k, on = (size_t)qc * k;
        e = cudaEventRecord(done, st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ds, done, 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_dist + o0, (float*)w->d_dist.p + o0, on * 4, cudaMemcpyDeviceToHost, ds);
        if (e == cudaSuccess && negatives)
          e = cudaMemcpyAsync(h_neg + o0, (float*)w->d_negdist.p + o0, on * 4, cudaMemcpyDeviceToHost, ds);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_row + o0, (long long*)w->d_row.p + o0, on * 8, cudaMemcpyDeviceToHost, ds);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_cnt + q0, (int*)w->d_count.p + q0, (size_t)qc * 4, cudaMemcpyDeviceToHost, ds);
        if (e == cudaSuccess) e = cudaEventRecord(landed, ds);
        if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, cudaGetErrorString(e)); break; }
      }
    }
    if (rc) break;
    publish_stats(idx, total);
    bool copied_out = false;
    if (n_chunks > 1) {
      // hand every chunk to the caller's buffers as soon as it has landed (later chunks are still in flight)
      for (int c = 0; c < n_chunks && e == cudaSuccess; ++c) {
        const int q0 = chunk_q0[(size_t)c], qc = (c + 1 < n_chunks ? chunk_q0[(size_t)c + 1] : q) - q0;
        const size_t o0 = (size_t)q0 * k, on = (size_t)qc * k;
        e = cudaEventSynchronize(w->chunk_ev[2 * n_chunks + c]);
        if (e != cudaSuccess) break;
        std::memcpy(out_dist + o0, h_dist + o0, on * 4);
        if (negatives) std::memcpy(out_negdist + o0, h_neg + o0, on * 4);
        std::memcpy(out_row + o0, h_row + o0, on * 8);
        std::memcpy(out_count + q0, h_cnt + q0, (size_t)qc * 4);
      }
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, std::string("search: ") + cudaGetErrorString(e)); break; }
      copied_out = true;
    }
    if (!copied_out) {
      e = cudaMemcpyAsync(h_dist, w->d_dist.p, obytes * 4, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess && negatives) e = cudaMemcpyAsync(h_neg, w->d_negdist.p, obytes * 4, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(h_row, w->d_row.p, obytes * 8, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(h_cnt, w->d_count.p, (size_t)q * 4, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, std::string("search: ") + cudaGetErrorString(e)); break; }
    }
    int escalations = 0;
    // queries the tensor-core regime could not certify: redo one by one with the flat scan
    if (total.path == 3) {
      const qg_scan_stats first = total;
      for (int i = 0; i < q && !rc; ++i) {
        if (h_cnt[i] >= 0) continue;
        ++escalations;
        SearchArgs a1{(const float*)w->d_q.p + (size_t)i * dim, 1, k, filter,
                      negatives ? (const float*)w->d_neg.p + (size_t)i * dim : nullptr,
                      (float*)w->d_dist.p + (size_t)i * k,
                      negatives ? (float*)w->d_negdist.p + (size_t)i * k : nullptr,
                      (long long*)w->d_row.p + (size_t)i * k, (int*)w->d_count.p + i, nullptr, 0, /*no_tc=*/true};
        if ((rc = search_enqueue(idx, w, a1, st))) break;
        e = cudaMemcpyAsync(h_dist + (size_t)i * k, (float*)w->d_dist.p + (size_t)i * k, (size_t)k * 4,
                            cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && negatives)
          e = cudaMemcpyAsync(h_neg + (size_t)i * k, (float*)w->d_negdist.p + (size_t)i * k, (size_t)k * 4,
                              cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess)
          e = cudaMemcpyAsync(h_row + (size_t)i * k, (long long*)w->d_row.p + (size_t)i * k, (size_t)k * 8,
                              cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_cnt + i, (int*)w->d_count.p + i, 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail(QG_ERR_CUDA, std::string("flat re-scan: ") + cudaGetErrorString(e));
      }
      if (rc) break;
      publish_stats(idx, first);
    }
    // queries the flat scan could not certify either: redo with the exhaustive path
    for (int i = 0; i < q && !rc; ++i) {
      if (h_cnt[i] >= 0) continue;
      ++escalations;
      const uint32_t* mask = filter ? (const uint32_t*)filter->comb_mask.p
                                    : (idx->n_live < idx->n_rows ? idx->live : nullptr);
      const float* qpad = (const float*)w->d_q.p + (size_t)i * idx->dp;
      const float* npad = negatives ? (const float*)w->d_neg.p + (size_t)i * idx->dp : nullptr;
      if (idx->dp != idx->dim) {
        // re-pad this one query (the padded batch may have been overwritten by a re-scan)
        const size_t rowb = (size_t)idx->dp * 4, srcb = (size_t)idx->dim * 4;
        if ((rc = w->qpad.ensure(rowb))) break;
        e = cudaMemsetAsync(w->qpad.p, 0, rowb, st);
        if (e == cudaSuccess)
          e = cudaMemcpyAsync(w->qpad.p, (const float*)w->d_q.p + (size_t)i * idx->dim, srcb, cudaMemcpyDeviceToDevice, st);
        qpad = (const float*)w->qpad.p;
        if (e == cudaSuccess && negatives) {
          if ((rc = w->negpad.ensure(rowb))) break;
          e = cudaMemsetAsync(w->negpad.p, 0, rowb, st);
          if (e == cudaSuccess)
            e = cudaMemcpyAsync(w->negpad.p, (const float*)w->d_neg.p + (size_t)i * idx->dim, srcb,
                                cudaMemcpyDeviceToDevice, st);
          npad = (const float*)w->negpad.p;
        }
        if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, cudaGetErrorString(e)); break; }
      }
      rc = exhaustive_search(w->ex, idx->vec, idx->n_rows, idx->dp, idx->dim, mask, qpad, npad, idx->metric,
                             idx->arith, k,
                             (float*)w->d_dist.p + (size_t)i * k,
                             negatives ? (float*)w->d_negdist.p + (size_t)i * k : nullptr,
                             (long long*)w->d_row.p + (size_t)i * k, (int*)w->d_count.p + i, nullptr, 0, st);
      if (rc) break;
      e = cudaMemcpyAsync(h_dist + (size_t)i * k, (float*)w->d_dist.p + (size_t)i * k, (size_t)k * 4,
                          cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess && negatives)
        e = cudaMemcpyAsync(h_neg + (size_t)i * k, (float*)w->d_negdist.p + (size_t)i * k, (size_t)k * 4,
                            cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(h_row + (size_t)i * k, (long long*)w->d_row.p + (size_t)i * k, (size_t)k * 8,
                            cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(h_cnt + i, (int*)w->d_count.p + i, 4, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) rc = fail(QG_ERR_CUDA, std::string("exhaustive search: ") + cudaGetErrorString(e));
    }
    if (rc) break;
    total.escalations = escalations;
    publish_stats(idx, total);
    if (!copied_out || escalations > 0) {  // (re-run queries changed their rows in the staging buffers)
      std::memcpy(out_dist, h_dist, obytes * 4);
      if (negatives) std::memcpy(out_negdist, h_neg, obytes * 4);
      std::memcpy(out_row, h_row, obytes * 8);
      std::memcpy(out_count, h_cnt, (size_t)q * 4);
    }
  } while (0);
  ws_release(idx, w);
  return rc;
}

int qg_search_exhaustive(qg_index* idx, const float* queries, int q, int dim, int k, qg_filter* filter,
                         float* out_dist, int64_t* out_row, int* out_count) {
  if (int rc = check_index(idx)) return rc;
  int err = 0;
  const int empty = validate_search(idx, q, dim, k, &err);
  if (err) return err;
  if (q == 0) return 0;
  if (!out_count) return fail(QG_ERR_INVALID, "out_count is null");
  if (empty) {
    for (int i = 0; i < q; ++i) out_count[i] = 0;
    for (long long i = 0; k > 0 && i < (long long)q * k; ++i) {
      if (out_dist) out_dist[i] = INFINITY;
      if (out_row) out_row[i] = -1;
    }
    return 0;
  }
  if (!queries || !out_dist || !out_row) return fail(QG_ERR_INVALID, "null buffer");
  if (filter && filter->owner != idx) return fail(QG_ERR_INVALID, "filter does not belong to this index");
  Workspace* w = ws_acquire(idx);
  if (!w) return fail(QG_ERR_CUDA, "could not create a stream");
  int rc = 0;
  do {
    cudaStream_t st = w->stream;
    const size_t rowb = (size_t)idx->dp * 4;
    if ((rc = w->qpad.ensure(rowb))) break;
    if ((rc = w->d_dist.ensure((size_t)k * 4))) break;
    if ((rc = w->d_row.ensure((size_t)k * 8))) break;
    if ((rc = w->d_count.ensure(4))) break;
    const uint32_t* mask = nullptr;
    if (filter) {
      if ((rc = filter_refresh(idx, filter, st))) break;
      mask = (const uint32_t*)filter->comb_mask.p;
    } else if (idx->n_live < idx->n_rows) {
      mask = idx->live;
    }
    for (int i = 0; i < q && !rc; ++i) {
      cudaError_t e = cudaMemsetAsync(w->qpad.p, 0, rowb, st);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(w->qpad.p, queries + (size_t)i * dim, (size_t)dim * 4, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, cudaGetErrorString(e)); break; }
      if ((rc = exhaustive_search(w->ex, idx->vec, idx->n_rows, idx->dp, idx->dim, mask, (const float*)w->qpad.p,
                                  nullptr, idx->metric, idx->arith, k, (float*)w->d_dist.p, nullptr,
                                  (long long*)w->d_row.p, (int*)w->d_count.p, nullptr, 0, st)))
        break;
      e = cudaMemcpyAsync(out_dist + (size_t)i * k, w->d_dist.p, (size_t)k * 4, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(out_row + (size_t)i * k, w->d_row.p, (size_t)k * 8, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(out_count + i, w->d_count.p, 4, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) rc = fail(QG_ERR_CUDA, std::string("exhaustive search: ") + cudaGetErrorString(e));
    }
  } while (0);
  ws_release(idx, w);
  return rc;
}

int qg_batch_distance_multi(qg_index* idx, const float* queries, int b, int dim, const uint32_t* rows, int m,
                            float* out) {
  if (int rc = check_index(idx)) return rc;
  if (b < 0 || m < 0) return fail(QG_ERR_INVALID, "negative batch size");
  if (dim != idx->dim)
    return fail(QG_ERR_DIM, "query dimension mismatch: expected " + std::to_string(idx->dim) + ", got " +
                                std::to_string(dim));
  if (b == 0 || m == 0) return 0;
  if (!queries || !rows || !out) return fail(QG_ERR_INVALID, "null buffer");
  Workspace* w = ws_acquire(idx);
  if (!w) return fail(QG_ERR_CUDA, "could not create a stream");
  const size_t qbytes = (size_t)b * dim * 4, rbytes = (size_t)b * m * 4;
  int rc = 0;
  do {
    if ((rc = w->h_in.ensure(qbytes + rbytes))) break;
    if ((rc = w->h_out.ensure(rbytes))) break;
    if ((rc = w->d_q.ensure(qbytes))) break;
    if ((rc = w->d_rows32.ensure(rbytes))) break;
    if ((rc = w->d_dist.ensure(rbytes))) break;
    cudaStream_t st = w->stream;
    std::memcpy(w->h_in.p, queries, qbytes);
    std::memcpy((char*)w->h_in.p + qbytes, rows, rbytes);
    cudaError_t e = cudaMemcpyAsync(w->d_q.p, w->h_in.p, qbytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(w->d_rows32.p, (char*)w->h_in.p + qbytes, rbytes, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, cudaGetErrorString(e)); break; }
    // unpadded queries are fine here: the pair kernel reads only the first `dim` elements
    if ((rc = launch_batch_distance(idx->vec, idx->dp, idx->dim, idx->n_rows, idx->metric, idx->arith,
                                    (const float*)w->d_q.p, dim, b, (const uint32_t*)w->d_rows32.p, m,
                                    (float*)w->d_dist.p, st)))
      break;
    e = cudaMemcpyAsync(w->h_out.p, w->d_dist.p, rbytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, cudaGetErrorString(e)); break; }
    std::memcpy(out, w->h_out.p, rbytes);
  } while (0);
  ws_release(idx, w);
  return rc;
}

struct qg_queries {
  qg_index* owner = nullptr;
  int b = 0, dim = 0;
  DevBuf d_q;
  int device = 0;  // the owner's device, for qg_queries_destroy
};

int qg_queries_upload(qg_index* idx, const float* queries, int b, int dim, qg_queries** out) {
  if (int rc = check_index(idx)) return rc;
  if (!out) return fail(QG_ERR_INVALID, "out is null");
  *out = nullptr;
  if (b <= 0 || !queries) return fail(QG_ERR_INVALID, "no queries");
  if (dim != idx->dim)
    return fail(QG_ERR_DIM, "query dimension mismatch: expected " + std::to_string(idx->dim) + ", got " +
                                std::to_string(dim));
  std::unique_ptr<qg_queries> qs(new qg_queries());
  qs->owner = idx;
  qs->device = idx->device;
  qs->b = b;
  qs->dim = dim;
  if (int rc = qs->d_q.ensure((size_t)b * dim * 4)) return rc;
  QG_CUDA_OK(cudaMemcpy(qs->d_q.p, queries, (size_t)b * dim * 4, cudaMemcpyHostToDevice));
  *out = qs.release();
  return 0;
}

int qg_queries_destroy(qg_queries* qs) {
  if (!qs) return 0;
  if (cudaSetDevice(qs->device) != cudaSuccess) (void)cudaGetLastError();  // the index may be gone already
  qs->d_q.release();
  delete qs;
  return 0;
}

int qg_batch_distance_queries(qg_index* idx, const qg_queries* qs, const uint32_t* rows, int m, float* out) {
  if (int rc = check_index(idx)) return rc;
  if (!qs || qs->owner != idx) return fail(QG_ERR_INVALID, "query set does not belong to this index");
  if (m < 0) return fail(QG_ERR_INVALID, "negative batch size");
  if (m == 0) return 0;
  if (!rows || !out) return fail(QG_ERR_INVALID, "null buffer");
  Workspace* w = ws_acquire(idx);
  if (!w) return fail(QG_ERR_CUDA, "could not create a stream");
  const size_t rbytes = (size_t)qs->b * m * 4;
  int rc = 0;
  do {
    if ((rc = w->h_in.ensure(rbytes))) break;
    if ((rc = w->h_out.ensure(rbytes))) break;
    if ((rc = w->d_rows32.ensure(rbytes))) break;
    if ((rc = w->d_dist.ensure(rbytes))) break;
    cudaStream_t st = w->stream;
    std::memcpy(w->h_in.p, rows, rbytes);
    cudaError_t e = cudaMemcpyAsync(w->d_rows32.p, w->h_in.p, rbytes, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, cudaGetErrorString(e)); break; }
    if ((rc = launch_batch_distance(idx->vec, idx->dp, idx->dim, idx->n_rows, idx->metric, idx->arith,
                                    (const float*)qs->d_q.p, qs->dim, qs->b, (const uint32_t*)w->d_rows32.p, m,
                                    (float*)w->d_dist.p, st)))
      break;
    e = cudaMemcpyAsync(w->h_out.p, w->d_dist.p, rbytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, cudaGetErrorString(e)); break; }
    std::memcpy(out, w->h_out.p, rbytes);
  } while (0);
  ws_release(idx, w);
  return rc;
}

int qg_batch_distance(qg_index* idx, const float* query, int dim, const uint32_t* rows, int n, float* out) {
  return qg_batch_distance_multi(idx, query, 1, dim, rows, n, out);
}

struct qg_hnsw {
  qg_index* owner = nullptr;
  int device = 0;  // the owner's device, for qg_hnsw_destroy (which may run after the index is gone)
  HnswDevGraph g{};
  DevBuf level, adj0, upper_off, upper_adj, work;
  std::mutex mu;  // one search at a time uses the workspace
};

int qg_hnsw_upload(qg_index* idx, int64_t n_nodes, int m, int max_m0, int entry_point, int current_level,
                   const int32_t* level, const uint32_t* adj0, const int64_t* upper_off, const uint32_t* upper_adj,
                   qg_hnsw** out) {
  if (int rc = check_index(idx)) return rc;
  if (!out) return fail(QG_ERR_INVALID, "out is null");
  *out = nullptr;
  if (n_nodes < 0 || m <= 0 || max_m0 <= 0) return fail(QG_ERR_INVALID, "bad graph shape");
  if (n_nodes > 0 && (!level || !adj0 || !upper_off)) return fail(QG_ERR_INVALID, "null graph array");
  if (n_nodes > idx->n_rows) return fail(QG_ERR_RANGE, "the graph has more nodes than the index has rows");
  std::unique_ptr<qg_hnsw> g(new qg_hnsw());
  g->owner = idx;
  g->device = idx->device;
  g->g.n_nodes = n_nodes;
  g->g.m = m;
  g->g.max_m0 = max_m0;
  g->g.entry_point = entry_point;
  g->g.current_level = current_level;
  if (n_nodes > 0) {
    const size_t n = (size_t)n_nodes;
    const size_t ulen = (size_t)upper_off[n];
    if (ulen > 0 && !upper_adj) return fail(QG_ERR_INVALID, "null graph array");
    if (int rc = g->level.ensure(n * 4)) return rc;
    if (int rc = g->adj0.ensure(n * (size_t)max_m0 * 4)) return rc;
    if (int rc = g->upper_off.ensure((n + 1) * 8)) return rc;
    if (int rc = g->upper_adj.ensure(std::max<size_t>(ulen, 1) * 4)) return rc;
    QG_CUDA_OK(cudaMemcpy(g->level.p, level, n * 4, cudaMemcpyHostToDevice));
    QG_CUDA_OK(cudaMemcpy(g->adj0.p, adj0, n * (size_t)max_m0 * 4, cudaMemcpyHostToDevice));
    QG_CUDA_OK(cudaMemcpy(g->upper_off.p, upper_off, (n + 1) * 8, cudaMemcpyHostToDevice));
    if (ulen > 0) QG_CUDA_OK(cudaMemcpy(g->upper_adj.p, upper_adj, ulen * 4, cudaMemcpyHostToDevice));
    const size_t wb = hnsw_workspace_bytes(n_nodes, idx->sm_count);
    if (int rc = g->work.ensure(wb)) return rc;
    QG_CUDA_OK(cudaMemset(g->work.p, 0, wb));  // the kernel leaves the visited bitsets clean
  }
  g->g.level = (const int32_t*)g->level.p;
  g->g.adj0 = (const uint32_t*)g->adj0.p;
  g->g.upper_off = (const long long*)g->upper_off.p;
  g->g.upper_adj = (const uint32_t*)g->upper_adj.p;
  *out = g.release();
  return 0;
}

int qg_hnsw_build(qg_index* idx, int m, int max_m0, int ef_construction, int max_level, uint64_t seed, int max_batch,
                  qg_hnsw** out) {
  if (int rc = check_index(idx)) return rc;
  if (!out) return fail(QG_ERR_INVALID, "out is null");
  *out = nullptr;
  if (m <= 0 || m > 32 || max_m0 <= 0 || max_m0 > 32 || ef_construction <= 0 || max_level <= 0)
    return fail(QG_ERR_INVALID, "hnsw build: need 0 < m, max_m0 <= 32, ef_construction > 0, max_level > 0");
  if (idx->n_live != idx->n_rows) return fail(QG_ERR_UNSUPPORTED, "hnsw build: compact the index first (tombstoned rows)");
  const long long n = idx->n_rows;
  if (n >= (1ll << 30)) return fail(QG_ERR_RANGE, "hnsw build: at most 2^30 - 1 nodes");
  if (max_batch <= 0) max_batch = 8192;
  std::unique_ptr<qg_hnsw> g(new qg_hnsw());
  g->owner = idx;
  g->device = idx->device;
  g->g.n_nodes = n;
  g->g.m = m;
  g->g.max_m0 = max_m0;
  if (n == 0) {
    *out = g.release();
    return 0;
  }
  // levels (randomLevel, hnsw.go:716-738) and the offsets of the upper-level blocks
  std::vector<int32_t> level((size_t)n);
  std::vector<int64_t> uoff((size_t)n + 1);
  const int attempts = std::min(max_level, 10);
  int64_t u = 0;
  for (long long i = 0; i < n; ++i) {
    int lv = 0;
    for (int a = 0; a < attempts; ++a) {
      if (synth_uniform(seed, (uint64_t)i, (uint32_t)a, 7) < 0.25f) lv++;
      else break;
    }
    if (lv >= max_level) lv = max_level - 1;
    level[(size_t)i] = lv;
    uoff[(size_t)i] = u;
    u += (int64_t)lv * m;
  }
  uoff[(size_t)n] = u;
  int rc = 0;
  if ((rc = g->level.ensure((size_t)n * 4)) || (rc = g->adj0.ensure((size_t)n * max_m0 * 4)) ||
      (rc = g->upper_off.ensure(((size_t)n + 1) * 8)) || (rc = g->upper_adj.ensure(std::max<size_t>((size_t)u, 1) * 4)))
    return rc;
  QG_CUDA_OK(cudaMemcpy(g->level.p, level.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
  QG_CUDA_OK(cudaMemcpy(g->upper_off.p, uoff.data(), ((size_t)n + 1) * 8, cudaMemcpyHostToDevice));
  QG_CUDA_OK(cudaMemset(g->adj0.p, 0xff, (size_t)n * max_m0 * 4));
  QG_CUDA_OK(cudaMemset(g->upper_adj.p, 0xff, std::max<size_t>((size_t)u, 1) * 4));
  const size_t wb = hnsw_workspace_bytes(n, idx->sm_count);
  if ((rc = g->work.ensure(wb))) return rc;
  QG_CUDA_OK(cudaMemset(g->work.p, 0, wb));
  g->g.level = (const int32_t*)g->level.p;
  g->g.adj0 = (const uint32_t*)g->adj0.p;
  g->g.upper_off = (const long long*)g->upper_off.p;
  g->g.upper_adj = (const uint32_t*)g->upper_adj.p;
  g->g.entry_point = 0;                 // the first node: entry point of its level (hnsw.go:317-323)
  g->g.current_level = level[0];
  // reverse-link records of a batch: at most max_m0 + levels * m per node
  const unsigned int rev_cap = (unsigned int)std::min<long long>((long long)max_batch * (max_m0 + 4 * m), 1ll << 31);
  DevBuf rev_key, rev_key2, rev_dist, rev_cnt, perm_a, perm_b;
  void* sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  auto cleanup = [&](int code) {
    rev_key.release(); rev_key2.release(); rev_dist.release(); rev_cnt.release(); perm_a.release(); perm_b.release();
    if (sort_tmp) cudaFree(sort_tmp);
    return code;
  };
  if ((rc = rev_key.ensure((size_t)rev_cap * 8)) || (rc = rev_key2.ensure((size_t)rev_cap * 8)) ||
      (rc = rev_dist.ensure((size_t)rev_cap * 4)) || (rc = rev_cnt.ensure(4)) || (rc = perm_a.ensure((size_t)rev_cap * 4)) ||
      (rc = perm_b.ensure((size_t)rev_cap * 4)))
    return cleanup(rc);
  Workspace* w = ws_acquire(idx);
  if (!w) return cleanup(fail(QG_ERR_CUDA, "could not create a stream"));
  cudaStream_t st = w->stream;
  long long committed = 1;
  while (committed < n && !rc) {
    // a batch never exceeds an eighth of the committed graph: the nodes of a batch do not see each other
    const int count = (int)std::min<long long>({(long long)max_batch, std::max<long long>(1, committed / 8), n - committed});
    if ((rc = launch_hnsw_insert_batch(g->g, (uint32_t*)g->adj0.p, (uint32_t*)g->upper_adj.p, idx->vec, idx->dp, idx->dim,
                                       idx->metric, idx->arith, ef_construction, committed, count, g->work.p, idx->sm_count,
                                       (unsigned long long*)rev_key.p, (float*)rev_dist.p, (unsigned int*)rev_cnt.p, rev_cap,
                                       st)))
      break;
    unsigned int n_rec = 0;
    cudaError_t e = cudaMemcpyAsync(&n_rec, rev_cnt.p, 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, std::string("hnsw build: ") + cudaGetErrorString(e)); break; }
    if (n_rec > rev_cap) { rc = fail(QG_ERR_CUDA, "hnsw build: reverse-link buffer overflow"); break; }
    if ((rc = hnsw_sort_records((const unsigned long long*)rev_key.p, (unsigned long long*)rev_key2.p,
                                (unsigned int*)perm_a.p, (unsigned int*)perm_b.p, n_rec, &sort_tmp, &sort_tmp_bytes, st)))
      break;
    if ((rc = launch_hnsw_link_batch(g->g, (uint32_t*)g->adj0.p, (uint32_t*)g->upper_adj.p, idx->vec, idx->dp, idx->dim,
                                     idx->metric, idx->arith, (const unsigned long long*)rev_key2.p,
                                     (const unsigned int*)perm_b.p, (const float*)rev_dist.p, n_rec, st)))
      break;
    // entry point / current level: the first node of the batch that reaches a new top level (hnsw.go:463-466)
    for (long long i = committed; i < committed + count; ++i)
      if (level[(size_t)i] > g->g.current_level) {
        g->g.current_level = level[(size_t)i];
        g->g.entry_point = (int)i;
      }
    committed += count;
  }
  if (!rc) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fail(QG_ERR_CUDA, std::string("hnsw build: ") + cudaGetErrorString(e));
  }
  ws_release(idx, w);
  if (rc) return cleanup(rc);
  cleanup(0);
  *out = g.release();
  return 0;
}

int64_t qg_hnsw_nodes(const qg_hnsw* g) { return g ? g->g.n_nodes : 0; }

int64_t qg_hnsw_upper_len(const qg_hnsw* g) {
  if (!g || g->g.n_nodes == 0) return 0;
  long long last = 0;
  cudaSetDevice(g->owner->device);
  if (cudaMemcpy(&last, (const long long*)g->upper_off.p + g->g.n_nodes, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return last;
}

int qg_hnsw_export(const qg_hnsw* g, int32_t* level, uint32_t* adj0, int64_t* upper_off, uint32_t* upper_adj,
                   int* entry_point, int* current_level) {
  if (!g) return fail(QG_ERR_INVALID, "graph handle is null");
  if (entry_point) *entry_point = g->g.entry_point;
  if (current_level) *current_level = g->g.current_level;
  const size_t n = (size_t)g->g.n_nodes;
  if (n == 0) return 0;
  if (!level || !adj0 || !upper_off) return fail(QG_ERR_INVALID, "null buffer");
  QG_CUDA_OK(cudaSetDevice(g->owner->device));
  QG_CUDA_OK(cudaMemcpy(level, g->level.p, n * 4, cudaMemcpyDeviceToHost));
  QG_CUDA_OK(cudaMemcpy(adj0, g->adj0.p, n * (size_t)g->g.max_m0 * 4, cudaMemcpyDeviceToHost));
  QG_CUDA_OK(cudaMemcpy(upper_off, g->upper_off.p, (n + 1) * 8, cudaMemcpyDeviceToHost));
  const size_t ulen = (size_t)upper_off[n];
  if (ulen > 0) {
    if (!upper_adj) return fail(QG_ERR_INVALID, "null buffer");
    QG_CUDA_OK(cudaMemcpy(upper_adj, g->upper_adj.p, ulen * 4, cudaMemcpyDeviceToHost));
  }
  return 0;
}

int qg_hnsw_destroy(qg_hnsw* g) {
  if (!g) return 0;
  if (cudaSetDevice(g->device) != cudaSuccess) (void)cudaGetLastError();
  g->level.release(); g->adj0.release(); g->upper_off.release(); g->upper_adj.release(); g->work.release();
  delete g;
  return 0;
}

int qg_hnsw_search_batch(qg_index* idx, const qg_hnsw* gc, const float* queries, int q, int dim, int k, int ef_search,
                         uint32_t* out_idx, float* out_dist, int* out_count, int64_t* out_evals) {
  if (int rc = check_index(idx)) return rc;
  qg_hnsw* g = const_cast<qg_hnsw*>(gc);
  if (!g || g->owner != idx) return fail(QG_ERR_INVALID, "graph does not belong to this index");
  if (q < 0) return fail(QG_ERR_INVALID, "negative query count");
  if (q == 0) return 0;
  if (!out_count) return fail(QG_ERR_INVALID, "out_count is null");
  if (g->g.n_nodes == 0) {  // hnsw.go:606-608: empty graph => no results, no error
    for (int i = 0; i < q; ++i) out_count[i] = 0;
    return 0;
  }
  if (dim != idx->dim)
    return fail(QG_ERR_DIM, "query dimension mismatch: expected " + std::to_string(idx->dim) + ", got " +
                                std::to_string(dim));
  if (k <= 0) return fail(QG_ERR_K, "k must be positive");
  if ((long long)k > g->g.n_nodes) return fail(QG_ERR_INVALID, "k must be clamped to the node count by the caller");
  if (!queries || !out_idx || !out_dist) return fail(QG_ERR_INVALID, "null buffer");
  if (g->g.entry_point < 0 || g->g.entry_point >= g->g.n_nodes)
    return fail(QG_ERR_RANGE, "entry point out of range (validate it first, hnsw.go:620-634)");
  std::lock_guard<std::mutex> lk(g->mu);
  Workspace* w = ws_acquire(idx);
  if (!w) return fail(QG_ERR_CUDA, "could not create a stream");
  int rc = 0;
  do {
    cudaStream_t st = w->stream;
    const size_t qbytes = (size_t)q * dim * 4, on = (size_t)q * k;
    if ((rc = w->d_q.ensure(qbytes))) break;
    if ((rc = w->d_rows32.ensure(on * 4))) break;
    if ((rc = w->d_dist.ensure(on * 4))) break;
    if ((rc = w->d_count.ensure((size_t)q * 4))) break;
    if ((rc = w->d_rows64.ensure((size_t)q * 8))) break;
    cudaError_t e = cudaMemcpyAsync(w->d_q.p, queries, qbytes, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, cudaGetErrorString(e)); break; }
    const int ef0 = std::max(ef_search, k);  // hnsw.go:660-663
    if ((rc = launch_hnsw_search(g->g, idx->vec, idx->dp, idx->dim, idx->metric, idx->arith, (const float*)w->d_q.p, q, k,
                                 ef0, g->work.p, idx->sm_count, (uint32_t*)w->d_rows32.p, (float*)w->d_dist.p,
                                 (int*)w->d_count.p, out_evals ? (long long*)w->d_rows64.p : nullptr, st)))
      break;
    e = cudaMemcpyAsync(out_idx, w->d_rows32.p, on * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_dist, w->d_dist.p, on * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_count, w->d_count.p, (size_t)q * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && out_evals) e = cudaMemcpyAsync(out_evals, w->d_rows64.p, (size_t)q * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fail(QG_ERR_CUDA, std::string("hnsw search: ") + cudaGetErrorString(e));
  } while (0);
  ws_release(idx, w);
  return rc;
}

int qg_index_set_profiling(qg_index* idx, int on) {
  if (!idx) return fail(QG_ERR_INVALID, "index handle is null");
  idx->profiling = on != 0;
  return 0;
}

int qg_index_read_profile(qg_index* idx, qg_profile* out) {
  if (int rc = check_index(idx)) return rc;
  if (!out) return fail(QG_ERR_INVALID, "out is null");
  *out = qg_profile{};
  std::lock_guard<std::mutex> lk(idx->ws_mu);
  auto drain = [&](Workspace* w) -> int {
    for (size_t i = 0; i + 1 < w->prof_used; i += 2) {
      float ms = 0.f;
      QG_CUDA_OK(cudaEventSynchronize(w->prof_ev[i + 1]));
      QG_CUDA_OK(cudaEventElapsedTime(&ms, w->prof_ev[i], w->prof_ev[i + 1]));
      if (w->prof_kind[i / 2] == 0) {
        out->scan_ms += ms;
        out->scan_launches++;
      } else if (w->prof_kind[i / 2] == 2) {
        out->prep_ms += ms;
        out->prep_launches++;
      } else {
        out->finalize_ms += ms;
        out->finalize_launches++;
      }
    }
    w->prof_used = 0;
    return 0;
  };
  for (auto& w : idx->ws_free)
    if (int rc = drain(w.get())) return rc;
  for (auto& w : idx->ws_async)
    if (int rc = drain(w.get())) return rc;
  return 0;
}

int qg_debug_tc_pass(qg_index* idx, const float* queries, int nq, int k, float* tau_out, int* cnt_out,
                     uint64_t* cand_out, int* cap_out) {
  if (int rc = check_index(idx)) return rc;
  if (!queries || nq <= 0 || nq > TC_MAX_COLS) return fail(QG_ERR_INVALID, "debug_tc_pass: 1..256 queries");
  if (idx->n_rows == 0) return fail(QG_ERR_INVALID, "debug_tc_pass: empty index");
  const int mode = scan_mode_of(idx->metric);
  TcPlan plan{};
  if (mode == MODE_L1 || tc_available() != 0 ||
      tc_plan(idx->dp, idx->dim + tc_extra_cols(idx->dim, mode == MODE_L2), nq, idx->use_bf16, &plan) != 0)
    return fail(QG_ERR_UNSUPPORTED, "tensor-core regime not available for this index");
  Workspace* w = ws_acquire(idx);
  if (!w) return fail(QG_ERR_CUDA, "could not create a stream");
  int rc = 0;
  bool dbg_raw = false;
  do {
    cudaStream_t st = w->stream;
    const int dp = idx->dp, d = idx->dim;
    if ((rc = w->qpad.ensure((size_t)nq * dp * 4))) break;
    cudaError_t e = cudaMemsetAsync(w->qpad.p, 0, (size_t)nq * dp * 4, st);
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync(w->qpad.p, (size_t)dp * 4, queries, (size_t)d * 4, (size_t)d * 4, (size_t)nq,
                            cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, cudaGetErrorString(e)); break; }
    const long long n_sample = tc_sample_tiles(plan, idx->n_rows, k);
    if ((rc = w->tc_sample.ensure((size_t)plan.n_cols * n_sample * plan.sample_vals * 4))) break;
    if ((rc = w->tc_tau.ensure((size_t)TC_MAX_COLS * 4))) break;
    if ((rc = w->tc_cand.ensure((size_t)plan.n_cols * TC_CAND_CAP * 8))) break;
    if ((rc = w->tc_cnt.ensure((size_t)(TC_MAX_COLS + 4) * 4))) break;
    if (plan.variant == 1) {
      if ((rc = w->tc_apack.ensure(tc_pack_bytes(plan, nq)))) break;
      dbg_raw = plan.bf16 && idx->n_live == idx->n_rows && idx->n_rows >= plan.tile_rows &&
                tc_raw_supported(d, mode == MODE_L2);
      if ((rc = launch_tc_pack(plan, (const float*)w->qpad.p, nq, dp, d, dbg_raw ? tc_extra_cols(d, mode == MODE_L2) : 0,
                               idx->metric == METRIC_COSINE, w->tc_apack.p, st)))
        break;
    }
    TcArgs ta{};
    ta.vec = idx->vec;
    ta.vec16 = idx->vec16;
    ta.dp16 = idx->dp16;
    ta.n_rows = idx->n_rows;
    ta.dp = dp;
    ta.row_norm2 = idx->norm2;
    ta.inv_norm = idx->metric == METRIC_COSINE ? idx->inv_norm : nullptr;
    ta.mask = idx->n_live < idx->n_rows ? idx->live : nullptr;
    ta.bias = mode == MODE_L2 ? idx->norm2 : idx->unit_bias;
    if (plan.variant == 1 && ta.mask != nullptr) {
      const long long n_pad = (idx->n_rows + 127) & ~127ll;
      if ((rc = w->tc_bias.ensure((size_t)n_pad * 4))) break;
      if ((rc = launch_tc_bias(ta.mask, mode == MODE_L2 ? idx->norm2 : nullptr, idx->n_rows, n_pad,
                               (float*)w->tc_bias.p, st)))
        break;
      ta.bias = (const float*)w->tc_bias.p;
    }
    ta.queries = (const float*)w->qpad.p;
    ta.nq = nq;
    ta.mode = mode;
    ta.cosine = idx->metric == METRIC_COSINE;
    ta.sample = (uint32_t*)w->tc_sample.p;
    ta.n_sample = (int)n_sample;
    ta.sample_rank = tc_sample_rank(k);
    ta.tau = (float*)w->tc_tau.p;
    ta.cand = (uint64_t*)w->tc_cand.p;
    ta.cand_cnt = (int*)w->tc_cnt.p;
    ta.apack = w->tc_apack.p;
    ta.raw = dbg_raw;
    ta.work_counter = (int*)w->tc_cnt.p + TC_MAX_COLS;
    int launches = 0;
    constexpr size_t kTraceWords = 64 + 5 * 2048;  // tc_scan.cu: TS_TRACE_ROLES x TS_TRACE_CAP events behind the counters
    if ((rc = w->counters.ensure(kTraceWords * 8))) break;
    cudaMemsetAsync(w->counters.p, 0, kTraceWords * 8, st);
    ta.dbg = (unsigned long long*)w->counters.p;
    if ((rc = launch_tc_pass(plan, ta, idx->sm_count, st, &launches))) break;
    e = cudaStreamSynchronize(st);
    if (e == cudaSuccess && std::getenv("QG_TC_TRACE")) {
      // event trace of CTA 0 (instrumented build): one line per event, "role index clock code"
      std::vector<unsigned long long> tr(kTraceWords);
      cudaMemcpy(tr.data(), w->counters.p, kTraceWords * 8, cudaMemcpyDeviceToHost);
      if (FILE* f = std::fopen(std::getenv("QG_TC_TRACE"), "w")) {
        for (int role = 0; role < 5; ++role)
          for (int i = 0; i < 2048 && tr[64 + role * 2048 + i] != 0; ++i)
            std::fprintf(f, "%d %d %llu %d\n", role, i, tr[64 + role * 2048 + i] >> 4, (int)(tr[64 + role * 2048 + i] & 15));
        std::fclose(f);
      }
    }
    if (e == cudaSuccess && std::getenv("QG_TC_TIMING")) {
      unsigned long long h[64];
      cudaMemcpy(h, w->counters.p, sizeof(h), cudaMemcpyDeviceToHost);
      fprintf(stderr, "tc timing (cycles, CTA 0): producer total %llu wait_xs %llu wait_empty %llu | mma total %llu "
                      "wait_tmem_empty %llu wait_full %llu\n", h[0], h[1], h[2], h[3], h[4], h[5]);
      for (int i = 0; i < 4; ++i)
        fprintf(stderr, "  epilogue warp %d: total %llu wait_xs %llu wait_tmem_full %llu tmem_ld %llu scores+vote %llu "
                        "publish+release %llu\n", i + 2, h[8 + i * 8], h[9 + i * 8], h[10 + i * 8], h[11 + i * 8],
                h[12 + i * 8], h[13 + i * 8]);
      for (int c = 0; c < 2; ++c) {
        const unsigned long long* t = h + 40 + c * 8;
        auto us = [&](int i) { return ((double)t[i] - (double)h[40]) * 1e-3; };
        fprintf(stderr, "  wall clock (us after CTA 0 entry), %s CTA: entry %.1f, tmem allocated %.1f, queries resident "
                        "%.1f, MMA starts %.1f, first tile drained %.1f, last MMA %.1f, exit %.1f\n",
                c ? "last" : "first", us(0), us(1), us(2), us(3), us(4), us(5), us(6));
      }
      // finalize of the same pass, with phase time stamps of CTA 0
      FinalizeCandParams cp{};
      cp.cand = (const uint64_t*)w->tc_cand.p;
      cp.cand_cnt = (const int*)w->tc_cnt.p;
      cp.cap = TC_CAND_CAP;
      cp.tau = (const float*)w->tc_tau.p;
      cp.kp = 32;
      cp.tc_gamma = (plan.bf16 ? 1.02 / 256.0 : 1.02 / 512.0) + (double)d / 4194304.0;
      cp.tc_norm_gamma = (dbg_raw && mode == MODE_L2) ? (double)(d + 16) / 4194304.0 : 0.0;
      cudaMemsetAsync(w->counters.p, 0, 64 * 8, st);
      cp.dbg = (unsigned long long*)w->counters.p;
      FinalizeParams& fb = cp.base;
      fb.vec = idx->vec; fb.dp = dp; fb.d = d; fb.metric = idx->metric; fb.arith = idx->arith; fb.mode = mode;
      fb.cosine = idx->metric == METRIC_COSINE; fb.k = k; fb.gamma = (float)((d + 16) * 5.9604645e-8);
      fb.max_norm2 = idx->max_norm2; fb.queries = (const float*)w->qpad.p;
      if (!w->d_dist.ensure((size_t)nq * k * 4) && !w->d_row.ensure((size_t)nq * k * 8) && !w->d_count.ensure((size_t)nq * 4)) {
        fb.out_dist = (float*)w->d_dist.p; fb.out_row = (long long*)w->d_row.p; fb.out_count = (int*)w->d_count.p;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        launch_finalize_cand(cp, nq, st);
        cudaEventRecord(e1, st);
        cudaStreamSynchronize(st);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h, w->counters.p, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "  finalize_cand: %.1f us; CTA 0 cycles: keys %llu, pivot %llu, selected %llu, re-ranked %llu, "
                        "certified %llu, done %llu (n %llu, nex %llu)\n", ms * 1e3, h[10], h[11], h[1], h[2], h[3], h[4],
                h[8], h[9]);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
      }
    }
    if (e == cudaSuccess && tau_out) e = cudaMemcpy(tau_out, w->tc_tau.p, (size_t)nq * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && cnt_out) e = cudaMemcpy(cnt_out, w->tc_cnt.p, (size_t)nq * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && cand_out)
      e = cudaMemcpy(cand_out, w->tc_cand.p, (size_t)nq * TC_CAND_CAP * 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { rc = fail(QG_ERR_CUDA, std::string("debug_tc_pass: ") + cudaGetErrorString(e)); break; }
    if (cap_out) *cap_out = TC_CAND_CAP;
  } while (0);
  ws_release(idx, w);
  return rc;
}

int qg_last_scan_stats(const qg_index* idx, qg_scan_stats* out) {
  if (!idx || !out) return fail(QG_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(const_cast<qg_index*>(idx)->stats_mu);
  *out = idx->stats;
  return 0;
}

}  // extern "C"
