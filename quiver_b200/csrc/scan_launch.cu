// scan_launch.cu — dispatch to the per-dimension fast kernels and the generic kernel.
#include "scan.cuh"

namespace qg {

#define QG_FAST_DIMS(X) X(32) X(64) X(96) X(128) X(192) X(256) X(384) X(512) X(768) X(1024) X(1536)

#define QG_DECL(DIM)                                                                           \
  int launch_scan_fast_##DIM(int qb, int mode, const ScanParams& p, int grid, cudaStream_t st); \
  int scan_fast_attr_##DIM();                                                                   \
  int launch_scan_dense_##DIM(int qb, int mode, const ScanParams& p, int grid, cudaStream_t st); \
  int scan_dense_geom_##DIM(int qb, int kp, int* tile_rows, int* warps, size_t* smem);           \
  int scan_fast_tile_rows_##DIM();                                                              \
  int scan_fast_max_qb_##DIM();                                                                 \
  int scan_fast_ring_##DIM();
QG_FAST_DIMS(QG_DECL)
#undef QG_DECL

int scan_fast_supported(int dp) {
  switch (dp) {
#define QG_CASE(DIM) case DIM:
    QG_FAST_DIMS(QG_CASE)
#undef QG_CASE
    return 1;
    default: return 0;
  }
}

int scan_fast_tile_rows(int dp) {
  switch (dp) {
#define QG_CASE(DIM) \
  case DIM: return scan_fast_tile_rows_##DIM();
    QG_FAST_DIMS(QG_CASE)
#undef QG_CASE
    default: return 0;
  }
}

int scan_fast_max_qb(int dp) {
  switch (dp) {
#define QG_CASE(DIM) \
  case DIM: return scan_fast_max_qb_##DIM();
    QG_FAST_DIMS(QG_CASE)
#undef QG_CASE
    default: return 0;
  }
}

int scan_fast_ring_bytes(int dp) {
  switch (dp) {
#define QG_CASE(DIM) \
  case DIM: return scan_fast_ring_##DIM();
    QG_FAST_DIMS(QG_CASE)
#undef QG_CASE
    default: return 0;
  }
}

int launch_scan_fast(int dp, int qb, int mode, const ScanParams& p, int grid, cudaStream_t st) {
  switch (dp) {
#define QG_CASE(DIM) \
  case DIM: return launch_scan_fast_##DIM(qb, mode, p, grid, st);
    QG_FAST_DIMS(QG_CASE)
#undef QG_CASE
    default: return -1;
  }
}

int launch_scan_dense(int dp, int qb, int mode, const ScanParams& p, int grid, cudaStream_t st) {
  switch (dp) {
#define QG_CASE(DIM) \
  case DIM: return launch_scan_dense_##DIM(qb, mode, p, grid, st);
    QG_FAST_DIMS(QG_CASE)
#undef QG_CASE
    default: return -1;
  }
}

int scan_dense_geometry(int dp, int qb, int kp, int* tile_rows, int* warps, size_t* smem) {
  switch (dp) {
#define QG_CASE(DIM) \
  case DIM: return scan_dense_geom_##DIM(qb, kp, tile_rows, warps, smem);
    QG_FAST_DIMS(QG_CASE)
#undef QG_CASE
    default: return -1;
  }
}

template <int QB, int MODE>
static int generic_one(const ScanParams& p, int grid, int nw, cudaStream_t st) {
  const size_t smem = scan_generic_smem<QB>(nw, p.stages, p.tile_rows, p.dp, p.kp);
  if (smem > 227 * 1024) return fail(6, "scan: row size / candidate pools do not fit shared memory");
  scan_generic_kernel<QB, MODE><<<grid, nw * 32, smem, st>>>(p);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int QB>
static int generic_mode(int mode, const ScanParams& p, int grid, int nw, cudaStream_t st) {
  if (mode == MODE_L2) return generic_one<QB, MODE_L2>(p, grid, nw, st);
  if (mode == MODE_DOT) return generic_one<QB, MODE_DOT>(p, grid, nw, st);
  return generic_one<QB, MODE_L1>(p, grid, nw, st);
}

int launch_scan_generic(int qb, int mode, const ScanParams& p, int grid, int nw, cudaStream_t st) {
  switch (qb) {
    case 1: return generic_mode<1>(mode, p, grid, nw, st);
    case 2: return generic_mode<2>(mode, p, grid, nw, st);
    case 4: return generic_mode<4>(mode, p, grid, nw, st);
    case 8: return generic_mode<8>(mode, p, grid, nw, st);
    default: return fail(1, "scan: query block must be 1, 2, 4 or 8");
  }
}

template <int QB>
static int generic_attr() {
  QG_CUDA_OK(cudaFuncSetAttribute(scan_generic_kernel<QB, MODE_L2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024));
  QG_CUDA_OK(cudaFuncSetAttribute(scan_generic_kernel<QB, MODE_DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024));
  QG_CUDA_OK(cudaFuncSetAttribute(scan_generic_kernel<QB, MODE_L1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024));
  return 0;
}

int scan_set_attributes() {
#define QG_ATTR(DIM) \
  if (int e = scan_fast_attr_##DIM()) return e;
  QG_FAST_DIMS(QG_ATTR)
#undef QG_ATTR
  if (int e = generic_attr<1>()) return e;
  if (int e = generic_attr<2>()) return e;
  if (int e = generic_attr<4>()) return e;
  if (int e = generic_attr<8>()) return e;
  return 0;
}

}  // namespace qg
