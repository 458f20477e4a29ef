// scan_fast_inst.cuh — instantiates the fast scan kernels for one compile-time dimension.
// Each scan_fast_<D>.cu includes this with QG_SCAN_D defined, so the dimensions compile in
// parallel (make -j).
#pragma once
#include "scan.cuh"

namespace qg {

template <int D>
constexpr int scan_fast_qb_limit() {
  return ScanGeom<D>::CPL <= 4 ? 8 : (ScanGeom<D>::CPL <= 8 ? 4 : 2);
}

template <int D, int QB, int MODE>
static int launch_one(const ScanParams& p, int grid, cudaStream_t st) {
  const size_t smem = scan_fast_smem<D, QB>(p.kp);
  if (smem > 227 * 1024) return fail(6, "scan: candidate pools do not fit shared memory");
  scan_fast_kernel<D, QB, MODE><<<grid, SCAN_NW * 32, smem, st>>>(p);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int D, int QB, int MODE>
static int attr_one() {
  QG_CUDA_OK(cudaFuncSetAttribute(scan_fast_kernel<D, QB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024));
  return 0;
}

template <int D, int QB>
static int launch_mode(int mode, const ScanParams& p, int grid, cudaStream_t st) {
  if constexpr (QB <= scan_fast_qb_limit<D>()) {
    if (mode == MODE_L2) return launch_one<D, QB, MODE_L2>(p, grid, st);
    if (mode == MODE_DOT) return launch_one<D, QB, MODE_DOT>(p, grid, st);
  }
  return fail(6, "scan: no fast kernel for this query-block / mode");
}

template <int D, int QB, int MODE>
static int launch_dense_one(const ScanParams& p, int grid, cudaStream_t st) {
  const size_t smem = scan_dense_smem<D, QB>(p.kp);
  if (smem > 227 * 1024) return fail(6, "scan: candidate pools do not fit shared memory");
  if (p.gather != nullptr)
    scan_dense_kernel<D, QB, MODE, true><<<grid, dense_nw<D, QB>() * 32, smem, st>>>(p);
  else
    scan_dense_kernel<D, QB, MODE, false><<<grid, dense_nw<D, QB>() * 32, smem, st>>>(p);
  QG_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int D, int QB>
static int launch_dense_mode(int mode, const ScanParams& p, int grid, cudaStream_t st) {
  if constexpr (QB <= scan_fast_qb_limit<D>()) {
    if (mode == MODE_L2) return launch_dense_one<D, QB, MODE_L2>(p, grid, st);
    if (mode == MODE_DOT) return launch_dense_one<D, QB, MODE_DOT>(p, grid, st);
  }
  return fail(6, "scan: no dense kernel for this query-block / mode");
}

template <int D>
int launch_scan_dense_d(int qb, int mode, const ScanParams& p, int grid, cudaStream_t st) {
  switch (qb) {
    case 1: return launch_dense_mode<D, 1>(mode, p, grid, st);
    case 2: return launch_dense_mode<D, 2>(mode, p, grid, st);
    case 4: return launch_dense_mode<D, 4>(mode, p, grid, st);
    case 8: return launch_dense_mode<D, 8>(mode, p, grid, st);
    default: return fail(1, "scan: query block must be 1, 2, 4 or 8");
  }
}

template <int D, int QB>
static int dense_geom_qb(int kp, int* tile_rows, int* warps, size_t* smem) {
  if constexpr (QB <= scan_fast_qb_limit<D>()) {
    constexpr int NW = dense_nw<D, QB>();
    *tile_rows = DenseGeom<D, NW>::RT;
    *warps = NW;
    *smem = scan_dense_smem<D, QB>(kp);
    return 0;
  }
  return -1;
}

template <int D>
int scan_dense_geom_d(int qb, int kp, int* tile_rows, int* warps, size_t* smem) {
  switch (qb) {
    case 1: return dense_geom_qb<D, 1>(kp, tile_rows, warps, smem);
    case 2: return dense_geom_qb<D, 2>(kp, tile_rows, warps, smem);
    case 4: return dense_geom_qb<D, 4>(kp, tile_rows, warps, smem);
    case 8: return dense_geom_qb<D, 8>(kp, tile_rows, warps, smem);
    default: return -1;
  }
}

template <int D>
int launch_scan_fast_d(int qb, int mode, const ScanParams& p, int grid, cudaStream_t st) {
  switch (qb) {
    case 1: return launch_mode<D, 1>(mode, p, grid, st);
    case 2: return launch_mode<D, 2>(mode, p, grid, st);
    case 4: return launch_mode<D, 4>(mode, p, grid, st);
    case 8: return launch_mode<D, 8>(mode, p, grid, st);
    default: return fail(1, "scan: query block must be 1, 2, 4 or 8");
  }
}

template <int D, int QB>
static int attr_qb() {
  if constexpr (QB <= scan_fast_qb_limit<D>()) {
    if (int e = attr_one<D, QB, MODE_L2>()) return e;
    if (int e = attr_one<D, QB, MODE_DOT>()) return e;
    QG_CUDA_OK(cudaFuncSetAttribute(scan_dense_kernel<D, QB, MODE_L2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    227 * 1024));
    QG_CUDA_OK(cudaFuncSetAttribute(scan_dense_kernel<D, QB, MODE_DOT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    227 * 1024));
    QG_CUDA_OK(cudaFuncSetAttribute(scan_dense_kernel<D, QB, MODE_L2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    227 * 1024));
    QG_CUDA_OK(cudaFuncSetAttribute(scan_dense_kernel<D, QB, MODE_DOT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    227 * 1024));
  }
  return 0;
}

template <int D>
int scan_fast_attr_d() {
  if (int e = attr_qb<D, 1>()) return e;
  if (int e = attr_qb<D, 2>()) return e;
  if (int e = attr_qb<D, 4>()) return e;
  if (int e = attr_qb<D, 8>()) return e;
  return 0;
}

}  // namespace qg

#define QG_DEFINE_SCAN_FAST(DIM)                                                                      \
  namespace qg {                                                                                      \
  int launch_scan_fast_##DIM(int qb, int mode, const ScanParams& p, int grid, cudaStream_t st) {      \
    return launch_scan_fast_d<DIM>(qb, mode, p, grid, st);                                            \
  }                                                                                                   \
  int scan_fast_attr_##DIM() { return scan_fast_attr_d<DIM>(); }                                      \
  int launch_scan_dense_##DIM(int qb, int mode, const ScanParams& p, int grid, cudaStream_t st) {     \
    return launch_scan_dense_d<DIM>(qb, mode, p, grid, st);                                           \
  }                                                                                                   \
  int scan_dense_geom_##DIM(int qb, int kp, int* tile_rows, int* warps, size_t* smem) {               \
    return scan_dense_geom_d<DIM>(qb, kp, tile_rows, warps, smem);                                    \
  }                                                                                                   \
  int scan_fast_tile_rows_##DIM() { return ScanGeom<DIM>::RT; }                                       \
  int scan_fast_max_qb_##DIM() { return scan_fast_qb_limit<DIM>(); }                                  \
  int scan_fast_ring_##DIM() { return SCAN_NW * ScanGeom<DIM>::STAGES * ScanGeom<DIM>::TILE_BYTES; }  \
  }
