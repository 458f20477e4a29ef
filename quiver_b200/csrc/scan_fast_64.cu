#include "scan_fast_inst.cuh"
QG_DEFINE_SCAN_FAST(64)
