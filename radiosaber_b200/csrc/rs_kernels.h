/* rs_kernels.h -- the TTI kernels are compiled in six translation units (rs_kernels.cu with -DRS_TU=1..6, built in
 * parallel: one unit took minutes); each hands out the addresses of its instantiations and fills its own copy of the
 * constant tables. */
#pragma once
#include <cuda_runtime.h>

/* 128 threads per cell, any shape: TU 1 = ids 9, 8, 10, 1; TU 2 = ids 7, 11, 101, 103.  512 threads per cell: TU 3 / TU 4,
 * same split.  TU 5 and TU 6: the compile-time-shape instantiations of the headline cell (which: 0 = one CQI byte per RBG,
 * 1 = two RBGs per byte). */
const void* rs_kernel_tu1(int algo, bool trace, bool queue);
const void* rs_kernel_tu2(int algo, bool trace, bool queue);
const void* rs_kernel_tu3(int algo, bool trace, bool queue);
const void* rs_kernel_tu4(int algo, bool trace, bool queue);
const void* rs_kernel_tu5(int algo, int which, bool trace);   /* ids 9, 8, 10, 101, 103 */
const void* rs_kernel_tu6(int algo, int which, bool trace);   /* ids 7, 11 */
/* ct: an rs::ConstTables (same struct in every unit) */
cudaError_t rs_tables_tu1(const void* ct);
cudaError_t rs_tables_tu2(const void* ct);
cudaError_t rs_tables_tu3(const void* ct);
cudaError_t rs_tables_tu4(const void* ct);
cudaError_t rs_tables_tu5(const void* ct);
cudaError_t rs_tables_tu6(const void* ct);
