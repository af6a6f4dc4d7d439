/* rs_kernels.cu -- one of the six translation units that hold the rs_tti_kernel instantiations (rs_kernels.h);
 * -DRS_TU=1..6 picks which.  No host logic here: rs_sched.cu asks for kernel addresses and launches them. */
#include "rs_kernels.h"

#ifndef RS_TU
#error "compile with -DRS_TU=1..6"
#endif
#ifndef RS_NARROW_THREADS
#define RS_NARROW_THREADS 128
#endif
#ifndef RS_NARROW_MIN_BLOCKS
#define RS_NARROW_MIN_BLOCKS 8
#endif
#ifndef RS_WIDE_THREADS
#define RS_WIDE_THREADS 512
#define RS_WIDE_MIN_BLOCKS 2
#endif

#define RS_CORE_ONLY
#if RS_TU == 3 || RS_TU == 4
#define RS_NS rsw
#define RS_THREADS RS_WIDE_THREADS
#define RS_MIN_BLOCKS RS_WIDE_MIN_BLOCKS
#else
#define RS_NS rs
#define RS_THREADS RS_NARROW_THREADS
#define RS_MIN_BLOCKS RS_NARROW_MIN_BLOCKS
#endif
#include "rs_device.cuh"

#define RS_CAT2(a, b) a##b
#define RS_CAT(a, b) RS_CAT2(a, b)

#define RS_TTI_PICK(A)                                                                                              \
  (queue ? (trace ? (const void*)RS_NS::rs_tti_kernel<A, true, true> : (const void*)RS_NS::rs_tti_kernel<A, false, true>) \
         : (trace ? (const void*)RS_NS::rs_tti_kernel<A, true, false> : (const void*)RS_NS::rs_tti_kernel<A, false, false>))

#if RS_TU == 1 || RS_TU == 3
const void* RS_CAT(rs_kernel_tu, RS_TU)(int algo, bool trace, bool queue) {
  switch (algo) {
    case 1: return RS_TTI_PICK(1);
    case 8: return RS_TTI_PICK(8);
    case 10: return RS_TTI_PICK(10);
    default: return RS_TTI_PICK(9);
  }
}
#elif RS_TU == 2 || RS_TU == 4
const void* RS_CAT(rs_kernel_tu, RS_TU)(int algo, bool trace, bool queue) {
  switch (algo) {
    case 7: return RS_TTI_PICK(7);
    case 11: return RS_TTI_PICK(11);
    case 101: return RS_TTI_PICK(101);
    default: return RS_TTI_PICK(103);
  }
}
#else
using FixedU8 = rs::FixedShape<20, 5, 64, 8, 0>;
using FixedNib = rs::FixedShape<20, 5, 64, 8, 2>;
#define RS_FIXED_PICK(A)                                                                                                   \
  (which == 0 ? (trace ? (const void*)rs::rs_tti_kernel<A, true, false, FixedU8> : (const void*)rs::rs_tti_kernel<A, false, false, FixedU8>) \
              : (trace ? (const void*)rs::rs_tti_kernel<A, true, false, FixedNib> : (const void*)rs::rs_tti_kernel<A, false, false, FixedNib>))
#if RS_TU == 5
const void* rs_kernel_tu5(int algo, int which, bool trace) {
  switch (algo) {
    case 8: return RS_FIXED_PICK(8);
    case 10: return RS_FIXED_PICK(10);
    case 101: return RS_FIXED_PICK(101);
    case 103: return RS_FIXED_PICK(103);
    default: return RS_FIXED_PICK(9);
  }
}
#else
const void* rs_kernel_tu6(int algo, int which, bool trace) {
  if (algo == 11) return RS_FIXED_PICK(11);
  return RS_FIXED_PICK(7);
}
#endif
#endif

cudaError_t RS_CAT(rs_tables_tu, RS_TU)(const void* ct) { return cudaMemcpyToSymbol(RS_NS::c_tab, ct, sizeof(RS_NS::ConstTables)); }
