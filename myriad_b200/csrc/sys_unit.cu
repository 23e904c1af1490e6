// One translation unit per system: compile with -DMYR_SYS_CLASS=SysCartpole (see myriad_b200/build.py).
#include "kernels.cuh"

#ifndef MYR_SYS_CLASS
#error "compile with -DMYR_SYS_CLASS=<generated system struct>"
#endif

#define MYR_CAT2(a, b) a##b
#define MYR_CAT(a, b) MYR_CAT2(a, b)

#ifdef MYR_NODE
// NodeSystem wrapper around the true system (node_mlp.cuh)
extern "C" const myr::SysVTable* MYR_CAT(myr_vtable_node_, MYR_SYS_CLASS)(void) {
  static const myr::SysVTable vt = myr::make_vtable<myr::SysNode<myr::MYR_SYS_CLASS>>();
  return &vt;
}
#else
extern "C" const myr::SysVTable* MYR_CAT(myr_vtable_, MYR_SYS_CLASS)(void) {
  static const myr::SysVTable vt = myr::make_vtable<myr::MYR_SYS_CLASS>();
  return &vt;
}
#endif
