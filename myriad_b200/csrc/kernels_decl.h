// Per-system function table shared by api.cu (dispatch) and sys_unit.cu (definitions).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/myriad_b200.h"

namespace myr {
struct SysVTable {
  int id;
  const char* name;
  int (*sizes)(const MyrDesc*, MyrSizes*);
  int (*eval)(const MyrDesc*, int, const double*, const double*, double*, double*, double*, double*, double*, void*);
  int (*host_eval)(const MyrDesc*, int, const double*, const double*, double*, double*, double*, double*, double*);
  int (*kkt)(const MyrDesc*, int, const double*, const double*, const double*, const double*, const double*, double, double, double*, double*,
             int32_t*, double*, size_t, void*);
  int (*host_kkt)(const MyrDesc*, int, const double*, const double*, const double*, const double*, const double*, double, double, double*,
                  double*, int32_t*, double*, size_t);
  int (*ipm)(const MyrDesc*, const MyrIpmOpts*, int, const double*, const double*, const double*, double*, double*, double*, double*, double*,
             double*, double*, int32_t*, int32_t*, double*, size_t, void*);
  int (*host_ipm)(const MyrDesc*, const MyrIpmOpts*, int, const double*, const double*, const double*, double*, double*, double*, double*,
                  double*, double*, double*, int32_t*, int32_t*, double*, size_t);
  int (*rollout)(const MyrDesc*, int, int, const double*, const double*, double*, double*, void*);
  int (*host_rollout)(const MyrDesc*, int, int, const double*, const double*, double*, double*);
  int (*dynamics)(const MyrDesc*, int, const double*, const double*, const double*, double*, double*, void*);
  int (*host_dynamics)(const MyrDesc*, int, const double*, const double*, const double*, double*, double*);
  int (*jtvec)(const MyrDesc*, int, const double*, const double*, double*, void*);
  int (*host_jtvec)(const MyrDesc*, int, const double*, const double*, double*);
};
}  // namespace myr
