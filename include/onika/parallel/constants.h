// onika/parallel/constants.h
#pragma once
namespace onika { namespace parallel {
  // kept for source compatibility: scheduling hint of host-executed (user-declared host-only) functors
  enum OMPScheduling { OMP_SCHED_DYNAMIC , OMP_SCHED_GUIDED , OMP_SCHED_STATIC };
  static inline constexpr int DEFAULT_EXECUTION_LANE   = -1;   // == ONK_LANE_DEFAULT
  static inline constexpr int UNDEFINED_EXECUTION_LANE = -2;   // == ONK_LANE_ALL when used as a filter
  static inline constexpr int MAX_EXECUTION_LANES      = 256;  // == ONK_MAX_LANES
} }
