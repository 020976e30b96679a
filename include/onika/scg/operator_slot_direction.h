// onika/scg/operator_slot_direction.h
#pragma once
namespace onika { namespace scg {
  enum SlotDirection { INPUT = 0 , OUTPUT = 1 , INPUT_OUTPUT = 2 };
  static inline constexpr SlotDirection PRIVATE = INPUT_OUTPUT;
} }
