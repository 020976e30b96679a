// onika/cuda/memmove.h -- see onika/cuda/memcpy.h (cu_block_memmove lives there)
#pragma once
#include <onika/cuda/memcpy.h>
