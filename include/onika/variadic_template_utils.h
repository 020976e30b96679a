// onika/variadic_template_utils.h -- kept for include compatibility (fold expressions replace the old helper macros)
#pragma once
#define TEMPLATE_LIST_BEGIN ( ... , (
#define TEMPLATE_LIST_END ) );
