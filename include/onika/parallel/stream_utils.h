// onika/parallel/stream_utils.h -- placeholder kept for include compatibility.
#pragma once
