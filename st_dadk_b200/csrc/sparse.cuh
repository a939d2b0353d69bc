// Large-knot-count regime (K_s in the 1e5 range): support-walking kernels.  Filled in below.
#pragma once
#include "common.cuh"
