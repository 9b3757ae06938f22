// spkdiar_gw.cu - translation unit of the growing-window search (K3): kernels of gw.cuh, host side abi_gw.inc.
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <limits>
#include <new>
#include <vector>

#include "common.cuh"
#include "score.cuh"
#include "gw.cuh"

using namespace spk;

#include "abi_gw.inc"
