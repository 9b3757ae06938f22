// spkdiar_cluster.cu - translation unit of the clustering engines (K5-K8): cluster.cuh, cluster_batch.cuh, host side abi_cluster.inc / abi_batch.inc.
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <limits>
#include <new>
#include <vector>

#include "common.cuh"
#include "score.cuh"
#include "cluster.cuh"
#include "cluster_small.cuh"
#include "cluster_batch.cuh"
#include "cluster_inorder.cuh"

using namespace spk;

#include "abi_cluster.inc"
#include "abi_batch.inc"
#include "abi_inorder.inc"
