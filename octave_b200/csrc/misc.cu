// ABI version + device queries.
#include "common.cuh"
#include "../../include/octave_b200.h"

extern "C" int octave_abi_version(void) { return OCTAVE_ABI_VERSION; }

extern "C" int octave_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return OCT_ERR_LAUNCH;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return OCT_ERR_LAUNCH;
  return n;
}

unsigned long long g_octave_launches = 0;
extern "C" unsigned long long octave_launch_count(void) { return g_octave_launches; }
