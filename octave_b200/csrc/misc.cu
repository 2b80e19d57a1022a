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

// Deterministic mode: every reduction that is normally split across CTAs and merged with fp32 atomics (split-K weight
// gradients, the K-split of the attention-branch linears) runs with a single writer per output element instead.
int g_octave_deterministic = 0;
extern "C" void octave_set_deterministic(int on) { g_octave_deterministic = on ? 1 : 0; }
extern "C" int octave_get_deterministic(void) { return g_octave_deterministic; }
