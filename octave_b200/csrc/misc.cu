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

// Statistics contract: by default every entry point that ACCUMULATES into a double-precision statistics output (stats of
// the conv forward, sums / sums2 of the BatchNorm passes, chan_sum of space-to-depth) zeroes it first with its own memset
// node.  A host that hands out slices of one pre-zeroed arena (octave_b200/ops.py) switches those ~250 memsets per step off.
int g_octave_stats_prezeroed = 0;
extern "C" void octave_set_stats_prezeroed(int on) { g_octave_stats_prezeroed = on ? 1 : 0; }
extern "C" int octave_get_stats_prezeroed(void) { return g_octave_stats_prezeroed; }

// Identity of the stream capture `stream` is part of (0: not capturing): an arena zeroed outside a capture must not be
// handed out inside it, nor one capture's arena inside the next.
extern "C" unsigned long long octave_stream_capture_id(void* stream) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  unsigned long long id = 0;
  if (cudaStreamGetCaptureInfo((cudaStream_t)stream, &st, &id) != cudaSuccess) return 0;
  return st == cudaStreamCaptureStatusActive ? id : 0;
}
