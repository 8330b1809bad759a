// prof.cuh — optional per-kernel-class CUDA-event timing inside the library (bench.py's roofline leg).
// Off by default: when off, a ProfScope is two predictable branches.  When on (satmvs_profile_begin),
// every launch site records an event pair on the launching stream; satmvs_profile_end sums them.
#pragma once
#include <vector>
#include "common.cuh"

namespace satmvs {

enum ProfPhase { kProfSweep = 0, kProfConvBatched, kProfGruGate, kProfGruOutput, kProfGruPointwise, kProfDecoder,
                 kProfCostReg, kProfHead, kProfFeature, kProfCount };

struct ProfState {
  bool on = false;
  std::vector<cudaEvent_t> ev[kProfCount];   // start/stop pairs
};
ProfState& prof_state();

struct ProfScope {
  cudaStream_t st; int phase; bool on;
  ProfScope(int phase_, cudaStream_t st_) : st(st_), phase(phase_), on(prof_state().on) {
    if (on) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); prof_state().ev[phase].push_back(e); }
  }
  ~ProfScope() {
    if (on) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); prof_state().ev[phase].push_back(e); }
  }
};

}  // namespace satmvs
