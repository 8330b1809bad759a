// prof.cuh — optional per-kernel-class CUDA-event timing inside the library (bench.py's roofline leg).
// Off by default: when off, a ProfScope is two predictable branches.  When on (satmvs_profile_begin),
// every launch site records an event pair on the launching stream; satmvs_profile_end sums them.
#pragma once
#include <mutex>
#include <vector>
#include "common.cuh"

namespace satmvs {

enum ProfPhase { kProfSweep = 0, kProfConvBatched, kProfGruGate, kProfGruOutput, kProfGruPointwise, kProfDecoder,
                 kProfCostReg, kProfHead, kProfFeature, kProfTrainConv, kProfTrainWgrad, kProfTrainNorm, kProfCount };

// One profiler per process (autograd runs the backward launches on its own thread, and they belong to the same step);
// scopes of concurrent host threads interleave safely: each scope owns its (start, stop) slot pair.
struct ProfState {
  bool on = false;
  unsigned generation = 0;
  std::mutex mu;
  std::vector<cudaEvent_t> ev[kProfCount];   // start/stop pairs
};
ProfState& prof_state();

struct ProfScope {
  cudaStream_t st; int phase; bool on; unsigned gen; size_t slot;
  ProfScope(int phase_, cudaStream_t st_) : st(st_), phase(phase_), on(prof_state().on), gen(0), slot(0) {
    if (on) {
      cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st);
      ProfState& s = prof_state();
      std::lock_guard<std::mutex> lk(s.mu);
      gen = s.generation; slot = s.ev[phase].size();
      s.ev[phase].push_back(e); s.ev[phase].push_back(nullptr);
    }
  }
  ~ProfScope() {
    if (on) {
      cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st);
      ProfState& s = prof_state();
      std::lock_guard<std::mutex> lk(s.mu);
      if (s.generation == gen && slot + 1 < s.ev[phase].size()) s.ev[phase][slot + 1] = e; else cudaEventDestroy(e);
    }
  }
};

}  // namespace satmvs
