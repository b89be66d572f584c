// prof.h -- optional device-time accounting by kernel family (ACE_B200_PROF=1).
//
// Every launch wrapper opens a prof::Scope; when profiling is on the scope records a CUDA event
// pair on the launching stream around the launches it covers.  The pairs are resolved once, in
// report(), so the run itself is not synchronised: the numbers are the device time of the real
// (warm-cache, back-to-back) execution, which ncu's serialised replay cannot give.  Off by
// default: one predictable branch per launch.
#pragma once
#include <cuda_runtime.h>

namespace ace {
namespace prof {

extern bool on;
void enable(bool v);
int  begin(const char* name, cudaStream_t s);  // returns a slot for end()
void end(int slot, cudaStream_t s);
void report(const char* title);                // synchronises, prints, resets
void reset();

struct Scope {
  cudaStream_t s;
  int          slot;
  Scope(const char* name, cudaStream_t st) : s(st), slot(on ? begin(name, st) : -1) {}
  ~Scope() {
    if (slot >= 0) end(slot, s);
  }
};

}  // namespace prof
}  // namespace ace
