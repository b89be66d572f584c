// prof.cu -- see prof.h
#include "prof.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace ace {
namespace prof {

bool on = false;

namespace {
struct Pair {
  const char* name;
  cudaEvent_t e0, e1;
};
std::vector<Pair>        g_pairs;
std::vector<cudaEvent_t> g_pool;

cudaEvent_t get_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

void enable(bool v) { on = v; }

int begin(const char* name, cudaStream_t s) {
  Pair p{name, get_event(), get_event()};
  cudaEventRecord(p.e0, s);
  g_pairs.push_back(p);
  return (int)g_pairs.size() - 1;
}

void end(int slot, cudaStream_t s) { cudaEventRecord(g_pairs[slot].e1, s); }

void reset() {
  cudaDeviceSynchronize();
  for (Pair& p : g_pairs) {
    g_pool.push_back(p.e0);
    g_pool.push_back(p.e1);
  }
  g_pairs.clear();
}

void report(const char* title) {
  cudaDeviceSynchronize();
  struct Acc { double ms = 0; size_t n = 0; };
  std::map<std::string, Acc> acc;
  for (Pair& p : g_pairs) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
      Acc& a = acc[p.name];
      a.ms += ms;
      a.n++;
    }
  }
  std::vector<std::pair<std::string, Acc>> v(acc.begin(), acc.end());
  std::sort(v.begin(), v.end(), [](auto& a, auto& b) { return a.second.ms > b.second.ms; });
  printf("[ace_b200 prof] %s: device time by scope (event pairs; api.* scopes contain the "
         "kernel scopes)\n", title);
  for (auto& kv : v)
    printf("[ace_b200 prof] %-28s %9zu scopes %10.3f ms %8.2f us/scope\n", kv.first.c_str(),
           kv.second.n, kv.second.ms, 1e3 * kv.second.ms / kv.second.n);
  fflush(stdout);
  reset();
}

}  // namespace prof
}  // namespace ace
