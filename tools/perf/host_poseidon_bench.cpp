// Latency of one Poseidon-12 permutation on one host core, per variant of csrc/host_poseidon.cpp (the Fiat-Shamir
// transcript is a chain of up to ~49 k of them per proof).  Minimum over many short batches: the boxes are shared.
//   g++ -O3 -std=c++17 -I starky_bls12_381_b200/csrc tools/perf/host_poseidon_bench.cpp -o tools/perf/bin/host_poseidon_bench
#include "host_poseidon.cpp"
#include <stdio.h>
#include <time.h>

static double now_ns() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return 1e9 * t.tv_sec + t.tv_nsec; }

int main() {
  u64 s[12];
  for (int i = 0; i < 12; i++) s[i] = i * 0x123456789ULL;
  const char* names[6] = {"0 scalar", "1 AVX2", "2 AVX-512 dense", "3 AVX-512 + sparse partial rounds", "4 hybrid IFMA/BMI2, look-ahead 1",
                          "5 hybrid IFMA/BMI2, look-ahead 2"};
  for (int pass = 0; pass < 2; pass++)
    for (int v = 0; v < 6; v++) {
      if (!sb_host_poseidon_permute_variant(s, v)) { printf("variant %s: not supported by this CPU\n", names[v]); continue; }
      double best = 1e30;
      for (int rep = 0; rep < 300; rep++) {
        const double t0 = now_ns();
        for (int i = 0; i < 500; i++) sb_host_poseidon_permute_variant(s, v);
        const double dt = (now_ns() - t0) / 500;
        if (dt < best) best = dt;
      }
      printf("variant %-40s %8.1f ns per permutation (min of 300 batches of 500)\n", names[v], best);
    }
  return (int)(s[0] & 1);
}
