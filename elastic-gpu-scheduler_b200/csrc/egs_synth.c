/*
 * egs_synth.c -- deterministic synthetic clusters and pod batches (SURVEY.md 8d,
 * BASELINE.json configs 0..4).  Harness code shared by bench.py and the tests;
 * plain C, no CUDA, built as libegs_synth.so.
 *
 * PRNG: splitmix64; u(m) = next() % m.  Cluster seed 0xB2000000+cfg, pod seed
 * 0xE6A50000+cfg.
 */
#include <stdint.h>

typedef struct { int32_t core, mem, count; } synth_unit; /* == egs_unit */

static uint64_t next(uint64_t *s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static uint32_t u(uint64_t *s, uint32_t m) { return (uint32_t)(next(s) % m); }

/* configs: nodes, gpus per node, memory per GPU, pods, policy (0 binpack, 1 spread) */
int egs_synth_config(int cfg, int *n_nodes, int *gpus, int *mem_total, int *n_pods, int *policy) {
  static const int N[5] = {4, 1000, 10000, 50000, 100000};
  static const int P[5] = {8, 10000, 100000, 500000, 1000000};
  static const int POL[5] = {0, 0, 1, 1, 0};
  if (cfg < 0 || cfg > 4) return -1;
  *n_nodes = N[cfg]; *gpus = cfg == 0 ? 2 : 8; *mem_total = cfg == 0 ? 16 : 81920;
  *n_pods = P[cfg]; *policy = POL[cfg];
  return 0;
}

/* free rows [n_nodes][gpus]: cfg 0 all free; else per GPU: u(100)<50 -> free,
 * otherwise core = 100 - 5*u(21), mem = M - 1024*u(81). */
void egs_synth_cluster(int cfg, int n_nodes, int gpus, int mem_total, int32_t *core, int32_t *mem) {
  uint64_t s = 0xB2000000ull + (uint64_t)cfg;
  for (long i = 0; i < (long)n_nodes * gpus; i++) {
    if (cfg == 0 || u(&s, 100) < 50) { core[i] = 100; mem[i] = mem_total; continue; }
    core[i] = 100 - 5 * (int32_t)u(&s, 21);
    mem[i] = mem_total - 1024 * (int32_t)u(&s, 81);
  }
}

/* c_off[n_pods+1], units[<= 3*n_pods]; returns total containers */
int egs_synth_pods(int cfg, int n_pods, int32_t *c_off, synth_unit *units) {
  uint64_t s = 0xE6A50000ull + (uint64_t)cfg;
  static const int KA0[8] = {4, 8, 4, 12, 16, 8, 4, 8};
  static const int C1[5] = {5, 10, 20, 25, 50};
  static const int C2[4] = {10, 20, 30, 50};
  static const int M2[5] = {1024, 2048, 4096, 8192, 16384};
  static const int T3[4][2] = {{10, 1024}, {20, 4096}, {30, 8192}, {50, 16384}};
  static const int C4[4] = {0, 10, 25, 50};
  static const int M4[4] = {4096, 8192, 16384, 40960};
  int k = 0;
  for (int p = 0; p < n_pods; p++) {
    c_off[p] = k;
    switch (cfg) {
      case 0: units[k].core = 0; units[k].mem = KA0[p % 8]; units[k].count = 0; k++; break;
      case 1: units[k].core = C1[u(&s, 5)]; units[k].mem = 0; units[k].count = 0; k++; break;
      case 2: units[k].core = C2[u(&s, 4)]; units[k].mem = M2[u(&s, 5)]; units[k].count = 0; k++; break;
      case 3: {
        int c = 2 + (int)u(&s, 2);
        for (int j = 0; j < c; j++) {
          int t = (int)u(&s, 4);
          units[k].core = T3[t][0]; units[k].mem = T3[t][1]; units[k].count = 0; k++;
        }
        break;
      }
      default: units[k].core = C4[u(&s, 4)]; units[k].mem = M4[u(&s, 4)]; units[k].count = 0; k++; break;
    }
  }
  c_off[n_pods] = k;
  return k;
}
