// Host-side sampling draws for one SMC step, bit-compatible with numpy.random.Generator.
//
// The particle loop of the reference draws, per non-retained particle and in swarm order
// (smc/kernels/semi_adapted.py:63-83, 143-183):
//     u = rng.random()
//     u < 0.5 : rng.multinomial(1, q).argmax()                      -> existing-node candidate
//     else    : k = rng.integers(0, R + 1); rng.choice(roots, k, replace=False)   -> new node
// ~5000 times per PG iteration; the Python call overhead of those Generator methods is a quarter of an
// iteration.  This file performs the same draws from the SAME bit generator (numpy's bitgen_t, reached
// through `Generator.bit_generator.ctypes.bit_generator`) with numpy's own distribution code
// (libnpyrandom.a: random_standard_uniform, random_multinomial, random_bounded_uint64[_fill]) plus a
// restatement of Generator.choice's sampling without replacement (Floyd's algorithm + final shuffle).
// Stream identity with the Python methods is asserted by tests/test_host_draws.py.
#include <cstdint>
#include <cstring>
#include <vector>

#include "numpy/random/bitgen.h"
#include "numpy/random/distributions.h"

namespace {

inline uint64_t gen_mask(uint64_t max_val) {
  uint64_t mask = max_val;
  mask |= mask >> 1;
  mask |= mask >> 2;
  mask |= mask >> 4;
  mask |= mask >> 8;
  mask |= mask >> 16;
  mask |= mask >> 32;
  return mask;
}

// Generator.choice(a, size, replace=False, p=None, shuffle=True) for pop_size <= 10000: indices into `a`.
void choice_no_replace(bitgen_t* bg, int64_t pop, int64_t size, int64_t* idx, std::vector<uint64_t>& hash_set) {
  uint64_t set_size = (uint64_t)(1.2 * (double)size);
  const uint64_t mask = gen_mask(set_size);
  set_size = 1 + mask;
  hash_set.assign(set_size, (uint64_t)-1);
  for (int64_t j = pop - size; j < pop; ++j) {
    const uint64_t val = random_bounded_uint64(bg, 0, (uint64_t)j, 0, 0);
    uint64_t loc = val & mask;
    while (hash_set[loc] != (uint64_t)-1 && hash_set[loc] != val) loc = (loc + 1) & mask;
    if (hash_set[loc] == (uint64_t)-1) {  // val not drawn yet
      hash_set[loc] = val;
      idx[j - pop + size] = (int64_t)val;
    } else {  // insert j instead
      loc = (uint64_t)j & mask;
      while (hash_set[loc] != (uint64_t)-1) loc = (loc + 1) & mask;
      hash_set[loc] = (uint64_t)j;
      idx[j - pop + size] = j;
    }
  }
  for (int64_t i = size - 1; i >= 1; --i) {  // _shuffle_int(bitgen, size, 1, idx)
    const int64_t j = (int64_t)random_bounded_uint64(bg, 0, (uint64_t)i, 0, 0);
    const int64_t t = idx[i];
    idx[i] = idx[j];
    idx[j] = t;
  }
}

}  // namespace

extern "C" {

int pcb_host_version(void) { return 1; }

// Draws of `n` particles.  Particle i uses proposal d = dist_of[i] whose existing-node probabilities are
// q_flat[q_off[d] .. q_off[d+1]) and whose parent has roots roots_flat[roots_off[d] .. roots_off[d+1]);
// empty[d] != 0 marks a proposal whose parent has no node (no coin flip, always a multinomial draw).
// out_kind[i]: 0 existing (out_index[i] = candidate index), 1 new node (children in
// out_children[out_child_off[i] .. out_child_off[i+1]) in draw order).  Returns 0, or -1 if a population
// exceeds the range this restatement of Generator.choice covers (callers then fall back to Python).
int pcb_smc_draws(void* bitgen, int64_t n, const int32_t* dist_of, const double* q_flat, const int64_t* q_off,
                  const int64_t* roots_flat, const int64_t* roots_off, const uint8_t* empty, int32_t* out_kind,
                  int64_t* out_index, int64_t* out_children, int64_t* out_child_off) {
  bitgen_t* bg = static_cast<bitgen_t*>(bitgen);
  binomial_t binomial;
  std::memset(&binomial, 0, sizeof(binomial));
  std::vector<RAND_INT_TYPE> counts;
  std::vector<double> pix;
  std::vector<uint64_t> hash_set;
  std::vector<int64_t> idx;
  int64_t nchild = 0;
  out_child_off[0] = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int d = dist_of[i];
    const int64_t qn = q_off[d + 1] - q_off[d];
    const int64_t R = roots_off[d + 1] - roots_off[d];
    if (R > 10000) return -1;
    bool existing = true;
    if (!empty[d]) existing = random_standard_uniform(bg) < 0.5;
    if (existing) {
      counts.assign((size_t)qn, 0);
      pix.assign(q_flat + q_off[d], q_flat + q_off[d + 1]);
      random_multinomial(bg, 1, counts.data(), pix.data(), (npy_intp)qn, &binomial);
      int64_t arg = 0;
      for (int64_t k = 0; k < qn; ++k)
        if (counts[k] > counts[arg]) arg = k;
      out_kind[i] = 0;
      out_index[i] = arg;
    } else {
      uint64_t k = 0;
      random_bounded_uint64_fill(bg, 0, (uint64_t)R, 1, false, &k);  // rng.integers(0, R + 1)
      out_kind[i] = 1;
      out_index[i] = (int64_t)k;
      if (k > 0) {
        idx.resize((size_t)k);
        choice_no_replace(bg, R, (int64_t)k, idx.data(), hash_set);
        const int64_t* roots = roots_flat + roots_off[d];
        for (uint64_t c = 0; c < k; ++c) out_children[nchild++] = roots[idx[c]];
      }
    }
    out_child_off[i + 1] = nchild;
  }
  return 0;
}

// Stand-alone wrappers used by the stream-identity tests.
double pcb_draw_random(void* bitgen) { return random_standard_uniform(static_cast<bitgen_t*>(bitgen)); }

int64_t pcb_draw_integers(void* bitgen, int64_t high_exclusive) {
  uint64_t k = 0;
  random_bounded_uint64_fill(static_cast<bitgen_t*>(bitgen), 0, (uint64_t)(high_exclusive - 1), 1, false, &k);
  return (int64_t)k;
}

void pcb_draw_choice(void* bitgen, int64_t pop, int64_t size, int64_t* idx) {
  std::vector<uint64_t> hash_set;
  choice_no_replace(static_cast<bitgen_t*>(bitgen), pop, size, idx, hash_set);
}

void pcb_draw_multinomial(void* bitgen, int64_t n, const double* p, int64_t d, int64_t* out) {
  binomial_t binomial;
  std::memset(&binomial, 0, sizeof(binomial));
  std::vector<RAND_INT_TYPE> counts((size_t)d, 0);
  std::vector<double> pix(p, p + d);
  random_multinomial(static_cast<bitgen_t*>(bitgen), (RAND_INT_TYPE)n, counts.data(), pix.data(), (npy_intp)d, &binomial);
  for (int64_t k = 0; k < d; ++k) out[k] = (int64_t)counts[k];
}

}  // extern "C"
