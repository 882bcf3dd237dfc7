/* Oracle C restatement of the label-noise sampler (TEST INFRASTRUCTURE ONLY --
 * see oracle/__init__.py; never linked into the product library).
 *
 * Follows numpy's legacy RandomState algorithms as reached from
 *   mnist/model.py:795-834   (seed 547, shuffle, per-sample multinomial/randint/multinomial)
 *   cifar10/common/data/cifar10.py:29-38
 * i.e. MT19937 init_genrand, random_double (53-bit from two words), masked-rejection
 * random_interval, multinomial(1,p) -> binomial(1, p_j/rem) by inversion.
 * Pinned against numpy itself in tests/test_oracle_sampler.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct { uint32_t key[624]; int pos; } mt_t;

static void mt_seed(mt_t *s, uint32_t seed) {
  for (int pos = 0; pos < 624; pos++) {
    s->key[pos] = seed;
    seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)pos + 1u;
  }
  s->pos = 624;
}
static void mt_gen(mt_t *s) {
  uint32_t *k = s->key;
  for (int i = 0; i < 624; i++) {
    uint32_t y = (k[i] & 0x80000000u) | (k[(i + 1) % 624] & 0x7fffffffu);
    k[i] = k[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  }
  s->pos = 0;
}
static uint32_t mt_next32(mt_t *s) {
  if (s->pos == 624) mt_gen(s);
  uint32_t y = s->key[s->pos++];
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
  return y;
}
static double mt_double(mt_t *s) {
  uint32_t a = mt_next32(s) >> 5, b = mt_next32(s) >> 6;
  return (a * 67108864.0 + b) / 9007199254740992.0;
}
static uint32_t mt_interval(mt_t *s, uint32_t mx) {
  if (mx == 0) return 0;
  uint32_t mask = mx, v;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  while ((v = (mt_next32(s) & mask)) > mx) {}
  return v;
}
static long inversion(mt_t *s, long n, double p) {
  double q = 1.0 - p, qn = exp(n * log(q));
  double np_ = n * p, bound = fmin((double)n, np_ + 10.0 * sqrt(np_ * q + 1));
  long X = 0; double px = qn, U = mt_double(s);
  while (U > px) {
    X++;
    if (X > bound) { X = 0; px = qn; U = mt_double(s); }
    else { U -= px; px = ((n - X + 1) * p * px) / (X * q); }
  }
  return X;
}
static long binomial(mt_t *s, long n, double p) {
  if (n == 0 || p == 0.0) return 0;
  if (p <= 0.5) return inversion(s, n, p);
  return n - inversion(s, n, 1.0 - p);
}
static int multinomial1(mt_t *s, const double *pv, int d) {
  double Sum = 1.0; long dn = 1;
  for (int j = 0; j < d - 1; j++) {
    long x = binomial(s, dn, pv[j] / Sum);
    dn -= x;
    if (dn <= 0) return j;
    Sum -= pv[j];
  }
  return d - 1;
}

int oracle_mnist_labels(uint32_t seed, int n, int shuffle, int real_match, const double *C,
                        int32_t *y, int32_t *perm, int32_t *real, int32_t *gen, int32_t *fake) {
  mt_t s; mt_seed(&s, seed);
  for (int i = 0; i < n; i++) perm[i] = i;
  if (shuffle) {
    for (int i = n - 1; i > 0; i--) { uint32_t j = mt_interval(&s, (uint32_t)i); int32_t t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
    mt_seed(&s, seed);
    for (int i = n - 1; i > 0; i--) { uint32_t j = mt_interval(&s, (uint32_t)i); int32_t t = y[i]; y[i] = y[j]; y[j] = t; }
  }
  for (int i = 0; i < n; i++) {
    int r = multinomial1(&s, C + 10 * y[i], 10);
    int g = (int)mt_interval(&s, 9);
    if (real_match) g = r;
    int f = multinomial1(&s, C + 10 * g, 10);
    real[i] = r; gen[i] = g; fake[i] = f;
  }
  return 0;
}

int oracle_cifar_labels(uint32_t seed, int n, const double *C, int32_t *labels, int32_t *rnd, int32_t *biased) {
  mt_t s; mt_seed(&s, seed);
  for (int i = 0; i < n; i++) rnd[i] = (int32_t)mt_interval(&s, 9);
  for (int i = 0; i < n; i++) {
    labels[i] = multinomial1(&s, C + 10 * labels[i], 10);
    biased[i] = multinomial1(&s, C + 10 * rnd[i], 10);
  }
  return 0;
}
