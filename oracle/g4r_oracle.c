/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the integer / bit-exact pieces of the path:
 *   k2_searchsorted : the control flow of the reference's GpuBinarySearchSorted kernel (custom_theano_ops.py:318-349)
 *   mrg31k3p_fill   : MRG31k3p as used by theano.sandbox.rng_mrg (third party; constants checked in
 *                     tests/test_host_logic.py::test_mrg_constants_self_consistency), stream i -> samples i, i+n_streams, ...
 * Built by oracle/Makefile into oracle/libg4r_oracle.so; only tests/ may load it (cross-check of the NumPy oracle). */
#include <stdint.h>

void k2_searchsorted(const float* d, long long ld, const float* x, long long n, long long* y) {
  for (long long i = 0; i < n; i++) {
    long long a = 0, b = ld - 1;
    const float minval = d[0], maxval = d[ld - 1], val = x[i];
    if (val > maxval) { a = ld; b = ld; }
    else if (val <= minval) { a = 0; b = 0; }
    while (b - a > 0) {
      const long long h = (b + a) / 2;
      const float t = d[h];
      if (val < t) b = h; else a = h + 1;
    }
    y[i] = b;
  }
}

#define M1 2147483647
#define M2 2147462579
#define MASK12 511
#define MASK13 16777215
#define MASK2 65535
#define MULT2 21069

/* int32 arithmetic exactly as in Theano's mrg_uniform C code (overflow wraps; corrected by the "< 0" tests) */
void mrg31k3p_fill(int32_t* state /* [n_streams][6], updated */, long long n_streams, float* out, long long n) {
  for (long long i = 0; i < n; i++) {
    int32_t* s = state + (i % n_streams) * 6;
    int32_t x11 = s[0], x12 = s[1], x13 = s[2], x21 = s[3], x22 = s[4], x23 = s[5];
    int32_t y1, y2;
    y1 = (int32_t)(((uint32_t)(x12 & MASK12) << 22) + (uint32_t)(x12 >> 9) + ((uint32_t)(x13 & MASK13) << 7) + (uint32_t)(x13 >> 24));
    if (y1 < 0 || y1 >= M1) y1 -= M1;
    y1 = (int32_t)((uint32_t)y1 + (uint32_t)x13);
    if (y1 < 0 || y1 >= M1) y1 -= M1;
    x13 = x12; x12 = x11; x11 = y1;
    y1 = (int32_t)(((uint32_t)(x21 & MASK2) << 15) + (uint32_t)MULT2 * (uint32_t)(x21 >> 16));
    if (y1 < 0 || y1 >= M2) y1 -= M2;
    y2 = (int32_t)(((uint32_t)(x23 & MASK2) << 15) + (uint32_t)MULT2 * (uint32_t)(x23 >> 16));
    if (y2 < 0 || y2 >= M2) y2 -= M2;
    y2 = (int32_t)((uint32_t)y2 + (uint32_t)x23);
    if (y2 < 0 || y2 >= M2) y2 -= M2;
    y2 = (int32_t)((uint32_t)y2 + (uint32_t)y1);
    if (y2 < 0 || y2 >= M2) y2 -= M2;
    x23 = x22; x22 = x21; x21 = y2;
    out[i] = (x11 <= x21) ? (float)(x11 - x21 + M1) * 4.6566126e-10f : (float)(x11 - x21) * 4.6566126e-10f;
    s[0] = x11; s[1] = x12; s[2] = x13; s[3] = x21; s[4] = x22; s[5] = x23;
  }
}
