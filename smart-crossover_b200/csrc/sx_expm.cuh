// sx_expm.cuh -- exp(x) for x <= 0, the only exponential the log-sum-exp inner loops of the Sinkhorn
// warm start need (every argument is "value minus running maximum").
//
// The streaming passes are bound by the fp64 pipe, so the exponential is built for few fp64 operations
// (11, against ~25 for libm's): knowing the sign removes the overflow path, and treating x <= -700 as 0
// (e^-700 ~ 1e-304, far below half an ulp of a sum that holds the maximum's exp(0) = 1) removes the
// denormal path, so the power of two is always a normal number assembled with integer adds.
//   k = rint(x 32 / ln2)          by the 1.5 * 2^52 rounding constant;  k = 32 e + j, 0 <= j < 32
//   r = x - k ln2 / 32            in two fused steps (ln2 split hi / lo, k ln2_hi / 32 exact for |k| < 2^21)
//   e^r                           Taylor to degree 6 on |r| <= ln2 / 64: truncation 3.5e-18 relative
//   e^x = e^r 2^(j/32) 2^e        2^(j/32) from a 32-entry table (shared memory on the device)
// The range test and the table index are integer operations (other pipes).  NaN goes through.
// Host-compilable so tests bound its error against libm without a GPU (tests/test_host_mirror.py).
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define SX_HD __host__ __device__ __forceinline__
#else
#define SX_HD inline
#endif

namespace sx {

constexpr int kExpTabSize = 32;
// 2^(j/32), j = 0..31, correctly rounded
#define SX_EXP_TAB_VALUES                                                                       \
    1, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237,                              \
    1.0905077326652577, 1.1143867425958924, 1.1387886347566916, 1.1637248587775775,             \
    1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332,                \
    1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,              \
    1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228,             \
    1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.6457554781539649,             \
    1.681792830507429, 1.7186192981224779, 1.7562521603732995, 1.7947090750031072,              \
    1.8340080864093424, 1.8741676341103, 1.9152065613971474, 1.9571441241754002

SX_HD void expm_split(double t, int &hi, int &lo) {
#ifdef __CUDA_ARCH__
    hi = __double2hiint(t);
    lo = __double2loint(t);
#else
    uint64_t u;
    memcpy(&u, &t, 8);
    hi = (int)(uint32_t)(u >> 32);
    lo = (int)(uint32_t)u;
#endif
}

SX_HD double expm_join(int hi, int lo) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(hi, lo);
#else
    const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}

SX_HD double expm_fma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return fma(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

// `tab` = the 32 values above (shared memory on the device).
SX_HD double exp_nonpos(double x, const double *tab) {
    const double kMagic = 6755399441055744.0;       // 1.5 * 2^52
    const double t = expm_fma(x, 46.16624130844682838, kMagic);          // 32 / ln2
    int t_hi, k, x_hi, x_lo;
    expm_split(t, t_hi, k);
    expm_split(x, x_hi, x_lo);
    const double kd = t - kMagic;
    double r = expm_fma(kd, -2.16608493865351192653e-02, x);             // ln2_hi / 32
    r = expm_fma(kd, -5.96317165397058656257e-12, r);                    // ln2_lo / 32
    double p = 1.3888888888888889419e-03;           // 1 / 6!
    p = expm_fma(p, r, 8.3333333333333332177e-03);  // 1 / 5!
    p = expm_fma(p, r, 4.1666666666666664354e-02);  // 1 / 4!
    p = expm_fma(p, r, 1.6666666666666665741e-01);  // 1 / 3!
    p = expm_fma(p, r, 0.5);
    p = expm_fma(p, r, 1.0);
    p = expm_fma(p, r, 1.0);
    int s_hi, s_lo;
    expm_split(tab[k & (kExpTabSize - 1)], s_hi, s_lo);
    const double scale = expm_join(s_hi + ((k >> 5) << 20), s_lo);       // 2^(j/32) 2^e, e >= -1010: normal
    // x <= -700 (as unsigned high words: more negative = larger): below anything a sum with a 1 can see
    return (unsigned)x_hi >= 0xC085E000u ? 0.0 : p * scale;
}

}  // namespace sx
