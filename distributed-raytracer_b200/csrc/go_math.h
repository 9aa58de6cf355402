// go_math.h -- host-side restatement of Go's math.Tan (Go stdlib math/tan.go, the pure-Go Cephes
// routine used on amd64), needed because tracer.go:17 evaluates math.Tan(cam.Fov / 2.0) and the
// result is a frame constant of the primary-ray arithmetic.  Product code (not the oracle's copy).
// Host translation units including this file are compiled with -ffp-contract=off (csrc/Makefile):
// Go never fuses a*b+c on amd64.
#pragma once
#include <math.h>
#include <stdint.h>

static inline double rt_go_tan(double x) {
    const double PI4A = 7.85398125648498535156E-1, PI4B = 3.77489470793079817668E-8, PI4C = 2.69515142907905952645E-15;
    const double P0 = -1.30936939181383777646E4, P1 = 1.15351664838587416140E6, P2 = -1.79565251976484877988E7;
    const double Q1 = 1.36812963470692954678E4, Q2 = -1.32089234440210967447E6, Q3 = 2.50083801823357915839E7,
                 Q4 = -5.38695755929454629881E7;
    if (x == 0 || x != x) return x;
    if (isinf(x)) return NAN;
    bool sign = false;
    if (x < 0) { x = -x; sign = true; }
    if (x >= 536870912.0) { double t = tan(x); return sign ? -t : t; }   // Payne-Hanek branch not restated
    uint64_t j = (uint64_t)(x * 1.27323954473516268615107010698011489627567716592365);   // Go constant 4/Pi
    double y = (double)j;
    if (j & 1) { j++; y++; }
    double z = ((x - y * PI4A) - y * PI4B) - y * PI4C;
    double zz = z * z;
    if (zz > 1e-14) y = z + z * (zz * (((P0 * zz) + P1) * zz + P2) / ((((zz + Q1) * zz + Q2) * zz + Q3) * zz + Q4));
    else y = z;
    if (j & 2) y = -1 / y;
    if (sign) y = -y;
    return y;
}
