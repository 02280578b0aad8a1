// exp_glibc.cuh - double-precision exp with the bits of glibc 2.39's exp() (FMA variant) for arguments <= 0.
//
// The DSS density features (dss.cpp:217-244, 339-372) sum exp(-d/20) in double precision; a sum that differs in its last bit can
// move a value across a bin threshold and flip a feature letter, which changes alignments.  glibc's exp is not correctly
// rounded, so "the same bits" means the same algorithm: x = k*ln2/128 + r, 2^(k/128) from a 128-entry table (value + tail),
// degree-5 polynomial in r, operations fused exactly where the compiled __exp_fma fuses them (read from the disassembly of
// libm.so.6; see tools/extract_glibc_exp.py for the data and tools/check_exp_glibc.c for the bit-for-bit check against libm:
// every float-derived density argument on a fine grid + 2e8 random doubles, 0 mismatches).
// Only x <= 0 is needed (and supported): the overflow side of the special case is not restated.
#pragma once

#include <stdint.h>
#include <string.h>

#include "exp_glibc_data.inc"

#ifndef RSK_HOSTDEV
#define RSK_HOSTDEV __device__ __forceinline__
#define RSK_FMA(a, b, c) __fma_rn((a), (b), (c))
#define RSK_EXP_TABLE_SPACE __device__
#else
#define RSK_EXP_TABLE_SPACE static
#endif

RSK_EXP_TABLE_SPACE const unsigned long long rsk_exp_table[256] = RSK_EXP_TABLE_INIT;

#ifdef __CUDA_ARCH__
#define RSK_D2U(x) ((unsigned long long)__double_as_longlong(x))
#define RSK_U2D(u) __longlong_as_double((long long)(u))
#else
static inline unsigned long long rsk_d2u_(double x) { unsigned long long u; memcpy(&u, &x, 8); return u; }
static inline double rsk_u2d_(unsigned long long u) { double x; memcpy(&x, &u, 8); return x; }
#define RSK_D2U(x) rsk_d2u_(x)
#define RSK_U2D(u) rsk_u2d_(u)
#endif

RSK_HOSTDEV double rsk_exp_glibc(double x)
{
	const unsigned long long ix = RSK_D2U(x);
	unsigned abstop = (unsigned)(ix >> 52) & 0x7ffu;
	if (abstop - 0x3c9u >= 0x3fu) {
		if (abstop < 0x3c9u)
			return 1.0 + x;              // |x| < 2^-54
		if (abstop >= 0x409u)            // |x| >= 1024: exp(-inf) = 0, finite negative underflows to 0 (round to nearest)
			return 0.0;
		abstop = 0;                      // 512 <= |x| < 1024: the scale may leave the normal range
	}
	const double kd0 = RSK_FMA(x, RSK_EXP_INVLN2N, RSK_EXP_SHIFT);
	const unsigned long long ki = RSK_D2U(kd0);
	const double kd = kd0 - RSK_EXP_SHIFT;
	double r = RSK_FMA(kd, RSK_EXP_NEGLN2HIN, x);
	r = RSK_FMA(kd, RSK_EXP_NEGLN2LON, r);
	const unsigned idx = 2u * (unsigned)(ki & 127u);
	const unsigned long long top = ki << 45;
	const double p23 = RSK_FMA(RSK_EXP_C3, r, RSK_EXP_C2);
	const double tr = r + RSK_U2D(rsk_exp_table[idx]);
	unsigned long long sbits = rsk_exp_table[idx + 1] + top;
	const double r2 = r * r;
	const double p45 = RSK_FMA(r, RSK_EXP_C5, RSK_EXP_C4);
	const double t = RSK_FMA(p23, r2, tr);
	const double r4 = r2 * r2;
	const double tmp = RSK_FMA(r4, p45, t);
	if (abstop != 0) {
		const double scale = RSK_U2D(sbits);
		return RSK_FMA(scale, tmp, scale);
	}
	// specialcase(), k < 0 side: the result may be subnormal
	sbits += 0x3feull << 52;
	const double scale = RSK_U2D(sbits);
	double y = scale + scale * tmp;   // NOT fused in the compiled code (vmulsd + vaddsd)
	if (y < 1.0) {
		// round to the subnormal precision without double rounding
		const double lo = scale - y + scale * tmp;
		const double hi = 1.0 + y;
		const double lo2 = 1.0 - hi + y + lo;
		y = (hi + lo2) - 1.0;
		if (y == 0.0)
			y = 0.0;
	}
	return 0x1p-1022 * y;
}
